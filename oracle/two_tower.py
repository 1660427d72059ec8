"""Oracle: the whole unsharded two-tower train step on the CPU, built from stock
``torch`` ops -- per-table ``nn.EmbeddingBag(include_last_offset=True)`` +
``relu(Linear)`` towers + loss + row-wise Adagrad applied to the dense embedding
gradient + ``torch.optim.Adam`` on the tower parameters.  This is the set of
ops unsharded TorchRec runs on CPU for
/root/reference/03_model_training.py:770-829 with ``device="cpu"`` and
/root/reference/utils/model_training.py:79-143, so it doubles as the CPU
baseline that ``bench.py`` times.  Test infrastructure only.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn
import torch.nn.functional as F

from .ebc import TableSpec, output_layout, rowwise_adagrad_dense, rowwise_adam_sparse
from .kjt import lengths_to_offsets


class OracleTwoTower(nn.Module):
    def __init__(
        self, tables: Sequence[TableSpec], layer_sizes: Sequence[int],
        loss: str = "bce", sparse_optimizer: str = "rowwise_adagrad",
        sparse_lr: float = 1e-2, sparse_eps: Optional[float] = None,
        dense_lr: float = 1e-2, temperature: float = 1.0,
        query_features: Optional[List[str]] = None,
        candidate_features: Optional[List[str]] = None, seed: int = 0,
        dense_optimizer: str = "adam", dense_index: Optional[int] = None, dense_dim: int = 0,
        column_blocks: Optional[Dict[str, int]] = None,
    ) -> None:
        """``layer_sizes`` may be ``[user_layers, item_layers]`` and ``dense_index`` / ``dense_dim`` concatenate
        ``dense[:, :dense_index]`` / ``dense[:, dense_index:dense_dim]`` to the tower inputs, as the Ray-Tune variant of
        the reference does (/root/reference/ray_tune_optuna_tuning_alex_test.py:227-306)."""
        super().__init__()
        self.dense_index, self.dense_dim = dense_index, dense_dim
        # column-wise sharded tables (TorchRec: each column shard is its own fused table): row-wise Adagrad runs per block of
        # D / n columns, each block with its own per-row accumulator -- {table name: number of column shards}
        self.column_blocks = dict(column_blocks or {})
        self.column_state: Dict[str, torch.Tensor] = {}
        self.tables = list(tables)
        self.loss_kind = loss
        self.sparse_optimizer = sparse_optimizer
        self.sparse_lr = sparse_lr
        self.sparse_eps = sparse_eps if sparse_eps is not None else (1e-10 if sparse_optimizer == "rowwise_adagrad" else 1e-8)
        self.temperature = temperature
        if query_features is None:
            # utils/model_training.py:88-93 -- exactly two tables, same dim
            assert len(self.tables) == 2, "Expected two EmbeddingBags in the two tower model"
            assert self.tables[0].embedding_dim == self.tables[1].embedding_dim
            query_features = list(self.tables[0].feature_names)
            candidate_features = list(self.tables[1].feature_names)
        self.query_features = query_features
        self.candidate_features = candidate_features
        g = torch.Generator().manual_seed(seed)
        self.embedding_bags = nn.ModuleDict()
        for t in self.tables:
            eb = nn.EmbeddingBag(t.num_embeddings, t.embedding_dim, mode=t.pooling, include_last_offset=True)
            bound = (1.0 / t.num_embeddings) ** 0.5  # TorchRec EmbeddingBagConfig default init
            with torch.no_grad():
                eb.weight.copy_((torch.rand(eb.weight.shape, generator=g) * 2 - 1) * bound)
            self.embedding_bags[t.name] = eb
        keys, dims = output_layout(self.tables)
        dim_of = dict(zip(keys, dims))
        q_in = sum(dim_of[f] for f in self.query_features)
        c_in = sum(dim_of[f] for f in self.candidate_features)
        per_tower = any(isinstance(x, (list, tuple)) for x in layer_sizes)
        q_layers, c_layers = (layer_sizes[0], layer_sizes[1]) if per_tower else (layer_sizes, layer_sizes)
        if dense_index is not None:
            q_in, c_in = q_in + dense_index, c_in + (dense_dim - dense_index)
        self.query_proj = self._make_mlp(q_in, q_layers, g)
        self.candidate_proj = self._make_mlp(c_in, c_layers, g)
        self.sparse_state: Dict[str, Dict[str, torch.Tensor]] = {}
        for t in self.tables:
            st = {"sum": torch.zeros(t.num_embeddings)}
            if sparse_optimizer == "rowwise_adam":
                st = {"m": torch.zeros(t.num_embeddings, t.embedding_dim), "v": torch.zeros(t.num_embeddings)}
            self.sparse_state[t.name] = st
        self.step_count = 0
        dense_params = list(self.query_proj.parameters()) + list(self.candidate_proj.parameters())
        # Adam is the reference's choice (03_model_training.py:826-829).  "sgd" exists for
        # parity tests of the in-batch softmax: there the last-layer bias gradient of any
        # always-active unit is structurally zero (softmax rows sum to one), so it is pure
        # rounding noise, and Adam's g/sqrt(v) turns that noise into +-lr steps that no two
        # implementations (or summation orders) agree on.
        self.dense_opt = (torch.optim.Adam(dense_params, lr=dense_lr) if dense_optimizer == "adam"
                          else torch.optim.SGD(dense_params, lr=dense_lr))

    @staticmethod
    def _make_mlp(in_size: int, layer_sizes: Sequence[int], g: torch.Generator) -> nn.ModuleList:
        layers = nn.ModuleList()
        for out in layer_sizes:
            lin = nn.Linear(in_size, out)
            bound = 1.0 / in_size ** 0.5
            with torch.no_grad():
                lin.weight.copy_((torch.rand(lin.weight.shape, generator=g) * 2 - 1) * bound)
                lin.bias.copy_((torch.rand(lin.bias.shape, generator=g) * 2 - 1) * bound)
            layers.append(lin)
            in_size = out
        return layers

    # ---- TorchRec-named state dict (N03:1030,1143: ``two_tower.`` prefix stripped by the caller)
    def torchrec_state_dict(self) -> Dict[str, torch.Tensor]:
        sd = {}
        for t in self.tables:
            sd[f"ebc.embedding_bags.{t.name}.weight"] = self.embedding_bags[t.name].weight.detach().clone()
        for tower, mods in (("query_proj", self.query_proj), ("candidate_proj", self.candidate_proj)):
            for i, lin in enumerate(mods):
                sd[f"{tower}._mlp.{i}._linear.weight"] = lin.weight.detach().clone()
                sd[f"{tower}._mlp.{i}._linear.bias"] = lin.bias.detach().clone()
        return sd

    def load_torchrec_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        with torch.no_grad():
            for t in self.tables:
                self.embedding_bags[t.name].weight.copy_(sd[f"ebc.embedding_bags.{t.name}.weight"])
            for tower, mods in (("query_proj", self.query_proj), ("candidate_proj", self.candidate_proj)):
                for i, lin in enumerate(mods):
                    lin.weight.copy_(sd[f"{tower}._mlp.{i}._linear.weight"])
                    lin.bias.copy_(sd[f"{tower}._mlp.{i}._linear.bias"])

    # ---- forward
    def pooled(self, keys: Sequence[str], values: torch.Tensor, lengths: torch.Tensor) -> Dict[str, torch.Tensor]:
        Fk = len(keys)
        B = lengths.numel() // Fk
        offsets = lengths_to_offsets(lengths).to(torch.int64)
        out = {}
        for t in self.tables:
            for feat in t.feature_names:
                f = list(keys).index(feat)
                s, e = int(offsets[f * B]), int(offsets[(f + 1) * B])
                out[feat] = self.embedding_bags[t.name](values[s:e], offsets[f * B:(f + 1) * B + 1] - s)
        return out

    def towers(self, pooled: Dict[str, torch.Tensor], dense: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        q = torch.cat([pooled[f] for f in self.query_features], dim=1)
        c = torch.cat([pooled[f] for f in self.candidate_features], dim=1)
        if self.dense_index is not None:
            q = torch.cat([q, dense[:, :self.dense_index].float()], dim=1)
            c = torch.cat([c, dense[:, self.dense_index:self.dense_dim].float()], dim=1)
        for lin in self.query_proj:
            q = torch.relu(lin(q))
        for lin in self.candidate_proj:
            c = torch.relu(lin(c))
        return q, c

    def forward(self, keys, values, lengths, dense: Optional[torch.Tensor] = None):
        return self.towers(self.pooled(keys, values, lengths), dense)

    def loss(self, q: torch.Tensor, c: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.loss_kind == "bce":
            logits = (q * c).sum(dim=1).squeeze()
            return F.binary_cross_entropy_with_logits(logits, labels.float()), logits
        if q.shape[0] > 16384:
            # same loss / gradients, row-chunked: the [B, B] logits (17 GB at B = 65 536) never exist at once
            from .dense import in_batch_softmax_loss_chunked
            return in_batch_softmax_loss_chunked(q, c, self.temperature)
        s = (q @ c.t()) / self.temperature
        return F.cross_entropy(s, torch.arange(q.shape[0])), s.diagonal()

    def _rowwise_adagrad(self, t: TableSpec, w: torch.Tensor, g: torch.Tensor) -> None:
        n = self.column_blocks.get(t.name, 1)
        if n == 1:
            rowwise_adagrad_dense(w, self.sparse_state[t.name]["sum"], g, lr=self.sparse_lr, eps=self.sparse_eps)
            return
        st = self.column_state.setdefault(t.name, torch.zeros(t.num_embeddings, n))
        dw = t.embedding_dim // n
        for j in range(n):
            blk, acc = w[:, j * dw:(j + 1) * dw].clone(), st[:, j].clone()
            rowwise_adagrad_dense(blk, acc, g[:, j * dw:(j + 1) * dw], lr=self.sparse_lr, eps=self.sparse_eps)
            w[:, j * dw:(j + 1) * dw] = blk
            st[:, j] = acc

    # ---- one data-parallel step as W ranks would do it: every rank's loss is the mean over ITS
    # batch; tower gradients are AVERAGED (DDP); embedding gradients of all ranks meet in the
    # (model-parallel) tables and are divided by W as well -- TorchRec's pooled all-to-all /
    # reduce-scatter backward divides by the world size (comm_ops GRADIENT_DIVISION, default on) --
    # i.e. both see the gradient of (1/W) * sum_r loss_r; then both optimizers step once.
    # ``gradient_division=False`` reproduces set_gradient_division(False): embedding gradients summed.
    # ``negatives="global"`` (softmax loss): rank r's queries see the candidates of EVERY rank,
    # loss_r = CE(q_r @ cat(c_0..c_{W-1})^T / T, r*B + arange(B)) -- the gradient w.r.t. another rank's
    # candidates flows back into the shared towers / tables (all-gather forward, reduce-scatter backward).
    def train_step_ranks(self, keys, batches, gradient_division: bool = True, negatives: str = "local") -> List[torch.Tensor]:
        self.dense_opt.zero_grad(set_to_none=True)
        for eb in self.embedding_bags.values():
            eb.weight.grad = None
        losses = []
        if negatives == "global":
            assert self.loss_kind != "bce"
            qs, cs = zip(*[self.forward(keys, v, l) for v, l, _ in batches])
            c_all = torch.cat(cs)
            total = 0
            for r, q in enumerate(qs):
                B = q.shape[0]
                loss = F.cross_entropy((q @ c_all.t()) / self.temperature, r * B + torch.arange(B))
                losses.append(loss.detach())
                total = total + loss
            total.backward()
        for values, lengths, labels in (batches if negatives != "global" else []):
            q, c = self.forward(keys, values, lengths)
            loss, _ = self.loss(q, c, labels)
            loss.backward()
            losses.append(loss.detach())
        W = len(batches)
        self.step_count += 1
        with torch.no_grad():
            for group in self.dense_opt.param_groups:
                for p in group["params"]:
                    if p.grad is not None:
                        p.grad.div_(W)
            for t in self.tables:
                w = self.embedding_bags[t.name].weight
                if w.grad is None:
                    continue
                assert self.sparse_optimizer == "rowwise_adagrad"
                g = w.grad / W if gradient_division else w.grad
                self._rowwise_adagrad(t, w, g)
                w.grad = None
        self.dense_opt.step()
        return losses

    # ---- one train step: fwd, bwd, row-wise sparse update "in backward", dense Adam
    def train_step(self, keys, values, lengths, labels) -> Tuple[torch.Tensor, torch.Tensor]:
        self.dense_opt.zero_grad(set_to_none=True)
        for eb in self.embedding_bags.values():
            eb.weight.grad = None
        q, c = self.forward(keys, values, lengths)
        loss, logits = self.loss(q, c, labels)
        loss.backward()
        self.step_count += 1
        Fk = len(keys)
        B = lengths.numel() // Fk
        offsets = lengths_to_offsets(lengths).to(torch.int64)
        with torch.no_grad():
            for t in self.tables:
                w = self.embedding_bags[t.name].weight
                g = w.grad
                if g is None:
                    continue
                if self.sparse_optimizer == "rowwise_adagrad":
                    self._rowwise_adagrad(t, w, g)
                else:
                    ids = torch.cat([values[int(offsets[list(keys).index(f) * B]):int(offsets[(list(keys).index(f) + 1) * B])]
                                     for f in t.feature_names])
                    rows = torch.unique(ids, sorted=True)
                    st = self.sparse_state[t.name]
                    rowwise_adam_sparse(w, st["m"], st["v"], rows, g[rows], self.step_count, lr=self.sparse_lr, eps=self.sparse_eps)
                w.grad = None
        self.dense_opt.step()
        return loss.detach(), logits.detach()
