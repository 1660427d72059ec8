"""``TwoTower`` and ``TwoTowerTrainTask`` with the constructor / forward contract of
/root/reference/utils/model_training.py:79-143, plus the two extensions
BASELINE.json names: explicit per-tower feature lists (multi-feature towers as in
ray_tune_optuna_tuning_alex_test.py:227-306) and the in-batch softmax loss.

Defaults reproduce the reference: exactly two tables of equal dim, query tower
reads table 0's features, candidate tower table 1's, loss =
``BCEWithLogitsLoss((q*c).sum(1), labels.float())``, forward returns
``(loss, (loss.detach(), logits.detach(), labels.detach()))``.
"""
from typing import List, Optional, Tuple

import torch
from torch import nn

from .datasets.utils import Batch
from .functional import FusedTowersTC, dot_bce_loss, fused_towers_supported, in_batch_softmax_loss
from .modules.embedding_modules import EmbeddingBagCollection
from .modules.mlp import MLP
from .sparse.jagged_tensor import KeyedJaggedTensor, KeyedTensor


class TwoTower(nn.Module):
    def __init__(self, embedding_bag_collection: EmbeddingBagCollection, layer_sizes,
                 device: Optional[torch.device] = None, query_features: Optional[List[str]] = None,
                 candidate_features: Optional[List[str]] = None, precision: str = "fp32",
                 dense_index: Optional[int] = None, dense_dim: int = 0) -> None:
        """``layer_sizes``: one list for both towers or ``[query_layers, candidate_layers]``;
        ``dense_index`` / ``dense_dim``: the Ray-Tune variant of the reference
        (ray_tune_optuna_tuning_alex_test.py:227-306) concatenates ``batch.dense_features[:, :dense_index]`` to the
        query tower's input and ``[:, dense_index:dense_dim]`` to the candidate tower's; call ``forward(batch)`` then."""
        super().__init__()
        cfgs = embedding_bag_collection.embedding_bag_configs()
        if query_features is None and candidate_features is None:
            assert len(cfgs) == 2, "Expected two EmbeddingBags in the two tower model"
            assert cfgs[0].embedding_dim == cfgs[1].embedding_dim, "Both EmbeddingBagConfigs must have the same dimension"
            query_features = list(cfgs[0].feature_names)
            candidate_features = list(cfgs[1].feature_names)
        elif query_features is None or candidate_features is None:
            raise ValueError("give both query_features and candidate_features, or neither")
        dim_of = {f: c.embedding_dim for c in cfgs for f in c.feature_names}
        self._feature_names_query: List[str] = list(query_features)
        self._candidate_feature_names: List[str] = list(candidate_features)
        self.ebc = embedding_bag_collection
        per_tower = any(isinstance(x, (list, tuple)) for x in layer_sizes)
        q_layers, c_layers = (list(layer_sizes[0]), list(layer_sizes[1])) if per_tower else (list(layer_sizes), list(layer_sizes))
        self.dense_index = dense_index
        self._dense_dim = dense_dim if dense_index is not None else 0
        q_dense = dense_index if dense_index is not None else 0
        c_dense = self._dense_dim - q_dense
        if dense_index is not None and (q_dense < 0 or c_dense < 0):
            raise ValueError("dense_index must lie inside [0, dense_dim]")
        self.query_proj = MLP(in_size=sum(dim_of[f] for f in self._feature_names_query) + q_dense,
                              layer_sizes=q_layers, device=device, precision=precision)
        self.candidate_proj = MLP(in_size=sum(dim_of[f] for f in self._candidate_feature_names) + c_dense,
                                  layer_sizes=c_layers, device=device, precision=precision)

    @staticmethod
    def _tower_input(pooled: KeyedTensor, features: List[str]) -> torch.Tensor:
        # Adjacent features are one column window of the pooled matrix: no copy,
        # the GEMM reads it with the pooled row pitch.
        col, width = pooled.columns(features)
        if col >= 0:
            return pooled.values().narrow(1, col, width)
        return torch.cat([pooled[f] for f in features], dim=1)

    def _fused_plan(self, pooled: KeyedTensor):
        """(column offsets, in_dim, parameters) when both towers fit the one-launch fused kernels, else None."""
        towers = (self.query_proj, self.candidate_proj)
        if self.dense_index is not None:
            return None
        if any(m.precision != "bf16" or not m.all_relu() for m in towers) or pooled.values().shape[0] == 0:
            return None
        layers = [list(m._mlp) for m in towers]
        wins = [pooled.columns(f) for f in (self._feature_names_query, self._candidate_feature_names)]
        dims = [(l[0]._in_size, l[0]._out_size, l[-1]._out_size) for l in layers]
        v = pooled.values()
        if (any(c < 0 for c, _ in wins) or dims[0] != dims[1] or not fused_towers_supported([w for _, w in wins], dims[0][1], dims[0][2], len(layers[0]))
                or len(layers[1]) != 2 or any(c % 4 for c, _ in wins) or v.stride(0) % 4 or v.stride(1) != 1 or v.data_ptr() % 16):
            return None
        params = []
        for l in layers:
            for pc in l:
                params += [pc._linear.weight, pc._linear.bias]
        return [c for c, _ in wins], wins[0][1], params

    def forward(self, kjt) -> Tuple[torch.Tensor, torch.Tensor]:
        """``kjt``: the KeyedJaggedTensor (reference, utils/model_training.py:99) or the whole ``Batch`` (Ray-Tune variant)."""
        batch = None
        if isinstance(kjt, Batch):
            batch, kjt = kjt, kjt.sparse_features
        pooled_embeddings = self.ebc(kjt)
        if self.dense_index is not None:
            if batch is None:
                raise ValueError("this TwoTower concatenates dense features: call it with the Batch, not the KeyedJaggedTensor")
            dense = batch.dense_features.to(torch.float32)
            q_in = torch.cat([self._tower_input(pooled_embeddings, self._feature_names_query), dense[:, :self.dense_index]], dim=1)
            c_in = torch.cat([self._tower_input(pooled_embeddings, self._candidate_feature_names),
                              dense[:, self.dense_index:self._dense_dim]], dim=1)
            return self.query_proj(q_in), self.candidate_proj(c_in)
        plan = self._fused_plan(pooled_embeddings)
        if plan is not None:
            cols, in_dim, params = plan
            pv = pooled_embeddings.values()
            q, c, yb = FusedTowersTC.apply(pv, tuple(cols), in_dim, getattr(pv, "_tt_grad_dst", None), *params)
            # the towers' bf16 copies of their outputs ride along: the tensor-core loss takes them instead of casting again
            q._tt_bf16, q._tt_bf16_version = yb[0], q._version
            c._tt_bf16, c._tt_bf16_version = yb[1], c._version
            return q, c
        query_embedding = self.query_proj(self._tower_input(pooled_embeddings, self._feature_names_query))
        candidate_embedding = self.candidate_proj(self._tower_input(pooled_embeddings, self._candidate_feature_names))
        return query_embedding, candidate_embedding


class TwoTowerTrainTask(nn.Module):
    """``loss="bce"`` is the reference (utils/model_training.py:129-140).
    ``loss="in_batch_softmax"`` treats every other candidate of the batch as a
    negative: ``CE(q c^T / temperature, arange(B))``; ``logits`` returned in that
    mode are the positive-pair logits (the diagonal)."""

    def __init__(self, two_tower: TwoTower, loss: str = "bce", temperature: float = 1.0,
                 precision: str = "fp32", negatives: str = "local", pg=None) -> None:
        super().__init__()
        if loss not in ("bce", "in_batch_softmax"):
            raise ValueError(f"unknown loss {loss}")
        if precision not in ("fp32", "bf16"):
            raise ValueError(f"unknown precision {precision}")
        if negatives not in ("local", "global"):
            raise ValueError(f"unknown negatives {negatives}")
        # in-batch softmax inside a process group: "local" = per-rank negatives (no collective); "global" = the
        # candidates of every rank (all-gather + reduce-scatter), i.e. the single-GPU loss of the global batch
        self.negatives, self._pg = negatives, pg
        self.precision = precision  # "bf16": logits GEMM + softmax on tcgen05 (bf16 operands, fp32 accumulate)
        self.two_tower = two_tower
        self.loss_kind = loss
        self.temperature = temperature
        # the reference's attribute (utils/model_training.py:129); while it is the plain mean BCE-with-logits the fused kernel
        # computes it, a loss module put in its place (pos_weight, a weighted BCE as ray_tune_optuna_tuning_alex_test.py:309-330)
        # is CALLED on the row-wise dot products, as the reference's forward does (utils/model_training.py:136-140)
        self.loss_fn: nn.Module = nn.BCEWithLogitsLoss()

    def _loss_fn_is_plain_bce(self) -> bool:
        f = self.loss_fn
        return (type(f) is nn.BCEWithLogitsLoss and f.weight is None and f.pos_weight is None and f.reduction == "mean")

    def forward(self, batch: Batch) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        uses_dense = getattr(self.two_tower, "dense_index", None) is not None
        query_embedding, candidate_embedding = self.two_tower(batch if uses_dense else batch.sparse_features)
        if self.loss_kind == "bce" and not self._loss_fn_is_plain_bce():
            logits = (query_embedding * candidate_embedding).sum(dim=1).squeeze()
            loss = self.loss_fn(logits, batch.labels.float())
        elif self.loss_kind == "bce":
            loss, logits = dot_bce_loss(query_embedding, candidate_embedding, batch.labels)
        else:
            loss, logits = in_batch_softmax_loss(query_embedding, candidate_embedding, self.temperature, self.precision,
                                                 self.negatives, self._pg)
        return loss, (loss.detach(), logits.detach(), batch.labels.detach())
