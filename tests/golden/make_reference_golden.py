"""Generates tests/golden/reference_train.npz by EXECUTING THE REFERENCE'S OWN CODE.

    python tests/golden/make_reference_golden.py        (build container only: reads /root/reference)

/root/reference/utils/model_training.py is compiled unchanged and its own bodies are run on CPU:
``transform_to_torchrec_batch`` (U:43-69), ``TwoTower`` (U:79-120), ``TwoTowerTrainTask`` (U:123-143), ``train`` (U:255-317)
and ``evaluate`` (U:191-253), wired as 03_model_training.py:770-829 wires them (EmbeddingBagCollection on two tables,
``apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": ...})``, ``KeyedOptimizerWrapper(..., Adam)``).

The file's imports are torchrec names; torchrec / fbgemm_gpu (requirements.txt:1-3: torchrec==0.7.0, fbgemm-gpu==0.7.0)
are not installed and not installable here.  They are replaced, for this script only, by STOCK-TORCH stand-ins of what
unsharded TorchRec executes on CPU -- NOT by this repo's package and NOT by oracle/:

    EmbeddingBagCollection   nn.ModuleDict of nn.EmbeddingBag(mode="sum", include_last_offset=True)    (torchrec does exactly this)
    MLP                      Sequential of Perceptron = activation(nn.Linear(x)), relu on EVERY layer    (torchrec/modules/mlp.py)
    KeyedJaggedTensor        holder of keys / values / lengths / offsets (from_lengths_sync = cumsum)
    RowWiseAdagrad           torch.optim.Optimizer restating torchrec/optim/rowwise_adagrad.py: state_sum [rows] += mean(g*g, dim 1);
                             param -= lr * g / (sqrt(state_sum) + eps), eps 1e-10, lr_decay 0, weight_decay 0
    apply_optimizer_in_backward   the REAL torch.distributed.optim._apply_optimizer_in_backward (the file imports it from torch)
    TrainPipelineSparseDist  progress(it): next batch; train: zero_grad, forward, backward, step; eval: forward; returns output[1]
    DistributedModelParallel wrapper with .module (world size 1: no sharding, no communication)
    KeyedOptimizerWrapper    builds the optimizer over the parameters it is given (tables whose optimizer is fused in the
                             backward have grad None at step time and are skipped by Adam)

The fixture stores the raw batches, what the reference's transform made of them, the seeded initial weights, and -- all
produced by the reference's bodies on stock torch -- the loss / logits of every training step, the final weights, the
row-wise Adagrad accumulators, the evaluate() average loss, and the item / user embeddings the reference's own
``create_keyed_jagged_tensor`` / ``process_embeddings`` (03_model_training.py:1056-1122) compute from the trained model, with
their exact top-100 by stock ``torch.sort``.  tests/test_oracle_golden.py holds oracle/ to it on CPU,
tests/test_gpu_zz_reference_golden.py holds the CUDA path to it on the GPU box (where /root/reference does not exist).
"""
import itertools
import os
import sys
import types
from dataclasses import dataclass
from functools import partial

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/utils/model_training.py"

CAT = ["user_id", "product_id"]
EMB = [97, 331]
DIM, LAYERS, B, LR, STEPS = 32, [64, 32], 192, 0.01, 4


# ----------------------------------------------------------------------------- stock-torch stand-ins for the torchrec names
class KeyedJaggedTensor:
    def __init__(self, keys, values, lengths):
        self._keys, self._values, self._lengths = list(keys), values, lengths
        self._offsets = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(lengths.to(torch.int64), 0)])

    @staticmethod
    def from_lengths_sync(keys, values, lengths, weights=None):
        return KeyedJaggedTensor(keys, values, lengths)

    def keys(self):
        return self._keys

    def values(self):
        return self._values

    def lengths(self):
        return self._lengths

    def offsets(self):
        return self._offsets

    def length_per_key(self):
        n = len(self._keys)
        stride = self._lengths.numel() // n
        return [int(self._lengths[k * stride:(k + 1) * stride].sum()) for k in range(n)]

    def to(self, device, non_blocking=False):
        return self


class KeyedTensor(dict):
    pass


@dataclass
class Batch:
    dense_features: torch.Tensor
    sparse_features: KeyedJaggedTensor
    labels: torch.Tensor

    def to(self, device, non_blocking=False):
        return self


@dataclass
class EmbeddingBagConfig:
    name: str
    embedding_dim: int
    num_embeddings: int
    feature_names: list


class EmbeddingBagCollection(nn.Module):
    def __init__(self, tables, device=None):
        super().__init__()
        self._configs = list(tables)
        self.embedding_bags = nn.ModuleDict({c.name: nn.EmbeddingBag(c.num_embeddings, c.embedding_dim, mode="sum", include_last_offset=True)
                                             for c in tables})

    def embedding_bag_configs(self):
        return self._configs

    def forward(self, kjt):
        n = len(kjt.keys())
        stride = kjt.lengths().numel() // n
        out = KeyedTensor()
        for c in self._configs:
            for f in c.feature_names:
                k = kjt.keys().index(f)
                off = kjt.offsets()[k * stride:(k + 1) * stride + 1]
                vals = kjt.values()[int(off[0]):int(off[-1])]
                out[f] = self.embedding_bags[c.name](vals, off - off[0])
        return out


class Perceptron(nn.Module):
    def __init__(self, in_size, out_size):
        super().__init__()
        self._linear = nn.Linear(in_size, out_size)

    def forward(self, x):
        return torch.relu(self._linear(x))


class MLP(nn.Module):
    def __init__(self, in_size, layer_sizes, device=None):
        super().__init__()
        self._mlp = nn.Sequential(*[Perceptron(layer_sizes[i - 1] if i else in_size, layer_sizes[i]) for i in range(len(layer_sizes))])

    def forward(self, x):
        return self._mlp(x)


class RowWiseAdagrad(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-2, lr_decay=0.0, weight_decay=0.0, initial_accumulator_value=0.0, eps=1e-10, **_):
        super().__init__(params, dict(lr=lr, lr_decay=lr_decay, weight_decay=weight_decay, eps=eps))
        for group in self.param_groups:
            for p in group["params"]:
                self.state[p]["step"] = 0
                self.state[p]["sum"] = torch.full((p.shape[0],), initial_accumulator_value, dtype=p.dtype)

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                st["step"] += 1
                clr = group["lr"] / (1 + (st["step"] - 1) * group["lr_decay"])
                g = p.grad
                st["sum"].add_((g * g).mean(dim=1))
                std = st["sum"].sqrt().add_(group["eps"])
                p.addcdiv_(g, std.unsqueeze(1), value=-clr)


class DistributedModelParallel(nn.Module):
    def __init__(self, module, device=None, plan=None, **_):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


class KeyedOptimizerWrapper:
    def __init__(self, params, optim_factory):
        self._optimizer = optim_factory(list(params.values()))
        self.param_groups = self._optimizer.param_groups

    def zero_grad(self, set_to_none=True):
        self._optimizer.zero_grad(set_to_none=set_to_none)

    def step(self):
        self._optimizer.step()


class TrainPipelineSparseDist:
    def __init__(self, model, optimizer, device):
        self._model, self._optimizer, self._device = model, optimizer, device
        self.trace = []                    # (loss, logits) of every progress() call

    def progress(self, it):
        batch = next(it)
        if self._model.training:
            self._optimizer.zero_grad()
        loss, out = self._model(batch)
        if self._model.training:
            loss.backward()
            self._optimizer.step()
        self.trace.append((out[0].clone(), out[1].clone()))
        return out


def install_standins():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("torchrec")
    mod("torchrec.distributed", TrainPipelineSparseDist=TrainPipelineSparseDist)
    mod("torchrec.distributed.model_parallel", DistributedModelParallel=DistributedModelParallel, get_default_sharders=lambda: [])
    mod("torchrec.inference")
    mod("torchrec.inference.state_dict_transform", state_dict_gather=None, state_dict_to_device=None)
    mod("torchrec.modules")
    mod("torchrec.modules.embedding_configs", EmbeddingBagConfig=EmbeddingBagConfig)
    mod("torchrec.modules.embedding_modules", EmbeddingBagCollection=EmbeddingBagCollection)
    mod("torchrec.modules.mlp", MLP=MLP)
    mod("torchrec.optim")
    mod("torchrec.optim.keyed", KeyedOptimizerWrapper=KeyedOptimizerWrapper)
    mod("torchrec.optim.rowwise_adagrad", RowWiseAdagrad=RowWiseAdagrad)
    mod("torchrec.sparse")
    mod("torchrec.sparse.jagged_tensor", KeyedJaggedTensor=KeyedJaggedTensor)
    mod("torchrec.datasets")
    mod("torchrec.datasets.utils", Batch=Batch)
    mod("torchrec.distributed.comm", get_local_size=lambda: 1)
    mod("torchrec.distributed.planner", EmbeddingShardingPlanner=None, Topology=None)
    mod("torchrec.distributed.planner.storage_reservations", HeuristicalStorageReservation=None)
    # out of scope: imported at the top of the file, never reached by the bodies run here
    mod("streaming", StreamingDataset=object, StreamingDataLoader=object)
    mod("torchmetrics", AUROC=lambda task="binary": types.SimpleNamespace(
        to=lambda d: types.SimpleNamespace(__call__=None), compute=lambda: torch.tensor(0.0)))
    if "tqdm" not in sys.modules:
        try:
            import tqdm  # noqa: F401
        except ImportError:
            mod("tqdm", tqdm=lambda *a, **k: types.SimpleNamespace(update=lambda n: None))


class _Auroc:
    """torchmetrics.AUROC stand-in: evaluate() only calls it and reads .compute().item(); the value is not stored."""

    def __init__(self, task="binary"):
        pass

    def to(self, device):
        return self

    def __call__(self, preds, labels):
        return None

    def compute(self):
        return torch.tensor(0.0)


def _functions_of(path, names, env):
    """Executes ONLY the named top-level function definitions of a notebook source (the rest of the file talks to Spark /
    MLflow / a Databricks workspace); returns the namespace."""
    import ast
    with open(path) as f:
        tree = ast.parse(f.read(), path)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert sorted(n.name for n in keep) == sorted(names), [n.name for n in keep]
    ns = dict(env)
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns


def main():
    install_standins()
    sys.modules["torchmetrics"].AUROC = _Auroc
    import torch.distributed as dist
    from torch.distributed.optim import _apply_optimizer_in_backward as apply_optimizer_in_backward
    ns = {"__name__": "reference_model_training", "itertools": itertools, "cat_cols": list(CAT)}
    with open(REF) as f:
        exec(compile(f.read(), REF, "exec"), ns)
    ref = types.SimpleNamespace(**ns)
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29631", rank=0, world_size=1)

    g = torch.Generator().manual_seed(20261018)
    raws = []
    for s in range(STEPS + 2):                       # STEPS training batches + 2 evaluation batches
        raws.append({"user_id": torch.randint(0, 2 * EMB[0], (B,), generator=g).tolist(),       # ids >= rows: the modulo; id 0: empty bag
                     "product_id": torch.randint(0, 2 * EMB[1], (B,), generator=g).tolist(),
                     "label": torch.randint(0, 2, (B,), generator=g).tolist()})

    # 03_model_training.py:770-829 with the reference's classes
    eb_configs = [EmbeddingBagConfig(name=f"t_{c}", embedding_dim=DIM, num_embeddings=EMB[i], feature_names=[c]) for i, c in enumerate(CAT)]
    ebc = EmbeddingBagCollection(tables=eb_configs, device=torch.device("meta"))
    two_tower = ref.TwoTower(embedding_bag_collection=ebc, layer_sizes=LAYERS, device=torch.device("cpu"))
    task = ref.TwoTowerTrainTask(two_tower)
    # seeded initial weights, stored in the fixture (TorchRec's table init range; tower matrices 3x the nn.Linear range so
    # that the logits are O(0.1 - 1) and the loss moves in its leading digits)
    with torch.no_grad():
        for name, p in two_tower.named_parameters():
            if "embedding_bags" in name:
                bound = 1.0 / (p.shape[0] ** 0.5)
            else:
                bound = (3.0 / p.shape[1] ** 0.5) if p.dim() == 2 else (1.0 / p.shape[0] ** 0.5)
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
    init = {k: v.detach().clone() for k, v in two_tower.state_dict().items()}
    apply_optimizer_in_backward(RowWiseAdagrad, task.two_tower.ebc.parameters(), {"lr": LR})
    model = DistributedModelParallel(module=task, device=torch.device("cpu"))
    optimizer = KeyedOptimizerWrapper(dict(model.named_parameters()), lambda params: torch.optim.Adam(params, lr=LR))
    pipeline = TrainPipelineSparseDist(model, optimizer, torch.device("cpu"))
    transform_partial = partial(ref.transform_to_torchrec_batch, num_embeddings_per_feature=EMB)

    out = {}
    for i, raw in enumerate(raws):
        b = ref.transform_to_torchrec_batch(raw, EMB)
        out[f"raw{i}_user_id"] = np.asarray(raw["user_id"], dtype=np.int64)
        out[f"raw{i}_product_id"] = np.asarray(raw["product_id"], dtype=np.int64)
        out[f"batch{i}_values"] = b.sparse_features.values().numpy().astype(np.int64)
        out[f"batch{i}_lengths"] = b.sparse_features.lengths().numpy().astype(np.int32)
        out[f"batch{i}_labels"] = b.labels.numpy().astype(np.int32)

    # the reference's own train() over the first STEPS batches, then its evaluate() over the last two
    ref.train(pipeline, raws[:STEPS], raws[:STEPS], epoch=0, print_lr=False, validation_freq=None, limit_train_batches=None,
              limit_val_batches=None, transform_partial=transform_partial)
    assert len(pipeline.trace) == STEPS
    for i, (loss, logits) in enumerate(pipeline.trace):
        out[f"step{i}_loss"] = loss.numpy().astype(np.float32)
        out[f"step{i}_logits"] = logits.numpy().astype(np.float32)
    for name, p in two_tower.named_parameters():
        if "embedding_bags" in name:
            assert p.grad is None                                    # fused in the backward: the hook clears it
            (opt,) = p._in_backward_optimizers
            out["sum." + name.split(".")[2]] = opt.state[p]["sum"].numpy().astype(np.float32)
    final = {k: v.detach().clone() for k, v in two_tower.state_dict().items()}
    pipeline.trace.clear()
    avg_loss, _ = ref.evaluate(None, pipeline, raws[STEPS:], "val", transform_partial)
    for k, v in two_tower.state_dict().items():
        assert torch.equal(v, final[k])                              # evaluate() does not train
    out["eval_average_loss"] = np.float32(avg_loss)                  # U:246: summed batch losses / number of SAMPLES
    for i, (loss, logits) in enumerate(pipeline.trace):
        out[f"eval{i}_loss"] = loss.numpy().astype(np.float32)
        out[f"eval{i}_logits"] = logits.numpy().astype(np.float32)
    for k, v in init.items():
        out["init." + k] = v.numpy()
    for k, v in final.items():
        out["final." + k] = v.numpy()
    # 03_model_training.py:1056-1122: the reference's own corpus / query embedding functions on the trained model
    nb = _functions_of("/root/reference/03_model_training.py", ["create_keyed_jagged_tensor", "process_embeddings"],
                       {"torch": torch, "KeyedJaggedTensor": KeyedJaggedTensor})
    two_tower.eval()
    for key, n in (("product_id", EMB[1]), ("user_id", EMB[0])):
        kjt = nb["create_keyed_jagged_tensor"](n, list(CAT), key, device="cpu")
        emb = nb["process_embeddings"](two_tower, kjt, key)
        assert emb is not None and emb.shape == (n, LAYERS[-1])
        out[f"corpus_{key}_values"] = kjt.values().numpy().astype(np.int64)
        out[f"corpus_{key}_lengths"] = kjt.lengths().numpy().astype(np.int32)
        out[f"corpus_{key}_embeddings"] = emb.numpy().astype(np.float32)
    # 04_evaluate_retrieval.py:134-141 asks a remote Vector Search index for the 100 best items per user by score; the
    # service is not here -- stock torch gives the exact answer (descending score, stable: ties keep the lower id)
    users, items = torch.from_numpy(out["corpus_user_id_embeddings"]), torch.from_numpy(out["corpus_product_id_embeddings"])
    order = torch.sort(users @ items.t(), dim=1, descending=True, stable=True)
    out["top100_scores"] = order.values[:, :100].numpy().astype(np.float32)
    out["top100_ids"] = order.indices[:, :100].numpy().astype(np.int64)
    out["meta"] = np.asarray([EMB[0], EMB[1], DIM, LAYERS[0], LAYERS[1], B, STEPS], dtype=np.int64)
    out["lr"] = np.float32(LR)
    path = os.path.join(os.environ.get("TT_GOLDEN_OUT", HERE), "reference_train.npz")      # TT_GOLDEN_OUT: write elsewhere (tests/test_golden_provenance.py)
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)} bytes, losses {[float(out[f'step{i}_loss']) for i in range(STEPS)]}, eval {avg_loss:.6f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
