"""``EmbeddingBagCollection`` with the surface the reference uses
(/root/reference/03_model_training.py:781-784,1042-1045;
/root/reference/utils/model_training.py:88-94,101):

* ``EmbeddingBagCollection(tables=[EmbeddingBagConfig...], device=torch.device("meta") | real)``
* ``.embedding_bag_configs()``; ``__call__(kjt) -> KeyedTensor``; ``.parameters()``
* state-dict keys ``embedding_bags.<table name>.weight`` (TorchRec's layout, so a
  checkpoint written by either side loads into the other, 03_model_training.py:1052)
* ``apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": lr})``
  (03_model_training.py:791-795) is honoured: the tags torch puts on the
  parameters select the optimizer fused into the backward kernel.

Weights are fp32 ``[R, D]`` in HBM, one allocation per table, initialised
``U(-1/sqrt(R), 1/sqrt(R))`` like TorchRec.  Lookup, pooling, backward and the
row-wise optimizer all run in libtt_b200.so; nothing here computes on the CPU.
"""
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from .. import _native as N
from ..functional import EbcLookup
from ..sparse.jagged_tensor import KeyedJaggedTensor, KeyedTensor
from .embedding_configs import EmbeddingBagConfig, PoolingType

_TAG_ATTRS = ("_in_backward_optimizers", "_optimizer_classes", "_optimizer_kwargs", "_overlapped_optimizer")


class _EmbeddingTable(nn.Module):
    """Parameter holder named like ``nn.EmbeddingBag`` so state-dict keys match."""

    def __init__(self, cfg: EmbeddingBagConfig, device: Optional[torch.device]) -> None:
        super().__init__()
        self.num_embeddings = cfg.num_embeddings
        self.embedding_dim = cfg.embedding_dim
        self._init_min = cfg.get_weight_init_min()
        self._init_max = cfg.get_weight_init_max()
        self.weight = nn.Parameter(torch.empty(cfg.num_embeddings, cfg.embedding_dim, device=device, dtype=torch.float32))
        if self.weight.device.type != "meta":
            self.reset_parameters()

    def reset_parameters(self) -> None:
        with torch.no_grad():
            self.weight.uniform_(self._init_min, self._init_max)

    def materialize(self, device: torch.device) -> None:
        """Replaces a ``meta`` weight by a real one on ``device`` (what
        DistributedModelParallel does when it shards), keeping the
        optimizer-in-backward tags."""
        old = self.weight
        new = nn.Parameter(torch.empty(old.shape, device=device, dtype=torch.float32))
        for a in _TAG_ATTRS:
            if hasattr(old, a):
                setattr(new, a, getattr(old, a))
        self.weight = new
        self.reset_parameters()

    def extra_repr(self) -> str:
        return f"{self.num_embeddings}, {self.embedding_dim}"


class EmbeddingBagCollection(nn.Module):
    def __init__(self, tables: List[EmbeddingBagConfig], is_weighted: bool = False,
                 device: Optional[torch.device] = None) -> None:
        super().__init__()
        if is_weighted:
            raise NotImplementedError("weighted EmbeddingBagCollection is outside the reference's hot path")
        if len(tables) == 0:
            raise ValueError("EmbeddingBagCollection needs at least one table")
        self._is_weighted = False
        self._embedding_bag_configs = list(tables)
        self._device = torch.device(device) if device is not None else torch.device("cpu")
        self.embedding_bags = nn.ModuleDict()
        names, feats = set(), set()
        self._slot_feature: List[str] = []
        self._slot_table: List[int] = []
        for ti, cfg in enumerate(self._embedding_bag_configs):
            if cfg.name in names:
                raise ValueError(f"Duplicate table name {cfg.name}")
            names.add(cfg.name)
            if not cfg.feature_names:
                cfg.feature_names = [cfg.name]
            if cfg.pooling not in (PoolingType.SUM, PoolingType.MEAN):
                raise ValueError(f"Unsupported pooling {cfg.pooling}")
            self.embedding_bags[cfg.name] = _EmbeddingTable(cfg, self._device)
            for f in cfg.feature_names:
                if f in feats:
                    raise ValueError(f"Feature {f} is looked up by more than one table")
                feats.add(f)
                self._slot_feature.append(f)
                self._slot_table.append(ti)
        if len(self._slot_feature) > N.TT_MAX_FEATURES:
            raise ValueError(f"at most {N.TT_MAX_FEATURES} features are supported")
        self._slot_dim = [self._embedding_bag_configs[t].embedding_dim for t in self._slot_table]
        self._total_dim = sum(self._slot_dim)
        self._fused_state: Dict[str, Dict[str, torch.Tensor]] = {}
        self._fused_step = 0
        # Adam's step counter also lives on the device (float32 [1]): the fused backward increments it and derives
        # the bias corrections from it, so a captured CUDA graph keeps stepping (graph.py)
        self._fused_step_dev: Optional[torch.Tensor] = None
        # every gradient row is multiplied by this inside the fused backward; the sharded module sets 1/world
        # (TorchRec's gradient division in the pooled all-to-all / reduce-scatter backward)
        self._grad_scale = 1.0
        # False (default): state_dict() is weights-only, the reference's checkpoint format
        # (utils/model_training.py:161-189); True adds the fused optimizer state, see include_optimizer_state()
        self._state_dict_with_optimizer = False

    # ---- TorchRec surface
    def embedding_bag_configs(self) -> List[EmbeddingBagConfig]:
        return self._embedding_bag_configs

    def is_weighted(self) -> bool:
        return self._is_weighted

    @property
    def device(self) -> torch.device:
        return self._device

    def feature_names(self) -> List[str]:
        return list(self._slot_feature)

    def materialize(self, device: torch.device) -> None:
        for bag in self.embedding_bags.values():
            if bag.weight.device.type == "meta":
                bag.materialize(device)
        self._device = torch.device(device)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        w = next(iter(self.embedding_bags.values())).weight
        self._device = w.device
        for st in self._fused_state.values():
            for k in list(st.keys()):
                st[k] = fn(st[k])
        if self._fused_step_dev is not None:
            self._fused_step_dev = fn(self._fused_step_dev)
        return out

    def forward(self, features: KeyedJaggedTensor) -> KeyedTensor:
        weights = [self.embedding_bags[c.name].weight for c in self._embedding_bag_configs]
        if weights[0].device.type == "meta":
            raise RuntimeError("EmbeddingBagCollection is on the meta device; wrap it in "
                               "DistributedModelParallel or call .materialize(device) first")
        values = features.values()
        N.require_cuda(values, "KeyedJaggedTensor.values")
        if values.dtype != torch.int64:
            values = values.to(torch.int64)
        offsets = features.offsets()
        if offsets.dtype != torch.int32:
            offsets = offsets.to(torch.int32)
        if self._in_backward_kind() is not None and torch.is_grad_enabled():
            # Fused optimizer: the tables never receive a gradient, so they are not autograd inputs.
            # A fresh zero-size leaf keeps backward() reaching the lookup (TorchRec does the same);
            # unlike the parameters' long-lived AccumulateGrad nodes (pinned by
            # apply_optimizer_in_backward to the stream they were created on) it belongs to the
            # current stream, which keeps the step capturable in a CUDA graph.
            anchors = (torch.zeros(0, dtype=torch.float32, device=values.device, requires_grad=True),)
        else:
            anchors = tuple(weights)
        pooled = EbcLookup.apply(self, tuple(features.keys()), values.contiguous(), offsets.contiguous(),
                                 features.stride(), *anchors)
        return KeyedTensor(keys=self._slot_feature, length_per_key=self._slot_dim, values=pooled)

    # ---- plumbing for functional.EbcLookup
    def _build_plan(self, kjt_keys: Tuple[str, ...], batch: int, with_state: bool,
                    dense_grads: Optional[List[torch.Tensor]] = None,
                    out_layout: Optional[Tuple[int, Dict[str, int]]] = None) -> Tuple[N.EbcPlan, int]:
        """``out_layout`` = (row stride, first column per feature) places the pooled columns in a wider
        matrix than this collection's own (the sharded module's exchange buffer)."""
        plan = N.EbcPlan()
        nk = len(kjt_keys)
        if nk > N.TT_MAX_FEATURES:
            raise ValueError(f"KeyedJaggedTensor has more than {N.TT_MAX_FEATURES} keys")
        plan.num_slots = len(self._slot_feature)
        plan.batch_size = batch
        plan.num_kjt_keys = nk
        plan.out_stride = self._total_dim if out_layout is None else out_layout[0]
        key_pos = {k: i for i, k in enumerate(kjt_keys)}
        row_base, acc = [], 0
        for cfg in self._embedding_bag_configs:
            row_base.append(acc)
            acc += cfg.num_embeddings
        plan.total_rows = acc
        for i in range(N.TT_MAX_FEATURES):
            plan.slot_of_kjt[i] = -1
        spec_kind = self._in_backward_kind() if with_state else None
        col = 0
        for s, (feat, ti) in enumerate(zip(self._slot_feature, self._slot_table)):
            cfg = self._embedding_bag_configs[ti]
            if feat not in key_pos:
                raise KeyError(f"feature {feat} is not among the KeyedJaggedTensor keys {list(kjt_keys)}")
            w = self.embedding_bags[cfg.name].weight
            plan.weights[s] = w.data_ptr()
            plan.row_base[s] = row_base[ti]
            plan.num_rows[s] = cfg.num_embeddings
            plan.dim[s] = cfg.embedding_dim
            plan.kjt_index[s] = key_pos[feat]
            plan.slot_of_kjt[key_pos[feat]] = s
            plan.out_col[s] = col if out_layout is None else out_layout[1][feat]
            plan.pooling[s] = N.POOL_MEAN if cfg.pooling == PoolingType.MEAN else N.POOL_SUM
            col += cfg.embedding_dim
            if with_state:
                if dense_grads is not None:
                    plan.state1[s] = dense_grads[ti].data_ptr()
                elif spec_kind is not None:
                    st = self._state_for(cfg, w, spec_kind)
                    if spec_kind == N.OPT_ROWWISE_ADAGRAD:
                        plan.state0[s] = st["sum"].data_ptr()
                    elif spec_kind == N.OPT_ROWWISE_ADAM:
                        plan.state0[s] = st["exp_avg_sq"].data_ptr()
                        plan.state1[s] = st["exp_avg"].data_ptr()
        return plan, self._total_dim

    def _tagged(self):
        w = self.embedding_bags[self._embedding_bag_configs[0].name].weight
        classes = getattr(w, "_optimizer_classes", None)
        if not classes:
            return None, None
        return classes[0], dict(getattr(w, "_optimizer_kwargs")[0])

    def _in_backward_kind(self) -> Optional[int]:
        from ..optim.rowwise_adagrad import RowWiseAdagrad
        from ..optim.rowwise_adam import RowWiseAdam
        cls, _ = self._tagged()
        if cls is None:
            return None
        if issubclass(cls, RowWiseAdagrad):
            return N.OPT_ROWWISE_ADAGRAD
        if issubclass(cls, RowWiseAdam):
            return N.OPT_ROWWISE_ADAM
        if issubclass(cls, torch.optim.SGD):
            return N.OPT_SGD
        raise NotImplementedError(
            f"optimizer {cls.__name__} cannot be fused into the embedding backward; supported: "
            "RowWiseAdagrad, RowWiseAdam, torch.optim.SGD (plain)")

    def _sparse_optimizer_spec(self, advance_step: bool) -> Optional[N.SparseOptimizer]:
        from ..optim.rowwise_adagrad import RowWiseAdagrad
        from ..optim.rowwise_adam import RowWiseAdam
        kind = self._in_backward_kind()
        if kind is None:
            return None
        cls, kw = self._tagged()
        for name in self._embedding_bag_configs[1:]:
            c2 = getattr(self.embedding_bags[name.name].weight, "_optimizer_classes", None)
            if not c2 or c2[0] is not cls:
                raise NotImplementedError("all tables of one EmbeddingBagCollection must share the in-backward optimizer")
        if kw.get("weight_decay", 0.0) != 0.0 or kw.get("lr_decay", 0.0) != 0.0:
            raise NotImplementedError("weight_decay / lr_decay are not supported by the fused sparse optimizers")
        spec = N.SparseOptimizer(kind=kind)
        spec.grad_scale = float(self._grad_scale)
        if kind == N.OPT_ROWWISE_ADAGRAD:
            spec.lr = float(kw.get("lr", RowWiseAdagrad.DEFAULT_LR))
            spec.eps = float(kw.get("eps", RowWiseAdagrad.DEFAULT_EPS))
        elif kind == N.OPT_ROWWISE_ADAM:
            if advance_step:
                self._fused_step += 1
            b1, b2 = kw.get("betas", (0.9, 0.999))
            spec.lr = float(kw.get("lr", RowWiseAdam.DEFAULT_LR))
            spec.eps = float(kw.get("eps", RowWiseAdam.DEFAULT_EPS))
            spec.beta1, spec.beta2 = float(b1), float(b2)
            step = max(self._fused_step, 1)
            spec.bias_correction1 = 1.0 - float(b1) ** step
            spec.bias_correction2 = 1.0 - float(b2) ** step
            if advance_step:
                # the kernel increments the device counter itself and uses IT (the host values above are the
                # same numbers while the step runs eagerly; a replayed CUDA graph only has the device counter)
                w = self.embedding_bags[self._embedding_bag_configs[0].name].weight
                if self._fused_step_dev is None or self._fused_step_dev.device != w.device:
                    self._fused_step_dev = torch.full((1,), float(self._fused_step - 1), dtype=torch.float32, device=w.device)
                spec.step_dev = self._fused_step_dev.data_ptr()
        else:
            spec.lr = float(kw.get("lr", 1e-3))
        return spec

    def _state_for(self, cfg: EmbeddingBagConfig, w: torch.Tensor, kind: int) -> Dict[str, torch.Tensor]:
        st = self._fused_state.get(cfg.name)
        if st is None:
            st = {}
            self._fused_state[cfg.name] = st
        if kind == N.OPT_ROWWISE_ADAGRAD and "sum" not in st:
            _, kw = self._tagged()
            st["sum"] = torch.full((cfg.num_embeddings,), float(kw.get("initial_accumulator_value", 0.0)),
                                   dtype=torch.float32, device=w.device)
        if kind == N.OPT_ROWWISE_ADAM and "exp_avg" not in st:
            st["exp_avg"] = torch.zeros_like(w)
            st["exp_avg_sq"] = torch.zeros(cfg.num_embeddings, dtype=torch.float32, device=w.device)
        return st

    def _alloc_dense_grads(self) -> List[torch.Tensor]:
        return [torch.zeros_like(self.embedding_bags[c.name].weight) for c in self._embedding_bag_configs]

    def fused_optimizer_state(self) -> Dict[str, Dict[str, torch.Tensor]]:
        """Row-wise optimizer state per table (absent in the reference's
        checkpoints; exposed so a resume can be exact)."""
        return self._fused_state

    def fused_step(self) -> int:
        """Number of fused Adam steps taken (read from the device counter when a CUDA graph advanced it)."""
        if self._fused_step_dev is not None:
            return int(round(float(self._fused_step_dev.item())))
        return self._fused_step

    def load_fused_optimizer_state(self, state: Dict[str, Dict[str, torch.Tensor]], step: int = 0) -> None:
        for name, st in state.items():
            w = self.embedding_bags[name].weight
            for k, v in st.items():
                want = tuple(w.shape) if k == "exp_avg" else (w.shape[0],)
                if tuple(v.shape) != want:
                    raise ValueError(f"fused optimizer state {name}.{k}: shape {tuple(v.shape)}, table needs {want}")
            self._fused_state[name] = {k: v.to(w.device).clone() for k, v in st.items()}
        self._fused_step = int(step)
        self._fused_step_dev = None

    # ---- optimizer state inside the checkpoint (the reference's own checkpoints are weights-only)
    def include_optimizer_state(self, on: bool = True) -> "EmbeddingBagCollection":
        """``True``: ``state_dict()`` also carries the fused optimizer's state as
        ``embedding_bags.<table>.{sum | exp_avg | exp_avg_sq}`` plus ``fused_optimizer_step``, and
        ``load_state_dict`` restores it, so save -> reload -> next step is identical.  Default ``False`` keeps the
        reference's weights-only format (utils/model_training.py:161-189)."""
        self._state_dict_with_optimizer = bool(on)
        return self

    def _optimizer_state_entries(self) -> Dict[str, torch.Tensor]:
        out: Dict[str, torch.Tensor] = {}
        kind = self._in_backward_kind()
        if kind in (N.OPT_ROWWISE_ADAGRAD, N.OPT_ROWWISE_ADAM):
            for cfg in self._embedding_bag_configs:
                w = self.embedding_bags[cfg.name].weight
                if w.device.type == "meta":
                    continue
                for k, v in self._state_for(cfg, w, kind).items():
                    out[f"embedding_bags.{cfg.name}.{k}"] = v
            out["fused_optimizer_step"] = torch.tensor(float(self.fused_step()), dtype=torch.float32)
        return out

    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):
        destination = super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)
        if self._state_dict_with_optimizer:
            for k, v in self._optimizer_state_entries().items():
                destination[prefix + k] = v if keep_vars else v.detach()
        return destination

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # optimizer-state entries are optional on load (a weights-only checkpoint of the reference still loads)
        names = {f"{prefix}embedding_bags.{c.name}.{k}": (c, k) for c in self._embedding_bag_configs
                 for k in ("sum", "exp_avg", "exp_avg_sq")}
        for key in [k for k in state_dict if k in names]:
            cfg, k = names[key]
            w = self.embedding_bags[cfg.name].weight
            t = state_dict.pop(key)
            want = (cfg.num_embeddings, cfg.embedding_dim) if k == "exp_avg" else (cfg.num_embeddings,)
            if tuple(t.shape) != want:
                # the update kernel indexes the state by row: a shorter tensor would be written out of bounds
                error_msgs.append(f"size mismatch for {key}: checkpoint {tuple(t.shape)}, table needs {want}")
                continue
            self._fused_state.setdefault(cfg.name, {})[k] = t.detach().to(device=w.device, dtype=torch.float32).clone()
        step_key = prefix + "fused_optimizer_step"
        if step_key in state_dict:
            self._fused_step = int(round(float(state_dict.pop(step_key))))
            self._fused_step_dev = None
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
