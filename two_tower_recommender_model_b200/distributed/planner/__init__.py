"""``EmbeddingShardingPlanner`` / ``Topology`` / ``ParameterConstraints`` with the
call surface of /root/reference/03_model_training.py:798-811.

The reference passes no constraints and lets TorchRec's enumerator choose; here
the plan is a deterministic function of (tables, constraints, topology):

* an explicit ``ParameterConstraints(sharding_types=[...])`` for a table wins
  (``table_wise`` | ``row_wise`` | ``column_wise`` | ``data_parallel``: a replica on every rank, dense gradient
  all-reduce -- meant for SMALL tables such as configs[2]'s aisle / department features);
* otherwise tables go table-wise, largest first, each to the rank with the
  fewest bytes so far (greedy balance) -- unless a table (weights + row-wise
  optimizer state) does not fit in one GPU's budget, which makes it row-wise.

The budget is ``hbm_cap * (1 - storage_reservation.percentage)``; B200 = 180 GB.
Every rank computes the same plan from the same inputs, so ``collective_plan``
needs no broadcast (it still verifies agreement when a process group exists).
"""
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Dict, List, Optional

import torch
from torch import distributed as dist
from torch import nn

from .storage_reservations import HeuristicalStorageReservation  # noqa: F401

B200_HBM_BYTES = 180 * 1024 ** 3


class ShardingType(Enum):
    DATA_PARALLEL = "data_parallel"
    TABLE_WISE = "table_wise"
    ROW_WISE = "row_wise"
    COLUMN_WISE = "column_wise"
    TABLE_ROW_WISE = "table_row_wise"


@dataclass
class Topology:
    world_size: int
    compute_device: str = "cuda"
    hbm_cap: Optional[int] = None
    local_world_size: Optional[int] = None

    def __post_init__(self) -> None:
        if self.hbm_cap is None:
            self.hbm_cap = B200_HBM_BYTES
        if self.local_world_size is None:
            self.local_world_size = self.world_size


@dataclass
class ParameterConstraints:
    sharding_types: Optional[List[str]] = None
    compute_kernels: Optional[List[str]] = None
    pooling_factors: List[float] = field(default_factory=lambda: [1.0])


@dataclass
class ParameterSharding:
    sharding_type: str
    compute_kernel: str = "fused"
    ranks: Optional[List[int]] = None
    # row-wise: rows [r*block, (r+1)*block) live on ranks[r]
    block_size: Optional[int] = None
    num_embeddings: int = 0
    embedding_dim: int = 0

    def __repr__(self) -> str:
        extra = f", block_size={self.block_size}" if self.block_size else ""
        return (f"ParameterSharding(sharding_type='{self.sharding_type}', compute_kernel='{self.compute_kernel}', "
                f"ranks={self.ranks}{extra}, shape=[{self.num_embeddings}, {self.embedding_dim}])")


@dataclass
class ShardingPlan:
    plan: Dict[str, Dict[str, ParameterSharding]]

    def get_plan_for_module(self, module_path: str) -> Optional[Dict[str, ParameterSharding]]:
        return self.plan.get(module_path)

    def __str__(self) -> str:
        out = []
        for mod, tables in self.plan.items():
            out.append(f"module: {mod}")
            for name, ps in tables.items():
                out.append(f"  {name}: {ps}")
        return "\n".join(out)


def _find_ebcs(module: nn.Module):
    from ...modules.embedding_modules import EmbeddingBagCollection
    for path, m in module.named_modules():
        if isinstance(m, EmbeddingBagCollection):
            yield path, m


def _table_bytes(cfg, adam: bool = False) -> int:
    # weights + one fp32 of row-wise state per row (Adagrad's sum / Adam's second moment); row-wise Adam adds a
    # full-size first moment, i.e. the footprint of the table doubles
    nbytes = cfg.num_embeddings * cfg.embedding_dim * 4 + cfg.num_embeddings * 4
    return nbytes + (cfg.num_embeddings * cfg.embedding_dim * 4 if adam else 0)


def _uses_rowwise_adam(ebc) -> bool:
    """True when ``apply_optimizer_in_backward(RowWiseAdam, ebc.parameters(), ...)`` tagged this collection."""
    try:
        from ... import _native as N
        return ebc._in_backward_kind() == N.OPT_ROWWISE_ADAM
    except Exception:
        return False


class EmbeddingShardingPlanner:
    def __init__(self, topology: Optional[Topology] = None, batch_size: Optional[int] = None,
                 storage_reservation: Optional[HeuristicalStorageReservation] = None,
                 constraints: Optional[Dict[str, ParameterConstraints]] = None, **unused: Any) -> None:
        if topology is None:
            ws = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
            topology = Topology(world_size=ws)
        self._topology = topology
        self._batch_size = batch_size
        self._reservation = storage_reservation or HeuristicalStorageReservation(percentage=0.15)
        self._constraints = constraints or {}

    def plan(self, module: nn.Module, sharders: Optional[List[Any]] = None) -> ShardingPlan:
        W = self._topology.world_size
        budget = int(self._topology.hbm_cap * (1.0 - self._reservation.percentage))
        load = [0] * W
        plan: Dict[str, Dict[str, ParameterSharding]] = {}
        for path, ebc in _find_ebcs(module):
            tables: Dict[str, ParameterSharding] = {}
            adam = _uses_rowwise_adam(ebc)
            cfgs = sorted(ebc.embedding_bag_configs(), key=lambda c: (-_table_bytes(c, adam), c.name))
            for cfg in cfgs:
                want = None
                c = self._constraints.get(cfg.name)
                if c is not None and c.sharding_types:
                    want = c.sharding_types[0]
                nbytes = _table_bytes(cfg, adam)
                if want is None:
                    want = ShardingType.TABLE_WISE.value if (nbytes <= budget or W == 1) else ShardingType.ROW_WISE.value
                    # Fewer tables than ranks: table-wise would leave ranks without embedding work while the owners
                    # look up (and update) the global batch; row-wise spreads both evenly (TorchRec's planner reaches
                    # the same verdict through its perf estimates).
                    if W > len(cfgs) and cfg.num_embeddings >= W * 4096:
                        want = ShardingType.ROW_WISE.value
                if want == ShardingType.TABLE_WISE.value:
                    r = min(range(W), key=lambda i: (load[i], i))
                    if load[r] + nbytes > budget:
                        raise RuntimeError(f"table {cfg.name} ({nbytes / 2**30:.1f} GiB) does not fit rank {r}'s budget; "
                                           "constrain it to row_wise")
                    load[r] += nbytes
                    ps = ParameterSharding(want, ranks=[r])
                elif want == ShardingType.ROW_WISE.value:
                    block = -(-cfg.num_embeddings // W)
                    for r in range(W):
                        load[r] += nbytes // W
                    ps = ParameterSharding(want, ranks=list(range(W)), block_size=block)
                elif want == ShardingType.COLUMN_WISE.value:
                    # split D over as many ranks as divide it (at most W), least-loaded ranks first: shard j =
                    # columns [j * D/s, (j+1) * D/s) on ranks[j].  Spreads a table-wise owner's lookup / update of the
                    # global batch over s ranks; every shard keeps its own row-wise optimizer state,
                    # as TorchRec's column-wise shards (separate fused tables) do.
                    s_ = max(d for d in range(1, W + 1) if cfg.embedding_dim % d == 0)
                    ranks = sorted(sorted(range(W), key=lambda i: (load[i], i))[:s_])
                    for r in ranks:
                        load[r] += nbytes // s_
                    ps = ParameterSharding(want, ranks=ranks)
                elif want == ShardingType.DATA_PARALLEL.value:
                    # a full replica on every rank; its dense [R, D] gradient is all-reduced after the backward and the
                    # table's optimizer applied to the replica (sharding.py: sync_data_parallel) -- for small tables
                    if any(load[r] + nbytes > budget for r in range(W)):
                        raise RuntimeError(f"table {cfg.name} ({nbytes / 2**30:.1f} GiB) does not fit every rank's budget as a "
                                           "data_parallel replica; constrain it to row_wise")
                    for r in range(W):
                        load[r] += nbytes
                    ps = ParameterSharding(want, compute_kernel="dense", ranks=list(range(W)))
                else:
                    raise NotImplementedError(f"sharding type {want} is not implemented (table_wise, row_wise, column_wise, data_parallel are)")
                ps.num_embeddings, ps.embedding_dim = cfg.num_embeddings, cfg.embedding_dim
                tables[cfg.name] = ps
            # report in config order
            plan[path] = {c.name: tables[c.name] for c in ebc.embedding_bag_configs()}
        return ShardingPlan(plan)

    def collective_plan(self, module: nn.Module, sharders: Optional[List[Any]] = None, pg: Any = None) -> ShardingPlan:
        plan = self.plan(module, sharders)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(pg) > 1:
            mine = str(plan)
            gathered: List[Any] = [None] * dist.get_world_size(pg)
            dist.all_gather_object(gathered, mine, group=pg)
            if any(g != mine for g in gathered):
                raise RuntimeError("ranks disagree on the sharding plan (different tables or constraints?)")
        return plan
