"""``DistributedModelParallel(module=, device=)`` and ``get_default_sharders()``
(/root/reference/03_model_training.py:809-815): materialises ``meta`` embedding
tables according to the sharding plan, moves the dense part to ``device`` and
keeps it data-parallel.  ``.module``, ``._plan.plan`` (03_model_training.py:819),
``.named_parameters()`` and ``.state_dict()`` behave as the reference expects.

world_size == 1: tables are materialised whole on ``device``.
world_size  > 1: every EmbeddingBagCollection is replaced by a
``ShardedEmbeddingBagCollection`` (table-wise / row-wise / column-wise / data-parallel, see sharding.py) and the
dense parameters are all-reduced through one flat gradient buffer.
"""
from typing import Any, Dict, Iterator, List, Optional, Tuple

import torch
from torch import distributed as dist
from torch import nn

from ..modules.embedding_modules import EmbeddingBagCollection
from .planner import EmbeddingShardingPlanner, ShardingPlan, Topology


class EmbeddingBagCollectionSharder:
    """Marker returned by get_default_sharders(); the planner/DMP know the one module type."""
    module_type = EmbeddingBagCollection


def get_default_sharders() -> List[Any]:
    return [EmbeddingBagCollectionSharder()]


def _world(pg=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(pg), dist.get_world_size(pg)
    return 0, 1


class DistributedModelParallel(nn.Module):
    def __init__(self, module: nn.Module, env: Any = None, device: Optional[torch.device] = None,
                 plan: Optional[ShardingPlan] = None, sharders: Optional[List[Any]] = None,
                 init_data_parallel: bool = True, init_parameters: bool = True, pg: Any = None,
                 sharding_kwargs: Optional[Dict[str, Any]] = None) -> None:
        super().__init__()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._pg = pg
        rank, world = _world(pg)
        self._rank, self._world = rank, world
        if plan is None:
            plan = EmbeddingShardingPlanner(topology=Topology(world_size=world, compute_device=self.device.type)
                                            ).collective_plan(module, sharders or get_default_sharders(), pg)
        self._plan = plan
        self._dense_sync = None
        if world == 1:
            for _, m in module.named_modules():
                if isinstance(m, EmbeddingBagCollection):
                    m.materialize(self.device)
        else:
            from .sharding import shard_embedding_modules
            module = shard_embedding_modules(module, plan, self.device, pg, **(sharding_kwargs or {}))
        # move everything that is not already placed (dense towers, buffers)
        for p in module.parameters():
            if p.device.type == "meta":
                raise RuntimeError("a non-embedding parameter is on the meta device; only "
                                   "EmbeddingBagCollection tables may be constructed on meta")
        module.to(self.device)
        self._dmp_wrapped_module = module
        self._dp_modules = [m for m in module.modules() if getattr(m, "dp_ebc", None) is not None]
        if world > 1 and init_data_parallel:
            from .sharding import DenseGradSync
            self._dense_sync = DenseGradSync(module, pg)
            # start the towers' gradient all-reduce as soon as their backward is done, concurrent with the
            # embedding backward (TorchRec's DDP overlaps its buckets with the backward in the same way)
            for m in module.modules():
                if hasattr(m, "set_pre_backward"):
                    m.set_pre_backward(self._dense_sync.start_async)

    @property
    def module(self) -> nn.Module:
        return self._dmp_wrapped_module

    @property
    def plan(self) -> ShardingPlan:
        return self._plan

    def forward(self, *args, **kwargs) -> Any:
        return self._dmp_wrapped_module(*args, **kwargs)

    # TorchRec's DMP reports names relative to the wrapped module (no ``module.`` /
    # ``_dmp_wrapped_module.`` prefix): ``two_tower.ebc.embedding_bags.<t>.weight`` ...
    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True) -> Iterator[Tuple[str, nn.Parameter]]:
        yield from self._dmp_wrapped_module.named_parameters(prefix=prefix, recurse=recurse, remove_duplicate=remove_duplicate)

    def named_buffers(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        yield from self._dmp_wrapped_module.named_buffers(prefix=prefix, recurse=recurse, remove_duplicate=remove_duplicate)

    def state_dict(self, *args, **kwargs) -> Dict[str, Any]:
        return self._dmp_wrapped_module.state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict: bool = True):
        return self._dmp_wrapped_module.load_state_dict(state_dict, strict=strict)

    def start_sparse_data_dist(self, batch, ready_event=None):
        fn = getattr(self._dmp_wrapped_module, "start_sparse_data_dist", None)
        if fn is None:
            for m in self._dmp_wrapped_module.modules():
                if hasattr(m, "prefetch_input_dist"):
                    m.prefetch_input_dist(batch, ready_event)
            return batch
        return fn(batch, ready_event)

    def sync_dense_grads(self) -> None:
        """After the backward of a training step: joins (or runs) the all-reduce of the tower gradients, then averages the
        dense gradients of data_parallel embedding tables and applies their optimizer (sharding.py)."""
        if self._dense_sync is not None:
            self._dense_sync.all_reduce()
        for m in self._dp_modules:
            m.sync_data_parallel()
