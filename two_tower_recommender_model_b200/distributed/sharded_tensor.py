"""Helpers around ``torch.distributed._shard.sharded_tensor.ShardedTensor`` -- the
type /root/reference/utils/model_training.py:169-176 tests for and calls
``.gather(0, full_tensor)`` / ``.size()`` / ``.device`` on."""
from typing import List, Optional

import torch
from torch import distributed as dist

try:  # the import path the reference uses (deprecated alias) and the current one
    from torch.distributed._shard.sharded_tensor import Shard, ShardedTensor, ShardMetadata
except Exception:  # pragma: no cover
    Shard = ShardedTensor = ShardMetadata = None


def make_row_sharded(local_rows: Optional[torch.Tensor], row_offset: int, global_shape, pg=None):
    """Wraps this rank's block of rows (or nothing) of a ``[R, D]`` table -- or of a per-row ``[R]`` optimizer
    state -- as a ShardedTensor."""
    shards: List = []
    if local_rows is not None and local_rows.numel() > 0:
        rank = dist.get_rank(pg)
        dev = local_rows.device
        md = ShardMetadata(shard_offsets=[row_offset] + [0] * (local_rows.dim() - 1), shard_sizes=list(local_rows.shape),
                           placement=f"rank:{rank}/{dev}")
        shards.append(Shard(tensor=local_rows, metadata=md))
    return ShardedTensor._init_from_local_shards(shards, *global_shape, process_group=pg)


def make_col_sharded(local_cols: Optional[torch.Tensor], col_offset: int, global_shape, pg=None):
    """Wraps this rank's block of COLUMNS (or nothing) of a ``[R, D]`` table as a ShardedTensor (column-wise sharding);
    ``ShardedTensor.gather`` places shards by their N-d offsets, so utils/model_training.py:161-182 gathers it unchanged."""
    shards: List = []
    if local_cols is not None and local_cols.numel() > 0:
        rank = dist.get_rank(pg)
        dev = local_cols.device
        md = ShardMetadata(shard_offsets=[0, col_offset], shard_sizes=list(local_cols.shape), placement=f"rank:{rank}/{dev}")
        shards.append(Shard(tensor=local_cols, metadata=md))
    return ShardedTensor._init_from_local_shards(shards, *global_shape, process_group=pg)


def gather_if_sharded(t, dst_rank: int = 0) -> Optional[torch.Tensor]:
    if ShardedTensor is not None and isinstance(t, ShardedTensor):
        full = None
        if dist.get_rank() == dst_rank:
            full = torch.zeros(t.size(), device=t.local_shards()[0].tensor.device if t.local_shards() else None)
        t.gather(dst_rank, full)
        return full
    return t
