"""``torchrec.distributed.comm.get_local_size`` (/root/reference/03_model_training.py:800)."""
import os
from typing import Optional

import torch.distributed as dist


def get_local_size(world_size: Optional[int] = None) -> int:
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else int(os.environ.get("WORLD_SIZE", 1))
    local = os.environ.get("LOCAL_WORLD_SIZE") or os.environ.get("LOCAL_SIZE")
    if local is not None:
        return int(local)
    return world_size


def get_local_rank(world_size: Optional[int] = None, rank: Optional[int] = None) -> int:
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    return rank % get_local_size(world_size)
