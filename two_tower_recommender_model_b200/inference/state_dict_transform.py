"""``torchrec.inference.state_dict_transform`` names imported (but not called) at
/root/reference/utils/model_training.py:25-28."""
from typing import Dict, Union

import torch
from torch import distributed as dist

from ..distributed.sharded_tensor import gather_if_sharded


def state_dict_gather(src: Dict[str, torch.Tensor], dst: Dict[str, torch.Tensor]) -> None:
    """Gathers every (possibly sharded) entry of ``src`` into the full tensor ``dst[key]`` on rank 0."""
    for key, dst_tensor in dst.items():
        full = gather_if_sharded(src[key], dst_rank=0)
        if full is not None:
            dst_tensor.copy_(full)


def state_dict_to_device(state_dict: Dict[str, torch.Tensor], pg=None, device: Union[str, torch.device] = "cpu") -> Dict[str, torch.Tensor]:
    return {k: (v if not isinstance(v, torch.Tensor) else v.to(device)) for k, v in state_dict.items()}
