"""``torchrec.datasets.utils.Batch`` as built at
/root/reference/utils/model_training.py:65-69."""
from dataclasses import dataclass

import torch

from ..sparse.jagged_tensor import KeyedJaggedTensor


@dataclass
class Batch:
    dense_features: torch.Tensor
    sparse_features: KeyedJaggedTensor
    labels: torch.Tensor

    def to(self, device: torch.device, non_blocking: bool = False) -> "Batch":
        return Batch(
            dense_features=self.dense_features.to(device=device, non_blocking=non_blocking),
            sparse_features=self.sparse_features.to(device=device, non_blocking=non_blocking),
            labels=self.labels.to(device=device, non_blocking=non_blocking),
        )

    def record_stream(self, stream) -> None:
        if self.dense_features.is_cuda:
            self.dense_features.record_stream(stream)
        self.sparse_features.record_stream(stream)
        if self.labels.is_cuda:
            self.labels.record_stream(stream)

    def pin_memory(self) -> "Batch":
        return Batch(
            dense_features=self.dense_features.pin_memory(),
            sparse_features=self.sparse_features.pin_memory(),
            labels=self.labels.pin_memory(),
        )
