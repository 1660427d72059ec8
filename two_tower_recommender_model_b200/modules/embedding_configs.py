"""``torchrec.modules.embedding_configs`` surface used at
/root/reference/03_model_training.py:770-778 (``EmbeddingBagConfig(name=,
embedding_dim=, num_embeddings=, feature_names=)``) and
/root/reference/utils/model_training.py:88-93 (``.embedding_dim``,
``.feature_names``).  ``pooling`` defaults to SUM as in TorchRec; MEAN is what
BASELINE config 3 (user-history bags) needs."""
from dataclasses import dataclass, field
from enum import Enum, unique
from math import sqrt
from typing import Callable, List, Optional

import torch


@unique
class PoolingType(Enum):
    SUM = "SUM"
    MEAN = "MEAN"
    NONE = "NONE"


@unique
class DataType(Enum):
    FP32 = "FP32"


@dataclass
class BaseEmbeddingConfig:
    num_embeddings: int
    embedding_dim: int
    name: str = ""
    data_type: DataType = DataType.FP32
    feature_names: List[str] = field(default_factory=list)
    weight_init_max: Optional[float] = None
    weight_init_min: Optional[float] = None
    init_fn: Optional[Callable[[torch.Tensor], Optional[torch.Tensor]]] = None

    def get_weight_init_max(self) -> float:
        return sqrt(1 / self.num_embeddings) if self.weight_init_max is None else self.weight_init_max

    def get_weight_init_min(self) -> float:
        return -sqrt(1 / self.num_embeddings) if self.weight_init_min is None else self.weight_init_min

    def num_features(self) -> int:
        return len(self.feature_names)


@dataclass
class EmbeddingBagConfig(BaseEmbeddingConfig):
    pooling: PoolingType = PoolingType.SUM


def pooling_type_to_str(p: PoolingType) -> str:
    if p == PoolingType.SUM:
        return "sum"
    if p == PoolingType.MEAN:
        return "mean"
    raise ValueError(f"Unsupported pooling type {p}")
