"""B200-native two-tower training / retrieval hot path behind the TorchRec-shaped
API of /root/reference/utils/model_training.py.  All device work is hand-written
CUDA for sm_100a in libtt_b200.so (C ABI: include/tt_b200.h); there is no CPU,
Triton, FBGEMM or torch.compile fallback."""
from . import _native  # noqa: F401
from .datasets.utils import Batch  # noqa: F401
from .distributed import DistributedModelParallel, TrainPipelineSparseDist, get_default_sharders  # noqa: F401
from .distributed.comm import get_local_size  # noqa: F401
from .distributed.planner import EmbeddingShardingPlanner, ParameterConstraints, Topology  # noqa: F401
from .distributed.planner.storage_reservations import HeuristicalStorageReservation  # noqa: F401
from .graph import CudaGraphTrainStep  # noqa: F401
from .modules import MLP, EmbeddingBagCollection, EmbeddingBagConfig, PoolingType  # noqa: F401
from .optim import FlatAdam, KeyedOptimizerWrapper, RowWiseAdagrad, RowWiseAdam  # noqa: F401
from .retrieval import (BruteForceIndex, CorpusShardedIndex, create_keyed_jagged_tensor, embed_corpus, embed_corpus_sharded,  # noqa: F401
                        process_embeddings, retrieval_metrics)
from .shim import install_torchrec_shim  # noqa: F401
from .sparse import JaggedTensor, KeyedJaggedTensor, KeyedTensor  # noqa: F401
from .two_tower import TwoTower, TwoTowerTrainTask  # noqa: F401

__version__ = "0.1.0"
