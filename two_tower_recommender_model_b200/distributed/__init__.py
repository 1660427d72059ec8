from .train_pipeline import TrainPipelineSparseDist  # noqa: F401
from .model_parallel import DistributedModelParallel, get_default_sharders  # noqa: F401
