from .rowwise_adagrad import RowWiseAdagrad  # noqa: F401
from .rowwise_adam import RowWiseAdam  # noqa: F401
from .keyed import KeyedOptimizerWrapper, KeyedOptimizer  # noqa: F401
from .flat_adam import FlatAdam  # noqa: F401
