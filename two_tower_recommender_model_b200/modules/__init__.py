from .embedding_configs import EmbeddingBagConfig, PoolingType  # noqa: F401
from .embedding_modules import EmbeddingBagCollection  # noqa: F401
from .mlp import MLP, Perceptron  # noqa: F401
