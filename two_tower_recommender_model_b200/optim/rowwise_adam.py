"""Row-wise Adam for embedding tables (extension named by BASELINE.json config 4;
FBGEMM ``PARTIAL_ROWWISE_ADAM`` semantics): first moment per element, second
moment per row, bias-corrected with a global step; only touched rows advance.
Like RowWiseAdagrad this class is a tag consumed by EmbeddingBagCollection."""
from typing import Any, Iterable, Tuple

import torch
from torch.optim.optimizer import Optimizer


class RowWiseAdam(Optimizer):
    DEFAULT_LR = 1e-2
    DEFAULT_EPS = 1e-8

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-2,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 **unused: Any) -> None:
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure: Any = None) -> None:
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    raise RuntimeError("RowWiseAdam must be registered with apply_optimizer_in_backward; "
                                       "dense embedding gradients are never materialised")
