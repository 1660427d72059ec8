"""``torchrec.optim.rowwise_adagrad.RowWiseAdagrad`` -- the optimizer class the
reference hands to ``apply_optimizer_in_backward``
(/root/reference/03_model_training.py:791-795).

Inside this package the class is mostly a TAG: EmbeddingBagCollection reads
``param._optimizer_classes / _optimizer_kwargs`` and fuses the update into its
backward kernel (``s_r += mean_d(g_r^2)``; ``w_r -= lr * g_r / (sqrt(s_r) + eps)``,
TorchRec defaults lr=1e-2, eps=1e-10, initial_accumulator_value=0).  ``step()``
exists for API compatibility and refuses to run: a dense ``[R, D]`` embedding
gradient is exactly what this design never materialises."""
from typing import Any, Iterable

import torch
from torch.optim.optimizer import Optimizer


class RowWiseAdagrad(Optimizer):
    DEFAULT_LR = 1e-2
    DEFAULT_EPS = 1e-10

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-2, lr_decay: float = 0.0,
                 weight_decay: float = 0.0, initial_accumulator_value: float = 0.0, eps: float = 1e-10,
                 *, maximize: bool = False, **unused: Any) -> None:
        if lr < 0 or eps < 0 or initial_accumulator_value < 0:
            raise ValueError("invalid RowWiseAdagrad hyper-parameter")
        if maximize:
            raise NotImplementedError("maximize is not supported")
        defaults = dict(lr=lr, lr_decay=lr_decay, eps=eps, weight_decay=weight_decay,
                        initial_accumulator_value=initial_accumulator_value)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure: Any = None) -> None:
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is not None:
                    raise RuntimeError(
                        "RowWiseAdagrad.step() on a dense embedding gradient is not implemented: register it with "
                        "apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {...}) so the update is fused "
                        "into the EmbeddingBagCollection backward kernel")
