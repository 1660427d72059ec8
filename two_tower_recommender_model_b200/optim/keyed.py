"""``torchrec.optim.keyed.KeyedOptimizerWrapper`` as used at
/root/reference/03_model_training.py:826-829:
``KeyedOptimizerWrapper(dict(model.named_parameters()), lambda params: Adam(params, lr=...))``
and read at /root/reference/utils/model_training.py:303 (``.param_groups``).
Parameters that already carry an in-backward optimizer (the embedding tables)
are left out, as in TorchRec."""
from typing import Any, Callable, Dict, List, Mapping

import torch
from torch.optim.optimizer import Optimizer


class KeyedOptimizer(Optimizer):
    def __init__(self, params: Mapping[str, torch.Tensor], state: Mapping[Any, Any], param_groups: List[Dict[str, Any]]) -> None:
        self.params = dict(params)
        self.state = state
        self.param_groups = param_groups
        self.defaults = {}
        self._optimizer_step_pre_hooks = {}
        self._optimizer_step_post_hooks = {}

    def init_state(self, *args, **kwargs) -> None:
        pass


class KeyedOptimizerWrapper(KeyedOptimizer):
    def __init__(self, params: Mapping[str, torch.Tensor], optim_factory: Callable[[List[torch.Tensor]], Optimizer]) -> None:
        kept = {k: p for k, p in params.items() if not getattr(p, "_in_backward_optimizers", None)}
        self._optimizer = optim_factory(list(kept.values()))
        super().__init__(kept, self._optimizer.state, self._optimizer.param_groups)

    def zero_grad(self, set_to_none: bool = True) -> None:
        self._optimizer.zero_grad(set_to_none=set_to_none)

    def step(self, closure: Any = None) -> None:
        self._optimizer.step(closure=closure)

    def state_dict(self) -> Dict[str, Any]:
        return self._optimizer.state_dict()

    def load_state_dict(self, state_dict: Mapping[str, Any]) -> None:
        self._optimizer.load_state_dict(state_dict)
        # torch rebuilds param_groups / state on load: keep reading the live objects (utils/model_training.py:303 prints
        # ``pipeline._optimizer.param_groups[i]["lr"]``, schedulers write it)
        self.state, self.param_groups = self._optimizer.state, self._optimizer.param_groups
