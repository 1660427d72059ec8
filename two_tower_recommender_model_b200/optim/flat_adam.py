"""Adam over ONE flat buffer: the dense (tower) parameters are re-homed as views
of a single contiguous fp32 buffer, their gradients as views of a second one,
so ``step()`` is one kernel launch (``tt_adam_flat``) and a data-parallel
all-reduce is one collective on ``flat_grad``.  Same arithmetic and defaults as
``torch.optim.Adam`` (/root/reference/03_model_training.py:826-829)."""
from typing import Any, Iterable, Tuple

import torch
from torch.optim.optimizer import Optimizer

from .. import _native as N


class FlatAdam(Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8) -> None:
        params = [p for p in params]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdam updates ONE flat buffer with one (lr, betas, eps): give it a single parameter group")
        ps = [p for g in self.param_groups for p in g["params"]]
        if not ps:
            raise ValueError("FlatAdam got no parameters")
        dev = ps[0].device
        N.require_cuda(ps[0], "parameter")
        n = sum(p.numel() for p in ps)
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_count = 0
        # the same counter on the device: keeps step() free of host-computed arguments (CUDA graphs)
        self.step_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in ps:
                if p.dtype != torch.float32 or p.device != dev:
                    raise TypeError("FlatAdam needs float32 parameters on one CUDA device")
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_param[off:off + k].view(p.shape)
                p.grad = self.flat_grad[off:off + k].view(p.shape)
                off += k
        self._params = ps

    def zero_grad(self, set_to_none: bool = False) -> None:
        # gradients live in flat_grad permanently; autograd accumulates into the views
        self.flat_grad.zero_()
        off = 0
        for p in self._params:
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + off * 4:
                p.grad = self.flat_grad[off:off + k].view(p.shape)
            off += k

    def add_param_group(self, param_group) -> None:
        if getattr(self, "param_groups", None):
            raise ValueError("FlatAdam supports a single parameter group")
        super().add_param_group(param_group)

    # The moments and the step counter live in flat buffers outside ``Optimizer.state``; they ARE the optimizer
    # state, so they travel in state_dict() (a resume that silently restarted Adam would not be a resume).
    def state_dict(self):
        sd = super().state_dict()
        sd["flat_adam"] = {"exp_avg": self.exp_avg.detach().clone(), "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                           "step": float(self.step_dev.item())}
        return sd

    def load_state_dict(self, state_dict) -> None:
        state_dict = dict(state_dict)
        flat = state_dict.pop("flat_adam", None)
        super().load_state_dict(state_dict)
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdam supports a single parameter group")
        if flat is not None:
            if flat["exp_avg"].numel() != self.exp_avg.numel():
                raise ValueError("FlatAdam state does not match the parameters")
            self.exp_avg.copy_(flat["exp_avg"])
            self.exp_avg_sq.copy_(flat["exp_avg_sq"])
            self.step_dev.fill_(float(flat["step"]))
            self.step_count = int(flat["step"])

    @torch.no_grad()
    def step(self, closure: Any = None) -> None:
        g = self.param_groups[0]
        self.step_count += 1
        b1, b2 = g["betas"]
        N.call("tt_adam_flat_devstep", N.ptr(self.flat_param), N.ptr(self.flat_grad), N.ptr(self.exp_avg),
               N.ptr(self.exp_avg_sq), self.flat_param.numel(), float(g["lr"]), float(b1), float(b2),
               float(g["eps"]), N.ptr(self.step_dev), N.stream_ptr(self.flat_param.device))
