"""``torchrec.modules.mlp.MLP`` as used at
/root/reference/utils/model_training.py:95-96 (``MLP(in_size=, layer_sizes=,
device=)``) and /root/reference/03_model_training.py:1143
(``._mlp[-1]._linear.out_features``): a stack of ``Perceptron`` =
``activation(Linear(x))`` with the activation (default ``torch.relu``) after
EVERY layer, the last one included.  Parameter names match TorchRec
(``_mlp.<i>._linear.{weight,bias}``).  ``nn.Linear`` only holds the parameters
(and gives them torch's default init); the matmul, bias and ReLU run fused in
libtt_b200.so."""
from typing import Callable, List, Optional, Union

import torch
from torch import nn

from ..functional import MlpTC, linear_act


class Perceptron(nn.Module):
    def __init__(self, in_size: int, out_size: int, bias: bool = True,
                 activation: Union[nn.Module, Callable[[torch.Tensor], torch.Tensor]] = torch.relu,
                 device: Optional[torch.device] = None, dtype: torch.dtype = torch.float32) -> None:
        super().__init__()
        if dtype != torch.float32:
            raise NotImplementedError("tower parameters are kept in float32 (bf16 is a compute mode, not a storage dtype)")
        self._out_size = out_size
        self._in_size = in_size
        self._linear = nn.Linear(in_size, out_size, bias=bias, device=device, dtype=dtype)
        self._activation_fn = activation

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        act = self._activation_fn
        fused_relu = act is torch.relu or act is torch.nn.functional.relu or isinstance(act, nn.ReLU)
        y = linear_act(input, self._linear.weight, self._linear.bias, fused_relu)
        return y if fused_relu else act(y)


class MLP(nn.Module):
    def __init__(self, in_size: int, layer_sizes: List[int], bias: bool = True,
                 activation: Union[str, Callable[[], nn.Module], nn.Module, Callable[[torch.Tensor], torch.Tensor]] = torch.relu,
                 device: Optional[torch.device] = None, dtype: torch.dtype = torch.float32,
                 precision: str = "fp32") -> None:
        super().__init__()
        if precision not in ("fp32", "bf16"):
            raise ValueError(f"unknown precision {precision}")
        # "bf16": the tower runs on tcgen05 (bf16 operands, fp32 accumulate, fp32 master weights);
        # "fp32": exact-fp32 CUDA-core path (the reference's numerics).
        self.precision = precision
        if activation == "relu":
            activation = torch.relu
        elif activation == "sigmoid":
            activation = torch.sigmoid
        if isinstance(activation, str):
            raise ValueError(f"unsupported activation {activation}")
        sizes = [in_size] + list(layer_sizes)
        self._mlp = nn.Sequential(*[
            Perceptron(sizes[i], sizes[i + 1], bias=bias, activation=activation, device=device, dtype=dtype)
            for i in range(len(layer_sizes))
        ])

    def all_relu(self) -> bool:
        return all(p._activation_fn is torch.relu or p._activation_fn is torch.nn.functional.relu
                   or isinstance(p._activation_fn, nn.ReLU) for p in self._mlp)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        layers = list(self._mlp)
        relu_all = self.all_relu()
        if self.precision == "bf16" and relu_all and input.shape[0] > 0:
            params = []
            for p in layers:
                params += [p._linear.weight, p._linear.bias]
            return MlpTC.apply(input, *params)
        return self._mlp(input)
