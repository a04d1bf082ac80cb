"""Linear layers.  ``Linear`` == the reference's OptimizedLinear (weight (out,in), y = x W^T + b, he_uniform init,
optional fused tanh-GELU / ReLU; nn/optimized.py:26-325, exported as nn.Linear by nn/__init__.py:4-12);
``StandardLinear`` == nn/linear.py:13-207 (weight (in,out), y = x W + b, xavier_uniform init)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from ..core import Module, Parameter, Tensor
from ..functional import linear as F_linear


def _init_weight(scheme: str, fan_in: int, fan_out: int, shape) -> np.ndarray:
    """Draws from the GLOBAL NumPy RNG with the reference's exact calls so seeded models match."""
    if scheme == "xavier_uniform":
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return np.random.uniform(-lim, lim, shape).astype(np.float32)
    if scheme == "xavier_normal":
        return np.random.normal(0, np.sqrt(2.0 / (fan_in + fan_out)), shape).astype(np.float32)
    if scheme == "he_uniform":
        lim = np.sqrt(6.0 / fan_in)
        return np.random.uniform(-lim, lim, shape).astype(np.float32)
    if scheme == "he_normal":
        return np.random.normal(0, np.sqrt(2.0 / fan_in), shape).astype(np.float32)
    if scheme == "normal":
        return np.random.normal(0, 0.01, shape).astype(np.float32)
    if scheme == "zeros":
        return np.zeros(shape, dtype=np.float32)
    if scheme == "ones":
        return np.ones(shape, dtype=np.float32)
    raise ValueError(f"Unknown weight initialization: {scheme}")


def _init_bias(scheme: str, n: int) -> np.ndarray:
    if scheme == "zeros":
        return np.zeros(n, dtype=np.float32)
    if scheme == "ones":
        return np.ones(n, dtype=np.float32)
    if scheme == "normal":
        return np.random.normal(0, 0.01, n).astype(np.float32)
    if scheme == "uniform":
        return np.random.uniform(-0.1, 0.1, n).astype(np.float32)
    raise ValueError(f"Unknown bias initialization: {scheme}")


class Linear(Module):
    _layout = "out_in"
    _default_init = "he_uniform"

    def __init__(self, in_features: int, out_features: int, bias: bool = True, activation: Optional[str] = None, dtype=None,
                 device=None, enable_fusion: bool = True, enable_jit: bool = True, weight_init: Optional[str] = None,
                 bias_init: str = "zeros", name: Optional[str] = None):
        super().__init__()
        if in_features <= 0 or out_features <= 0:
            raise ValueError("in_features and out_features must be positive")
        self.in_features, self.out_features = in_features, out_features
        self.has_bias = bias
        self.use_bias = bias
        self.activation = activation
        self.enable_fusion = enable_fusion
        shape = (out_features, in_features) if self._layout == "out_in" else (in_features, out_features)
        self.weight = Parameter(_init_weight(weight_init or self._default_init, in_features, out_features, shape), name="weight", device=device)
        self.bias = Parameter(_init_bias(bias_init, out_features), name="bias", device=device) if bias else None

    def forward(self, x: Tensor, residual: Optional[Tensor] = None, out_dtype=None) -> Tensor:
        if not isinstance(x, Tensor):
            x = Tensor(x, device=self.weight.device)
        if x.shape[-1] != self.in_features:
            raise ValueError(f"Input features {x.shape[-1]} != expected {self.in_features}")
        act = self.activation.lower() if isinstance(self.activation, str) else None
        return F_linear(x, self.weight, self.bias, act if act in ("gelu", "relu") else None, residual, out_dtype, self._layout)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.has_bias}, activation={self.activation}"


OptimizedLinear = Linear


class StandardLinear(Linear):
    _layout = "in_out"
    _default_init = "xavier_uniform"
