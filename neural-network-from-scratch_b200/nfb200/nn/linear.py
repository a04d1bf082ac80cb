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
        self.weight = Parameter(self._draw_weight(weight_init or self._default_init, in_features, out_features, shape), name="weight", device=device)
        self.bias = Parameter(self._draw_bias(bias_init, out_features), name="bias", device=device) if bias else None

    # the two initialisation tables differ between the reference's nn/optimized.py (this class) and nn/linear.py (StandardLinear)
    def _draw_weight(self, scheme: str, fan_in: int, fan_out: int, shape) -> np.ndarray:
        return _init_weight(scheme, fan_in, fan_out, shape)

    def _draw_bias(self, scheme: str, n: int) -> np.ndarray:
        return _init_bias(scheme, n)

    def forward(self, x: Tensor, residual: Optional[Tensor] = None, out_dtype=None, rope=None) -> Tensor:
        if not isinstance(x, Tensor):
            x = Tensor(x, device=self.weight.device)
        if x.shape[-1] != self.in_features:
            raise ValueError(f"Input features {x.shape[-1]} != expected {self.in_features}")
        act = self.activation.lower() if isinstance(self.activation, str) else None
        return F_linear(x, self.weight, self.bias, act if act in ("gelu", "relu") else None, residual, out_dtype, self._layout, rope)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.has_bias}, activation={self.activation}"


OptimizedLinear = Linear


class StandardLinear(Linear):
    """nn/linear.py:13-275 of the reference: weight (in, out), y = x W + b, xavier_uniform, `LayerError` for invalid sizes or a
    wrong input width, a layer `name` that prefixes the parameter names, and the small monitoring API (`reset_parameters`,
    `weight_norm`, `bias_norm`, `get_weight_stats`).  Runs on the same fused Linear entry point as `Linear`."""
    _layout = "in_out"
    _default_init = "xavier_uniform"

    def __init__(self, in_features: int, out_features: int, bias: bool = True, *args, name: Optional[str] = None, **kwargs):
        from .layers import LayerError
        for what, n in (("in_features", in_features), ("out_features", out_features)):
            if n <= 0:
                raise LayerError(f"{what} must be positive, got {n}", layer_name=name, layer_type="Linear")
        super().__init__(in_features, out_features, bias, *args, name=name, **kwargs)
        self.name = name or f"Linear({in_features}, {out_features})"
        self.weight.name = f"{self.name}.weight"
        if self.bias is not None:
            self.bias.name = f"{self.name}.bias"

    def forward(self, x: Tensor, residual: Optional[Tensor] = None, out_dtype=None) -> Tensor:
        if isinstance(x, Tensor) and x.shape[-1] != self.in_features:
            from .layers import LayerError
            raise LayerError(f"Input feature dimension mismatch: expected {self.in_features}, got {x.shape[-1]}", layer_name=self.name,
                             layer_type="Linear")
        return super().forward(x, residual, out_dtype)

    def _draw_weight(self, scheme: str, fan_in: int, fan_out: int, shape) -> np.ndarray:
        """nn/linear.py:81-146: the shared schemes draw exactly as `_init_weight` does (seeded models stay identical), plus
        lecun_uniform / lecun_normal / uniform(+-0.1) / normal(0, 0.1); an unknown name is a `LayerError`."""
        if scheme in ("xavier_uniform", "xavier_normal", "he_uniform", "he_normal", "zeros", "ones"):
            return _init_weight(scheme, fan_in, fan_out, shape)
        if scheme == "lecun_uniform":
            lim = np.sqrt(3.0 / fan_in)
            return np.random.uniform(-lim, lim, shape).astype(np.float32)
        if scheme == "lecun_normal":
            return np.random.normal(0.0, np.sqrt(1.0 / fan_in), shape).astype(np.float32)
        if scheme == "uniform":
            return np.random.uniform(-0.1, 0.1, shape).astype(np.float32)
        if scheme == "normal":
            return np.random.normal(0.0, 0.1, shape).astype(np.float32)
        from .layers import LayerError
        raise LayerError(f"Unknown weight initialization scheme: {scheme}", layer_name=getattr(self, "name", None), layer_type="Linear")

    def _draw_bias(self, scheme: str, n: int) -> np.ndarray:
        """nn/linear.py:148-179."""
        if scheme == "zeros":
            return np.zeros(n, dtype=np.float32)
        if scheme == "ones":
            return np.ones(n, dtype=np.float32)
        if scheme == "uniform":
            return np.random.uniform(-0.1, 0.1, n).astype(np.float32)
        if scheme == "normal":
            return np.random.normal(0.0, 0.1, n).astype(np.float32)
        from .layers import LayerError
        raise LayerError(f"Unknown bias initialization scheme: {scheme}", layer_name=getattr(self, "name", None), layer_type="Linear")

    def reset_parameters(self, weight_init: Optional[str] = None, bias_init: Optional[str] = None) -> None:
        """Draw the parameters again with the given schemes (None keeps the current values), nn/linear.py:209-224."""
        if weight_init is not None:
            self.weight.data = self._draw_weight(weight_init, self.in_features, self.out_features, (self.in_features, self.out_features))
        if bias_init is not None and self.bias is not None:
            self.bias.data = self._draw_bias(bias_init, self.out_features)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.use_bias}"

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.extra_repr()})"

    @property
    def weight_norm(self) -> float:
        return float(np.linalg.norm(self.weight.numpy()))

    @property
    def bias_norm(self) -> float:
        return 0.0 if self.bias is None else float(np.linalg.norm(self.bias.numpy()))

    def get_weight_stats(self) -> dict:
        """mean / std / min / max / norm of the weight (and bias): one device->host copy each for cuda parameters."""
        out = {}
        for key, p in (("weight", self.weight), ("bias", self.bias)):
            if p is None:
                continue
            a = p.numpy()
            out.update({f"{key}_mean": float(np.mean(a)), f"{key}_std": float(np.std(a)), f"{key}_min": float(np.min(a)),
                        f"{key}_max": float(np.max(a)), f"{key}_norm": float(np.linalg.norm(a))})
        return out
