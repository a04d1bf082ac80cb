"""Gated feed-forward blocks of the reference (nn/modern_activations.py:25-470): ``SwiGLU`` / ``GeGLU``
(hidden = gate * act(up), gate = x W_gate + b_gate, up = x W_up + b_up, out = hidden W_down + b_down; note that the
ACTIVATION is on ``up``) and ``Swish`` (= SiLU).  Same parameter names, shapes ((in, hidden) / (hidden, in)), init order
(W_gate, W_up, W_down ~ N(0, 2/fan_in)) and default hidden size (8/3 * input_dim rounded up to a multiple of 8).

B200 path of SwiGLU: ONE tcgen05 GEMM against the concatenated [W_up | W_gate] (a differentiable concatenation, so each
parameter still receives its own gradient), the fused SwiGLU kernel on the packed projection, one GEMM for W_down.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .. import functional as F
from ..array import B200Array, bfloat16
from ..core import Module, Parameter, Tensor
from .layers import LayerError


def _act_dtype(x: Tensor):
    if isinstance(x.data, B200Array):
        from .. import kernels as K
        if K.get_precision() == "bf16":
            return bfloat16
    return None


def _default_hidden(input_dim: int) -> int:
    return ((int(2 * input_dim * 4 / 3) + 7) // 8) * 8


class _GatedFFN(Module):
    _tag = "glu"

    def __init__(self, input_dim: int, hidden_dim: Optional[int] = None, bias: bool = True, device: Optional[str] = None, dtype: Optional[str] = None):
        super().__init__()
        if input_dim <= 0:
            raise LayerError(f"input_dim must be positive, got {input_dim}")
        if hidden_dim is None:
            hidden_dim = _default_hidden(input_dim)
        self.input_dim, self.hidden_dim, self.use_bias = input_dim, hidden_dim, bias
        t = self._tag
        self.W_gate = Parameter(np.random.randn(input_dim, hidden_dim).astype(np.float32) * np.sqrt(2.0 / input_dim), name=f"{t}.W_gate")
        self.W_up = Parameter(np.random.randn(input_dim, hidden_dim).astype(np.float32) * np.sqrt(2.0 / input_dim), name=f"{t}.W_up")
        self.W_down = Parameter(np.random.randn(hidden_dim, input_dim).astype(np.float32) * np.sqrt(2.0 / hidden_dim), name=f"{t}.W_down")
        if bias:
            self.b_gate = Parameter(np.zeros(hidden_dim, dtype=np.float32), name=f"{t}.b_gate")
            self.b_up = Parameter(np.zeros(hidden_dim, dtype=np.float32), name=f"{t}.b_up")
            self.b_down = Parameter(np.zeros(input_dim, dtype=np.float32), name=f"{t}.b_down")
        else:
            self.b_gate = self.b_up = self.b_down = None

    def _check(self, x: Tensor):
        if x.shape[-1] != self.input_dim:
            raise LayerError(f"Input last dimension {x.shape[-1]} != expected {self.input_dim}")

    def extra_repr(self) -> str:
        return f"input_dim={self.input_dim}, hidden_dim={self.hidden_dim}, bias={self.use_bias}"


class SwiGLU(_GatedFFN):
    """(x W_gate + b_gate) * swish(x W_up + b_up), then W_down  (modern_activations.py:25-305)."""
    _tag = "swiglu"

    def forward(self, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        self._check(x)
        dt = _act_dtype(x)
        w = F.concatenate([self.W_up, self.W_gate], axis=1)                       # packed [activated | multiplier]
        b = F.concatenate([self.b_up, self.b_gate], axis=0) if self.use_bias else None
        packed = F.linear(x, w, b, out_dtype=dt, weight_layout="in_out")
        hidden = F.swiglu_packed(packed)                                           # silu(up) * gate
        out = F.linear(hidden, self.W_down, self.b_down, residual=residual, out_dtype=np.float32 if dt is not None else None,
                       weight_layout="in_out")
        out.name = f"swiglu({x.name or 'tensor'})"
        return out


class GeGLU(_GatedFFN):
    """(x W_gate + b_gate) * gelu_tanh(x W_up + b_up), then W_down  (modern_activations.py:308-396)."""
    _tag = "geglu"

    def forward(self, x: Tensor) -> Tensor:
        self._check(x)
        gate = F.linear(x, self.W_gate, self.b_gate, out_dtype=np.float32 if _act_dtype(x) is not None else None, weight_layout="in_out")
        up = F.linear(x, self.W_up, self.b_up, activation="gelu", out_dtype=np.float32 if _act_dtype(x) is not None else None,
                      weight_layout="in_out")                                     # tanh-GELU fused into the GEMM epilogue
        out = F.linear(F.mul(gate, up), self.W_down, self.b_down, weight_layout="in_out")
        out.name = f"geglu({x.name or 'tensor'})"
        return out


class Swish(Module):
    """x * sigmoid(x)  (modern_activations.py:399-441)."""

    def forward(self, x: Tensor) -> Tensor:
        return F.silu(x)


SiLU = Swish
