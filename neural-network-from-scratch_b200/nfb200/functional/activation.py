"""Activations with gradients (functional/activation.py:14-676 of the reference): relu, softmax, sigmoid, tanh,
gelu (exact erf by default, tanh form with approximate=True), silu, swiglu.  On cuda tensors each forward and
each backward is ONE vectorised kernel (csrc/elementwise.cu, csrc/softmax_ce.cu)."""
from __future__ import annotations

import math

import numpy as np

from ..array import B200Array
from ..core.tensor import Tensor, make_result

try:
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


def _dev(x: Tensor) -> bool:
    return isinstance(x.data, B200Array)


def _host_sigmoid(x):
    x = np.asarray(x)
    if x.dtype.kind != "f":          # integer / bool input: NumPy's own promotion (float64), not a truncating empty_like
        x = x.astype(np.float64)
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    e = np.exp(x[~pos])
    out[~pos] = e / (1.0 + e)
    return out


# name -> (host forward, host backward(g, src), src is output?)
_HOST = {
    "relu": (lambda x: np.maximum(0, x), lambda g, x: g * (x > 0), False),
    "tanh": (np.tanh, lambda g, y: g * (1.0 - y * y), True),
    "sigmoid": (_host_sigmoid, lambda g, y: g * y * (1.0 - y), True),
    "gelu_erf": (lambda x: (0.5 * x * (1.0 + _erf(x / np.sqrt(2.0)))).astype(x.dtype),
                 lambda g, x: (g * (0.5 * (1.0 + _erf(x / np.sqrt(2.0))) + x * np.exp(-0.5 * x * x) / np.sqrt(2.0 * np.pi))).astype(x.dtype), False),
    "gelu_tanh": (lambda x: (0.5 * x * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * x ** 3)))).astype(x.dtype),
                  None, False),
    "silu": (lambda x: x * _host_sigmoid(x), lambda g, x: g * (_host_sigmoid(x) * (1.0 + x * (1.0 - _host_sigmoid(x)))), False),   # int x: float64 result
}


def _gelu_tanh_bwd_host(g, x):
    c = np.sqrt(2.0 / np.pi)
    t = np.tanh(c * (x + 0.044715 * x ** 3))
    return (g * (0.5 * (1.0 + t) + 0.5 * x * (1.0 - t * t) * c * (1.0 + 3.0 * 0.044715 * x * x))).astype(x.dtype)


_HOST["gelu_tanh"] = (_HOST["gelu_tanh"][0], _gelu_tanh_bwd_host, False)


def _unary(x: Tensor, op: str, name: str) -> Tensor:
    fwd, bwd, src_is_out = _HOST[op]
    if _dev(x):
        from .. import array as A
        xd = x.data
        y = A.unary(op, xd)
        src = y if src_is_out else xd
        return make_result(y, [x], lambda g: (A.unary_bwd(op, src, g if g.dtype == src.dtype else g.astype(src.dtype)),), name)
    xd = x.data
    y = fwd(xd)
    src = y if src_is_out else xd
    return make_result(y, [x], lambda g: (bwd(g, src),), name)


def relu(x: Tensor) -> Tensor:
    return _unary(x, "relu", "relu")


def tanh(x: Tensor) -> Tensor:
    return _unary(x, "tanh", "tanh")


def sigmoid(x: Tensor) -> Tensor:
    return _unary(x, "sigmoid", "sigmoid")


def silu(x: Tensor) -> Tensor:
    return _unary(x, "silu", "silu")


def gelu(x: Tensor, approximate: bool = False) -> Tensor:
    return _unary(x, "gelu_tanh" if approximate else "gelu_erf", "gelu")


def softmax(x: Tensor, axis: int = -1) -> Tensor:
    """exp(x - max) / max(sum, 1e-8); backward p * (g - sum(g p))  (functional/activation.py:65-147)."""
    xd = x.data
    if _dev(x):
        be = x.backend
        p = be.softmax(xd, axis=axis)
        return make_result(p, [x], lambda g: (be.softmax_backward(p, g, axis=axis),), "softmax")
    e = np.exp(xd - np.max(xd, axis=axis, keepdims=True))
    p = e / np.maximum(np.sum(e, axis=axis, keepdims=True), 1e-8)
    return make_result(p, [x], lambda g: (p * (g - np.sum(g * p, axis=axis, keepdims=True)),), "softmax")


def swiglu(x: Tensor) -> Tensor:
    """functional/activation.py:512-572: split the last dim in two halves [x | gate]; silu(gate) * x, gate clipped to +-500."""
    from .shape import getitem
    from .arithmetic import mul
    h = x.shape[-1] // 2
    a = getitem(x, (Ellipsis, slice(0, h)))
    gate = getitem(x, (Ellipsis, slice(h, 2 * h)))
    gate = make_result(gate.data.clip(-500, 500) if _dev(gate) else np.clip(gate.data, -500, 500), [gate], lambda g: (g,), "clip")
    return mul(silu(gate), a)
