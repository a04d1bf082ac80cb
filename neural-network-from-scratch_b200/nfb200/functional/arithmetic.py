"""Broadcasting arithmetic + matmul with gradients (functional/arithmetic.py:14-575 of the reference):
add/sub/mul/div/neg/matmul, upstream gradients reduced to each operand's shape by reduce_gradient.
The same code runs on NumPy arrays (cpu tensors) and B200Arrays (cuda tensors): every operator below is
a kernel launch on the device for the latter."""
from __future__ import annotations

import numpy as np

from ..core.tensor import Tensor, make_result
from .utils import as_tensor, check_same_device, is_scalar, reduce_gradient


def add(a, b) -> Tensor:
    if not isinstance(a, Tensor):
        a, b = b, a
    if is_scalar(b):
        return make_result(a.data + b, [a], lambda g: (g,), "add")
    b = as_tensor(b, a)
    check_same_device(a, b, "add")
    sa, sb = a.shape, b.shape
    return make_result(a.data + b.data, [a, b], lambda g: (reduce_gradient(g, sa), reduce_gradient(g, sb)), "add")


def sub(a, b) -> Tensor:
    if is_scalar(a):
        return make_result(a - b.data, [b], lambda g: (-g,), "rsub")
    if is_scalar(b):
        return make_result(a.data - b, [a], lambda g: (g,), "sub")
    if not isinstance(a, Tensor):
        a = as_tensor(a, b)
    b = as_tensor(b, a)
    check_same_device(a, b, "sub")
    sa, sb = a.shape, b.shape
    return make_result(a.data - b.data, [a, b], lambda g: (reduce_gradient(g, sa), reduce_gradient(-g, sb)), "sub")


def mul(a, b) -> Tensor:
    if not isinstance(a, Tensor):
        a, b = b, a
    if is_scalar(b):
        return make_result(a.data * b, [a], lambda g: (g * b,), "mul")
    b = as_tensor(b, a)
    check_same_device(a, b, "mul")
    ad, bd = a.data, b.data
    return make_result(ad * bd, [a, b], lambda g: (reduce_gradient(g * bd, ad.shape), reduce_gradient(g * ad, bd.shape)), "mul")


def div(a, b) -> Tensor:
    if is_scalar(b):
        if b == 0:
            raise ValueError("Division by zero detected")
        return make_result(a.data / b, [a], lambda g: (g / b,), "div")
    if is_scalar(a):
        bd = b.data
        return make_result(a / bd, [b], lambda g: (-g * a / (bd * bd),), "rdiv")
    if not isinstance(a, Tensor):
        a = as_tensor(a, b)
    b = as_tensor(b, a)
    check_same_device(a, b, "div")
    ad, bd = a.data, b.data
    if isinstance(bd, np.ndarray) and np.any(bd == 0):
        raise ValueError("Division by zero detected")  # functional/arithmetic.py:319 (host tensors only: no device sync)
    return make_result(ad / bd, [a, b],
                       lambda g: (reduce_gradient(g / bd, ad.shape), reduce_gradient(-g * ad / (bd * bd), bd.shape)), "div")


def neg(a: Tensor) -> Tensor:
    return make_result(-a.data, [a], lambda g: (-g,), "neg")


def power(a: Tensor, p) -> Tensor:
    if not is_scalar(p):
        raise NotImplementedError("power: only scalar exponents are differentiable here")
    ad = a.data
    return make_result(ad ** p, [a], lambda g: (g * (p * ad ** (p - 1)),), "pow")


def _swap_last(x):
    return x.swapaxes(-1, -2)


def matmul(a, b) -> Tensor:
    """np.matmul semantics, ndim >= 2 required; dA = G.swap(B), dB = swap(A).G, broadcast-reduced."""
    if not isinstance(a, Tensor):
        a = as_tensor(a, b)
    b = as_tensor(b, a)
    check_same_device(a, b, "matmul")
    if a.ndim < 2 or b.ndim < 2:
        raise ValueError(f"matmul requires 2D+ tensors, got shapes {a.shape} and {b.shape}")
    if a.shape[-1] != b.shape[-2]:
        raise ValueError(f"Incompatible matrix dimensions: {a.shape} @ {b.shape}")
    ad, bd = a.data, b.data

    def backward(g):
        ga = reduce_gradient(g @ _swap_last(bd), ad.shape) if a.requires_grad else None
        gb = reduce_gradient(_swap_last(ad) @ g, bd.shape) if b.requires_grad else None
        return ga, gb

    return make_result(ad @ bd, [a, b], backward, "matmul")
