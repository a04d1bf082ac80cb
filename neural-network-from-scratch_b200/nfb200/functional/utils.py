"""Gradient plumbing helpers (functional/utils.py:33-96 of the reference)."""
from __future__ import annotations

import numbers

import numpy as np

from ..core.tensor import Tensor


def reduce_gradient(grad, shape):
    """Sum a broadcast gradient back to the operand's shape (column sums for bias grads, etc.)."""
    shape = tuple(shape)
    if tuple(grad.shape) == shape:
        return grad
    nd = len(grad.shape)
    lead = nd - len(shape)
    axes = list(range(lead))
    for i, s in enumerate(shape):
        if s == 1 and grad.shape[lead + i] != 1:
            axes.append(lead + i)
    if axes:
        grad = grad.sum(axis=tuple(axes), keepdims=True)
    return grad.reshape(shape)


def as_tensor(x, like: Tensor) -> Tensor:
    """Promote a scalar / ndarray operand to a Tensor on `like`'s device (no gradient)."""
    if isinstance(x, Tensor):
        return x
    if isinstance(x, (numbers.Number, np.generic)):
        t = Tensor.__new__(Tensor)
        Tensor.__init__(t, np.asarray(x, dtype=np.float32), device=like.device)
        return t
    return Tensor(x, device=like.device)


def is_scalar(x) -> bool:
    return isinstance(x, (numbers.Number, np.generic))


def check_same_device(a: Tensor, b: Tensor, op: str):
    if a.device != b.device:
        raise ValueError(f"{op}: tensors must be on the same device, got {a.device} and {b.device}")
