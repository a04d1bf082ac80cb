"""sum / mean with gradients."""
from __future__ import annotations

import numpy as np

from ..core.tensor import Tensor, make_result


def _expand(g, shape, axis, keepdims):
    if axis is None:
        return (g.reshape((1,) * len(shape)) if not keepdims else g)
    if not keepdims:
        axes = (axis,) if isinstance(axis, int) else tuple(axis)
        axes = sorted(a + len(shape) if a < 0 else a for a in axes)
        for a in axes:
            g = g.expand_dims(a) if hasattr(g, "expand_dims") else np.expand_dims(g, a)
    return g


def sum(x: Tensor, axis=None, keepdims=False) -> Tensor:  # noqa: A001
    shape = x.shape
    xd = x.data

    def backward(g):
        ge = _expand(g, shape, axis, keepdims)
        return (np.broadcast_to(ge, shape) + 0.0,)

    return make_result(xd.sum(axis=axis, keepdims=keepdims), [x], backward, "sum")


def mean(x: Tensor, axis=None, keepdims=False) -> Tensor:
    shape = x.shape
    xd = x.data
    out = xd.mean(axis=axis, keepdims=keepdims)
    n = x.size // max(int(out.size), 1)

    def backward(g):
        ge = _expand(g, shape, axis, keepdims)
        return (np.broadcast_to(ge, shape) * (1.0 / n),)

    return make_result(out, [x], backward, "mean")
