"""Differentiable shape ops: reshape / transpose / indexing / concatenate.  The reference performs these on raw
``.data`` and re-wraps the result in a fresh leaf Tensor, which severs the graph; here they are
ordinary graph nodes so the named models train end to end (the 'connected' mode of SURVEY section 0)."""
from __future__ import annotations

import numpy as np

from ..array import B200Array
from ..core.tensor import Tensor, make_result


def reshape(x: Tensor, shape) -> Tensor:
    old = x.shape
    return make_result(x.data.reshape(tuple(shape)), [x], lambda g: (g.reshape(old),), "reshape")


def transpose(x: Tensor, axes=None) -> Tensor:
    nd = x.ndim
    if axes is None:
        axes = tuple(reversed(range(nd)))
    axes = tuple(a + nd if a < 0 else a for a in axes)
    inv = tuple(int(i) for i in np.argsort(axes))
    return make_result(x.data.transpose(axes), [x], lambda g: (g.transpose(inv),), "transpose")


def swapaxes(x: Tensor, a: int, b: int) -> Tensor:
    ax = list(range(x.ndim))
    ax[a], ax[b] = ax[b], ax[a]
    return transpose(x, ax)


def _zeros_like_data(x: Tensor):
    if isinstance(x.data, B200Array):
        return x.backend.zeros(x.shape, np.float32)
    return np.zeros(x.shape, dtype=x.data.dtype)


def getitem(x: Tensor, key) -> Tensor:
    if isinstance(key, Tensor):
        key = key.data
    out = x.data[key]

    def backward(g):
        full = _zeros_like_data(x)
        if isinstance(full, np.ndarray):
            np.add.at(full, key, g)
        else:
            full[key] = g  # basic (slice/int) indexing only on device: positions are unique
        return (full,)

    return make_result(out, [x], backward, "getitem")


def concatenate(tensors, axis=0) -> Tensor:
    datas = [t.data for t in tensors]
    sizes = [t.shape[axis] for t in tensors]
    out = np.concatenate(datas, axis=axis)

    def backward(g):
        outs, pos = [], 0
        nd = len(g.shape)
        ax = axis + nd if axis < 0 else axis
        for s in sizes:
            sl = [slice(None)] * nd
            sl[ax] = slice(pos, pos + s)
            outs.append(g[tuple(sl)])
            pos += s
        return tuple(outs)

    return make_result(out, list(tensors), backward, "concatenate")
