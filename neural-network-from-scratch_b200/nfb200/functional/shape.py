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


def repeat_kv(x: Tensor, n_rep: int) -> Tensor:
    """(B, Hkv, S, D) -> (B, Hkv * n_rep, S, D), head h of the result = head h // n_rep of x -- the explicit key/value
    repetition of grouped-query attention (nn/modern_attention.py:137-161; there on the (B,S,Hkv,D) layout).  Backward sums
    each group.  Only the unfused attention composition needs it: the flash kernels read grouped K/V in place."""
    if n_rep == 1:
        return x
    B, Hkv, S, D = x.shape
    xd = x.data
    if isinstance(xd, B200Array):
        out = xd.reshape(B, Hkv, 1, S, D).broadcast_to((B, Hkv, n_rep, S, D)).contiguous().reshape(B, Hkv * n_rep, S, D)

        def backward(g):
            g32 = g if g.dtype == np.dtype(np.float32) else g.astype(np.float32)
            return (g32.reshape(B, Hkv, n_rep, S, D).sum(axis=2),)
    else:
        out = np.repeat(xd, n_rep, axis=1)

        def backward(g):
            return (g.reshape(B, Hkv, n_rep, S, D).sum(axis=2),)
    return make_result(out, [x], backward, "repeat_kv")


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
