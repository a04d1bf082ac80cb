"""Losses (functional/loss.py:15-230 of the reference): cross_entropy_loss, mse_loss.
cross_entropy_loss on a cuda tensor is ONE fused kernel producing the per-row loss and the gradient
(p - onehot) * scale in a single pass over the logits (csrc/softmax_ce.cu)."""
from __future__ import annotations

import numpy as np

from ..array import B200Array
from ..core.tensor import Tensor, make_result


def cross_entropy_loss(predictions: Tensor, targets, reduction: str = "mean") -> Tensor:
    """loss = -log(softmax(logits)[target] + 1e-8); 2-D logits; targets are class ids or one-hot rows."""
    if reduction not in ("mean", "sum", "none"):
        raise ValueError(f"Unknown reduction: {reduction}")
    if predictions.ndim != 2:
        raise ValueError("cross_entropy_loss expects logits of shape (batch_size, num_classes)")
    tdata = targets.data if isinstance(targets, Tensor) else targets
    B = predictions.shape[0]
    if isinstance(predictions.data, B200Array):
        be = predictions.backend
        from .. import kernels as K
        gdt = be.bfloat16 if K.get_precision() == "bf16" else None
        loss, dl = be.cross_entropy(predictions.data, tdata, reduction, need_grad=predictions.requires_grad, grad_dtype=gdt)

        def backward(g):
            # dl already carries 1/B for 'mean'; scale by the upstream gradient (a scalar unless reduction='none').
            # loss.backward() seeds with ones: skipping that multiply saves a full pass over the (T,V) gradient.
            from ..core.tensor import is_unit_seed
            if is_unit_seed(g):
                return (dl,)
            d32 = dl if dl.dtype == np.dtype(np.float32) else dl.astype(np.float32)
            if reduction == "none":
                return (d32 * g.reshape(B, 1),)
            return (d32 * g.reshape(()) if g.size == 1 else d32 * g,)

        return make_result(loss, [predictions], backward, "cross_entropy_loss")
    z = predictions.data
    e = np.exp(z - np.max(z, axis=-1, keepdims=True))
    probs = e / np.maximum(np.sum(e, axis=-1, keepdims=True), 1e-8)
    t = np.asarray(tdata)
    t = t.astype(int) if t.ndim == 1 else np.argmax(t, axis=1)
    per_row = -np.log(probs[np.arange(B), t] + 1e-8)
    loss = np.mean(per_row) if reduction == "mean" else (np.sum(per_row) if reduction == "sum" else per_row)

    def backward(g):
        d = probs.copy()
        d[np.arange(B), t] -= 1.0
        if reduction == "mean":
            d = d / B
        if reduction == "none":
            return (d * np.reshape(g, (B, 1)),)
        return (d * (g.item() if np.size(g) == 1 else g),)

    # (host path: a targets Tensor is listed among the graph inputs, as upstream; it never receives a gradient)
    inputs = [predictions, targets] if isinstance(targets, Tensor) else [predictions]
    return make_result(np.asarray(loss, dtype=z.dtype), inputs, backward, "cross_entropy_loss")


def mse_loss(predictions: Tensor, targets, reduction: str = "mean") -> Tensor:
    """mean/sum of (p - t)^2 (functional/loss.py:116-180)."""
    t = targets.data if isinstance(targets, Tensor) else targets
    if isinstance(predictions.data, B200Array) and not isinstance(t, B200Array):
        t = predictions.backend.from_numpy(np.asarray(t, np.float32))
    d = predictions.data - t
    sq = d * d
    n = predictions.size
    if reduction == "mean":
        out, coef = sq.mean(), 2.0 / n
    elif reduction == "sum":
        out, coef = sq.sum(), 2.0
    elif reduction == "none":
        out, coef = sq, 2.0
    else:
        raise ValueError(f"Unknown reduction: {reduction}")
    inputs = [predictions, targets] if isinstance(targets, Tensor) else [predictions]   # a targets Tensor is a graph input too

    def backward(g):
        gp = d * coef * g
        return (gp, -gp) if len(inputs) == 2 and targets.requires_grad else (gp,)

    return make_result(out if not isinstance(out, np.generic) else np.asarray(out), inputs, backward, "mse_loss")
