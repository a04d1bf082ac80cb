"""Global-norm gradient clipping (examples/translation/train_conversational.py:38-54;
distributed/advanced.py:362-388): total = sqrt(sum g^2); if total > max_norm: g *= max_norm / total.
On cuda the sum of squares, the coefficient and the scaling all stay on the device (no host sync): the
coefficient is a device scalar consumed either by a scaling kernel or directly by the fused Adam kernel."""
from __future__ import annotations

import numpy as np

from .._lib import call
from ..array import B200Array


def clip_grad_norm(parameters, max_norm: float = 1.0, optimizer=None, comm=None):
    """Returns the total norm (a host float for cpu tensors, a 1-element device array for cuda tensors).
    With `optimizer` (an Adam/AdamW/SGD of this package) the scaling is deferred into the optimizer kernel.
    With `comm` (a ProcessGroup: distributed.get_backend()) the squared norm is all-reduced (SUM) over the ranks first --
    for gradients that are SHARDED over ranks; data-parallel replicas hold identical averaged gradients and need none."""
    params = list(parameters.values()) if hasattr(parameters, "values") else list(parameters)
    dev = [p for p in params if p.grad is not None and isinstance(p.grad, B200Array)]
    host = [p for p in params if p.grad is not None and not isinstance(p.grad, B200Array)]
    norm = None
    if host:
        total = 0.0
        for p in host:
            total += float(np.sum(np.asarray(p.grad, dtype=np.float64) ** 2))
        if comm is not None:   # every rank holds a shard of the gradients: the global norm needs the sum over ranks
            total = float(np.asarray(comm.all_reduce(np.asarray([total], np.float64)))[0])
        total = float(np.sqrt(total))
        if total > max_norm:
            for p in host:
                p.grad = p.grad * np.float32(max_norm / total)
        norm = total
    if dev:
        total = B200Array.empty((1,), np.float64)
        coef, nrm = B200Array.empty((1,), np.float32), B200Array.empty((1,), np.float32)
        arena = getattr(dev[0], "_arena", None)
        if arena is not None and all(getattr(p, "_arena", None) is arena for p in dev) and len(dev) == len(arena.params):
            call("nfb_sumsq", arena.flat_g.ptr, arena.total, total.ptr, 0)
            spans = [(arena.flat_g.ptr, arena.total)]
        else:
            spans = []
            for i, p in enumerate(dev):
                g = p.grad if p.grad.is_contiguous else p.grad.contiguous()
                call("nfb_sumsq", g.ptr, g.size, total.ptr, 1 if i else 0)
                spans.append((p.grad.ptr, p.grad.size))
        if comm is not None:
            comm.allreduce_buffer(total, 1)    # SUM of the per-rank squared norms, on the compute stream, in place
        call("nfb_clip_coef", total.ptr, float(max_norm), 0.0, coef.ptr, nrm.ptr)
        if optimizer is not None and hasattr(optimizer, "grad_scale"):
            optimizer.grad_scale = coef
        else:
            for ptr, n in spans:
                call("nfb_scale_by_device_scalar", ptr, n, coef.ptr)
        norm = nrm
    return norm
