"""Flat parameter / gradient / moment arenas in HBM.

The reference optimizers loop over parameters in Python and make ~12 NumPy temporaries per tensor
(optim/adam.py:165-252).  Here every cuda parameter of an optimizer is re-homed into ONE contiguous fp32
buffer (parameter.data becomes a view), with matching flat gradient and moment buffers, so that
  * the fused Adam/AdamW kernel updates all parameters in one launch (28 B/param, or 44 B/param with fp64 moments),
  * zero_grad is one memset,
  * DistributedDataParallel all-reduces contiguous ~25 MB slices of the gradient arena (buckets) with no packing copy,
  * the bf16 shadow weights used by the tcgen05 GEMMs are refreshed by the same optimizer kernel.
Tied parameters are de-duplicated by identity.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .._lib import call
from ..array import B200Array, bfloat16

ALIGN = 64  # elements: every parameter starts on a 256-byte boundary


class ParamArena:
    def __init__(self, params: List, moment_dtype=np.float32, n_moments: int = 2, with_bf16: bool = False):
        uniq, seen = [], set()
        for p in params:
            if id(p) not in seen and isinstance(p.data, B200Array):
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        self.offsets: Dict[int, int] = {}
        off = 0
        for p in uniq:
            self.offsets[id(p)] = off
            off += (p.size + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        homes = {id(a): a for a in (getattr(p, "_arena", None) for p in uniq)}
        existing = next(iter(homes.values())) if len(homes) == 1 else None
        self.subset_of = None
        if len(homes) > 1:
            # some parameters already live in an arena (a DistributedDataParallel wrapper's, or another optimizer's) and others
            # elsewhere: re-homing them would silently cut the buckets / the other optimizer off from the live data
            raise RuntimeError("ParamArena: the parameters belong to different flat arenas (or only some of them to one); build the "
                               "optimizer(s) over parameters of ONE DistributedDataParallel module / one earlier optimizer, or before wrapping")
        if existing is not None:
            # parameters already live in an arena (e.g. created by DistributedDataParallel): share its buffers and offsets.  A
            # SUBSET (decay / no-decay groups given to two optimizers) works on the same flat buffers: it only ever touches
            # the runs of its own parameters; its moment buffers span the whole arena so the offsets stay valid.
            self.flat_p, self.flat_g, self.flat_bf16 = existing.flat_p, existing.flat_g, existing.flat_bf16
            self.offsets = existing.offsets
            self.total = existing.total
            if len(existing.params) != len(uniq):
                self.subset_of = existing
        else:
            self.flat_p = B200Array.full((max(off, 1),), 0.0, np.float32)
            self.flat_g = B200Array.full((max(off, 1),), 0.0, np.float32)
            self.flat_bf16 = B200Array.full((max(off, 1),), 0.0, bfloat16) if with_bf16 else None
            for p in uniq:
                o = self.offsets[id(p)]
                view = self.flat_p[o:o + p.size].reshape(p.shape)
                view._assign(p.data if p.data.dtype == np.dtype(np.float32) else p.data.astype(np.float32))
                old_grad = p._grad
                p._data = view
                p._arena_bound = True
                p._arena = self
                p._grad_slot = self.flat_g[o:o + p.size].reshape(p.shape)
                p._slot_zero = True
                p._grad = None
                if old_grad is not None:
                    p._grad_slot._assign(old_grad)
                    p._grad = p._grad_slot
                    p._slot_zero = False
                if self.flat_bf16 is not None:
                    p._bf16 = self.flat_bf16[o:o + p.size].reshape(p.shape)
                    p._bf16_live = True
            if self.flat_bf16 is not None and self.total:
                call("nfb_cast", 0, 5, self.flat_p.ptr, self.flat_bf16.ptr, self.total)
        if with_bf16 and self.flat_bf16 is None:
            self.enable_bf16()
        self.moments = [B200Array.full((max(self.total, 1),), 0.0, moment_dtype) for _ in range(n_moments)]

    def enable_bf16(self):
        if self.flat_bf16 is None:
            owner = self.subset_of if self.subset_of is not None else self
            if owner.flat_bf16 is None:
                owner.flat_bf16 = B200Array.empty((max(self.total, 1),), bfloat16)
                call("nfb_cast", 0, 5, self.flat_p.ptr, owner.flat_bf16.ptr, self.total)
            self.flat_bf16 = owner.flat_bf16
            for p in owner.params:
                o = self.offsets[id(p)]
                p._bf16 = self.flat_bf16[o:o + p.size].reshape(p.shape)
                p._bf16_live = True

    def view(self, flat: B200Array, p) -> B200Array:
        o = self.offsets[id(p)]
        return flat[o:o + p.size].reshape(p.shape)

    def zero_grad(self):
        if self.total and self.subset_of is None:
            call("nfb_memset", self.flat_g.ptr, 0, self.total * 4)
        elif self.total:
            for off, n in self.runs(lambda q: True):   # a subset clears only its own runs of the shared gradient buffer
                call("nfb_memset", self.flat_g.ptr + off * 4, 0, n * 4)
        for p in self.params:
            p._grad = None
            p._slot_zero = True

    def runs(self, predicate):
        """Merge parameters that satisfy `predicate` into maximal contiguous (offset, length) runs."""
        out = []
        for p in self.params:
            if not predicate(p):
                continue
            o = self.offsets[id(p)]
            n = (p.size + ALIGN - 1) // ALIGN * ALIGN
            if out and out[-1][0] + out[-1][1] == o:
                out[-1][1] += n
            else:
                out.append([o, n])
        return out
