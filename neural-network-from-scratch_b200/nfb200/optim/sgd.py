"""SGD with momentum / dampening / nesterov / L2 (optim/sgd.py:20-115 of the reference)."""
from __future__ import annotations

import numpy as np

from .._lib import call
from ..array import B200Array
from ..core import Optimizer
from .arena import ParamArena


class SGD(Optimizer):
    _lr = None
    _lr_dev = None

    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        """Host value + device mirror (capturable mode), as in optim/adam.py: schedulers act on CUDA-graph replays."""
        self._lr = value
        if self._lr_dev is not None:
            from ..runtime import is_capturing
            if not is_capturing():
                call("nfb_fill", 1, self._lr_dev.ptr, 1, float(value))   # 1 = NFB_F64

    def __init__(self, parameters, lr: float = 0.01, momentum: float = 0.0, weight_decay: float = 0.0, dampening: float = 0.0,
                 nesterov: bool = False, maximize: bool = False):
        if hasattr(parameters, "items"):
            param_dict = parameters if isinstance(parameters, dict) else dict(parameters.items())   # the caller's dict itself, as upstream
        else:
            param_dict = {f"param_{i}": p for i, p in enumerate(parameters)}
        super().__init__(param_dict, lr=lr)
        self.lr, self.momentum, self.weight_decay, self.dampening, self.nesterov, self.maximize = lr, momentum, weight_decay, dampening, nesterov, maximize
        self.step_count = 0
        self.capturable = False
        dev = [p for p in self.parameters.values() if isinstance(p.data, B200Array)]
        self.arena = None
        if dev:
            from .. import kernels as K
            self.arena = ParamArena(dev, np.float32, 1, with_bf16=(K.get_precision() == "bf16"))
        self.momentum_buffer = {}
        self._started = set()
        for name, p in self.parameters.items():
            if momentum != 0:
                if self.arena is not None and id(p) in self.arena.offsets:
                    self.momentum_buffer[name] = self.arena.view(self.arena.moments[0], p)
                else:
                    self.momentum_buffer[name] = None

    def step(self) -> None:
        self.step_count += 1
        done = set()
        if self.arena is not None:
            a = self.arena
            if self.capturable and self._lr_dev is None:
                self._lr_dev = B200Array.full((1,), float(self._lr), np.float64)
            for first in (True, False):
                ids = {id(p) for p in a.params if p.grad is not None and ((id(p) not in self._started) == first)}
                for off, n in a.runs(lambda q: id(q) in ids):
                    call("nfb_sgd_step_ex", a.flat_p.ptr + off * 4, a.flat_g.ptr + off * 4, a.moments[0].ptr + off * 4,
                         (a.flat_bf16.ptr + off * 2) if a.flat_bf16 is not None else None, n, float(self.lr), float(self.momentum),
                         float(self.dampening), float(self.weight_decay), 1 if self.nesterov else 0, 1 if first else 0,
                         1 if self.maximize else 0, self._lr_dev.ptr if self._lr_dev is not None else None)
                done |= ids
            self._started |= done
            for p in a.params:
                if id(p) in done:
                    p._version += 1
        for name, p in self.parameters.items():
            if id(p) in done or p.grad is None:
                continue
            done.add(id(p))
            g, pd = np.asarray(p.grad), np.asarray(p.data)
            if self.maximize:
                g = -g
            if self.weight_decay != 0:
                g = g + self.weight_decay * pd
            if self.momentum != 0:
                buf = self.momentum_buffer.get(name)
                buf = g.copy() if buf is None else self.momentum * buf + (1 - self.dampening) * g
                self.momentum_buffer[name] = buf
                g = g + self.momentum * buf if self.nesterov else buf
            p.data = (pd - self.lr * g).astype(pd.dtype)

    def zero_grad(self) -> None:
        if self.arena is not None:
            self.arena.zero_grad()
        for p in self.parameters.values():
            if self.arena is None or id(p) not in self.arena.offsets:
                p.zero_grad()


SGDMomentum = SGD   # alias kept by the reference (optim/sgd.py:123)
