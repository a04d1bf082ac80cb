"""Lion with the reference's semantics (optim/lion.py:25-315): c = b1*m + (1-b1)*g; update = lr*sign(c) + lr*wd*p formed in
fp64, cast to fp32, clipped to +-1 and subtracted in fp32; m = b2*m + (1-b2)*g kept in fp64; non-finite parameters scrubbed.
Same constructor arguments / validation, ``state[name]['momentum_buffer']``, ``momentum``, ``get_lr / set_lr``,
``get_state_dict / load_state_dict``, ``get_statistics``.

cuda parameters live in the flat arena (optim/arena.py) and are updated by ONE fused kernel launch (csrc/optim.cu::k_lion*:
22 B/param with ``moment_dtype="float32"``, 30 B/param with the reference's fp64 momentum); the learning rate is a device
scalar in capturable mode, like Adam's, so schedulers act on CUDA-graph replays.  cpu parameters run the reference's NumPy
formulas."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .._lib import F32 as C_F32, F64 as C_F64, call
from ..array import B200Array
from ..core import Optimizer
from .adam import OptimizerError
from .arena import ParamArena


class Lion(Optimizer):
    _lr = None
    _lr_dev = None

    def __init__(self, parameters, lr: float = 1e-4, beta1: float = 0.9, beta2: float = 0.99, weight_decay: float = 0.0, maximize: bool = False,
                 betas: Optional[tuple] = None, moment_dtype: str = "float64"):
        if betas is not None:
            if len(betas) != 2:
                raise OptimizerError(f"betas must be a tuple of length 2, got {len(betas)}")
            beta1, beta2 = betas
        param_dict = dict(parameters.items()) if hasattr(parameters, "items") else {f"param_{i}": p for i, p in enumerate(parameters)}
        super().__init__(param_dict, lr=lr, beta1=beta1, beta2=beta2, weight_decay=weight_decay, maximize=maximize)
        if not 0.0 <= lr:
            raise OptimizerError(f"Invalid learning rate: {lr}")
        if not 0.0 <= beta1 < 1.0:
            raise OptimizerError(f"Invalid beta1: {beta1}")
        if not 0.0 <= beta2 < 1.0:
            raise OptimizerError(f"Invalid beta2: {beta2}")
        if not 0.0 <= weight_decay:
            raise OptimizerError(f"Invalid weight decay: {weight_decay}")
        self.lr, self.beta1, self.beta2, self.weight_decay, self.maximize = lr, beta1, beta2, weight_decay, maximize
        self.step_count = 0
        self.moment_dtype = np.dtype(moment_dtype)
        self.grad_scale = None       # optional device scalar multiplied into every gradient (clip_grad_norm(optimizer=...))
        self.capturable = False
        dev = [p for p in self.parameters.values() if isinstance(p.data, B200Array)]
        self.arena = None
        if dev:
            from .. import kernels as K
            self.arena = ParamArena(dev, self.moment_dtype, 1, with_bf16=(K.get_precision() == "bf16"))
        self.state: Dict[str, dict] = {}
        self.momentum = {}
        seen = {}
        for name, p in self.parameters.items():
            if id(p) in seen:
                self.state[name] = self.state[seen[id(p)]]
                continue
            seen[id(p)] = name
            if self.arena is not None and id(p) in self.arena.offsets:
                buf = self.arena.view(self.arena.moments[0], p)
            else:
                buf = np.zeros(p.shape, dtype=np.float64)
            self.state[name] = {"step": 0, "momentum_buffer": buf}
            self.momentum[name] = buf

    # -- learning rate: host value + device mirror (see optim/adam.py)
    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        self._lr = value
        if self._lr_dev is not None:
            from ..runtime import is_capturing
            if not is_capturing():
                call("nfb_fill", C_F64, self._lr_dev.ptr, 1, float(value))

    def get_lr(self) -> float:
        return self.lr

    def set_lr(self, lr: float) -> None:
        if not 0.0 <= lr:
            raise OptimizerError(f"Invalid learning rate: {lr}")
        self.lr = lr

    def step(self) -> None:
        try:
            self._step()
        finally:
            self.grad_scale = None   # the clip coefficient of clip_grad_norm(optimizer=...) applies to this step only

    def _step(self) -> None:
        self.step_count += 1
        done = set()
        if self.arena is not None:
            a = self.arena
            if self.capturable and self._lr_dev is None:
                self._lr_dev = B200Array.full((1,), float(self._lr), np.float64)
            ids = {id(p) for p in a.params if p.grad is not None}
            mcode = C_F64 if self.moment_dtype == np.dtype(np.float64) else C_F32
            msz = self.moment_dtype.itemsize
            for off, n in a.runs(lambda q: id(q) in ids):
                call("nfb_lion_step", mcode, a.flat_p.ptr + off * 4, a.flat_g.ptr + off * 4, a.moments[0].ptr + off * msz,
                     (a.flat_bf16.ptr + off * 2) if a.flat_bf16 is not None else None, n, float(self._lr), float(self.beta1), float(self.beta2),
                     float(self.weight_decay), 1 if self.maximize else 0, self.grad_scale.ptr if self.grad_scale is not None else None,
                     self._lr_dev.ptr if self._lr_dev is not None else None)
            done |= ids
            for p in a.params:
                if id(p) in done:
                    p._version += 1
        for name, p in self.parameters.items():
            if p.grad is None:
                continue
            st = self.state[name]
            if id(p) in done:
                if not st.get("_counted") == self.step_count:
                    st["step"] += 1
                    st["_counted"] = self.step_count
                continue
            done.add(id(p))
            g = np.asarray(p.grad if not isinstance(p.grad, B200Array) else p.grad.get()).astype(np.float64)
            pd = np.asarray(p.data if not isinstance(p.data, B200Array) else p.data.get())
            if self.maximize:
                g = -g
            m = st["momentum_buffer"]
            st["step"] += 1
            c = self.beta1 * m + (1 - self.beta1) * g
            upd = self.lr * np.sign(c)
            if self.weight_decay != 0:
                upd = upd + self.lr * self.weight_decay * pd.astype(np.float64)
            upd = np.clip(upd.astype(pd.dtype), -1.0, 1.0)
            new = np.asarray(pd - upd, dtype=pd.dtype)
            st["momentum_buffer"] = self.momentum[name] = self.beta2 * m + (1 - self.beta2) * g
            if not np.all(np.isfinite(new)):
                new = np.nan_to_num(new, nan=0.0, posinf=1.0, neginf=-1.0)
            p.data = new

    def zero_grad(self) -> None:
        if self.arena is not None:
            self.arena.zero_grad()
        for p in self.parameters.values():
            if self.arena is None or id(p) not in self.arena.offsets:
                p.zero_grad()

    # -- state (optim/lion.py:215-250): momentum buffers as fp64 host arrays
    def get_state_dict(self) -> Dict:
        def host(x):
            return (x.get() if isinstance(x, B200Array) else np.asarray(x)).astype(np.float64)
        return {"state": {n: {"step": s["step"], "momentum_buffer": host(s["momentum_buffer"])} for n, s in self.state.items()},
                "param_groups": [{"lr": self.lr, "beta1": self.beta1, "beta2": self.beta2, "weight_decay": self.weight_decay, "maximize": self.maximize}]}

    state_dict = get_state_dict

    def load_state_dict(self, state_dict: Dict) -> None:
        for n, s in state_dict["state"].items():
            if n not in self.state:
                continue
            cur = self.state[n]
            cur["step"] = int(s.get("step", 0))
            if isinstance(cur["momentum_buffer"], B200Array):
                cur["momentum_buffer"]._assign(B200Array.from_numpy(np.asarray(s["momentum_buffer"]).astype(self.moment_dtype)))
            else:
                cur["momentum_buffer"] = self.momentum[n] = np.asarray(s["momentum_buffer"], dtype=np.float64)
        g = state_dict["param_groups"][0]
        self.lr, self.beta1, self.beta2 = g["lr"], g["beta1"], g["beta2"]
        self.weight_decay, self.maximize = g["weight_decay"], g["maximize"]

    def __repr__(self) -> str:
        return f"Lion(lr={self.lr}, beta1={self.beta1}, beta2={self.beta2}, weight_decay={self.weight_decay})"

    def get_statistics(self) -> Dict:
        def host(x):
            return x.get() if isinstance(x, B200Array) else np.asarray(x)
        stats = {"lr": self.lr, "num_parameters": len(self.parameters), "total_steps": max((s["step"] for s in self.state.values()), default=0)}
        gn = [np.linalg.norm(host(p.grad)) for p in self.parameters.values() if p.grad is not None]
        pn = [np.linalg.norm(host(p.data)) for p in self.parameters.values()]
        mn = [np.linalg.norm(host(s["momentum_buffer"])) for s in self.state.values()]
        for tag, v in (("grad", gn), ("param", pn), ("momentum", mn)):
            if v:
                stats.update({f"avg_{tag}_norm": np.mean(v), f"max_{tag}_norm": np.max(v), f"min_{tag}_norm": np.min(v)})
        return stats
