"""Adam / AdamW with the reference's exact (quirky) semantics -- optim/adam.py:27-252, optim/adamw.py:41-372 and
SURVEY F7: fp64 moment accumulation, epsilon inside the square root, Adam's small-LR boost min(5, 0.1/lr), AdamW's
double bias correction, update clip to +-10, nan_to_num scrub of the parameters.  cuda parameters are updated by
one fused kernel launch over the flat arena (csrc/optim.cu); cpu parameters run the same formulas in NumPy.

`moment_dtype`: "float64" reproduces the reference bit-for-bit-ish (44 B/param of HBM traffic); "float32" is the
throughput mode (28 B/param) whose 10-step trajectories stay within 1e-5 of it (tests/test_gpu_train_parity.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .. import exceptions as _exc
from .._lib import F32 as C_F32, F64 as C_F64, call
from ..array import B200Array
from ..core import Optimizer
from .arena import ParamArena


class OptimizerError(_exc.OptimizerError, ValueError):
    """The reference's OptimizerError (exceptions.py:267-291); also a ValueError, as it has always been here."""


class _AdamBase(Optimizer):
    _variant = 0  # 0 Adam, 1 AdamW
    _lr = None
    _lr_dev = None

    def __init__(self, parameters, lr=0.001, beta1=0.9, beta2=0.999, eps=1e-5, weight_decay=0.0, amsgrad=False, maximize=False,
                 betas: Optional[tuple] = None, moment_dtype: str = "float64"):
        if betas is not None:
            if len(betas) != 2:
                raise OptimizerError(f"betas must be a tuple of length 2, got {len(betas)}")
            beta1, beta2 = betas
        if hasattr(parameters, "items"):
            param_dict = dict(parameters.items())
        else:
            param_dict = {f"param_{i}": p for i, p in enumerate(parameters)}
        super().__init__(param_dict, lr=lr, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad, maximize=maximize)
        if not 0.0 <= lr:
            raise OptimizerError(f"Invalid learning rate: {lr}")
        if not 0.0 <= beta1 < 1.0:
            raise OptimizerError(f"Invalid beta1: {beta1}")
        if not 0.0 <= beta2 < 1.0:
            raise OptimizerError(f"Invalid beta2: {beta2}")
        if not 0.0 <= eps:
            raise OptimizerError(f"Invalid epsilon: {eps}")
        if not 0.0 <= weight_decay:
            raise OptimizerError(f"Invalid weight decay: {weight_decay}")
        if amsgrad and any(isinstance(p.data, B200Array) for p in param_dict.values()):
            raise NotImplementedError("amsgrad is not implemented on the b200 backend (not used by the named configs); host tensors only")
        self._lr_dev = None      # device copy of lr (capturable mode): schedulers act on replays of a captured step
        self._replay_lag = 0     # replays of a captured step not yet reflected in the per-parameter "step" entries
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.weight_decay, self.amsgrad, self.maximize = weight_decay, amsgrad, maximize
        self.step_count = 0
        self.moment_dtype = np.dtype(moment_dtype)
        self.grad_scale = None  # optional device scalar multiplied into every gradient (global-norm clipping)
        self.capturable = False  # True: keep the step counter on the device (CUDA-graph capture of the training step)
        self._step_dev = None
        self._sib = False        # step-in-backward mode (enable_step_in_backward)
        self.state: Dict[str, dict] = {}
        self.m, self.v = {}, {}
        dev_params = [p for p in self.parameters.values() if isinstance(p.data, B200Array)]
        self.arena = None
        if dev_params:
            from .. import kernels as K
            self.arena = ParamArena(dev_params, self.moment_dtype, 2, with_bf16=(K.get_precision() == "bf16"))
        seen = {}
        for name, p in self.parameters.items():
            if id(p) in seen:  # tied parameter registered under a second name: share the state
                self.state[name] = self.state[seen[id(p)]]
                continue
            seen[id(p)] = name
            if self.arena is not None and id(p) in self.arena.offsets:
                ea, es = self.arena.view(self.arena.moments[0], p), self.arena.view(self.arena.moments[1], p)
            else:
                ea, es = np.zeros(p.shape, dtype=np.float64), np.zeros(p.shape, dtype=np.float64)
            self.state[name] = {"step": 0, "exp_avg": ea, "exp_avg_sq": es}
            if amsgrad:   # host tensors only (checked above); float32 like the parameter, as upstream (optim/adam.py:119-120)
                self.state[name]["max_exp_avg_sq"] = np.zeros_like(np.asarray(p.data))
            self.m[name], self.v[name] = ea, es

    # ------------------------------------------------------------------ learning rate (host value + device mirror)
    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        self._lr = value
        if self._lr_dev is not None:
            from ..runtime import is_capturing
            if not is_capturing():   # inside a capture the write would be frozen into the graph
                call("nfb_fill", C_F64, self._lr_dev.ptr, 1, float(value))

    def _device_scalars(self):
        """capturable mode: step counter and learning rate live on the device and are advanced / rewritten between
        replays (GraphedStep), so one captured step stays valid for the whole run."""
        if not self.capturable:
            return
        if self._step_dev is None:
            self._step_dev = B200Array.full((1,), self.step_count - 1, np.int32)
        if self._lr_dev is None:
            self._lr_dev = B200Array.full((1,), float(self._lr), np.float64)
        call("nfb_add_i32", self._step_dev.ptr, 1)

    def _launch(self, lo, n, step, use_grad_scale=True):
        a = self.arena
        mcode = C_F64 if self.moment_dtype == np.dtype(np.float64) else C_F32
        msz = self.moment_dtype.itemsize
        call("nfb_adam_step_ex", self._variant, mcode, a.flat_p.ptr + lo * 4, a.flat_g.ptr + lo * 4,
             a.moments[0].ptr + lo * msz, a.moments[1].ptr + lo * msz,
             (a.flat_bf16.ptr + lo * 2) if a.flat_bf16 is not None else None, n, float(self._lr), float(self.beta1),
             float(self.beta2), float(self.eps), float(self.weight_decay), int(step), 1 if self.maximize else 0,
             self.grad_scale.ptr if (use_grad_scale and self.grad_scale is not None) else None,
             self._step_dev.ptr if self._step_dev is not None else None,
             self._lr_dev.ptr if self._lr_dev is not None else None)

    def _graph_replayed(self, before: bool):
        """GraphedStep bookkeeping around one replay: the host-side step count advances and the built-in schedule (AdamW)
        computes this step's learning rate BEFORE the replay reads it."""
        if before:
            self.step_count += 1
            self._replay_lag += 1
            self._update_learning_rate()

    # ------------------------------------------------------------------ optimizer fused into backward
    def enable_step_in_backward(self, min_run_elems: int = 1 << 21):
        """Update every parameter as soon as its gradient is FINAL, from inside backward(): the autograd engine reports
        when the last consumer of a parameter has run its backward (the data-parallel bucket machinery,
        core/tensor.py), and contiguous runs of finished parameters get their fused Adam/AdamW launch on the side
        stream (kernels.set_wgrad_overlap) right behind their weight-gradient GEMMs.  The HBM-bound optimizer pass
        (30 B/param) then overlaps the tensor-core-bound backward of the earlier layers instead of trailing the step.
        Same arithmetic and the same per-parameter update as step(); only valid when nothing needs ALL gradients
        before the first update (global-norm clipping, gradient accumulation, DistributedDataParallel -- which has its
        own bucket-pipelined step_optimizer).  step() must still be called after backward(): it updates whatever
        is left (unused / host parameters) and closes the step."""
        if self.arena is None:
            return
        for p in self.arena.params:
            if "_ddp" in p.__dict__ and p._ddp is not self:
                raise RuntimeError("enable_step_in_backward: parameters are wrapped by DistributedDataParallel (use ddp.step_optimizer)")
        self._sib = True
        self._sib_min = int(min_run_elems)
        self._sib_names = {}
        for name, p in self.parameters.items():
            self._sib_names.setdefault(id(p), name)
        for p in self.arena.params:
            p._ddp = self
        self._sib_reset()

    def _sib_reset(self):
        self._sib_begun = False
        self._sib_ready = set()
        self._sib_done = set()
        for p in self.arena.params:
            p._ddp_uses = 0

    def _begin_backward(self, counts):   # autograd: {id(p): (p, consumer nodes of p in the graph this backward runs)}
        if not self._sib:
            return
        for p, n in counts.values():
            p._ddp_uses = getattr(p, "_ddp_uses", 0) + n

    def _note_use(self, p):      # (counting happens in _begin_backward)
        pass

    def _note_grad_done(self, p):   # autograd: a consumer node of p has run its backward
        if not self._sib:
            return
        p._ddp_uses -= 1
        if p._ddp_uses == 0 and p._grad is not None and self.grad_scale is None:
            self._sib_ready.add(id(p))
            self._sib_flush(final=False)

    def _sib_begin(self):
        if self._sib_begun:
            return
        self._sib_begun = True
        self.step_count += 1
        self._update_learning_rate()
        self._device_scalars()

    def _sib_flush(self, final: bool):
        """Launch the fused kernel for maximal contiguous runs of ready, not yet updated arena parameters."""
        a = self.arena
        from .arena import ALIGN

        def padded(q):
            return (q.size + ALIGN - 1) // ALIGN * ALIGN
        runs, cur, prev_end = [], [], None
        for p in a.params:
            ok = id(p) in self._sib_ready and id(p) not in self._sib_done
            o = a.offsets[id(p)]
            if ok and cur and prev_end == o:
                cur.append(p)
            else:
                if cur:
                    runs.append(cur)
                cur = [p] if ok else []
            prev_end = o + padded(p)
        if cur:
            runs.append(cur)
        launches = [rn for rn in runs if final or sum(padded(q) for q in rn) >= self._sib_min]
        if not launches:
            return
        from .. import kernels as K
        K.flush_wgrad()   # the queued weight-gradient GEMMs of these parameters precede their update
        self._sib_begin()
        for rn in launches:
            steps = set()
            for p in rn:
                st = self.state[self._sib_names[id(p)]]
                st["step"] += 1
                steps.add(st["step"])
                self._sib_done.add(id(p))
                p._version += 1
            if len(steps) != 1:
                raise RuntimeError("enable_step_in_backward: parameters of one arena run have different step counts")
            step = steps.pop()
            lo = a.offsets[id(rn[0])]
            hi = a.offsets[id(rn[-1])] + (rn[-1].size + ALIGN - 1) // ALIGN * ALIGN

            def launch(lo=lo, hi=hi, step=step):
                self._launch(lo, hi - lo, step, use_grad_scale=False)
            if final or not K._WGRAD_OVERLAP:
                launch()
            else:
                K.run_on_side_stream(launch)

    # ------------------------------------------------------------------ step
    def _update_learning_rate(self):
        pass

    def step(self, segments=None) -> None:
        try:
            self._step(segments)
        finally:
            # a clip coefficient set by clip_grad_norm(optimizer=...) scales the gradients of THIS step only (a CUDA-graph
            # capture has already baked the device pointer into the captured launch)
            self.grad_scale = None

    def _step(self, segments=None) -> None:
        """One optimizer step.  `segments` (optional, from DistributedDataParallel.step_optimizer): a list of
        (lo, hi, event) slices of the gradient arena in the order their all-reduces were launched; the fused kernel
        then runs slice by slice, each fenced only by ITS bucket's event, so the update of early buckets overlaps the
        all-reduce of the late ones."""
        if self._sib and self.arena is not None and self.grad_scale is None:
            # step-in-backward: most (usually all) arena parameters were updated from inside backward()
            from .. import kernels as K
            K.join_side_stream()
            for p in self.arena.params:
                if p._grad is not None and id(p) not in self._sib_done:
                    self._sib_ready.add(id(p))
            self._sib_begin()
            self._sib_flush(final=True)
            done = set(self._sib_done)
            self._sib_reset()
            for name, p in self.parameters.items():
                if id(p) in done or p.grad is None:
                    continue
                done.add(id(p))
                self._host_step(name, p)
            return
        self.step_count += 1
        self._update_learning_rate()
        done = set()
        if self.arena is not None:
            # group the arena parameters that have a gradient by their per-parameter step count (all equal unless some
            # parameters were skipped earlier: the reference skips parameters whose grad is None)
            name_of = {}
            for name, p in self.parameters.items():
                name_of.setdefault(id(p), name)
            steps = {}
            for p in self.arena.params:
                if p.grad is None:
                    continue
                st = self.state[name_of[id(p)]]
                st["step"] += 1
                steps.setdefault(st["step"], set()).add(id(p))
                done.add(id(p))
            # graph-capturable mode: the step count lives on the device and is advanced by a kernel, so a captured
            # step replays with the right bias corrections (all arena parameters share one count in this mode)
            self._device_scalars()
            a = self.arena
            if segments is not None and len(steps) == 1 and len(next(iter(steps.values()))) == len(a.params):
                from ..runtime import compute_stream_wait
                (step, _ids), = steps.items()
                for lo, hi, ev in segments:
                    if ev is not None:
                        compute_stream_wait(ev)
                    self._launch(lo, hi - lo, step)
                steps = {}
            elif segments is not None:
                from ..runtime import compute_stream_wait
                for _lo, _hi, ev in segments:
                    if ev is not None:
                        compute_stream_wait(ev)
            for step, ids in steps.items():
                for off, n in a.runs(lambda q: id(q) in ids):
                    self._launch(off, n, step)
            for p in a.params:
                if id(p) in done:
                    p._version += 1
        for name, p in self.parameters.items():
            if id(p) in done or p.grad is None:
                continue
            done.add(id(p))
            self._host_step(name, p)

    def _host_step(self, name, p):
        st = self.state[name]
        g = p.grad if not isinstance(p.grad, B200Array) else p.grad.get()
        pd = p.data if not isinstance(p.data, B200Array) else p.data.get()
        if self.maximize:
            g = -g
        if self._variant == 0 and self.weight_decay != 0:
            g = g + self.weight_decay * pd
        st["step"] += 1
        t = st["step"]
        st["exp_avg"] = self.beta1 * st["exp_avg"] + (1 - self.beta1) * g.astype(np.float64)
        st["exp_avg_sq"] = self.beta2 * st["exp_avg_sq"] + (1 - self.beta2) * (g.astype(np.float64) ** 2)
        self.m[name], self.v[name] = st["exp_avg"], st["exp_avg_sq"]
        bc1, bc2 = 1 - self.beta1 ** t, 1 - self.beta2 ** t
        if self.amsgrad:   # maximum of the bias-corrected second moments so far, epsilon OUTSIDE the root (optim/adam.py:222-226)
            np.maximum(st["max_exp_avg_sq"], st["exp_avg_sq"] / bc2, out=st["max_exp_avg_sq"], casting="unsafe")
            denom = np.sqrt(st["max_exp_avg_sq"]) + self.eps
        else:
            denom = np.sqrt(st["exp_avg_sq"] / bc2 + self.eps)
        if self._variant == 0:
            eff = self.lr * min(5.0, 0.1 / self.lr) if (0 < self.lr <= 0.02) else self.lr
        else:
            eff = self.lr / bc1
        upd = np.clip((eff * (st["exp_avg"] / bc1) / denom).astype(pd.dtype), -10.0, 10.0)
        if self._variant == 1 and self.weight_decay != 0:
            new = pd * (1 - self.lr * self.weight_decay) - upd
        else:
            new = pd - upd
        if not np.all(np.isfinite(new)):
            new = np.nan_to_num(new, nan=0.0, posinf=1.0, neginf=-1.0)
        p.data = np.asarray(new, dtype=pd.dtype)

    # ------------------------------------------------------------------ small host-side API of optim/adam.py:263-372
    def get_lr(self) -> float:
        return self.lr

    def set_lr(self, lr: float) -> None:
        if not 0.0 <= lr:
            raise OptimizerError(f"Invalid learning rate: {lr}", learning_rate=lr)
        self.lr = lr

    def __repr__(self) -> str:
        return (f"{type(self).__name__}(lr={self.lr}, beta1={self.beta1}, beta2={self.beta2}, eps={self.eps}, "
                f"weight_decay={self.weight_decay}, amsgrad={self.amsgrad})")

    def get_statistics(self) -> Dict:
        """lr, parameter count, steps taken and the avg / max / min L2 norms of the gradients and parameters (one device->host
        copy per tensor for cuda parameters: a monitoring call, not part of the step)."""
        def host(x):
            return x.get() if isinstance(x, B200Array) else np.asarray(x)
        stats = {"lr": self.lr, "num_parameters": len(self.parameters),
                 "total_steps": max((st["step"] for st in self.state.values()), default=0) + self._replay_lag}
        gn = [float(np.linalg.norm(host(p.grad))) for p in self.parameters.values() if p.grad is not None]
        pn = [float(np.linalg.norm(host(p.data))) for p in self.parameters.values()]
        for key, vals in (("grad", gn), ("param", pn)):
            if vals:
                stats.update({f"avg_{key}_norm": float(np.mean(vals)), f"max_{key}_norm": float(np.max(vals)), f"min_{key}_norm": float(np.min(vals))})
        return stats

    def zero_grad(self) -> None:
        if self.arena is not None:
            self.arena.zero_grad()
            if self._sib:
                self._sib_reset()
        for p in self.parameters.values():
            if self.arena is None or id(p) not in self.arena.offsets:
                p.zero_grad()

    # ------------------------------------------------------------------ state (optim/adam.py:279-311)
    def get_state_dict(self) -> dict:
        def host(x):
            return x.get() if isinstance(x, B200Array) else np.asarray(x)
        lag = self._replay_lag
        return {"state": {n: {"step": s["step"] + lag, "exp_avg": host(s["exp_avg"]).astype(np.float64),
                              "exp_avg_sq": host(s["exp_avg_sq"]).astype(np.float64)} for n, s in self.state.items()},
                "param_groups": [{"lr": self.lr, "beta1": self.beta1, "beta2": self.beta2, "eps": self.eps,
                                  "weight_decay": self.weight_decay, "amsgrad": self.amsgrad, "maximize": self.maximize}],
                "step_count": self.step_count}

    state_dict = get_state_dict

    def load_state_dict(self, sd: dict) -> None:
        for n, s in sd.get("state", {}).items():
            if n not in self.state:
                continue
            cur = self.state[n]
            cur["step"] = int(s["step"])
            for k in ("exp_avg", "exp_avg_sq"):
                if isinstance(cur[k], B200Array):
                    cur[k]._assign(B200Array.from_numpy(np.asarray(s[k]).astype(self.moment_dtype)))
                else:
                    cur[k] = np.asarray(s[k], dtype=np.float64)
        g = (sd.get("param_groups") or [{}])[0]
        for k in ("lr", "beta1", "beta2", "eps", "weight_decay"):
            if k in g:
                setattr(self, k, g[k])
        self.step_count = sd.get("step_count", self.step_count)
        self._replay_lag = 0
        if self._step_dev is not None:   # capturable: the device counter follows the loaded state
            call("nfb_fill", 2, self._step_dev.ptr, 1, float(self.step_count))  # 2 = NFB_I32


class Adam(_AdamBase):
    _variant = 0


class AdamW(_AdamBase):
    _variant = 1

    def __init__(self, parameters, lr=0.001, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.01, amsgrad=False, maximize=False,
                 betas=None, lr_scheduler: Optional[str] = None, lr_scheduler_params: Optional[dict] = None,
                 moment_dtype: str = "float64"):
        super().__init__(parameters, lr, beta1, beta2, eps, weight_decay, amsgrad, maximize, betas, moment_dtype)
        self.initial_lr = lr
        self.lr_scheduler = lr_scheduler
        self.lr_scheduler_params = lr_scheduler_params or {}

    def _update_learning_rate(self):
        """Built-in schedules of optim/adamw.py:232-275 (host-side scalars)."""
        s, pr = self.lr_scheduler, self.lr_scheduler_params
        if not s:
            return
        t = self.step_count
        if s == "cosine":
            T, eta = pr.get("T_max", 1000), pr.get("eta_min", 0)
            self.lr = eta + (self.initial_lr - eta) * 0.5 * (1 + np.cos(np.pi * min(t, T) / T))
        elif s == "linear":
            T = pr.get("total_steps", 1000)
            self.lr = self.initial_lr * max(0.0, 1.0 - t / T)
        elif s == "exponential":
            self.lr = self.initial_lr * pr.get("gamma", 0.95) ** t
        elif s == "step":
            self.lr = self.initial_lr * pr.get("gamma", 0.1) ** (t // pr.get("step_size", 30))
        elif s == "warmup_cosine":
            W, T, eta = pr.get("warmup_steps", 100), pr.get("T_max", 1000), pr.get("eta_min", 0)
            if t <= W:
                self.lr = self.initial_lr * t / W
            else:
                prog = min((t - W) / (T - W), 1.0) if T > W else 1.0
                self.lr = eta + (self.initial_lr - eta) * 0.5 * (1 + np.cos(np.pi * prog))
