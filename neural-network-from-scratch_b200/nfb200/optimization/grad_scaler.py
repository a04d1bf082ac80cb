"""Loss / gradient scaler with the reference's policy (optimization/grad_scaler.py:27-637): ``ScalerConfig``,
``ScalingStrategy`` (dynamic / fixed / adaptive / conservative growth), ``AdvancedGradScaler.scale / unscale_ / step``,
the statistics and state dictionaries, and the helper functions -- same names, arguments and step-by-step behaviour.

What ``unscale_`` does per parameter, in optimizer order (grad_scaler.py:160-183): g / scale; stop at the first parameter
holding a non-finite value (the step is then skipped and the scale backs off); otherwise clip that parameter's own L2
norm to ``gradient_clip_threshold`` (coef = thr / (norm + 1e-6)) and write the gradient back.

On the B200 the gradients of an optimizer live in one flat arena (optim/arena.py), so the whole pass is one
``nfb_grad_unscale`` call (csrc/optim.cu: a statistics kernel -- per-parameter sum of squares in double + first bad
parameter index -- and an apply kernel, both over a per-parameter chunk table at HBM speed) followed by one small
device->host read of (first_bad, norms): the reference API returns a bool the caller branches on, so one
synchronisation per step is inherent.  Host (cpu) tensors run the same arithmetic in NumPy.

Deviation (PARITY.md #15): ``scale(loss)`` returns ``loss * scale`` WITH its autograd history; the reference wraps the
product in a fresh Tensor, which severs the graph (backward through the scaled loss reaches no parameter there).
"""
from __future__ import annotations

import logging
import time
from dataclasses import dataclass
from enum import Enum
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .. import exceptions as _exc
from .._lib import call
from ..array import B200Array
from ..core import Tensor

logger = logging.getLogger(__name__)

_CHUNK = 1 << 14   # elements per block of the unscale kernels (64 KB of fp32)


class NumericalError(_exc.NumericalError, ArithmeticError):
    pass


class ScalingStrategy(Enum):
    DYNAMIC = "dynamic"
    FIXED = "fixed"
    ADAPTIVE = "adaptive"
    CONSERVATIVE = "conservative"


@dataclass
class ScalerConfig:
    init_scale: float = 65536.0
    growth_factor: float = 2.0
    backoff_factor: float = 0.5
    growth_interval: int = 2000
    max_loss_scale: float = 2 ** 24
    min_loss_scale: float = 1.0
    strategy: ScalingStrategy = ScalingStrategy.DYNAMIC
    enabled: bool = True
    consecutive_success_threshold: int = 2000
    overflow_cooldown: int = 500
    gradient_clip_threshold: float = 1.0
    stability_check_interval: int = 100


def _param_list(optimizer):
    if not hasattr(optimizer, "parameters"):
        return None
    params = optimizer.parameters
    return list(params.values()) if hasattr(params, "values") else list(params)


def _grad_array(p):
    """The gradient array of a parameter (ours are arrays; reference-style Tensor-valued grads expose .data)."""
    g = getattr(p, "grad", None)
    if g is None:
        return None
    return g.data if isinstance(g, Tensor) or (hasattr(g, "data") and not isinstance(g, (np.ndarray, B200Array))) else g


def _set_grad(p, new):
    g = p.grad
    if isinstance(g, Tensor) or (hasattr(g, "data") and not isinstance(g, (np.ndarray, B200Array))):
        g.data = new
    else:
        p.grad = new


def _fresh_timing():
    return {"unscale_time": 0.0, "overflow_check_time": 0.0, "step_time": 0.0}


class AdvancedGradScaler:
    device_loss_check = True   # scale(): read a device loss back to raise on non-finite values like the reference does

    def __init__(self, config: Optional[ScalerConfig] = None):
        self.config = config or ScalerConfig()
        self._tables: Dict[tuple, tuple] = {}
        self.reset(log=False)

    # ------------------------------------------------------------------ loss scaling
    def scale(self, loss: Tensor) -> Tensor:
        if not isinstance(loss, Tensor):
            raise TypeError(f"Expected Tensor, got {type(loss)}")
        if not self.config.enabled:
            return loss
        on_device = isinstance(loss.data, B200Array)
        if not on_device or self.device_loss_check:
            host = loss.data.get() if on_device else np.asarray(loss.data)
            if not np.all(np.isfinite(host)):
                raise NumericalError("Loss contains non-finite values before scaling")
            if not np.all(np.isfinite(host * np.asarray(self._scale, dtype=host.dtype))):
                logger.warning(f"Overflow detected in scaled loss with scale {self._scale}")
                self._scale = max(self._scale * self.config.backoff_factor, self.config.min_loss_scale)
        out = loss * float(self._scale)
        out.name = f"scaled_{loss.name or 'loss'}"
        return out

    # ------------------------------------------------------------------ unscale + per-parameter clip
    def _arena_table(self, arena, params):
        """Chunk table {seg, len, start} over the gradient arena for `params` (cached per parameter set)."""
        key = (id(arena), tuple(id(p) for p in params))
        hit = self._tables.get(key)
        if hit is None:
            rec = []
            for seg, p in enumerate(params):
                off, n = arena.offsets[id(p)], p.size
                for s in range(0, n, _CHUNK):
                    rec.append((seg, min(_CHUNK, n - s), off + s))
            tab = np.zeros(len(rec), dtype=np.dtype([("seg", np.int32), ("len", np.int32), ("start", np.int64)]))
            if rec:
                tab["seg"], tab["len"], tab["start"] = zip(*rec)
            dev = B200Array.from_numpy(tab.view(np.int64).reshape(-1)) if rec else None
            hit = (dev, len(rec), B200Array.empty((max(len(params), 1),), np.float64), B200Array.empty((1,), np.int32))
            self._tables = {key: hit}   # one live table: parameter sets do not change between steps
        return hit

    def _unscale_device(self, gbase_ptr, table, nseg):
        dev, nchunks, norms2, first_bad = table
        call("nfb_grad_unscale", gbase_ptr, dev.ptr if dev is not None else None, nchunks, nseg, float(self._scale),
             float(max(self.config.gradient_clip_threshold, 0.0)), norms2.ptr, first_bad.ptr)
        fb = int(first_bad.get()[0])
        return fb, np.sqrt(norms2.get()[:nseg]).astype(np.float32)

    def unscale_(self, optimizer) -> bool:
        t0 = time.time()
        if not self.config.enabled:
            return True
        self._found_inf = False
        params = _param_list(optimizer)
        if params is None:
            raise ValueError("Optimizer must have parameters attribute")
        live, seen = [], set()
        for p in params:
            if id(p) not in seen and _grad_array(p) is not None:
                seen.add(id(p))
                live.append(p)
        norms = []
        arena = getattr(optimizer, "arena", None)
        clip = self.config.gradient_clip_threshold
        if live and arena is not None and all(id(p) in arena.offsets and p._grad is getattr(p, "_grad_slot", None) for p in live):
            fb, nrm = self._unscale_device(arena.flat_g.ptr, self._arena_table(arena, live), len(live))
            self._found_inf = fb < len(live)
            if clip > 0:
                norms = list(nrm[:fb])
        else:
            for p in live:
                g = _grad_array(p)
                if isinstance(g, B200Array):
                    if g.dtype != np.dtype(np.float32) or not g.is_contiguous:
                        g = g.astype(np.float32) if g.dtype != np.dtype(np.float32) else g.contiguous()
                        _set_grad(p, g)
                        g = _grad_array(p)
                    tab = np.zeros(-(-g.size // _CHUNK), dtype=np.dtype([("seg", np.int32), ("len", np.int32), ("start", np.int64)]))
                    for c in range(len(tab)):
                        tab[c] = (0, min(_CHUNK, g.size - c * _CHUNK), c * _CHUNK)
                    dev = B200Array.from_numpy(tab.view(np.int64).reshape(-1)) if len(tab) else None
                    fb, nrm = self._unscale_device(g.ptr, (dev, len(tab), B200Array.empty((1,), np.float64), B200Array.empty((1,), np.int32)), 1)
                    if fb == 0:
                        self._found_inf = True
                        break
                    if clip > 0:
                        norms.append(nrm[0])
                else:
                    u = np.asarray(g) / self._scale
                    if not np.all(np.isfinite(u)):
                        self._found_inf = True
                        break
                    if clip > 0:
                        n = np.linalg.norm(u)
                        norms.append(n)
                        if n > clip:
                            u = u * (clip / (n + 1e-6))
                    _set_grad(p, u)
        if self._found_inf:
            logger.warning("Found inf/nan in gradients during unscaling")
        if norms:
            self._gradient_norms_history.append(np.mean(norms))
            if len(self._gradient_norms_history) > 1000:
                self._gradient_norms_history = self._gradient_norms_history[-1000:]
        self._timing_stats["unscale_time"] += time.time() - t0
        return not self._found_inf

    # ------------------------------------------------------------------ step + scale policy
    def step(self, optimizer) -> bool:
        t0 = time.time()
        self._step_count += 1
        if not self.config.enabled:
            optimizer.step()
            self._total_successful_steps += 1
            return True
        taken = self.unscale_(optimizer)
        if taken:
            optimizer.step()
            self._total_successful_steps += 1
            self._growth_tracker += 1
            self._update_scale_on_success()
        else:
            self._handle_overflow()
        self._scale_history.append(self._scale)
        if len(self._scale_history) > 1000:
            self._scale_history = self._scale_history[-1000:]
        if self._step_count % self.config.stability_check_interval == 0:
            self._check_gradient_stability()
        self._timing_stats["step_time"] += time.time() - t0
        return taken

    def _grow(self, factor):
        self._scale = min(self._scale * factor, self.config.max_loss_scale)
        self._growth_tracker = 0

    def _update_scale_on_success(self):
        cfg, strat = self.config, self.config.strategy
        if strat == ScalingStrategy.DYNAMIC:
            if self._growth_tracker >= cfg.growth_interval:
                self._grow(cfg.growth_factor)
        elif strat == ScalingStrategy.CONSERVATIVE:
            since = self._step_count - self._last_overflow_step
            if self._growth_tracker >= cfg.consecutive_success_threshold and since > cfg.overflow_cooldown:
                self._grow(cfg.growth_factor)
        elif strat == ScalingStrategy.ADAPTIVE:
            if len(self._gradient_norms_history) >= 100:
                recent = self._gradient_norms_history[-100:]
                avg, std = np.mean(recent), np.std(recent)
                if std < 0.1 * avg and avg < 0.1 and self._growth_tracker >= cfg.growth_interval // 2:
                    self._grow(1.5)

    def _handle_overflow(self):
        self._total_overflows += 1
        self._last_overflow_step = self._step_count
        self._growth_tracker = 0
        old = self._scale
        self._scale = max(self._scale * self.config.backoff_factor, self.config.min_loss_scale)
        logger.warning(f"Gradient overflow detected. Reduced scale from {old} to {self._scale}")
        if self._total_overflows % 10 == 0:
            logger.warning(f"Frequent overflows detected ({self._total_overflows} total). Consider adjusting model or learning rate.")

    def _check_gradient_stability(self):
        if len(self._gradient_norms_history) < 50:
            return
        recent = self._gradient_norms_history[-50:]
        ratio = np.std(recent) / (np.mean(recent) + 1e-8)
        if ratio > 2.0:
            logger.warning(f"High gradient instability detected (ratio: {ratio:.2f}). "
                           f"Consider reducing learning rate or using more conservative scaling.")

    def update(self):
        pass

    # ------------------------------------------------------------------ accessors / state
    def get_scale(self) -> float:
        return self._scale

    def set_scale(self, scale: float):
        self._scale = max(scale, self.config.min_loss_scale)

    def get_growth_tracker(self) -> int:
        return self._growth_tracker

    def is_enabled(self) -> bool:
        return self.config.enabled

    def enable(self):
        self.config.enabled = True

    def disable(self):
        self.config.enabled = False

    def get_statistics(self) -> Dict[str, Any]:
        n = max(self._step_count, 1)
        stats = {"current_scale": self._scale, "total_steps": self._step_count, "successful_steps": self._total_successful_steps,
                 "total_overflows": self._total_overflows, "overflow_rate": self._total_overflows / n,
                 "success_rate": self._total_successful_steps / n, "growth_tracker": self._growth_tracker,
                 "strategy": self.config.strategy.value, "enabled": self.config.enabled}
        if self._gradient_norms_history:
            r = self._gradient_norms_history[-100:]
            stats.update(avg_gradient_norm=np.mean(r), gradient_norm_std=np.std(r), max_gradient_norm=np.max(r), min_gradient_norm=np.min(r))
        if self._scale_history:
            stats.update(scale_history_length=len(self._scale_history), max_scale_used=np.max(self._scale_history),
                         min_scale_used=np.min(self._scale_history), scale_changes=len(set(self._scale_history)))
        if self._step_count > 0:
            stats.update(avg_unscale_time=self._timing_stats["unscale_time"] / self._step_count,
                         avg_step_time=self._timing_stats["step_time"] / self._step_count,
                         total_overhead_time=sum(self._timing_stats.values()))
        return stats

    _CONFIG_KEYS = ("init_scale", "growth_factor", "backoff_factor", "growth_interval", "max_loss_scale", "min_loss_scale", "enabled",
                    "consecutive_success_threshold", "overflow_cooldown", "gradient_clip_threshold", "stability_check_interval")

    def state_dict(self) -> Dict[str, Any]:
        cfg = {k: getattr(self.config, k) for k in self._CONFIG_KEYS}
        cfg["strategy"] = self.config.strategy.value
        return {"scale": self._scale, "growth_tracker": self._growth_tracker, "overflow_tracker": self._overflow_tracker,
                "step_count": self._step_count, "total_overflows": self._total_overflows,
                "total_successful_steps": self._total_successful_steps, "last_overflow_step": self._last_overflow_step,
                "gradient_norms_history": self._gradient_norms_history[-100:], "scale_history": self._scale_history[-100:],
                "timing_stats": self._timing_stats.copy(), "config": cfg}

    def load_state_dict(self, state_dict: Dict[str, Any]):
        self._scale = state_dict["scale"]
        self._growth_tracker = state_dict["growth_tracker"]
        self._overflow_tracker = state_dict.get("overflow_tracker", 0)
        self._step_count = state_dict["step_count"]
        self._total_overflows = state_dict["total_overflows"]
        self._total_successful_steps = state_dict["total_successful_steps"]
        self._last_overflow_step = state_dict.get("last_overflow_step", -1)
        self._gradient_norms_history = state_dict.get("gradient_norms_history", [])
        self._scale_history = state_dict.get("scale_history", [])
        self._timing_stats = state_dict.get("timing_stats", _fresh_timing())
        if "config" in state_dict:
            c = dict(state_dict["config"])
            strategy = ScalingStrategy(c.pop("strategy", "dynamic"))
            defaults = ScalerConfig()
            self.config = ScalerConfig(strategy=strategy, **{k: c.get(k, getattr(defaults, k)) for k in self._CONFIG_KEYS})

    def reset(self, log: bool = True):
        self._scale = self.config.init_scale
        self._growth_tracker = 0
        self._overflow_tracker = 0
        self._found_inf = False
        self._last_overflow_step = -1
        self._step_count = 0
        self._total_overflows = 0
        self._total_successful_steps = 0
        self._gradient_norms_history = []
        self._scale_history = []
        self._timing_stats = _fresh_timing()
        if log:
            logger.info("Reset gradient scaler to initial state")


# ---------------------------------------------------------------------- helpers (grad_scaler.py:501-637)
def check_gradients_finite(optimizer) -> Tuple[bool, Dict[str, float]]:
    stats = {"total_params": 0, "params_with_grad": 0, "finite_grads": 0, "infinite_grads": 0, "nan_grads": 0,
             "max_grad_norm": 0.0, "avg_grad_norm": 0.0}
    params = _param_list(optimizer)
    if params is None:
        return False, stats
    norms, ok = [], True
    for p in params:
        stats["total_params"] += 1
        g = _grad_array(p)
        if g is None:
            continue
        stats["params_with_grad"] += 1
        if isinstance(g, B200Array):     # two tiny device reductions per parameter; diagnostics, not the hot path
            s2 = float(np.asarray((g.astype(np.float64) ** 2).sum().get()))
            if np.isnan(s2) or np.isinf(s2):
                h = g.get()
                key = "infinite_grads" if np.any(np.isinf(h)) else "nan_grads"
                stats[key] += 1
                ok = False
                continue
            stats["finite_grads"] += 1
            norms.append(np.float32(np.sqrt(s2)))
        else:
            h = np.asarray(g)
            if np.any(np.isinf(h)):
                stats["infinite_grads"] += 1
                ok = False
            elif np.any(np.isnan(h)):
                stats["nan_grads"] += 1
                ok = False
            else:
                stats["finite_grads"] += 1
                norms.append(np.linalg.norm(h))
    if norms:
        stats["max_grad_norm"] = float(np.max(norms))
        stats["avg_grad_norm"] = float(np.mean(norms))
    return ok, stats


def clip_gradients_by_norm(optimizer, max_norm: float) -> float:
    """Global-norm clipping, coef = max_norm / (total + 1e-6) when total > max_norm; returns the norm before clipping."""
    params = _param_list(optimizer)
    if params is None:
        return 0.0
    dev = [p for p in params if isinstance(_grad_array(p), B200Array)]
    total = 0.0
    if dev:
        acc = B200Array.empty((1,), np.float64)
        for i, p in enumerate(dev):
            g = _grad_array(p)
            g = g if g.is_contiguous else g.contiguous()
            call("nfb_sumsq", g.ptr, g.size, acc.ptr, 1 if i else 0)
        total += float(acc.get()[0])
    for p in params:
        g = _grad_array(p)
        if g is not None and not isinstance(g, B200Array):
            total += np.linalg.norm(np.asarray(g)) ** 2
    total = np.sqrt(total)
    if total > max_norm:
        scale_gradients(optimizer, max_norm / (total + 1e-6))
    return float(total)


def scale_gradients(optimizer, scale_factor: float):
    params = _param_list(optimizer)
    if params is None:
        return
    seen = set()
    for p in params:
        g = _grad_array(p)
        if g is None or id(p) in seen:
            continue
        seen.add(id(p))
        if isinstance(g, B200Array):
            g *= np.float32(scale_factor)      # in place: arena-bound gradients stay in their slot
        else:
            _set_grad(p, np.asarray(g) * scale_factor)


def create_scaler(strategy: str = "dynamic", **kwargs) -> AdvancedGradScaler:
    try:
        strat = ScalingStrategy(strategy.lower())
    except ValueError:
        logger.warning(f"Unknown strategy '{strategy}', using 'dynamic'")
        strat = ScalingStrategy.DYNAMIC
    return AdvancedGradScaler(ScalerConfig(strategy=strat, **kwargs))
