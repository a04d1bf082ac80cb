"""Learning-rate schedulers with the reference's semantics (optim/lr_scheduler.py:19-606): same class names, constructor
arguments, validation errors and -- epoch by epoch -- the same learning rates, quirks included:

* constructing a scheduler already performs ``step(last_epoch + 1)``, i.e. writes ``optimizer.lr`` (lr_scheduler.py:48);
* ``base_lrs`` is whatever ``optimizer.lr`` was at construction time, so the second scheduler of a chain starts from the
  learning rate the first one's constructor left behind (the warm-up factor 0.1: lr_scheduler.py:566-575, 596-604);
* ``ChainedScheduler`` forwards the GLOBAL epoch to the active scheduler (no offset by the milestone, lr_scheduler.py:528-545).

Everything here is host-side scalar arithmetic in Python floats, evaluated in the reference's operation order so the
values agree bit for bit (tests/test_train_glue.py pins them against traces of the unmodified reference).  The only
B200-specific part is where the number goes: ``optimizer.lr`` is a property whose setter also refreshes the device
scalar the fused Adam/AdamW kernel reads (optim/adam.py), so a scheduler keeps working when the training step is
replayed from a CUDA graph (training.GraphedStep): call ``scheduler.step()`` between replays as in an eager loop.
(Adam, AdamW, Lion and SGD all read the device scalar in capturable mode: nfb_adam_step_ex, nfb_lion_step, nfb_sgd_step_ex;
tests/test_gpu_train_glue.py replays captured steps under StepLR / cosine / plateau schedules.)
"""
from __future__ import annotations

import logging
import math
from typing import List, Optional

from ..core import Optimizer
from .adam import OptimizerError

logger = logging.getLogger(__name__)


def _require_optimizer(optimizer):
    if not isinstance(optimizer, Optimizer):
        raise OptimizerError(f"Expected Optimizer, got {type(optimizer)}")


class LRScheduler:
    """Base: subclasses provide ``_rate(epoch, base_lr)``; ``step()`` advances ``last_epoch`` and writes ``optimizer.lr``."""

    def __init__(self, optimizer, last_epoch: int = -1, verbose: bool = False):
        _require_optimizer(optimizer)
        self.optimizer = optimizer
        self.last_epoch = last_epoch
        self.verbose = verbose
        self.base_lrs = [optimizer.lr] if hasattr(optimizer, "lr") else [0.001]
        self.step(last_epoch + 1)

    # -- state ---------------------------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        return {"last_epoch": self.last_epoch}

    def load_state_dict(self, state_dict: dict) -> None:
        self.last_epoch = state_dict["last_epoch"]

    def get_last_lr(self) -> List[float]:
        return [self.optimizer.lr] if hasattr(self.optimizer, "lr") else [0.001]

    # -- schedule ------------------------------------------------------------------------------------------------
    def _rate(self, epoch: int, base_lr: float) -> float:
        raise NotImplementedError("Subclasses must implement get_lr()")

    def get_lr(self) -> List[float]:
        if type(self)._rate is LRScheduler._rate:
            raise NotImplementedError("Subclasses must implement get_lr()")
        return [self._rate(self.last_epoch, b) for b in self.base_lrs]

    def step(self, epoch: Optional[int] = None) -> None:
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        lrs = self.get_lr()
        if hasattr(self.optimizer, "lr"):
            old = self.optimizer.lr
            self.optimizer.lr = lrs[0]
            if self.verbose:
                logger.info(f"Epoch {self.last_epoch}: lr {old:.8f} -> {lrs[0]:.8f}")


class StepLR(LRScheduler):
    """lr = base * gamma ** (epoch // step_size)   (lr_scheduler.py:96-135)."""

    def __init__(self, optimizer, step_size: int, gamma: float = 0.1, last_epoch: int = -1, verbose: bool = False):
        if step_size <= 0:
            raise OptimizerError(f"step_size must be positive, got {step_size}")
        if gamma <= 0 or gamma > 1:
            raise OptimizerError(f"gamma must be in (0, 1], got {gamma}")
        self.step_size, self.gamma = step_size, gamma
        super().__init__(optimizer, last_epoch, verbose)

    def _rate(self, epoch, base_lr):
        return base_lr if epoch == 0 else base_lr * self.gamma ** (epoch // self.step_size)


class ExponentialLR(LRScheduler):
    """lr = base * gamma ** epoch   (lr_scheduler.py:138-168)."""

    def __init__(self, optimizer, gamma: float, last_epoch: int = -1, verbose: bool = False):
        if gamma <= 0 or gamma > 1:
            raise OptimizerError(f"gamma must be in (0, 1], got {gamma}")
        self.gamma = gamma
        super().__init__(optimizer, last_epoch, verbose)

    def _rate(self, epoch, base_lr):
        return base_lr if epoch == 0 else base_lr * self.gamma ** epoch


class CosineAnnealingLR(LRScheduler):
    """lr = eta_min + (base - eta_min) * (1 + cos(pi * epoch / T_max)) / 2, not clamped past T_max (lr_scheduler.py:171-216)."""

    def __init__(self, optimizer, T_max: int, eta_min: float = 0, last_epoch: int = -1, verbose: bool = False):
        if T_max <= 0:
            raise OptimizerError(f"T_max must be positive, got {T_max}")
        if eta_min < 0:
            raise OptimizerError(f"eta_min must be non-negative, got {eta_min}")
        self.T_max, self.eta_min = T_max, eta_min
        super().__init__(optimizer, last_epoch, verbose)

    def _rate(self, epoch, base_lr):
        if epoch == 0:
            return base_lr
        half_cos = (1 + math.cos(math.pi * (epoch / self.T_max))) / 2
        return self.eta_min + (base_lr - self.eta_min) * half_cos


class LinearLR(LRScheduler):
    """factor goes linearly from start_factor to end_factor over total_iters epochs (lr_scheduler.py:219-266)."""

    def __init__(self, optimizer, start_factor: float = 1.0 / 3, end_factor: float = 1.0, total_iters: int = 5, last_epoch: int = -1,
                 verbose: bool = False):
        if start_factor <= 0 or end_factor <= 0:
            raise OptimizerError("start_factor and end_factor must be positive")
        if total_iters <= 0:
            raise OptimizerError(f"total_iters must be positive, got {total_iters}")
        self.start_factor, self.end_factor, self.total_iters = start_factor, end_factor, total_iters
        super().__init__(optimizer, last_epoch, verbose)

    def _rate(self, epoch, base_lr):
        if epoch == 0:
            return base_lr * self.start_factor
        if epoch >= self.total_iters:
            return base_lr * self.end_factor
        return base_lr * (self.start_factor + (self.end_factor - self.start_factor) * (epoch / self.total_iters))


class WarmupLR(LRScheduler):
    """factor = warmup_factor + (epoch / warmup_epochs) * (1 - warmup_factor) until warmup_epochs, then 1 (lr_scheduler.py:269-311)."""

    def __init__(self, optimizer, warmup_epochs: int, warmup_factor: float = 0.1, last_epoch: int = -1, verbose: bool = False):
        if warmup_epochs <= 0:
            raise OptimizerError(f"warmup_epochs must be positive, got {warmup_epochs}")
        if warmup_factor <= 0 or warmup_factor > 1:
            raise OptimizerError(f"warmup_factor must be in (0, 1], got {warmup_factor}")
        self.warmup_epochs, self.warmup_factor = warmup_epochs, warmup_factor
        super().__init__(optimizer, last_epoch, verbose)

    def _rate(self, epoch, base_lr):
        if epoch >= self.warmup_epochs:
            return base_lr
        return base_lr * (self.warmup_factor + (epoch / self.warmup_epochs) * (1 - self.warmup_factor))


class PolynomialLR(LRScheduler):
    """lr = (base - end_lr) * (1 - epoch / total_iters) ** power + end_lr   (lr_scheduler.py:314-361)."""

    def __init__(self, optimizer, total_iters: int, power: float = 1.0, end_lr: float = 0.0, last_epoch: int = -1, verbose: bool = False):
        if total_iters <= 0:
            raise OptimizerError(f"total_iters must be positive, got {total_iters}")
        if power <= 0:
            raise OptimizerError(f"power must be positive, got {power}")
        if end_lr < 0:
            raise OptimizerError(f"end_lr must be non-negative, got {end_lr}")
        self.total_iters, self.power, self.end_lr = total_iters, power, end_lr
        super().__init__(optimizer, last_epoch, verbose)

    def _rate(self, epoch, base_lr):
        if epoch == 0:
            return base_lr
        if epoch >= self.total_iters:
            return self.end_lr
        return (base_lr - self.end_lr) * (1 - epoch / self.total_iters) ** self.power + self.end_lr


class ReduceLROnPlateau(LRScheduler):
    """Multiply lr by `factor` after more than `patience` epochs without improvement of the monitored metric
    (lr_scheduler.py:364-499).  Its constructor does NOT touch optimizer.lr; ``step(metrics)`` takes the metric."""

    def __init__(self, optimizer, mode: str = "min", factor: float = 0.1, patience: int = 10, threshold: float = 1e-4,
                 threshold_mode: str = "rel", cooldown: int = 0, min_lr: float = 0, eps: float = 1e-8, verbose: bool = False):
        if mode not in ("min", "max"):
            raise OptimizerError(f"mode must be 'min' or 'max', got {mode}")
        if factor >= 1.0 or factor <= 0:
            raise OptimizerError(f"factor must be in (0, 1), got {factor}")
        if patience < 0:
            raise OptimizerError(f"patience must be non-negative, got {patience}")
        if threshold_mode not in ("rel", "abs"):
            raise OptimizerError(f"threshold_mode must be 'rel' or 'abs', got {threshold_mode}")
        _require_optimizer(optimizer)
        self.mode, self.factor, self.patience = mode, factor, patience
        self.threshold, self.threshold_mode = threshold, threshold_mode
        self.cooldown, self.min_lr, self.eps = cooldown, min_lr, eps
        self.best = None
        self.num_bad_epochs = 0
        self.cooldown_counter = 0
        self.mode_worse = float("inf") if mode == "min" else -float("inf")
        self.optimizer, self.last_epoch, self.verbose = optimizer, -1, verbose
        self.base_lrs = [optimizer.lr] if hasattr(optimizer, "lr") else [0.001]

    def step(self, metrics, epoch: Optional[int] = None) -> None:
        value = float(metrics)
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        if self.is_better(value, self.best):
            self.best, self.num_bad_epochs = value, 0
        else:
            self.num_bad_epochs += 1
        if self.in_cooldown:
            self.cooldown_counter -= 1
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            self._reduce_lr(self.last_epoch)
            self.cooldown_counter = self.cooldown
            self.num_bad_epochs = 0

    def _reduce_lr(self, epoch: int) -> None:
        if not hasattr(self.optimizer, "lr"):
            return
        old = self.optimizer.lr
        new = max(old * self.factor, self.min_lr)
        if old - new > self.eps:
            self.optimizer.lr = new
            if self.verbose:
                logger.info(f"Epoch {epoch}: reducing lr to {new:.8f}")

    def is_better(self, a: float, best: Optional[float]) -> bool:
        if best is None:
            return True
        if self.threshold_mode == "rel":
            return a < best * (1.0 - self.threshold) if self.mode == "min" else a > best * (self.threshold + 1.0)
        return a < best - self.threshold if self.mode == "min" else a > best + self.threshold

    @property
    def in_cooldown(self) -> bool:
        return self.cooldown_counter > 0


class ChainedScheduler(LRScheduler):
    """schedulers[i] is active while epoch < milestones[i]; the last one afterwards (lr_scheduler.py:502-550)."""

    def __init__(self, schedulers: List[LRScheduler], milestones: List[int], verbose: bool = False):
        if len(schedulers) != len(milestones) + 1:
            raise OptimizerError("Number of schedulers must be one more than milestones")
        if not all(a < b for a, b in zip(milestones, milestones[1:])):
            raise OptimizerError("Milestones must be in ascending order")
        self.schedulers, self.milestones = schedulers, milestones
        self.current_scheduler_idx = 0
        super().__init__(schedulers[0].optimizer, -1, verbose)

    def step(self, epoch: Optional[int] = None) -> None:
        if epoch is None:
            epoch = self.last_epoch + 1
        self.current_scheduler_idx = next((i for i, m in enumerate(self.milestones) if epoch < m), len(self.schedulers) - 1)
        self.schedulers[self.current_scheduler_idx].step(epoch)
        self.last_epoch = epoch

    def get_lr(self) -> List[float]:
        return self.schedulers[self.current_scheduler_idx].get_lr()


def get_linear_schedule_with_warmup(optimizer, num_warmup_steps: int, num_training_steps: int, last_epoch: int = -1) -> ChainedScheduler:
    """WarmupLR then LinearLR(1.0 -> 0.001) over the remaining steps (lr_scheduler.py:553-576)."""
    warm = WarmupLR(optimizer, num_warmup_steps, last_epoch=last_epoch)
    decay = LinearLR(optimizer, start_factor=1.0, end_factor=0.001, total_iters=num_training_steps - num_warmup_steps, last_epoch=last_epoch)
    return ChainedScheduler([warm, decay], [num_warmup_steps])


def get_cosine_schedule_with_warmup(optimizer, num_warmup_steps: int, num_training_steps: int, num_cycles: float = 0.5,
                                    last_epoch: int = -1) -> ChainedScheduler:
    """WarmupLR then CosineAnnealingLR(T_max = int(remaining * num_cycles)) (lr_scheduler.py:579-606)."""
    warm = WarmupLR(optimizer, num_warmup_steps, last_epoch=last_epoch)
    cosine = CosineAnnealingLR(optimizer, T_max=int((num_training_steps - num_warmup_steps) * num_cycles), last_epoch=last_epoch)
    return ChainedScheduler([warm, cosine], [num_warmup_steps])
