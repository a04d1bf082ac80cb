from .adam import Adam, AdamW, OptimizerError
from .arena import ParamArena
from .clip import clip_grad_norm
from . import lr_scheduler
from .lr_scheduler import (ChainedScheduler, CosineAnnealingLR, ExponentialLR, LinearLR, LRScheduler, PolynomialLR, ReduceLROnPlateau,
                           StepLR, WarmupLR, get_cosine_schedule_with_warmup, get_linear_schedule_with_warmup)
from .lion import Lion
from .sgd import SGD, SGDMomentum
