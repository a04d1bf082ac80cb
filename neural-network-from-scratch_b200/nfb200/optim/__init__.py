from .adam import Adam, AdamW, OptimizerError
from .arena import ParamArena
from .clip import clip_grad_norm
from .sgd import SGD
