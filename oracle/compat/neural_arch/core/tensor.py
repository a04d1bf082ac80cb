from nfb200.core.tensor import *  # noqa: F401,F403
from nfb200.core.tensor import GradientFunction, Shape, Tensor, TensorLike, _grad_enabled, enable_grad, is_grad_enabled, logger, no_grad
