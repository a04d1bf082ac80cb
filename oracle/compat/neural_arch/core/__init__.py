from nfb200.core import *  # noqa: F401,F403
from nfb200.core import (DType, Device, DeviceType, GradientFunction, Module, Optimizer, Parameter, Tensor, enable_grad,
                         get_default_device, get_default_dtype, is_grad_enabled, no_grad, set_default_device, set_default_dtype)
