from .base import Module, Optimizer, Parameter, ParameterView
from .device import Device, DeviceType, get_default_device, set_default_device
from .dtype import DType, get_default_dtype, set_default_dtype
from .tensor import (ACCUMULATED, GradientFunction, Shape, Tensor, TensorLike, enable_grad, get_grad_clip, is_grad_enabled,
                     make_result, no_grad, set_grad_clip)

__all__ = ["Tensor", "Parameter", "Module", "Optimizer", "GradientFunction", "Device", "DeviceType", "DType", "no_grad",
           "enable_grad", "is_grad_enabled", "get_default_device", "set_default_device", "get_default_dtype",
           "set_default_dtype", "set_grad_clip", "get_grad_clip", "make_result", "ACCUMULATED", "ParameterView"]
