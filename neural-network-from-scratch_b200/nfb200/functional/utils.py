"""Gradient plumbing helpers (functional/utils.py:33-96 of the reference) and the small host-side utilities that live in the
same file upstream (broadcast_tensors, get_broadcast_shape, validate_tensor_operation, ensure_tensor, compute_output_shape,
check_finite_gradients, apply_gradient_clipping, memory_efficient_operation: functional/utils.py:13-31,99-275)."""
from __future__ import annotations

import logging
import numbers

import numpy as np

from ..core.tensor import Tensor

logger = logging.getLogger(__name__)


def reduce_gradient(grad, target_shape, broadcast_shape=None):
    """Sum a broadcast gradient back to the operand's shape (column sums for bias grads, etc.).  The reference passes the
    broadcast shape as a third argument and documents that it may be ignored (functional/utils.py:33-48); so it is."""
    shape = tuple(target_shape)
    if tuple(grad.shape) == shape:
        return grad
    nd = len(grad.shape)
    lead = nd - len(shape)
    axes = list(range(lead))
    for i, s in enumerate(shape):
        if s == 1 and grad.shape[lead + i] != 1:
            axes.append(lead + i)
    if axes:
        grad = grad.sum(axis=tuple(axes), keepdims=True)
    return grad.reshape(shape)


def as_tensor(x, like: Tensor) -> Tensor:
    """Promote a scalar / ndarray operand to a Tensor on `like`'s device (no gradient)."""
    if isinstance(x, Tensor):
        return x
    if isinstance(x, (numbers.Number, np.generic)):
        t = Tensor.__new__(Tensor)
        Tensor.__init__(t, np.asarray(x, dtype=np.float32), device=like.device)
        return t
    return Tensor(x, device=like.device)


def is_scalar(x) -> bool:
    return isinstance(x, (numbers.Number, np.generic))


def check_same_device(a: Tensor, b: Tensor, op: str):
    if a.device != b.device:
        raise ValueError(f"{op}: tensors must be on the same device, got {a.device} and {b.device}")


# ---------------------------------------------------------------------- host-side utilities of functional/utils.py
def _host(t):
    d = t.data if isinstance(t, Tensor) else t
    return d.get() if hasattr(d, "get") and not isinstance(d, np.ndarray) else np.asarray(d)


def broadcast_tensors(*tensors):
    """Host views of the tensors' data broadcast against each other (functional/utils.py:13-31)."""
    try:
        return list(np.broadcast_arrays(*[_host(t) for t in tensors]))
    except ValueError as e:
        raise ValueError(f"Cannot broadcast tensors with shapes {[tuple(t.shape) for t in tensors]}") from e


def get_broadcast_shape(*shapes):
    """NumPy's broadcasting rule on shapes (functional/utils.py:99-135)."""
    if not shapes:
        return ()
    try:
        return tuple(int(s) for s in np.broadcast_shapes(*[tuple(s) for s in shapes]))
    except ValueError as e:
        raise ValueError(f"Cannot broadcast shapes: incompatible at dimension (shapes {list(shapes)})") from e


def validate_tensor_operation(a, b, operation: str) -> None:
    """TypeError unless both operands are Tensors; a device mismatch is only logged, as upstream (functional/utils.py:138-158)."""
    if not isinstance(a, Tensor):
        raise TypeError(f"{operation} requires Tensor inputs, got {type(a)} for first argument")
    if not isinstance(b, Tensor):
        raise TypeError(f"{operation} requires Tensor inputs, got {type(b)} for second argument")
    if a.device != b.device:
        logger.warning(f"{operation}: tensors on different devices ({a.device} vs {b.device})")


def ensure_tensor(value, name: str = "tensor") -> Tensor:
    if isinstance(value, Tensor):
        return value
    try:
        return Tensor(value)
    except Exception as e:
        raise TypeError(f"Cannot convert {type(value)} to Tensor for {name}") from e


def compute_output_shape(input_shape, operation: str, **kwargs):
    """Result shape of the three operations the reference knows here (functional/utils.py:183-216)."""
    input_shape = tuple(input_shape)
    if operation in ("softmax", "relu"):
        return input_shape
    if operation == "mean_pool":
        axis = kwargs.get("axis", 1)
        if axis < 0:
            axis += len(input_shape)
        if axis >= len(input_shape):
            raise ValueError(f"Cannot pool over axis {axis} for shape {input_shape}")
        return input_shape[:axis] + input_shape[axis + 1:]
    raise ValueError(f"Unknown operation: {operation}")


def check_finite_gradients(tensor: Tensor, operation: str) -> None:
    """Log (never raise) when a tensor's gradient holds NaN / inf; one device->host copy for a cuda gradient."""
    if tensor.grad is None:
        return
    g = _host(tensor.grad)
    if np.all(np.isfinite(g)):
        return
    if np.any(np.isnan(g)):
        logger.warning(f"NaN gradients detected in {operation} for tensor {tensor.name}")
    if np.any(np.isinf(g)):
        logger.warning(f"Infinite gradients detected in {operation} for tensor {tensor.name}")


def apply_gradient_clipping(grad, max_norm: float = 10.0):
    """L2-norm clipping of ONE gradient array (functional/utils.py:236-250); optim.clip_grad_norm is the device-side, global form."""
    norm = np.linalg.norm(_host(grad))
    if norm > max_norm:
        logger.debug(f"Clipping gradient: norm {norm:.4f} -> {max_norm}")
        return grad * (max_norm / norm)
    return grad


def memory_efficient_operation(func):
    """Upstream's placeholder decorator (functional/utils.py:253-275): debug logging around the call, nothing else (and,
    like upstream's, it does not copy the wrapped function's name: tests/test_functional_utils_complete.py checks that)."""

    def wrapper(*args, **kwargs):
        logger.debug(f"Starting operation: {func.__name__}")
        try:
            result = func(*args, **kwargs)
        except Exception as e:
            logger.error(f"Error in operation {func.__name__}: {e}")
            raise
        logger.debug(f"Completed operation: {func.__name__}")
        return result

    return wrapper
