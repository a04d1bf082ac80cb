"""Host (CPU) backend registered as ``"numpy"``: the reference's default backend
(backends/numpy_backend.py:15-299), kept so that ``Device.cpu()`` tensors behave as they do upstream.

It is NOT a fallback for the B200 path: a ``Device.cuda`` tensor never routes here, and the b200
backend raises when its shared library or a GPU is missing (see b200_backend.py).
"""
from __future__ import annotations

import numpy as np

from .plugin import Backend, register_backend


class NumpyBackend(Backend):
    name = "numpy"
    is_available = True
    supports_gradients = False
    available = True  # backwards-compat alias used upstream (numpy_backend.py:25-28)
    float32, float64, int32, int64, bool = np.float32, np.float64, np.int32, np.int64, np.bool_

    # creation -- float32 unless told otherwise (numpy_backend.py:55-69)
    def array(self, data, dtype=None):
        return np.array(data, dtype=dtype)

    def zeros(self, shape, dtype=None):
        return np.zeros(shape, dtype=dtype or np.float32)

    def ones(self, shape, dtype=None):
        return np.ones(shape, dtype=dtype or np.float32)

    def full(self, shape, fill_value, dtype=None):
        return np.full(shape, fill_value, dtype=dtype or np.float32)

    def arange(self, start, stop, step=1.0, dtype=None):
        return np.arange(start, stop, step, dtype=dtype or np.float32)

    # the reference always hands back float32 from random_* (numpy_backend.py:71-93)
    def random_normal(self, shape, mean=0.0, std=1.0, dtype=None):
        return np.random.normal(mean, std, shape).astype(np.float32)

    def random_uniform(self, shape, low=0.0, high=1.0, dtype=None):
        return np.random.uniform(low, high, shape).astype(np.float32)

    # shape
    def reshape(self, x, shape):
        return x.reshape(shape)

    def flatten(self, x):
        return x.flatten()

    def transpose(self, x, axes=None):
        return np.transpose(x, axes)

    def squeeze(self, x, axis=None):
        return np.squeeze(x, axis)

    def expand_dims(self, x, axis):
        return np.expand_dims(x, axis)

    # math
    def add(self, x, y):
        return np.add(x, y)

    def subtract(self, x, y):
        return np.subtract(x, y)

    def multiply(self, x, y):
        return np.multiply(x, y)

    mul = multiply

    def divide(self, x, y):
        return np.divide(x, y)

    def power(self, x, y):
        return np.power(x, y)

    def matmul(self, x, y):
        return np.matmul(x, y)

    def dot(self, x, y):
        return np.dot(x, y)

    def maximum(self, x, y):
        return np.maximum(x, y)

    def minimum(self, x, y):
        return np.minimum(x, y)

    # reductions
    def sum(self, x, axis=None, keepdims=False):
        return np.sum(x, axis=axis, keepdims=keepdims)

    def mean(self, x, axis=None, keepdims=False):
        return np.mean(x, axis=axis, keepdims=keepdims)

    def max(self, x, axis=None, keepdims=False):
        return np.max(x, axis=axis, keepdims=keepdims)

    def min(self, x, axis=None, keepdims=False):
        return np.min(x, axis=axis, keepdims=keepdims)

    def argmax(self, x, axis=None):
        return np.argmax(x, axis=axis)

    def argmin(self, x, axis=None):
        return np.argmin(x, axis=axis)

    # unary
    def exp(self, x):
        return np.exp(x)

    def log(self, x):
        return np.log(x)

    def sqrt(self, x):
        return np.sqrt(x)

    def abs(self, x):
        return np.abs(x)

    def sign(self, x):
        return np.sign(x)

    def tanh(self, x):
        return np.tanh(x)

    def clip(self, x, min_val, max_val):
        return np.clip(x, min_val, max_val)

    def softmax(self, x, axis=-1):
        e = np.exp(x - np.max(x, axis=axis, keepdims=True))
        return e / np.sum(e, axis=axis, keepdims=True)

    # comparisons
    def equal(self, x, y):
        return np.equal(x, y)

    def not_equal(self, x, y):
        return np.not_equal(x, y)

    def less(self, x, y):
        return np.less(x, y)

    def less_equal(self, x, y):
        return np.less_equal(x, y)

    def greater(self, x, y):
        return np.greater(x, y)

    def greater_equal(self, x, y):
        return np.greater_equal(x, y)

    # manipulation
    def concatenate(self, arrays, axis=0):
        return np.concatenate(arrays, axis=axis)

    def stack(self, arrays, axis=0):
        return np.stack(arrays, axis=axis)

    def split(self, x, indices_or_sections, axis=0):
        return np.split(x, indices_or_sections, axis=axis)

    # conversion
    def astype(self, x, dtype):
        return x.astype(dtype)

    def to_numpy(self, x):
        return np.asarray(x)

    def from_numpy(self, x, dtype=None):
        return x.astype(dtype) if dtype else x

    # device
    def to_device(self, x, device):
        if device != "cpu":
            raise ValueError(f"NumPy backend only supports CPU, got {device}")
        return x

    def device_of(self, x):
        return "cpu"

    # utility
    def is_array(self, x):
        return isinstance(x, np.ndarray)

    def shape(self, x):
        return x.shape

    def size(self, x):
        return x.size

    def dtype(self, x):
        return x.dtype

    # advanced
    def einsum(self, equation, *operands):
        return np.einsum(equation, *operands)

    def where(self, condition, x, y):
        return np.where(condition, x, y)

    def unique(self, x, return_counts=False):
        return np.unique(x, return_counts=True) if return_counts else np.unique(x)


register_backend("numpy", NumpyBackend)
