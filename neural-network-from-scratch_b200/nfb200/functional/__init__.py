"""Functional ops with automatic differentiation (mirror of neural_arch.functional)."""
from .activation import gelu, relu, sigmoid, silu, softmax, swiglu, tanh
from .arithmetic import add, div, matmul, mul, neg, power, sub
from .fused import (embedding, layer_norm, linear, packed_qkv_attention, rms_norm, rope, scaled_dot_product_attention, to_float32,
                    swiglu_packed)
from .loss import cross_entropy_loss, mse_loss
from .reduction import mean, sum  # noqa: A004
from .shape import concatenate, getitem, repeat_kv, reshape, swapaxes, transpose
from .utils import reduce_gradient
