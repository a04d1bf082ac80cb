"""Embedding, LayerNorm, RMSNorm, activations, Dropout, containers (mirrors of nn/embedding.py, nn/normalization.py,
nn/activation.py, nn/dropout.py, nn/container.py of the reference)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .. import exceptions as _exc
from .. import functional as F
from ..core import Module, Parameter, Tensor


class LayerError(_exc.LayerError, ValueError):
    """Shape / argument errors raised by layers (exceptions.py LayerError in the reference)."""


class Embedding(Module):
    def __init__(self, vocab_size: int, embed_dim: int, name: Optional[str] = None, device=None):
        super().__init__()
        self.vocab_size, self.embed_dim = vocab_size, embed_dim
        self.name = name or f"Embedding({vocab_size}, {embed_dim})"
        scale = 1.0 / np.sqrt(embed_dim)
        self.weight = Parameter(np.random.uniform(-scale, scale, (vocab_size, embed_dim)).astype(np.float32),
                                name=f"{self.name}.weight", device=device)

    def forward(self, indices) -> Tensor:
        return F.embedding(self.weight, indices)


class LayerNorm(Module):
    def __init__(self, normalized_shape: int, eps: float = 1e-5, elementwise_affine: bool = True, bias: bool = True, device=None):
        super().__init__()
        if normalized_shape <= 0:
            raise LayerError(f"normalized_shape must be positive, got {normalized_shape}")
        if eps <= 0:
            raise LayerError(f"eps must be positive, got {eps}")
        self.normalized_shape, self.eps, self.elementwise_affine = normalized_shape, eps, elementwise_affine
        if elementwise_affine:
            self.weight = Parameter(np.ones(normalized_shape, dtype=np.float32), name="layernorm.weight", device=device)
            self.bias = Parameter(np.zeros(normalized_shape, dtype=np.float32), name="layernorm.bias", device=device) if bias else None
        else:
            self.weight = None
            self.bias = None

    @property
    def gamma(self):
        return self.weight

    @property
    def beta(self):
        return self.bias

    def forward(self, x: Tensor, out_dtype=None) -> Tensor:
        if x.shape[-1] != self.normalized_shape:
            raise LayerError(f"Input last dimension {x.shape[-1]} != normalized_shape {self.normalized_shape}")
        return F.layer_norm(x, self.weight, self.bias, self.eps, out_dtype)

    def extra_repr(self):
        return f"normalized_shape={self.normalized_shape}, eps={self.eps}, elementwise_affine={self.elementwise_affine}"


class RMSNorm(Module):
    def __init__(self, normalized_shape: int, eps: float = 1e-8, elementwise_affine: bool = True, device=None):
        super().__init__()
        if normalized_shape <= 0:
            raise LayerError(f"normalized_shape must be positive, got {normalized_shape}")
        if eps <= 0:
            raise LayerError(f"eps must be positive, got {eps}")
        self.normalized_shape, self.eps, self.elementwise_affine = normalized_shape, eps, elementwise_affine
        self.weight = Parameter(np.ones(normalized_shape, dtype=np.float32), name="rmsnorm.weight", device=device) if elementwise_affine else None

    def forward(self, x: Tensor, out_dtype=None) -> Tensor:
        if x.shape[-1] != self.normalized_shape:
            raise LayerError(f"Input last dimension {x.shape[-1]} != normalized_shape {self.normalized_shape}")
        return F.rms_norm(x, self.weight, self.eps, out_dtype)


class ReLU(Module):
    def forward(self, x):
        return F.relu(x)


class GELU(Module):
    """True GELU.  (The reference's nn.GELU returns relu(x) -- a placeholder; see PARITY.md.)"""

    def __init__(self, approximate: bool = False):
        super().__init__()
        self.approximate = approximate

    def forward(self, x):
        return F.gelu(x, self.approximate)


class Tanh(Module):
    def forward(self, x):
        return F.tanh(x)


class Sigmoid(Module):
    def forward(self, x):
        return F.sigmoid(x)


class Softmax(Module):
    def __init__(self, axis: int = -1):
        super().__init__()
        self.axis = axis

    def forward(self, x):
        return F.softmax(x, self.axis)


class Dropout(Module):
    """Identity, as in the reference (nn/dropout.py:13-16): training is deterministic."""

    def __init__(self, p: float = 0.5):
        super().__init__()
        self.p = p

    def forward(self, x):
        return x


class ModuleList(Module):
    def __init__(self, modules=None):
        super().__init__()
        self._list = []
        for m in modules or []:
            self.append(m)

    def append(self, module: Module):
        setattr(self, str(len(self._list)), module)
        self._list.append(module)
        return self

    def __iter__(self):
        return iter(self._list)

    def __len__(self):
        return len(self._list)

    def __getitem__(self, i):
        return self._list[i]


class Sequential(Module):
    def __init__(self, *modules):
        super().__init__()
        self._list = []
        for i, m in enumerate(modules):
            setattr(self, str(i), m)
            self._list.append(m)

    def forward(self, x):
        for m in self._list:
            x = m(x)
        return x

    def __iter__(self):
        return iter(self._list)

    def __len__(self):
        return len(self._list)

    def __getitem__(self, i):
        return self._list[i]
