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
    """nn/activation.py:14-22 of the reference spells the axis `dim`; `axis` is kept as a synonym."""

    def __init__(self, dim: int = -1, axis: int = None):
        super().__init__()
        self.dim = self.axis = dim if axis is None else axis

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
    """A list of sub-modules that the parent sees (nn/container.py:8-171): registered under "0", "1", ... so parameters(),
    state_dict() and .to() reach them; the whole list protocol (index / slice / assign / delete / + / += / append / extend /
    insert) with the reference's type and range errors.  Entries are re-registered after every structural change."""

    def __init__(self, modules=None):
        super().__init__()
        object.__setattr__(self, "_modules_list", [])
        if modules is not None:
            self.extend(modules)

    # ---- bookkeeping
    @staticmethod
    def _check(module):
        if not isinstance(module, Module):
            raise TypeError(f"ModuleList can only contain Module instances, got {type(module)}")
        return module

    def _index(self, idx) -> int:
        if isinstance(idx, bool) or not isinstance(idx, int):
            raise TypeError(f"ModuleList indices must be integers, not {type(idx).__name__}")
        n = len(self._modules_list)
        if not -n <= idx < n:
            raise IndexError("ModuleList index out of range")
        return idx + n if idx < 0 else idx

    def _reregister(self):
        mods = self.__dict__["_modules"]
        for key in [k for k in mods if k.isdigit()]:
            mods.pop(key)
            self.__dict__.pop(key, None)
        for i, m in enumerate(self._modules_list):
            setattr(self, str(i), m)

    # ---- list protocol
    def __len__(self) -> int:
        return len(self._modules_list)

    def __iter__(self):
        return iter(self._modules_list)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return type(self)(self._modules_list[idx]) if type(self) is ModuleList else ModuleList(self._modules_list[idx])
        return self._modules_list[self._index(idx)]

    def __setitem__(self, idx, module) -> None:
        self._check(module)
        i = self._index(idx)
        self._modules_list[i] = module
        setattr(self, str(i), module)

    def __delitem__(self, idx) -> None:
        if isinstance(idx, slice):
            del self._modules_list[idx]
        else:
            del self._modules_list[self._index(idx)]
        self._reregister()

    def __iadd__(self, modules):
        return self.extend(modules)

    def __add__(self, other):
        out = ModuleList(self._modules_list)
        out.extend(other)
        return out

    def append(self, module):
        self._check(module)
        setattr(self, str(len(self._modules_list)), module)
        self._modules_list.append(module)
        return self

    def extend(self, modules):
        if isinstance(modules, Module) and not isinstance(modules, ModuleList):
            raise TypeError(f"ModuleList.extend should be called with an iterable, but got {type(modules).__name__}")
        try:
            items = list(modules)
        except TypeError:
            raise TypeError(f"ModuleList.extend should be called with an iterable, but got {type(modules).__name__}") from None
        for m in items:
            self.append(m)
        return self

    def insert(self, index: int, module) -> None:
        self._check(module)
        if isinstance(index, bool) or not isinstance(index, int):
            raise TypeError(f"ModuleList indices must be integers, not {type(index).__name__}")
        self._modules_list.insert(index, module)
        self._reregister()

    def forward(self, x):
        """Chains the entries (nn/container.py:163-171); a ModuleList is normally iterated by its owner instead."""
        for m in self._modules_list:
            x = m(x)
        return x


class Sequential(ModuleList):
    """Modules applied one after the other (nn/container.py:174-197); a ModuleList, as upstream."""

    def __init__(self, *modules):
        super().__init__()
        for m in modules:
            self.append(m)
