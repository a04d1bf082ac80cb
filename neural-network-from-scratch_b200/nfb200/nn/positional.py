"""Positional modules of the reference (nn/positional.py:23-529): SinusoidalPositionalEncoding, RotaryPositionalEmbedding
(interleaved pairs), LearnedPositionalEmbedding -- same constructor arguments, attributes (``pe``, ``inv_freq``,
``cos_cached`` / ``sin_cached``, ``embedding``), validation errors and forward values.  On the B200 the additive encodings
are one broadcast-add kernel against a device-resident table and the rotation is one in-place kernel per tensor
(csrc/attention.cu::k_rope_pairs); all three are ordinary autograd nodes.

Deviation (PARITY.md #17): the reference back-propagates through RoPE by applying the SAME rotation to the gradient
(positional.py:312-361); the gradient of y = R x is R^T g, which is what runs here.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .. import functional as F
from ..array import B200Array
from ..core import Module, Parameter, Tensor
from ..core.tensor import make_result
from .layers import Dropout, LayerError


def _on_device(x: Tensor) -> bool:
    return isinstance(x.data, B200Array)


class _DeviceTableCache:
    """Host table -> device copy, uploaded once per (table identity, row range)."""

    def __init__(self):
        self._cache = {}

    def get(self, key, make_host):
        hit = self._cache.get(key)
        if hit is None:
            hit = B200Array.from_numpy(np.ascontiguousarray(make_host(), dtype=np.float32))
            if len(self._cache) > 16:
                self._cache.clear()
            self._cache[key] = hit
        return hit


class SinusoidalPositionalEncoding(Module):
    """x + PE[start_pos : start_pos + S];  PE(pos, 2i) = sin(pos / base^(2i/d)), PE(pos, 2i+1) = cos(...)  (positional.py:23-137)."""

    def __init__(self, d_model: int, max_len: int = 5000, dropout: float = 0.0, base: float = 10000.0):
        super().__init__()
        if d_model % 2 != 0:
            raise LayerError(f"d_model must be even for sinusoidal PE, got {d_model}")
        if d_model <= 0:
            raise LayerError(f"d_model must be positive, got {d_model}")
        if max_len <= 0:
            raise LayerError(f"max_len must be positive, got {max_len}")
        self.d_model, self.max_len, self.dropout_rate, self.base = d_model, max_len, dropout, base
        position = np.arange(0, max_len, dtype=np.float32).reshape(-1, 1)
        div_term = np.exp(np.arange(0, d_model, 2, dtype=np.float32) * -(np.log(base) / d_model))
        pe = np.zeros((max_len, d_model), dtype=np.float32)
        pe[:, 0::2] = np.sin(position * div_term)
        pe[:, 1::2] = np.cos(position * div_term)
        self.pe = pe
        self.dropout = Dropout(dropout) if dropout > 0 else None
        self._dev = _DeviceTableCache()

    def forward(self, x: Tensor, start_pos: int = 0) -> Tensor:
        _, seq_len, d_model = x.shape
        if d_model != self.d_model:
            raise LayerError(f"Input d_model {d_model} != expected {self.d_model}")
        if start_pos + seq_len > self.max_len:
            raise LayerError(f"Sequence length {start_pos + seq_len} exceeds max_len {self.max_len}")
        if _on_device(x):
            table = self._dev.get((start_pos, seq_len), lambda: self.pe[start_pos:start_pos + seq_len][None])
            out = F.add(x, Tensor(table))
        else:
            out = F.add(x, Tensor(self.pe[start_pos:start_pos + seq_len][None]))
        return self.dropout(out) if self.dropout is not None else out


class RotaryPositionalEmbedding(Module):
    """Interleaved-pair RoPE on q and k: (x[2j], x[2j+1]) rotated by pos * base^(-2j/dim)  (positional.py:140-366)."""

    def __init__(self, dim: int, max_seq_len: int = 2048, base: float = 10000.0, precision: str = "float32"):
        super().__init__()
        if dim % 2 != 0:
            raise LayerError(f"RoPE dimension must be even, got {dim}")
        if dim <= 0:
            raise LayerError(f"RoPE dimension must be positive, got {dim}")
        if max_seq_len <= 0:
            raise LayerError(f"max_seq_len must be positive, got {max_seq_len}")
        self.dim, self.max_seq_len, self.base, self.precision = dim, max_seq_len, base, precision
        dtype = np.float64 if precision == "float64" else np.float32
        self.inv_freq = 1.0 / (base ** (np.arange(0, dim, 2, dtype=dtype) / dim))
        self._precompute_rotation_matrices(max_seq_len, dtype)
        self._dev = _DeviceTableCache()

    def _precompute_rotation_matrices(self, max_seq_len: int, dtype) -> None:
        freqs = np.outer(np.arange(max_seq_len, dtype=dtype), self.inv_freq)
        self.cos_cached = np.cos(freqs).astype(np.float32)
        self.sin_cached = np.sin(freqs).astype(np.float32)
        self._dev = _DeviceTableCache()

    @staticmethod
    def rotate_half(x: np.ndarray) -> np.ndarray:
        out = np.empty_like(x)
        out[..., ::2] = -x[..., 1::2]
        out[..., 1::2] = x[..., ::2]
        return out

    def _tables(self, seq_len: int, start_pos: int):
        if start_pos + seq_len > self.cos_cached.shape[0]:
            self._precompute_rotation_matrices(start_pos + seq_len, np.float32)
        return self.cos_cached[start_pos:start_pos + seq_len], self.sin_cached[start_pos:start_pos + seq_len]

    def apply_rope(self, x, seq_len: int, start_pos: int = 0, inverse: bool = False):
        """Array in, array out (NumPy or device); x: (..., seq_len, dim)."""
        cos, sin = self._tables(seq_len, start_pos)
        if isinstance(x, B200Array):
            from .. import kernels as K
            ct = self._dev.get(("c", start_pos, seq_len), lambda: cos)
            st = self._dev.get(("s", start_pos, seq_len), lambda: sin)
            return K.rope_interleaved_inplace(x.copy(), ct, st, inverse)
        c, s = np.repeat(cos, 2, axis=-1), np.repeat(sin, 2, axis=-1)
        if inverse:
            s = -s
        return x * c + self.rotate_half(x) * s

    def _rotate(self, t: Tensor, seq_len: int, start_pos: int, name: str) -> Tensor:
        return make_result(self.apply_rope(t.data, seq_len, start_pos), [t],
                           lambda g: (self.apply_rope(g, seq_len, start_pos, inverse=True),), name)

    def forward(self, q: Tensor, k: Tensor, start_pos: int = 0) -> Tuple[Tensor, Tensor]:
        if q.shape != k.shape:
            raise LayerError(f"Query and Key shapes must match: {q.shape} vs {k.shape}")
        if q.shape[-1] != self.dim:
            raise LayerError(f"Last dimension {q.shape[-1]} != RoPE dim {self.dim}")
        if q.ndim not in (3, 4):
            raise LayerError(f"RoPE expects (batch, seq, dim) or (batch, heads, seq, dim), got {q.shape}")
        seq_len = q.shape[-2]
        return self._rotate(q, seq_len, start_pos, "rope_q"), self._rotate(k, seq_len, start_pos, "rope_k")

    def extra_repr(self) -> str:
        return f"dim={self.dim}, max_seq_len={self.max_seq_len}, base={self.base}"


class LearnedPositionalEmbedding(Module):
    """x + embedding[start_pos : start_pos + S] with a trainable (max_len, d_model) table ~ N(0, 0.02^2)  (positional.py:369-497)."""

    def __init__(self, max_len: int, d_model: int, dropout: float = 0.0):
        super().__init__()
        if max_len <= 0:
            raise LayerError(f"max_len must be positive, got {max_len}")
        if d_model <= 0:
            raise LayerError(f"d_model must be positive, got {d_model}")
        self.max_len, self.d_model, self.dropout_rate = max_len, d_model, dropout
        self.embedding = Parameter(np.random.normal(0, 0.02, (max_len, d_model)).astype(np.float32), name="learned_pe.embedding")
        self.dropout = Dropout(dropout) if dropout > 0 else None

    def forward(self, x: Tensor, start_pos: int = 0) -> Tensor:
        _, seq_len, d_model = x.shape
        if d_model != self.d_model:
            raise LayerError(f"Input d_model {d_model} != expected {self.d_model}")
        if start_pos + seq_len > self.max_len:
            raise LayerError(f"Sequence length {start_pos + seq_len} exceeds max_len {self.max_len}")
        out = F.add(x, F.getitem(self.embedding, (None, slice(start_pos, start_pos + seq_len))))
        return self.dropout(out) if self.dropout is not None else out

    def extra_repr(self) -> str:
        return f"max_len={self.max_len}, d_model={self.d_model}"


RoPE = RotaryPositionalEmbedding
SinusoidalPE = SinusoidalPositionalEncoding
LearnedPE = LearnedPositionalEmbedding


def create_rope(dim: int, max_seq_len: int = 2048, base: float = 10000.0) -> RoPE:
    return RoPE(dim, max_seq_len, base)


def create_sinusoidal_pe(d_model: int, max_len: int = 5000) -> SinusoidalPE:
    return SinusoidalPE(d_model, max_len)
