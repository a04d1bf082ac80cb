"""Grouped-query / multi-query attention of the reference (nn/modern_attention.py:26-358): ``GroupedQueryAttention``
(``W_q`` (d, H*hd), ``W_k`` / ``W_v`` (d, Hkv*hd), ``W_o`` (H*hd, d), optional biases ``b_q/b_k/b_v/b_o``; same names, init
order and scale sqrt(2/fan_in)), ``MultiQueryAttention`` (= GQA with one key/value head) and ``FlashAttentionConcept``.

Forward, as the reference computes it: q/k/v projections, softmax(q k^T / sqrt(hd) [+ mask]) v with key/value head
h // n_rep serving query head h, output projection.  The reference builds a ``rope`` sub-module but never applies it
(modern_attention.py:208-221 assigns the un-rotated tensors back); that is kept: ``self.rope`` exists, the forward does
not rotate -- outputs match the reference's.

B200 path (cuda + bf16 precision, head_dim 64): three tcgen05 GEMMs write bf16 q / k / v, the flash kernels read the
Hkv-head K/V IN PLACE for every query head of the group (no repeat_kv copy, csrc/attention*.cu ``kv_group``) and the
backward accumulates dK / dV over the group in TMEM.  Anything else (fp32 mode, other head dims, a full (S,S) additive
mask) runs the fp32 composition with an explicit differentiable ``repeat_kv``.

Deviation (PARITY.md #18): the reference's backward is a placeholder (grad_x = grad_attn @ W_q^T, no parameter
gradients, modern_attention.py:281-318); here every input and parameter gets its true gradient.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from .. import functional as F
from ..array import B200Array, bfloat16
from ..core import Module, Parameter, Tensor
from .layers import LayerError
from .linear import StandardLinear as Linear   # the reference imports nn/linear.py here (xavier_uniform, (in, out) weight)


def _act_dtype(x: Tensor):
    if isinstance(x.data, B200Array):
        from .. import kernels as K
        if K.get_precision() == "bf16":
            return bfloat16
    return None


class GroupedQueryAttention(Module):
    def __init__(self, d_model: int, n_heads: int = 8, n_kv_heads: Optional[int] = None, dropout: float = 0.0, bias: bool = False,
                 rope_base: float = 10000.0, max_seq_len: int = 2048):
        super().__init__()
        if d_model % n_heads != 0:
            raise LayerError(f"d_model ({d_model}) must be divisible by n_heads ({n_heads})")
        if n_kv_heads is None:
            n_kv_heads = n_heads
        if n_heads % n_kv_heads != 0:
            raise LayerError(f"n_heads ({n_heads}) must be divisible by n_kv_heads ({n_kv_heads})")
        self.d_model, self.n_heads, self.n_kv_heads = d_model, n_heads, n_kv_heads
        self.n_rep = n_heads // n_kv_heads
        self.head_dim = d_model // n_heads
        self.scale = 1.0 / math.sqrt(self.head_dim)
        hd = self.head_dim

        def init(rows, cols, fan_in):
            return np.random.randn(rows, cols).astype(np.float32) * np.sqrt(2.0 / fan_in)
        self.W_q = Parameter(init(d_model, n_heads * hd, d_model), name="gqa.W_q")
        self.W_k = Parameter(init(d_model, n_kv_heads * hd, d_model), name="gqa.W_k")
        self.W_v = Parameter(init(d_model, n_kv_heads * hd, d_model), name="gqa.W_v")
        self.W_o = Parameter(init(n_heads * hd, d_model, n_heads * hd), name="gqa.W_o")
        if bias:
            self.b_q = Parameter(np.zeros(n_heads * hd, dtype=np.float32), name="gqa.b_q")
            self.b_k = Parameter(np.zeros(n_kv_heads * hd, dtype=np.float32), name="gqa.b_k")
            self.b_v = Parameter(np.zeros(n_kv_heads * hd, dtype=np.float32), name="gqa.b_v")
            self.b_o = Parameter(np.zeros(d_model, dtype=np.float32), name="gqa.b_o")
        else:
            self.b_q = self.b_k = self.b_v = self.b_o = None
        self.use_rope = rope_base > 0
        if self.use_rope:
            from .positional import RotaryPositionalEmbedding
            self.rope = RotaryPositionalEmbedding(hd, max_seq_len, rope_base)
        else:
            self.rope = None

    def repeat_kv(self, x):
        """(batch, seq, n_kv_heads, head_dim) array -> (batch, seq, n_heads, head_dim), as modern_attention.py:137-161."""
        if self.n_rep == 1:
            return x
        b, s, hkv, hd = x.shape
        if isinstance(x, B200Array):
            return x.reshape(b, s, hkv, 1, hd).broadcast_to((b, s, hkv, self.n_rep, hd)).contiguous().reshape(b, s, hkv * self.n_rep, hd)
        return np.repeat(x[:, :, :, None, :], self.n_rep, axis=3).reshape(b, s, hkv * self.n_rep, hd)

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, start_pos: int = 0, use_cache: bool = False):
        B, S, d_model = x.shape
        if d_model != self.d_model:
            raise LayerError(f"Input d_model {d_model} != expected {self.d_model}")
        dt = _act_dtype(x)
        H, Hkv, hd = self.n_heads, self.n_kv_heads, self.head_dim
        q = F.linear(x, self.W_q, self.b_q, out_dtype=dt, weight_layout="in_out")
        k = F.linear(x, self.W_k, self.b_k, out_dtype=dt, weight_layout="in_out")
        v = F.linear(x, self.W_v, self.b_v, out_dtype=dt, weight_layout="in_out")
        q = F.transpose(F.reshape(q, (B, S, H, hd)), (0, 2, 1, 3))
        k = F.transpose(F.reshape(k, (B, S, Hkv, hd)), (0, 2, 1, 3))
        v = F.transpose(F.reshape(v, (B, S, Hkv, hd)), (0, 2, 1, 3))
        ctx = F.scaled_dot_product_attention(q, k, v, self.scale, causal=False, mask=mask)
        ctx = F.reshape(F.transpose(ctx, (0, 2, 1, 3)), (B, S, H * hd))
        out = F.linear(ctx, self.W_o, self.b_o, out_dtype=np.float32 if dt is not None else None, weight_layout="in_out")
        out.name = f"gqa({x.name or 'tensor'})"
        if use_cache:
            # the reference hands back the REPEATED (batch, n_heads, seq, head_dim) arrays
            rep = lambda t: self.repeat_kv(t.data.transpose(0, 2, 1, 3)).transpose(0, 2, 1, 3)   # noqa: E731
            return out, (rep(k), rep(v))
        return out

    def extra_repr(self) -> str:
        return f"d_model={self.d_model}, n_heads={self.n_heads}, n_kv_heads={self.n_kv_heads}, n_rep={self.n_rep}"


class MultiQueryAttention(Module):
    """One key/value head for all query heads (modern_attention.py:327-358)."""

    def __init__(self, d_model: int, n_heads: int = 8, dropout: float = 0.0, bias: bool = False):
        super().__init__()
        self.gqa = GroupedQueryAttention(d_model=d_model, n_heads=n_heads, n_kv_heads=1, dropout=dropout, bias=bias)

    def forward(self, x: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        return self.gqa(x, mask)


class FlashAttentionConcept(Module):
    """Fused qkv projection -> softmax(q k^T / sqrt(hd)) v -> output projection (modern_attention.py:360-415).  The
    reference names the tiling idea and computes plain attention; here the flash kernels are what runs."""

    def __init__(self, d_model: int, n_heads: int = 8):
        super().__init__()
        self.d_model, self.n_heads = d_model, n_heads
        self.head_dim = d_model // n_heads
        self.scale = 1.0 / math.sqrt(self.head_dim)
        self.qkv_proj = Linear(d_model, 3 * d_model, bias=False)
        self.out_proj = Linear(d_model, d_model, bias=False)

    def forward(self, x: Tensor) -> Tensor:
        qkv = self.qkv_proj(x, out_dtype=_act_dtype(x))
        ctx = F.packed_qkv_attention(qkv, self.n_heads, causal=False, use_rope=False, scale=float(self.scale))
        return self.out_proj(ctx)
