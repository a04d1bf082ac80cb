"""Differential attention of the reference (nn/differential_attention.py:32-531): two softmax attention maps per head
pair, out_h = (1 + lambda_h) * A1_h V1_h - lambda_h * A2_h V2_h with a learnable per-head lambda, the result duplicated over
both head halves before the output projection; ``DifferentialTransformerBlock`` (pre-norm RMSNorm + SwiGLU FFN) and
``DifferentialTransformer`` (embedding -> blocks -> RMSNorm -> vocabulary projection; like the reference it builds a RoPE
module it never applies).  Same parameter names (``lambda_param``, ``W_q1 .. W_v2``, ``W_o``, optional biases), init order and
forward values.

B200 mapping: attention heads are independent, so the two groups are simply the first and second half of ONE H-head
attention.  The six projections run as one tcgen05 GEMM against the (differentiably) concatenated
[W_q1|W_q2|W_k1|W_k2|W_v1|W_v2], whose output IS the packed (B,S,3,H,hd) qkv layout the flash kernels read in place; one
flash forward / backward serves both maps; the lambda combination is a handful of elementwise kernels on (B,S,d/2).

Deviation (PARITY.md #20): the reference's backward is a placeholder whose matmul shapes do not even agree
(differential_attention.py:294-318); every input and parameter, lambda included, gets its true gradient here.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from .. import functional as F
from ..array import B200Array, bfloat16
from ..core import Module, Parameter, Tensor
from .layers import Embedding, LayerError, RMSNorm
from .linear import StandardLinear as Linear   # the reference imports nn/linear.py here (xavier_uniform, (in, out) weight)
from .modern_activations import SwiGLU


def _act_dtype(x: Tensor):
    if isinstance(x.data, B200Array):
        from .. import kernels as K
        if K.get_precision() == "bf16":
            return bfloat16
    return None


class DifferentialAttention(Module):
    def __init__(self, d_model: int, n_heads: int = 8, dropout: float = 0.0, bias: bool = False, lambda_init: float = 0.5, max_seq_len: int = 4096):
        super().__init__()
        if d_model % n_heads != 0:
            raise LayerError(f"d_model ({d_model}) must be divisible by n_heads ({n_heads})")
        if n_heads % 2 != 0:
            raise LayerError(f"n_heads ({n_heads}) must be even for differential attention")
        self.d_model, self.n_heads = d_model, n_heads
        self.n_heads_per_group = n_heads // 2
        self.head_dim = d_model // n_heads
        self.scale = 1.0 / math.sqrt(self.head_dim)
        half = self.n_heads_per_group * self.head_dim
        self.lambda_param = Parameter(np.full(self.n_heads_per_group, lambda_init, dtype=np.float32), name="diff_attn.lambda")
        for n in ("W_q1", "W_k1", "W_v1", "W_q2", "W_k2", "W_v2"):          # the reference's RNG call order
            setattr(self, n, Parameter(np.random.randn(d_model, half).astype(np.float32) * np.sqrt(2.0 / d_model), name=f"diff_attn.{n}"))
        self.W_o = Parameter(np.random.randn(n_heads * self.head_dim, d_model).astype(np.float32) * np.sqrt(2.0 / (n_heads * self.head_dim)),
                             name="diff_attn.W_o")
        if bias:
            for n in ("b_q1", "b_k1", "b_v1", "b_q2", "b_k2", "b_v2"):
                setattr(self, n, Parameter(np.zeros(half, dtype=np.float32)))
            self.b_o = Parameter(np.zeros(d_model, dtype=np.float32))
        else:
            self.b_q1 = self.b_k1 = self.b_v1 = self.b_q2 = self.b_k2 = self.b_v2 = self.b_o = None

    def _attention_maps(self, qkv: Tensor, mask):
        """(attn_weights1, attn_weights2) as host arrays -- analysis only (return_attention_weights=True)."""
        B, S, _ = qkv.shape
        H, hd = self.n_heads, self.head_dim
        d = np.asarray(qkv.data.get() if isinstance(qkv.data, B200Array) else qkv.data, dtype=np.float32).reshape(B, S, 3, H, hd)
        q, k = d[:, :, 0].transpose(0, 2, 1, 3), d[:, :, 1].transpose(0, 2, 1, 3)
        s = np.matmul(q, k.transpose(0, 1, 3, 2)) * self.scale
        if mask is not None:
            m = mask.data if isinstance(mask, Tensor) else mask
            s = s + (m.get() if isinstance(m, B200Array) else np.asarray(m))
        e = np.exp(s - s.max(axis=-1, keepdims=True))
        a = e / e.sum(axis=-1, keepdims=True)
        return a[:, :H // 2], a[:, H // 2:]

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, return_attention_weights: bool = False):
        B, S, d_model = x.shape
        if d_model != self.d_model:
            raise LayerError(f"Input d_model {d_model} != expected {self.d_model}")
        dt = _act_dtype(x)
        Hg, hd = self.n_heads_per_group, self.head_dim
        w = F.concatenate([self.W_q1, self.W_q2, self.W_k1, self.W_k2, self.W_v1, self.W_v2], axis=1)
        b = F.concatenate([self.b_q1, self.b_q2, self.b_k1, self.b_k2, self.b_v1, self.b_v2], axis=0) if self.b_q1 is not None else None
        qkv = F.linear(x, w, b, out_dtype=dt, weight_layout="in_out")            # (B, S, 3, H, hd) packed: heads [group 1 | group 2]
        ctx = F.packed_qkv_attention(qkv, self.n_heads, causal=False, use_rope=False, scale=float(self.scale), mask=mask)
        ctx = F.reshape(F.to_float32(ctx), (B, S, 2, Hg, hd))
        o1, o2 = F.getitem(ctx, (slice(None), slice(None), 0)), F.getitem(ctx, (slice(None), slice(None), 1))
        lam = F.reshape(self.lambda_param, (1, 1, Hg, 1))
        diff = F.sub(F.mul(F.add(lam, 1.0), o1), F.mul(lam, o2))                 # (1 + lambda) o1 - lambda o2, the reference's order
        diff = F.reshape(diff, (B, S, Hg * hd))
        full = F.concatenate([diff, diff], axis=2)
        out = F.linear(full, self.W_o, self.b_o, weight_layout="in_out")
        out.name = f"diff_attn({x.name or 'tensor'})"
        if return_attention_weights:
            return out, self._attention_maps(qkv, mask)
        return out

    def get_attention_stats(self) -> dict:
        lam = self.lambda_param.data
        lam = lam.get() if isinstance(lam, B200Array) else np.asarray(lam)
        return {"lambda_mean": float(np.mean(lam)), "lambda_std": float(np.std(lam)), "lambda_min": float(np.min(lam)), "lambda_max": float(np.max(lam))}

    def extra_repr(self) -> str:
        return (f"d_model={self.d_model}, n_heads={self.n_heads}, n_heads_per_group={self.n_heads_per_group}, "
                f"lambda_mean={self.get_attention_stats()['lambda_mean']:.3f}")


class DifferentialTransformerBlock(Module):
    """h = x + DiffAttn(RMSNorm(x)); out = h + SwiGLU(RMSNorm(h))  (differential_attention.py:365-427)."""

    def __init__(self, d_model: int, n_heads: int = 8, d_ff: Optional[int] = None, dropout: float = 0.0, lambda_init: float = 0.5):
        super().__init__()
        self.attention = DifferentialAttention(d_model=d_model, n_heads=n_heads, dropout=dropout, lambda_init=lambda_init)
        self.norm1 = RMSNorm(d_model)
        self.norm2 = RMSNorm(d_model)
        self.ffn = SwiGLU(input_dim=d_model, hidden_dim=d_ff)

    def forward(self, x: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        dt = _act_dtype(x)
        h = F.add(x, self.attention(self.norm1(x, out_dtype=dt), mask))
        return self.ffn(self.norm2(h, out_dtype=dt), residual=h)          # residual add folded into the W_down GEMM epilogue

    def get_attention_stats(self) -> dict:
        return self.attention.get_attention_stats()


class DifferentialTransformer(Module):
    """token embedding -> n_layers blocks -> RMSNorm -> Linear(d_model, vocab)  (differential_attention.py:430-531)."""

    def __init__(self, n_layers: int = 12, d_model: int = 768, n_heads: int = 12, d_ff: Optional[int] = None, vocab_size: int = 50257,
                 max_seq_len: int = 2048, dropout: float = 0.0, lambda_init: float = 0.5):
        super().__init__()
        self.n_layers, self.d_model = n_layers, d_model
        self.token_embedding = Embedding(vocab_size, d_model)
        from .positional import RotaryPositionalEmbedding
        self.pos_encoding = RotaryPositionalEmbedding(dim=d_model // n_heads, max_seq_len=max_seq_len)   # built, never applied (as the reference)
        self.blocks = []
        for i in range(n_layers):
            block = DifferentialTransformerBlock(d_model=d_model, n_heads=n_heads, d_ff=d_ff, dropout=dropout, lambda_init=lambda_init)
            self.blocks.append(block)
            setattr(self, f"block_{i}", block)
        self.final_norm = RMSNorm(d_model)
        self.output_proj = Linear(d_model, vocab_size, bias=False)

    def forward(self, input_ids: Tensor, mask: Optional[Tensor] = None) -> Tensor:
        x = self.token_embedding(input_ids)
        for block in self.blocks:
            x = block(x, mask)
        return self.output_proj(self.final_norm(x, out_dtype=_act_dtype(x)))

    def get_all_lambda_stats(self) -> dict:
        stats = {f"layer_{i}": blk.get_attention_stats() for i, blk in enumerate(self.blocks)}
        means = [v["lambda_mean"] for v in stats.values()]
        stats["global"] = {"lambda_mean": float(np.mean(means)), "lambda_std": float(np.std(means)), "lambda_min": float(np.min(means)),
                           "lambda_max": float(np.max(means))}
        return stats
