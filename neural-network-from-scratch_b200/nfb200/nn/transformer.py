"""MultiHeadAttention / TransformerBlock / TransformerDecoderBlock as the reference defines them
(nn/attention.py:13-80, nn/transformer.py:20-260): the reference MultiHeadAttention is four chained Linear
layers (no QK^T/softmax: SURVEY F5) and the decoder block ignores `memory`; both are reproduced because the
translation example (BASELINE config 1) is built from them.  Residuals are real graph nodes here (the
reference re-wraps them in fresh leaves, severing backward: F5/F6; PARITY.md 'connected graph')."""
from __future__ import annotations

import math
from typing import Optional

from ..core import Module, Tensor
from .layers import Dropout, LayerNorm
from .linear import StandardLinear


class MultiHeadAttention(Module):
    def __init__(self, d_model: int, num_heads: int = 8, dropout: float = 0.1, bias: bool = True):
        super().__init__()
        if d_model % num_heads != 0:
            raise ValueError(f"d_model ({d_model}) must be divisible by num_heads ({num_heads})")
        self.d_model, self.num_heads = d_model, num_heads
        self.d_k = d_model // num_heads
        self.scale = 1.0 / math.sqrt(self.d_k)
        self.query_proj = StandardLinear(d_model, d_model, bias=bias)
        self.key_proj = StandardLinear(d_model, d_model, bias=bias)
        self.value_proj = StandardLinear(d_model, d_model, bias=bias)
        self.out_proj = StandardLinear(d_model, d_model, bias=bias)

    def forward(self, x: Tensor, mask=None, residual: Optional[Tensor] = None) -> Tensor:
        return self.out_proj(self.value_proj(self.key_proj(self.query_proj(x))), residual=residual)


class TransformerBlock(Module):
    def __init__(self, d_model: int, num_heads: int = 8, d_ff: int = 2048, dropout: float = 0.1, activation: str = "relu"):
        super().__init__()
        self.d_model, self.num_heads, self.activation = d_model, num_heads, activation
        self.self_attn = MultiHeadAttention(d_model, num_heads, dropout=dropout)
        self.ffn1 = StandardLinear(d_model, d_ff, activation="relu" if activation == "relu" else None)
        self.ffn2 = StandardLinear(d_ff, d_model)
        self.norm1 = LayerNorm(d_model)
        self.norm2 = LayerNorm(d_model)
        self.dropout = Dropout(dropout)

    def forward(self, x: Tensor, mask=None) -> Tensor:
        x = self.self_attn(self.norm1(x), mask=mask, residual=x)          # x + attn(norm1(x)), residual fused in the GEMM epilogue
        return self.ffn2(self.ffn1(self.norm2(x)), residual=x)            # x + ffn2(relu(ffn1(norm2(x))))


class TransformerDecoderBlock(Module):
    """nn/transformer.py:172-260: self-attention, 'cross-attention' applied to x only (memory is ignored upstream), FFN."""

    def __init__(self, d_model: int, num_heads: int = 8, d_ff: int = 2048, dropout: float = 0.1):
        super().__init__()
        self.self_attn = MultiHeadAttention(d_model, num_heads, dropout=dropout)
        self.cross_attn = MultiHeadAttention(d_model, num_heads, dropout=dropout)
        self.ffn1 = StandardLinear(d_model, d_ff, activation="relu")
        self.ffn2 = StandardLinear(d_ff, d_model)
        self.norm1 = LayerNorm(d_model)
        self.norm2 = LayerNorm(d_model)
        self.norm3 = LayerNorm(d_model)
        self.dropout = Dropout(dropout)

    def forward(self, x: Tensor, memory: Optional[Tensor] = None, tgt_mask=None, memory_mask=None) -> Tensor:
        x = self.self_attn(self.norm1(x), mask=tgt_mask, residual=x)
        x = self.cross_attn(self.norm2(x), residual=x)
        return self.ffn2(self.ffn1(self.norm3(x)), residual=x)


class TransformerEncoder(Module):
    """Stack of TransformerBlocks + final LayerNorm (nn/transformer.py:119-183).  As in the reference the blocks live in a plain
    list and ``parameters()`` is overridden to expose them as ``layer{i}_{name}`` / ``final_norm_{name}``."""

    def __init__(self, d_model: int, num_layers: int = 6, num_heads: int = 8, d_ff: int = 2048, dropout: float = 0.1):
        super().__init__()
        self.d_model, self.num_layers = d_model, num_layers
        self.layers = []
        for _ in range(num_layers):
            self.layers.append(TransformerBlock(d_model, num_heads, d_ff, dropout))
        self.norm = LayerNorm(d_model)

    def forward(self, x: Tensor, mask=None) -> Tensor:
        for layer in self.layers:
            x = layer(x, mask=mask)
        return self.norm(x)

    def parameters(self, recurse: bool = True):
        from collections import OrderedDict
        from ..core.base import ParameterView
        params = OrderedDict()
        for i, layer in enumerate(self.layers):
            # the reference block exposes attn_* / ffn1_* / ffn2_* / norm1_* / norm2_* (nn/transformer.py:95-116)
            for prefix, sub in (("attn", layer.self_attn), ("ffn1", layer.ffn1), ("ffn2", layer.ffn2), ("norm1", layer.norm1), ("norm2", layer.norm2)):
                for name, p in sub.parameters().items():
                    params[f"layer{i}_{prefix}_{name}"] = p
        for name, p in self.norm.parameters().items():
            params[f"final_norm_{name}"] = p
        return ParameterView(params)

    def to(self, device):
        for layer in self.layers:
            layer.to(device)
        return super().to(device)


class SelfAttention(Module):
    """Identity placeholder, as in the reference (nn/attention.py:83-92)."""

    def __init__(self, d_model: int):
        super().__init__()
        self.d_model = d_model

    def forward(self, x: Tensor) -> Tensor:
        return x
