"""Pre-norm Transformer with RoPE of the reference (models/language/modern_transformer.py:29-687): token embedding
(optionally scaled by sqrt(d) and tied to the output projection) -> N x [x + Attn(norm1(x)); x + FFN(norm2(x))] -> final
norm -> logits.  Attention = separate q/k/v/out Linear layers (N(0, 2/(2d)) init), interleaved-pair RoPE on q and k
(nn/positional.py), BIDIRECTIONAL softmax(q k^T / sqrt(hd)) with padded keys replaced by -1e9; FFN = Linear + fused
GELU / ReLU, or the [x | gate] SwiGLU of functional.swiglu; LayerNorm or RMSNorm.  Same module tree, parameter names,
configuration validation and initialisation order.

B200 path (cuda, bf16 precision, head_dim 64): the three projections run as ONE tcgen05 GEMM against the differentiably
concatenated [W_q; W_k; W_v], whose output is the packed (B,S,3,H,hd) layout; RoPE rotates the q and k parts of that
buffer in place (k_rope_pairs); the flash kernels read it in place with the padding mask as an additive per-key mask
(-1e9 absorbs any score, so "replace" and "add" give the same probabilities); residual adds live in the out_proj /
down_proj GEMM epilogues; the tied output projection reuses the embedding's bf16 shadow.

Deviation (PARITY.md #22): the reference severs its autograd graph at every reshape / residual / scaling step, so its
backward reaches almost nothing; here the model is one graph (forward values identical).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .. import exceptions as _exc
from .. import functional as F
from ..array import B200Array, bfloat16
from ..core import Module, Parameter, Tensor
from ..core.tensor import make_result
from ..nn import Dropout, LayerNorm, Linear, RMSNorm
from ..nn.positional import RotaryPositionalEmbedding


class ModelError(_exc.ModelError, ValueError):
    """exceptions.py ModelError of the reference."""


def _act_dtype(x: Tensor):
    if isinstance(x.data, B200Array):
        from .. import kernels as K
        if K.get_precision() == "bf16":
            return bfloat16
    return None


class PreNormTransformerConfig:
    def __init__(self, d_model: int = 512, num_layers: int = 6, num_heads: int = 8, d_ff: int = 2048, max_seq_len: int = 2048,
                 vocab_size: int = 30000, dropout: float = 0.1, activation: str = "gelu", normalization: str = "layernorm",
                 use_rope: bool = True, rope_base: float = 10000.0, eps: float = 1e-5, tie_embeddings: bool = True, scale_embeddings: bool = True):
        if d_model <= 0:
            raise ModelError(f"d_model must be positive, got {d_model}")
        if d_model % num_heads != 0:
            raise ModelError(f"d_model {d_model} must be divisible by num_heads {num_heads}")
        for name, v in (("num_layers", num_layers), ("num_heads", num_heads), ("d_ff", d_ff), ("max_seq_len", max_seq_len), ("vocab_size", vocab_size)):
            if v <= 0:
                raise ModelError(f"{name} must be positive, got {v}")
        self.d_model, self.num_layers, self.num_heads, self.d_ff = d_model, num_layers, num_heads, d_ff
        self.max_seq_len, self.vocab_size, self.dropout = max_seq_len, vocab_size, dropout
        self.activation, self.normalization, self.use_rope, self.rope_base = activation, normalization, use_rope, rope_base
        self.eps, self.tie_embeddings, self.scale_embeddings = eps, tie_embeddings, scale_embeddings
        self.head_dim = d_model // num_heads


def _rope_packed(qkv: Tensor, rope: RotaryPositionalEmbedding, n_heads: int, start_pos: int) -> Tensor:
    """Interleaved RoPE on the q and k parts of a packed (B, S, 3*d) projection output (a copy is rotated in place); the
    backward rotates the q / k parts of the gradient back."""
    B, S, three_d = qkv.shape
    hd = three_d // 3 // n_heads

    def rotate(arr, inverse):
        out = arr.copy() if isinstance(arr, B200Array) else np.array(arr, copy=True)
        p5 = out.reshape(B, S, 3, n_heads, hd)
        for part in (0, 1):
            view = p5[:, :, part].transpose(0, 2, 1, 3)              # logical (B, H, S, hd), strided
            if isinstance(out, B200Array):
                from .. import kernels as K
                cos, sin = rope._tables(S, start_pos)
                ct = rope._dev.get(("c", start_pos, S), lambda: cos)
                st = rope._dev.get(("s", start_pos, S), lambda: sin)
                K.rope_interleaved_inplace(view, ct, st, inverse)
            else:
                p5[:, :, part] = rope.apply_rope(view, S, start_pos, inverse=inverse).transpose(0, 2, 1, 3)
        return out
    return make_result(rotate(qkv.data, False), [qkv], lambda g: (rotate(g, True),), "rope_packed")


class RoPEMultiHeadAttention(Module):
    def __init__(self, config: PreNormTransformerConfig):
        super().__init__()
        self.config = config
        self.d_model, self.num_heads, self.head_dim = config.d_model, config.num_heads, config.head_dim
        self.scale = 1.0 / np.sqrt(self.head_dim)
        self.q_proj = Linear(config.d_model, config.d_model, bias=True)
        self.k_proj = Linear(config.d_model, config.d_model, bias=True)
        self.v_proj = Linear(config.d_model, config.d_model, bias=True)
        self.out_proj = Linear(config.d_model, config.d_model, bias=True)
        self.rope = RotaryPositionalEmbedding(dim=self.head_dim, max_seq_len=config.max_seq_len, base=config.rope_base) if config.use_rope else None
        self.dropout = Dropout(config.dropout)
        self._init_weights()

    def _init_weights(self):
        std = np.sqrt(2.0 / (self.d_model + self.d_model))
        for layer in (self.q_proj, self.k_proj, self.v_proj, self.out_proj):
            layer.weight.data = np.random.normal(0, std, layer.weight.shape).astype(np.float32)
            if layer.bias is not None:
                layer.bias.data = np.zeros(layer.bias.shape, dtype=np.float32)

    def forward(self, x: Tensor, attention_mask=None, start_pos: int = 0, use_cache: bool = False, residual: Optional[Tensor] = None):
        B, S, d_model = x.shape
        if d_model != self.d_model:
            raise ModelError(f"Input d_model {d_model} != expected {self.d_model}")
        w = F.concatenate([self.q_proj.weight, self.k_proj.weight, self.v_proj.weight], axis=0)          # (3d, d), (out, in) layout
        b = F.concatenate([self.q_proj.bias, self.k_proj.bias, self.v_proj.bias], axis=0)
        qkv = F.linear(x, w, b, out_dtype=_act_dtype(x))
        if self.rope is not None:
            qkv = _rope_packed(qkv, self.rope, self.num_heads, start_pos)
        mask = None
        if attention_mask is not None:
            m = attention_mask.data if isinstance(attention_mask, Tensor) else attention_mask
            keep = (m != 0)
            keep = keep.astype(np.float32) if isinstance(keep, (np.ndarray, B200Array)) else np.asarray(keep, dtype=np.float32)
            add = (1.0 - keep.reshape(B, 1, 1, S)) * -1e9                                                  # "where(mask, s, -1e9)" as an additive key mask
            mask = Tensor(add, device=x.device) if not isinstance(add, B200Array) else Tensor(add)
        ctx = F.packed_qkv_attention(qkv, self.num_heads, causal=False, use_rope=False, scale=float(self.scale), mask=mask)
        out = self.out_proj(ctx, residual=residual)
        return out, None


class PreNormFeedForward(Module):
    def __init__(self, config: PreNormTransformerConfig):
        super().__init__()
        self.config = config
        self.activation_type = config.activation
        if config.activation == "swiglu":
            self.up_proj = Linear(config.d_model, 2 * config.d_ff, bias=False)
            self.down_proj = Linear(config.d_ff, config.d_model, bias=False)
        else:
            act = config.activation.lower() if config.activation.lower() in ("gelu", "relu") else None
            if act is None:
                raise ModelError(f"Unknown activation: {config.activation}")
            self.up_proj = Linear(config.d_model, config.d_ff, bias=True, activation=act, enable_fusion=True)
            self.down_proj = Linear(config.d_ff, config.d_model, bias=True)
        self.gate_proj = None
        self.dropout = Dropout(config.dropout)
        self._init_weights()

    def _init_weights(self):
        std = np.sqrt(2.0 / self.config.d_model)
        for layer in (self.up_proj, self.down_proj):
            layer.weight.data = np.random.normal(0, std, layer.weight.shape).astype(np.float32)
            if layer.bias is not None:
                layer.bias.data = np.zeros(layer.bias.shape, dtype=np.float32)

    def forward(self, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        hidden = self.up_proj(x, out_dtype=_act_dtype(x))
        if self.activation_type == "swiglu":
            hidden = F.swiglu_packed(hidden, gate_last=True)        # functional.swiglu: x, gate = split(up); silu(gate) * x
        return self.down_proj(hidden, residual=residual)


class PreNormTransformerLayer(Module):
    def __init__(self, config: PreNormTransformerConfig):
        super().__init__()
        self.config = config
        self.self_attention = RoPEMultiHeadAttention(config)
        self.feed_forward = PreNormFeedForward(config)
        norm = RMSNorm if config.normalization == "rmsnorm" else LayerNorm
        self.norm1 = norm(config.d_model, eps=config.eps)
        self.norm2 = norm(config.d_model, eps=config.eps)

    def forward(self, x: Tensor, attention_mask=None, start_pos: int = 0) -> Tensor:
        dt = _act_dtype(x)
        x, _ = self.self_attention(self.norm1(x, out_dtype=dt), attention_mask, start_pos, residual=x)
        return self.feed_forward(self.norm2(x, out_dtype=dt), residual=x)


class PreNormTransformer(Module):
    def __init__(self, config: Optional[PreNormTransformerConfig] = None, **kwargs):
        super().__init__()
        self.config = config = config or PreNormTransformerConfig(**kwargs)
        self.token_embedding = Parameter(np.random.normal(0, 0.02, (config.vocab_size, config.d_model)).astype(np.float32), name="token_embedding")
        self.layers = []
        for i in range(config.num_layers):
            layer = PreNormTransformerLayer(config)
            self.layers.append(layer)
            setattr(self, f"layer_{i}", layer)
        norm = RMSNorm if config.normalization == "rmsnorm" else LayerNorm
        self.final_norm = norm(config.d_model, eps=config.eps)
        self.output_projection = None if config.tie_embeddings else Linear(config.d_model, config.vocab_size, bias=False)
        self.dropout = Dropout(config.dropout)
        self._init_weights()

    def _init_weights(self):
        std = 0.02
        self.token_embedding.data = np.random.normal(0, std, self.token_embedding.shape).astype(np.float32)
        if self.output_projection is not None:
            self.output_projection.weight.data = np.random.normal(0, std, self.output_projection.weight.shape).astype(np.float32)

    def _count_parameters(self) -> int:
        return sum(int(p.size) for p in self.parameters())

    def get_embeddings(self, input_ids: Tensor) -> Tensor:
        emb = F.embedding(self.token_embedding, input_ids)
        return F.mul(emb, float(np.sqrt(self.config.d_model))) if self.config.scale_embeddings else emb

    def forward(self, input_ids: Tensor, attention_mask=None, start_pos: int = 0, output_hidden_states: bool = False,
                labels=None) -> Dict[str, Tensor]:
        x = self.get_embeddings(input_ids)
        hidden = [] if output_hidden_states else None
        for layer in self.layers:
            if output_hidden_states:
                hidden.append(x)
            x = layer(x, attention_mask, start_pos)
        x = self.final_norm(x, out_dtype=_act_dtype(x))
        if output_hidden_states:
            hidden.append(x)
        if self.config.tie_embeddings:
            logits = F.linear(x, self.token_embedding, None, out_dtype=np.float32)        # x @ E^T: E is (vocab, d) = (out, in)
        else:
            logits = self.output_projection(x, out_dtype=np.float32)
        out = {"logits": logits, "last_hidden_state": x}
        if output_hidden_states:
            out["hidden_states"] = hidden
        if labels is not None:                                                             # convenience beyond the reference: LM loss
            lab = labels.data if isinstance(labels, Tensor) else labels
            B, S, V = logits.shape
            out["loss"] = F.cross_entropy_loss(F.reshape(logits, (B * S, V)), lab.reshape(-1))
        return out

    def extra_repr(self) -> str:
        c = self.config
        return f"d_model={c.d_model}, num_layers={c.num_layers}, num_heads={c.num_heads}, vocab_size={c.vocab_size}"


def prenorm_transformer_small(**kw):
    return PreNormTransformer(**{**dict(d_model=512, num_layers=6, num_heads=8, d_ff=2048), **kw})


def prenorm_transformer_base(**kw):
    return PreNormTransformer(**{**dict(d_model=768, num_layers=12, num_heads=12, d_ff=3072), **kw})


def prenorm_transformer_large(**kw):
    return PreNormTransformer(**{**dict(d_model=1024, num_layers=24, num_heads=16, d_ff=4096), **kw})


ModernTransformer = PreNormTransformer
RoPETransformer = PreNormTransformer
