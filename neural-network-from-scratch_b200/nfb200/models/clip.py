"""CLIP as the reference defines it (models/multimodal/clip.py:28-536): a ViT image tower (class-token feature ->
Linear projection), a causal pre-LN text transformer (token + learned positional embedding -> N x [x + Attn(LN(x));
x + MLP(LN(x))] -> LayerNorm -> LAST token -> projection), L2-normalised embeddings, logits = img @ txt^T * exp(logit_scale)
and the symmetric cross-entropy over the batch diagonal.  Same module tree, parameter names and initialisation order.

On the B200 both towers run on the hot-path kernels: packed-qkv flash attention (plain for the image tower, causal for the
text tower -- the reference masks the text scores twice, by replacement and by an additive -1e4 triangle; both underflow to
probability 0 exactly like the single replacement here), tanh-GELU in the fc1 GEMM epilogue, residual adds in the GEMM
epilogues, bf16 tensor-core GEMMs in bf16 mode.

Deviation (PARITY.md #21): the reference re-wraps ``.data`` after almost every step (residual adds, GELU, normalisation,
logits), so no gradient reaches any parameter there; here the whole model is one autograd graph (contrastive loss ->
both towers), with forward values identical to the reference's.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .. import functional as F
from ..array import B200Array, bfloat16
from ..core import Module, Parameter, Tensor
from ..core.tensor import make_result
from ..nn import Embedding, LayerNorm, Linear, ModuleList
from .vit import VisionTransformer


def _act_dtype(x: Tensor):
    if isinstance(x.data, B200Array):
        from .. import kernels as K
        if K.get_precision() == "bf16":
            return bfloat16
    return None


def _xp_sqrt(a):
    return np.sqrt(a)


def l2_normalize(x: Tensor, eps: float = 1e-12) -> Tensor:
    """y = x / max(||x||_2, eps) over the last axis (clip.py:341-355), differentiable: dx = (g - y <g, y>) / max(||x||, eps)."""
    xd = x.data
    nrm = np.maximum(_xp_sqrt((xd * xd).sum(axis=-1, keepdims=True)), eps)
    y = xd / nrm

    def backward(g):
        return ((g - y * (g * y).sum(axis=-1, keepdims=True)) / nrm,)
    return make_result(y, [x], backward, "l2_normalize")


def scale_by_exp(x: Tensor, log_scale) -> Tensor:
    """x * exp(log_scale) for a one-element learnable log-temperature (clip.py:375-381)."""
    if not isinstance(log_scale, Tensor):
        return F.mul(x, float(np.exp(log_scale)))
    e = np.exp(log_scale.data)                       # one element, stays on the tensor's device
    y = x.data * e

    def backward(g):
        return g * e, (g * y).sum().reshape(log_scale.shape)
    return make_result(y, [x, log_scale], backward, "scale_by_exp")


class MultiHeadCausalAttention(Module):
    def __init__(self, d_model: int, n_head: int, dropout: float = 0.0):
        super().__init__()
        self.d_model, self.n_head = d_model, n_head
        self.head_dim = d_model // n_head
        self.scale = self.head_dim ** -0.5
        self.qkv_proj = Linear(d_model, d_model * 3)
        self.out_proj = Linear(d_model, d_model)
        self.dropout = None

    def forward(self, x: Tensor, attn_mask: Optional[Tensor] = None, residual: Optional[Tensor] = None) -> Tensor:
        qkv = self.qkv_proj(x, out_dtype=_act_dtype(x))
        ctx = F.packed_qkv_attention(qkv, self.n_head, causal=True, use_rope=False, scale=float(self.scale))
        return self.out_proj(ctx, residual=residual)


class MLP(Module):
    def __init__(self, d_model: int, d_ff: int, dropout: float = 0.0):
        super().__init__()
        self.fc1 = Linear(d_model, d_ff, activation="gelu", enable_fusion=True)     # tanh-GELU (clip.py:133-139) in the GEMM epilogue
        self.fc2 = Linear(d_ff, d_model)
        self.dropout = None

    def forward(self, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        return self.fc2(self.fc1(x, out_dtype=_act_dtype(x)), residual=residual)


class TextTransformerBlock(Module):
    def __init__(self, d_model: int, n_head: int, mlp_ratio: float = 4.0, dropout: float = 0.0, layer_norm_eps: float = 1e-5):
        super().__init__()
        self.d_model, self.n_head = d_model, n_head
        self.head_dim = d_model // n_head
        self.ln_1 = LayerNorm(d_model, eps=layer_norm_eps)
        self.attn = MultiHeadCausalAttention(d_model, n_head, dropout)
        self.ln_2 = LayerNorm(d_model, eps=layer_norm_eps)
        self.mlp = MLP(d_model, int(d_model * mlp_ratio), dropout)

    def forward(self, x: Tensor, attn_mask: Optional[Tensor] = None) -> Tensor:
        dt = _act_dtype(x)
        x = self.attn(self.ln_1(x, out_dtype=dt), attn_mask, residual=x)
        return self.mlp(self.ln_2(x, out_dtype=dt), residual=x)


class CLIPVisionTransformer(Module):
    def __init__(self, input_resolution: int = 224, patch_size: int = 16, width: int = 768, layers: int = 12, heads: int = 12,
                 output_dim: int = 512, dropout: float = 0.0):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.transformer = VisionTransformer(img_size=input_resolution, patch_size=patch_size, embed_dim=width, depth=layers, num_heads=heads,
                                             drop_rate=dropout, num_classes=0, global_pool="token")
        self.proj = Linear(width, output_dim)

    def forward(self, x: Tensor) -> Tensor:
        return self.proj(self.transformer.forward_features(x), out_dtype=np.float32)


class CLIPTextTransformer(Module):
    def __init__(self, context_length: int = 77, vocab_size: int = 49408, width: int = 512, layers: int = 12, heads: int = 8,
                 output_dim: int = 512, dropout: float = 0.0):
        super().__init__()
        self.context_length, self.vocab_size, self.width, self.output_dim = context_length, vocab_size, width, output_dim
        self.token_embedding = Embedding(vocab_size, width)
        self.positional_embedding = Parameter(np.random.randn(context_length, width).astype(np.float32) * 0.02)
        self.transformer = ModuleList()
        for _ in range(layers):
            self.transformer.append(TextTransformerBlock(width, heads, dropout=dropout))
        self.ln_final = LayerNorm(width)
        self.text_projection = Linear(width, output_dim)
        self._init_weights()

    def _init_weights(self):
        std = 0.02
        self.token_embedding.weight.data = np.random.randn(*self.token_embedding.weight.shape).astype(np.float32) * std
        self.text_projection.weight.data = np.random.randn(*self.text_projection.weight.shape).astype(np.float32) * (std / np.sqrt(self.width))

    def build_attention_mask(self, seq_len: int) -> Tensor:
        return Tensor(np.triu(np.full((seq_len, seq_len), -1e4), k=1), requires_grad=False)

    def forward(self, text: Tensor) -> Tensor:
        seq_len = text.shape[1]
        x = self.token_embedding(text)
        x = F.add(x, F.getitem(self.positional_embedding, (None, slice(0, seq_len))))
        for block in self.transformer:
            x = block(x)
        x = self.ln_final(x)
        last = F.getitem(x, (slice(None), seq_len - 1))              # the reference pools the LAST position (clip.py:277)
        return self.text_projection(last, out_dtype=np.float32)


class CLIP(Module):
    def __init__(self, embed_dim: int = 512, image_resolution: int = 224, vision_layers: int = 12, vision_width: int = 768,
                 vision_patch_size: int = 16, context_length: int = 77, vocab_size: int = 49408, transformer_width: int = 512,
                 transformer_heads: int = 8, transformer_layers: int = 12, temperature_init: float = 0.07, learnable_temperature: bool = True):
        super().__init__()
        self.visual = CLIPVisionTransformer(input_resolution=image_resolution, patch_size=vision_patch_size, width=vision_width,
                                            layers=vision_layers, heads=vision_width // 64, output_dim=embed_dim)
        self.transformer = CLIPTextTransformer(context_length=context_length, vocab_size=vocab_size, width=transformer_width,
                                               layers=transformer_layers, heads=transformer_heads, output_dim=embed_dim)
        if learnable_temperature:
            self.logit_scale = Parameter(np.array([np.log(1 / temperature_init)], dtype=np.float32))
        else:
            self.logit_scale = np.log(1 / temperature_init)

    def encode_image(self, image: Tensor) -> Tensor:
        return l2_normalize(self.visual(image))

    def encode_text(self, text: Tensor) -> Tensor:
        return l2_normalize(self.transformer(text))

    get_image_features = encode_image
    get_text_features = encode_text

    def forward(self, image: Optional[Tensor] = None, text: Optional[Tensor] = None, return_loss: bool = True) -> Dict[str, Tensor]:
        out: Dict[str, Tensor] = {}
        if image is not None:
            out["image_embeds"] = img = self.encode_image(image)
        if text is not None:
            out["text_embeds"] = txt = self.encode_text(text)
        if image is not None and text is not None and return_loss:
            logits = scale_by_exp(F.matmul(img, F.transpose(txt, (1, 0))), self.logit_scale)
            out["logits_per_image"], out["logits_per_text"] = logits, F.transpose(logits, (1, 0))
            B = image.shape[0]
            labels = np.arange(B, dtype=np.int64)
            if isinstance(logits.data, B200Array):
                # a constant of the batch size: uploaded once and kept, so a training step does no host->device copy (graph-capturable)
                cache = self.__dict__.setdefault("_diag_labels", {})
                if B not in cache:
                    cache[B] = B200Array.from_numpy(labels)
                labels = cache[B]
            li = F.cross_entropy_loss(out["logits_per_image"], labels)
            lt = F.cross_entropy_loss(out["logits_per_text"], labels)
            out["loss"] = F.mul(F.add(li, lt), 0.5)
        return out


CLIP_CONFIGS = {
    "base": dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=32, context_length=77,
                 vocab_size=49408, transformer_width=512, transformer_heads=8, transformer_layers=12),
    "large": dict(embed_dim=768, image_resolution=224, vision_layers=24, vision_width=1024, vision_patch_size=14, context_length=77,
                  vocab_size=49408, transformer_width=768, transformer_heads=12, transformer_layers=12),
}


class CLIPModel(CLIP):
    pass


def clip_base(**kw):
    return CLIP(**{**CLIP_CONFIGS["base"], **kw})


def clip_large(**kw):
    return CLIP(**{**CLIP_CONFIGS["large"], **kw})
