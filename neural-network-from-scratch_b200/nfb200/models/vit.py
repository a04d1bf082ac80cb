"""Vision Transformer as the reference defines it (models/vision/vision_transformer.py:27-458): patchify by
reshape/transpose -> Linear(3*16*16 -> d) -> [cls] + pos_embed -> N x [x + Attn(LN(x)); x + MLP(LN(x))] -> LN -> token 0
-> Linear(d -> classes).  Attention = fused qkv Linear, (q k^T) * hd^-0.5, softmax, @ v, proj; MLP = Linear + tanh-GELU
(fused), Linear.  Same module tree / parameter names as the reference."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .. import functional as F
from ..array import B200Array, bfloat16
from ..core import Module, Parameter, Tensor
from ..nn import LayerNorm, Linear, ModuleList


def _act_dtype(x: Tensor):
    if isinstance(x.data, B200Array):
        from .. import kernels as K
        if K.get_precision() == "bf16":
            return bfloat16
    return None


class PatchEmbed(Module):
    def __init__(self, img_size=224, patch_size=16, in_channels=3, embed_dim=768, norm_layer=None):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size // patch_size) ** 2
        self.embed_dim = embed_dim
        self.proj = Linear(patch_size * patch_size * in_channels, embed_dim)
        self.norm = norm_layer(embed_dim) if norm_layer else None

    def forward(self, x: Tensor) -> Tensor:
        B, C, H, W = x.shape
        assert H == W == self.img_size, f"Input size {H}x{W} != {self.img_size}x{self.img_size}"
        p = self.patch_size
        nh, nw = H // p, W // p
        x = F.reshape(x, (B, C, nh, p, nw, p))
        x = F.transpose(x, (0, 2, 4, 1, 3, 5))            # B, nh, nw, C, ph, pw  (vision_transformer.py:62-64)
        x = F.reshape(x, (B, nh * nw, C * p * p))
        x = self.proj(x)
        return self.norm(x) if self.norm else x


class Attention(Module):
    def __init__(self, dim, num_heads=8, qkv_bias=True, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = Linear(dim, dim)

    def forward(self, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        qkv = self.qkv(x)                                                                     # (B, N, 3, H, hd) packed
        ctx = F.packed_qkv_attention(qkv, self.num_heads, causal=False, use_rope=False, scale=float(self.scale))
        return self.proj(ctx, residual=residual)


class MLP(Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = Linear(in_features, hidden_features, activation="gelu", enable_fusion=True)
        self.fc2 = Linear(hidden_features, out_features)

    def forward(self, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        return self.fc2(self.fc1(x, out_dtype=_act_dtype(x)), residual=residual)


class Block(Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=True, drop=0.0, attn_drop=0.0, drop_path=0.0, layer_scale_init=None):
        super().__init__()
        self.norm1 = LayerNorm(dim)
        self.attn = Attention(dim, num_heads, qkv_bias, attn_drop, drop)
        self.drop_path_rate = drop_path
        self.norm2 = LayerNorm(dim)
        self.mlp = MLP(dim, int(dim * mlp_ratio), drop=drop)
        if layer_scale_init is not None:
            self.layer_scale_1 = Parameter(np.ones(dim, dtype=np.float32) * layer_scale_init)
            self.layer_scale_2 = Parameter(np.ones(dim, dtype=np.float32) * layer_scale_init)
        else:
            self.layer_scale_1 = None
            self.layer_scale_2 = None

    def forward(self, x: Tensor) -> Tensor:
        dt = _act_dtype(x)
        if self.layer_scale_1 is None:
            x = self.attn(self.norm1(x, out_dtype=dt), residual=x)       # residual adds folded into the GEMM epilogues
            return self.mlp(self.norm2(x, out_dtype=dt), residual=x)
        x = x + self.attn(self.norm1(x)) * self.layer_scale_1
        return x + self.mlp(self.norm2(x)) * self.layer_scale_2


class VisionTransformer(Module):
    def __init__(self, img_size=224, patch_size=16, in_channels=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 qkv_bias=True, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0, layer_scale_init=None, class_token=True,
                 global_pool="token"):
        super().__init__()
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_tokens = 1 if class_token else 0
        self.global_pool = global_pool
        self.patch_embed = PatchEmbed(img_size, patch_size, in_channels, embed_dim)
        n = self.patch_embed.num_patches
        if class_token:
            self.cls_token = Parameter(np.zeros((1, 1, embed_dim), dtype=np.float32))
        self.pos_embed = Parameter(np.zeros((1, n + self.num_tokens, embed_dim), dtype=np.float32))
        dpr = np.linspace(0, drop_path_rate, depth).tolist()
        self.blocks = ModuleList()
        for i in range(depth):
            self.blocks.append(Block(embed_dim, num_heads, mlp_ratio, qkv_bias, drop_rate, attn_drop_rate, dpr[i], layer_scale_init))
        self.norm = LayerNorm(embed_dim)
        self.head = Linear(embed_dim, num_classes) if num_classes > 0 else None
        self._init_weights()

    def _init_weights(self):
        """vision_transformer.py:310-325: patch projection, pos_embed and cls_token ~ N(0, 0.02^2), in that order."""
        std = 0.02
        self.patch_embed.proj.weight.data = np.random.randn(*self.patch_embed.proj.weight.shape).astype(np.float32) * std
        self.pos_embed.data = np.random.randn(*self.pos_embed.shape).astype(np.float32) * std
        if "cls_token" in self._parameters:
            self.cls_token.data = np.random.randn(*self.cls_token.shape).astype(np.float32) * std

    def forward_features(self, x: Tensor) -> Tensor:
        B = x.shape[0]
        x = self.patch_embed(x)
        if "cls_token" in self._parameters:
            cls = F.add(F.reshape(self.cls_token, (1, 1, self.embed_dim)), _zeros_like_rows(x, B, self.embed_dim))
            x = F.concatenate([cls, x], axis=1)
        x = x + self.pos_embed[:, :x.shape[1]]
        for blk in self.blocks:
            x = blk(x)
        x = self.norm(x)
        if self.global_pool == "avg":
            return F.mean(x[:, self.num_tokens:], axis=1)
        return x[:, 0]

    def forward(self, x: Tensor, labels=None):
        x = self.forward_features(x)
        if self.head is not None:
            x = self.head(x, out_dtype=np.float32)
        if labels is not None:
            lab = labels.data if isinstance(labels, Tensor) else labels
            return {"logits": x, "loss": F.cross_entropy_loss(x, lab.reshape(-1))}
        return x


def _zeros_like_rows(x: Tensor, B: int, d: int) -> Tensor:
    """(B, 1, d) zeros on x's device: broadcasting the class token over the batch as a differentiable add."""
    if isinstance(x.data, B200Array):
        return Tensor(x.backend.zeros((B, 1, d), np.float32))
    return Tensor(np.zeros((B, 1, d), dtype=np.float32))


def vit_b_16(**kw):
    return VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, **kw)


def vit_l_16(**kw):
    return VisionTransformer(patch_size=16, embed_dim=1024, depth=24, num_heads=16, **kw)
