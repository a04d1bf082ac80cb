"""GPT-2 as the reference defines it (models/language/gpt2.py:25-609): GPT-2 shapes with rotary position embeddings,
RMSNorm, SwiGLU MLP (w1,w2: d->2d, w3: 2d->d, no bias), tied lm_head/wte, causal mask constant -1e4, identity dropout.

Same module tree and parameter names as the reference (transformer.wte / h.N.{ln_1,attn.{c_attn,c_proj},ln_2,mlp.swiglu.{w1,w2,w3}}
/ ln_f / lm_head) so state_dicts interchange.  Differences, all in PARITY.md: the graph is CONNECTED (the reference re-wraps
intermediates in fresh leaf tensors, so its own models do not train: SURVEY F6), and on a cuda device the block runs on
fused kernels: residual adds live in GEMM epilogues, RoPE + causal flash attention consume the packed qkv projection in
place, and w1/w2 (adjacent in the parameter arena) run as one GEMM."""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np

from .. import functional as F
from ..array import B200Array, bfloat16
from ..core import Module, Parameter, Tensor
from ..nn import Dropout, Embedding, Linear, ModuleList


class RMSNorm(Module):
    """models/language/gpt2.py:98-116 (its own RMSNorm class, eps from layer_norm_epsilon)."""

    def __init__(self, hidden_size: int, eps: float = 1e-6):
        super().__init__()
        self.weight = Parameter(np.ones(hidden_size, dtype=np.float32))
        self.variance_epsilon = eps

    def forward(self, hidden_states: Tensor, out_dtype=None) -> Tensor:
        return F.rms_norm(hidden_states, self.weight, self.variance_epsilon, out_dtype)


class SwiGLU(Module):
    def __init__(self, dim: int):
        super().__init__()
        self.w1 = Linear(dim, dim * 2, bias=False)  # gate
        self.w2 = Linear(dim, dim * 2, bias=False)  # up
        self.w3 = Linear(dim * 2, dim, bias=False)  # down
        self._fused_w12 = None

    def _fused_gate_up(self):
        """A (4d, d) Parameter view spanning w1.weight and w2.weight when they are adjacent in the flat arena."""
        w1, w2 = self.w1.weight, self.w2.weight
        a1, a2 = getattr(w1, "_arena", None), getattr(w2, "_arena", None)
        if a1 is None or a1 is not a2:
            return None
        o1, o2 = a1.offsets[id(w1)], a1.offsets[id(w2)]
        if o2 != o1 + w1.size:
            return None
        f = self._fused_w12
        if f is None or f._arena is not a1:
            n = w1.size + w2.size
            f = Parameter.__new__(Parameter)
            Tensor.__init__(f, a1.flat_p[o1:o1 + n].reshape(2 * w1.shape[0], w1.shape[1]), requires_grad=True)
            f._arena = a1
            f._arena_bound = True
            f._grad_slot = a1.flat_g[o1:o1 + n].reshape(f.shape)
            if a1.flat_bf16 is not None:
                f._bf16 = a1.flat_bf16[o1:o1 + n].reshape(f.shape)
                f._bf16_live = True
            f._ddp_alias = (w1, w2)
            object.__setattr__(self, "_fused_w12", f)
        # the fused view shares w1/w2's gradient slots: it is "zero" exactly when they are untouched
        f._grad = f._grad_slot if (w1._grad is not None or w2._grad is not None) else None
        f._slot_zero = w1._slot_zero and w2._slot_zero and w1._grad is None and w2._grad is None
        return f

    def forward(self, x: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        fused = self._fused_gate_up() if isinstance(x.data, B200Array) else None
        if fused is not None:
            gu = F.linear(x, fused, None)                       # one GEMM for gate|up
            w1, w2 = self.w1.weight, self.w2.weight
            # mark the two real parameters as having a (shared-slot) gradient once backward has run
            gu_fn = gu._grad_fn
            if gu_fn is not None:
                inner = gu_fn.backward_fn

                def wrapped(g, _inner=inner):
                    out = _inner(g)
                    for w in (w1, w2):
                        if w._grad is None:
                            w._grad = w._grad_slot
                        w._slot_zero = False
                    return out

                gu_fn.backward_fn = wrapped
            return self.w3(F.swiglu_packed(gu), residual=residual)
        from ..functional.fused import to_float32
        gate = F.silu(to_float32(self.w1(x)))   # unfused route (no flat arena yet, or host tensors): fp32 elementwise math
        up = to_float32(self.w2(x))
        return self.w3(F.mul(gate, up), residual=residual)


class GPT2Attention(Module):
    def __init__(self, config: Dict[str, Any], layer_idx: Optional[int] = None):
        super().__init__()
        self.embed_dim = config["n_embd"]
        self.num_heads = config["n_head"]
        self.head_dim = self.embed_dim // self.num_heads
        self.scale_attn_weights = config.get("scale_attn_weights", True)
        self.layer_idx = layer_idx
        if self.head_dim * self.num_heads != self.embed_dim:
            raise ValueError(f"embed_dim must be divisible by num_heads (got embed_dim={self.embed_dim} and num_heads={self.num_heads})")
        self.c_attn = Linear(self.embed_dim, 3 * self.embed_dim)
        self.c_proj = Linear(self.embed_dim, self.embed_dim)
        self.attn_dropout = Dropout(config.get("attn_pdrop", 0.1))
        self.resid_dropout = Dropout(config.get("resid_pdrop", 0.1))

    def forward(self, hidden_states: Tensor, attention_mask=None, residual: Optional[Tensor] = None) -> Tensor:
        from ..functional.fused import fused_rope_available
        fused = fused_rope_available(hidden_states, self.head_dim) and hidden_states.ndim == 3
        # bf16 tensor-core path: the rotary embedding of q and k rides in the projection's epilogue (before the rounding to bf16)
        rope = (hidden_states.shape[1], 2 * self.embed_dim, self.head_dim) if fused else None
        qkv = self.c_attn(hidden_states, rope=rope) if fused else self.c_attn(hidden_states)   # (B,S,3d) laid out (B,S,3,H,hd)
        scale = (1.0 / np.sqrt(self.head_dim)) if self.scale_attn_weights else 1.0
        ctx = F.packed_qkv_attention(qkv, self.num_heads, causal=True, use_rope=True, scale=float(scale), mask=attention_mask, rope_applied=fused)
        return self.c_proj(ctx, residual=residual)                              # + residual in the GEMM epilogue


class GPT2MLP(Module):
    def __init__(self, intermediate_size: int, config: Dict[str, Any]):
        super().__init__()
        self.swiglu = SwiGLU(config["n_embd"])
        self.dropout = Dropout(config.get("resid_pdrop", 0.1))

    def forward(self, hidden_states: Tensor, residual: Optional[Tensor] = None) -> Tensor:
        return self.swiglu(hidden_states, residual=residual)


class GPT2Block(Module):
    def __init__(self, config: Dict[str, Any], layer_idx: Optional[int] = None):
        super().__init__()
        d = config["n_embd"]
        eps = config.get("layer_norm_epsilon", 1e-5)
        self.ln_1 = RMSNorm(d, eps=eps)
        self.attn = GPT2Attention(config, layer_idx)
        self.ln_2 = RMSNorm(d, eps=eps)
        self.mlp = GPT2MLP(4 * d, config)

    def forward(self, hidden_states: Tensor, attention_mask=None) -> Tensor:
        act_dtype = _activation_dtype(hidden_states)
        h = self.attn(self.ln_1(hidden_states, out_dtype=act_dtype), attention_mask, residual=hidden_states)
        return self.mlp(self.ln_2(h, out_dtype=act_dtype), residual=h)


def _activation_dtype(x: Tensor):
    """bf16 activations feed the tensor-core GEMMs in bf16 mode; the residual stream stays fp32."""
    if isinstance(x.data, B200Array):
        from .. import kernels as K
        if K.get_precision() == "bf16":
            return bfloat16
    return None


class GPT2Model(Module):
    def __init__(self, config: Dict[str, Any]):
        super().__init__()
        self.embed_dim = config["n_embd"]
        self.wte = Embedding(config["vocab_size"], config["n_embd"])
        self.drop = Dropout(config.get("embd_pdrop", 0.1))
        self.h = ModuleList()
        for i in range(config["n_layer"]):
            self.h.append(GPT2Block(config, layer_idx=i))
        self.ln_f = RMSNorm(config["n_embd"], eps=config.get("layer_norm_epsilon", 1e-5))
        self._init_weights()

    def _init_weights(self):
        self.wte.weight.data = np.random.randn(*self.wte.weight.shape).astype(np.float32) * 0.02  # gpt2.py:369-373

    def get_input_embeddings(self):
        return self.wte

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, **_ignored) -> Tensor:
        x = self.wte(input_ids) if inputs_embeds is None else inputs_embeds
        for block in self.h:
            x = block(x, attention_mask)
        return self.ln_f(x, out_dtype=_activation_dtype(x))


GPT2_CONFIGS = {
    "small": {"vocab_size": 50257, "n_positions": 1024, "n_embd": 768, "n_layer": 12, "n_head": 12},
    "medium": {"vocab_size": 50257, "n_positions": 1024, "n_embd": 1024, "n_layer": 24, "n_head": 16},
    "large": {"vocab_size": 50257, "n_positions": 1024, "n_embd": 1280, "n_layer": 36, "n_head": 20},
    "xl": {"vocab_size": 50257, "n_positions": 1024, "n_embd": 1600, "n_layer": 48, "n_head": 25},
}


class GPT2LMHead(Module):
    def __init__(self, config: Dict[str, Any]):
        super().__init__()
        self.config = dict(config)
        self.transformer = GPT2Model(config)
        self.lm_head = Linear(config["n_embd"], config["vocab_size"], bias=False)
        self.tie_weights()

    def tie_weights(self):
        self.lm_head.weight = self.transformer.wte.weight  # gpt2.py:435-440: one Parameter, two names

    def forward(self, input_ids=None, attention_mask=None, inputs_embeds=None, labels=None, **kw):
        hidden = self.transformer(input_ids, attention_mask=attention_mask, inputs_embeds=inputs_embeds)
        B, S, d = hidden.shape
        # training step on the bf16 tensor-core path: the (T, V) logits stay bf16 (412 MB instead of 825 MB at 8x512x50257:
        # the LM-head GEMM is HBM-write-bound and the loss kernel HBM-read-bound); the softmax / loss math is fp32 either way
        from .. import kernels as K
        from ..array import B200Array, bfloat16
        train_bf16 = labels is not None and isinstance(hidden.data, B200Array) and K.get_precision() == "bf16"
        logits2d = self.lm_head(F.reshape(hidden, (B * S, d)), out_dtype=bfloat16 if train_bf16 else np.float32)
        out = {"logits": F.reshape(logits2d, (B, S, logits2d.shape[-1]))}
        if labels is not None:
            lab = labels.data if isinstance(labels, Tensor) else labels
            out["loss"] = F.cross_entropy_loss(logits2d, lab.reshape(-1))  # examples/training/gpt2_training.py:355-361
        return out


class GPT2(GPT2LMHead):
    pass


def gpt2_small(**kw):
    return GPT2({**GPT2_CONFIGS["small"], **kw})


def gpt2_medium(**kw):
    return GPT2({**GPT2_CONFIGS["medium"], **kw})


def gpt2_large(**kw):
    return GPT2({**GPT2_CONFIGS["large"], **kw})


def gpt2_xl(**kw):
    return GPT2({**GPT2_CONFIGS["xl"], **kw})
