"""BERT as the reference defines it (models/language/bert.py:18-729): word + token-type (+ position) embeddings ->
LayerNorm(eps 1e-12) -> N x [self-attention with additive (1 - mask) * -1e4 key mask, dense + LayerNorm(h + x),
Linear(d -> 4d) fused with tanh-GELU, Linear(4d -> d) + LayerNorm(h + x)] -> pooler (tanh) ; MLM head = untied
Linear(d -> vocab) with cross-entropy over ALL tokens (bert.py:516-529).  Same module tree / parameter names as the
reference.  On a cuda device in bf16 mode: q/k/v projections feed the fused flash kernel (additive per-key mask), the
GELU lives in the GEMM epilogue, LayerNorm(h + x) takes the residual add from the preceding GEMM's epilogue."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .. import functional as F
from ..array import B200Array, bfloat16
from ..core import Module, Tensor
from ..nn import Dropout, Embedding, LayerNorm, Linear, ModuleList, Tanh


class BERTConfig:
    def __init__(self, vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=512,
                 type_vocab_size=2, initializer_range=0.02, layer_norm_eps=1e-12, pad_token_id=0, position_embedding_type="absolute",
                 use_cache=True, classifier_dropout=None):
        self.vocab_size, self.hidden_size, self.num_hidden_layers = vocab_size, hidden_size, num_hidden_layers
        self.num_attention_heads, self.intermediate_size, self.hidden_act = num_attention_heads, intermediate_size, hidden_act
        self.hidden_dropout_prob, self.attention_probs_dropout_prob = hidden_dropout_prob, attention_probs_dropout_prob
        self.max_position_embeddings, self.type_vocab_size = max_position_embeddings, type_vocab_size
        self.initializer_range, self.layer_norm_eps, self.pad_token_id = initializer_range, layer_norm_eps, pad_token_id
        self.position_embedding_type, self.use_cache = position_embedding_type, use_cache
        self.classifier_dropout = classifier_dropout or hidden_dropout_prob


def _act_dtype(x: Tensor):
    if isinstance(x.data, B200Array):
        from .. import kernels as K
        if K.get_precision() == "bf16":
            return bfloat16
    return None


class BERTEmbeddings(Module):
    def __init__(self, config: BERTConfig):
        super().__init__()
        self.word_embeddings = Embedding(config.vocab_size, config.hidden_size)
        self.position_embeddings = Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = Embedding(config.type_vocab_size, config.hidden_size)
        self.pad_token_id = config.pad_token_id
        self.LayerNorm = LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = Dropout(config.hidden_dropout_prob)
        self.position_embedding_type = config.position_embedding_type

    def forward(self, input_ids=None, token_type_ids=None, position_ids=None, inputs_embeds=None) -> Tensor:
        shape = input_ids.shape if input_ids is not None else inputs_embeds.shape[:-1]
        dev = self.word_embeddings.weight.device
        # default position / segment ids are constants of the input shape: built once per (shape, device) and kept on the
        # device, so a training step makes no host->device copy (and can be captured into a CUDA graph)
        cache = self.__dict__.setdefault("_const_ids", {})
        if position_ids is None:
            key = ("pos", int(shape[1]), str(dev))
            if key not in cache:
                cache[key] = Tensor(np.arange(shape[1], dtype=np.int64).reshape(1, -1), device=dev)
            position_ids = cache[key]
        if token_type_ids is None:
            key = ("seg", tuple(int(v) for v in shape), str(dev))
            if key not in cache:
                cache[key] = Tensor(np.zeros(shape, dtype=np.int64), device=dev)
            token_type_ids = cache[key]
        if inputs_embeds is None:
            inputs_embeds = self.word_embeddings(input_ids)
        emb = inputs_embeds + self.token_type_embeddings(token_type_ids)
        if self.position_embedding_type == "absolute":
            emb = emb + self.position_embeddings(position_ids)
        return self.LayerNorm(emb)


class BERTSelfAttention(Module):
    def __init__(self, config: BERTConfig):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError(f"The hidden size ({config.hidden_size}) is not a multiple of the number of attention heads ({config.num_attention_heads})")
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = Linear(config.hidden_size, self.all_head_size)
        self.key = Linear(config.hidden_size, self.all_head_size)
        self.value = Linear(config.hidden_size, self.all_head_size)
        self.dropout = Dropout(config.attention_probs_dropout_prob)

    def _heads(self, x: Tensor) -> Tensor:
        B, S, _ = x.shape
        return F.transpose(F.reshape(x, (B, S, self.num_attention_heads, self.attention_head_size)), (0, 2, 1, 3))

    def forward(self, hidden_states: Tensor, attention_mask=None) -> Tensor:
        dt = _act_dtype(hidden_states)
        q = self._heads(self.query(hidden_states, out_dtype=dt))
        k = self._heads(self.key(hidden_states, out_dtype=dt))
        v = self._heads(self.value(hidden_states, out_dtype=dt))
        scale = 1.0 / float(np.sqrt(self.attention_head_size))
        ctx = F.scaled_dot_product_attention(q, k, v, scale, causal=False, mask=attention_mask)   # bert.py:186-213
        B, H, S, D = ctx.shape
        return F.reshape(F.transpose(ctx, (0, 2, 1, 3)), (B, S, H * D))


class BERTSelfOutput(Module):
    def __init__(self, config: BERTConfig):
        super().__init__()
        self.dense = Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states: Tensor, input_tensor: Tensor) -> Tensor:
        return self.LayerNorm(self.dense(hidden_states, residual=input_tensor))   # LayerNorm(dense(h) + x)


class BERTAttention(Module):
    def __init__(self, config: BERTConfig):
        super().__init__()
        self.self = BERTSelfAttention(config)
        self.output = BERTSelfOutput(config)

    def forward(self, hidden_states: Tensor, attention_mask=None) -> Tensor:
        return self.output(self.self(hidden_states, attention_mask), hidden_states)


class BERTIntermediate(Module):
    def __init__(self, config: BERTConfig):
        super().__init__()
        self.dense = Linear(config.hidden_size, config.intermediate_size, activation="gelu", enable_fusion=True)

    def forward(self, hidden_states: Tensor) -> Tensor:
        return self.dense(hidden_states, out_dtype=_act_dtype(hidden_states))


class BERTOutput(Module):
    def __init__(self, config: BERTConfig):
        super().__init__()
        self.dense = Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states: Tensor, input_tensor: Tensor) -> Tensor:
        return self.LayerNorm(self.dense(hidden_states, residual=input_tensor))


class BERTLayer(Module):
    def __init__(self, config: BERTConfig):
        super().__init__()
        self.attention = BERTAttention(config)
        self.intermediate = BERTIntermediate(config)
        self.output = BERTOutput(config)

    def forward(self, hidden_states: Tensor, attention_mask=None) -> Tensor:
        a = self.attention(hidden_states, attention_mask)
        return self.output(self.intermediate(a), a)


class BERTEncoder(Module):
    def __init__(self, config: BERTConfig):
        super().__init__()
        self.layer = ModuleList([BERTLayer(config) for _ in range(config.num_hidden_layers)])

    def forward(self, hidden_states: Tensor, attention_mask=None) -> Tensor:
        for layer in self.layer:
            hidden_states = layer(hidden_states, attention_mask)
        return hidden_states


class BERTPooler(Module):
    def __init__(self, config: BERTConfig):
        super().__init__()
        self.dense = Linear(config.hidden_size, config.hidden_size)
        self.activation = Tanh()

    def forward(self, hidden_states: Tensor) -> Tensor:
        return self.activation(self.dense(hidden_states[:, 0]))   # bert.py:375-388


class BERT(Module):
    def __init__(self, config: Optional[BERTConfig] = None, **kwargs):
        super().__init__()
        self.config = config or BERTConfig(**kwargs)
        self.embeddings = BERTEmbeddings(self.config)
        self.encoder = BERTEncoder(self.config)
        self.pooler = BERTPooler(self.config)
        self.init_weights()

    def init_weights(self):
        """N(0, 0.02^2) for every Linear / Embedding weight, biases zero, in module-traversal order (bert.py:406-426)."""
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, Linear):
                m.weight.data = np.random.normal(0, std, m.weight.shape).astype(np.float32)
                if m.bias is not None:
                    m.bias.data = np.zeros(m.bias.shape, dtype=np.float32)
            elif isinstance(m, Embedding):
                m.weight.data = np.random.normal(0, std, m.weight.shape).astype(np.float32)

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, inputs_embeds=None,
                with_pooler: bool = True, **_ignored) -> Dict[str, Tensor]:
        shape = input_ids.shape if input_ids is not None else inputs_embeds.shape[:-1]
        ext = None
        if attention_mask is not None:
            m = attention_mask.data if isinstance(attention_mask, Tensor) else attention_mask
            if not isinstance(m, B200Array):
                m = np.asarray(m, dtype=np.float32)
            else:
                m = m.astype(np.float32)
            ext = (1.0 - m.reshape(shape[0], 1, 1, shape[1])) * -10000.0        # bert.py:456-457
            ext = Tensor(ext, device=self.embeddings.word_embeddings.weight.device)
        h = self.embeddings(input_ids, token_type_ids, position_ids, inputs_embeds)
        seq = self.encoder(h, ext)
        return {"last_hidden_state": seq, "pooler_output": self.pooler(seq) if with_pooler else None, "hidden_states": None, "attentions": None}


class BERTForMaskedLM(Module):
    def __init__(self, config: Optional[BERTConfig] = None, **kwargs):
        super().__init__()
        self.config = config or BERTConfig(**kwargs)
        self.bert = BERT(self.config)
        self.cls = Linear(self.config.hidden_size, self.config.vocab_size)

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, labels=None, **kw) -> Dict[str, Tensor]:
        out = self.bert(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids, with_pooler=False, **kw)
        seq = out["last_hidden_state"]
        B, S, d = seq.shape
        logits2d = self.cls(F.reshape(seq, (B * S, d)), out_dtype=np.float32)
        res = {"logits": F.reshape(logits2d, (B, S, logits2d.shape[-1])), "hidden_states": None, "attentions": None}
        if labels is not None:
            lab = labels.data if isinstance(labels, Tensor) else labels
            res["loss"] = F.cross_entropy_loss(logits2d, lab.reshape(-1))   # all tokens, as the reference does
        return res


def bert_base(**kw):
    return BERT(BERTConfig(**kw))


def bert_base_mlm(**kw):
    return BERTForMaskedLM(BERTConfig(**kw))
