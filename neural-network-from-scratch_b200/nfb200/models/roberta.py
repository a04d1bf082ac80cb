"""RoBERTa as the reference defines it (models/language/roberta.py:35-415): BERT's encoder and pooler under
RoBERTa's configuration (vocab 50265, 514 positions, one token type, LayerNorm eps 1e-5, pad id 1) with position ids
counted from ``pad_token_id + 1`` (roberta.py:108-114).  Same module tree / parameter names (``roberta.embeddings...``,
``lm_head``, ``classifier``) and the same initialisation -- including its quirk: ``init_weights`` redraws EVERY module that
has both a ``weight`` and a ``bias`` attribute from N(0, 0.02^2), which covers the LayerNorm gains as well as the Linear
layers, while the embedding tables keep their constructor values (roberta.py:149-163).

Everything runs on the BERT path of this package (models/bert.py): fused flash attention with the additive per-key mask,
GELU in the GEMM epilogue, residual adds folded into the LayerNorm kernels, bf16 tensor-core GEMMs in bf16 mode.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .. import functional as F
from ..core import Module, Tensor
from ..nn import Dropout, Linear
from .bert import BERT, BERTConfig, BERTEmbeddings, BERTEncoder, BERTPooler


class RoBERTaConfig(BERTConfig):
    def __init__(self, vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=514,
                 type_vocab_size=1, initializer_range=0.02, layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0, eos_token_id=2,
                 position_embedding_type="absolute", use_cache=True, classifier_dropout=None):
        super().__init__(vocab_size=vocab_size, hidden_size=hidden_size, num_hidden_layers=num_hidden_layers,
                         num_attention_heads=num_attention_heads, intermediate_size=intermediate_size, hidden_act=hidden_act,
                         hidden_dropout_prob=hidden_dropout_prob, attention_probs_dropout_prob=attention_probs_dropout_prob,
                         max_position_embeddings=max_position_embeddings, type_vocab_size=type_vocab_size,
                         initializer_range=initializer_range, layer_norm_eps=layer_norm_eps, pad_token_id=pad_token_id,
                         position_embedding_type=position_embedding_type, use_cache=use_cache, classifier_dropout=classifier_dropout)
        self.bos_token_id, self.eos_token_id = bos_token_id, eos_token_id


class RoBERTaEmbeddings(BERTEmbeddings):
    def __init__(self, config: RoBERTaConfig):
        super().__init__(config)
        self.padding_idx = config.pad_token_id

    def forward(self, input_ids=None, token_type_ids=None, position_ids=None, inputs_embeds=None) -> Tensor:
        if position_ids is None:
            shape = input_ids.shape if input_ids is not None else inputs_embeds.shape[:-1]
            dev = self.word_embeddings.weight.device
            cache = self.__dict__.setdefault("_const_ids", {})
            key = ("rpos", int(shape[1]), str(dev))
            if key not in cache:   # constant of the sequence length: built once, stays on the device (graph-capturable step)
                lo = self.padding_idx + 1
                cache[key] = Tensor(np.arange(lo, shape[1] + lo, dtype=np.int64).reshape(1, -1), device=dev)
            position_ids = cache[key]
        return super().forward(input_ids, token_type_ids, position_ids, inputs_embeds)


class RoBERTa(BERT):
    def __init__(self, config: Optional[RoBERTaConfig] = None, **kwargs):
        Module.__init__(self)
        self.config = config or RoBERTaConfig(**kwargs)
        self.embeddings = RoBERTaEmbeddings(self.config)
        self.encoder = BERTEncoder(self.config)
        self.pooler = BERTPooler(self.config)
        self.init_weights()

    def init_weights(self):
        """roberta.py:149-163: every module exposing BOTH ``weight`` and ``bias`` (Linear and LayerNorm; not Embedding)."""
        std = self.config.initializer_range
        for m in self.modules():
            if hasattr(m, "weight") and hasattr(m, "bias"):
                if hasattr(m.weight, "data"):
                    m.weight.data = np.random.normal(0, std, m.weight.shape).astype(np.float32)
                if m.bias is not None and hasattr(m.bias, "data"):
                    m.bias.data = np.zeros(m.bias.shape, dtype=np.float32)

    def set_input_embeddings(self, value):
        self.embeddings.word_embeddings = value

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, inputs_embeds=None,
                with_pooler: bool = True, **kw) -> Dict[str, Tensor]:
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        return super().forward(input_ids, attention_mask, token_type_ids, position_ids, inputs_embeds, with_pooler=with_pooler, **kw)


class RoBERTaForMaskedLM(Module):
    def __init__(self, config: Optional[RoBERTaConfig] = None, **kwargs):
        super().__init__()
        self.config = config or RoBERTaConfig(**kwargs)
        self.roberta = RoBERTa(self.config)
        self.lm_head = Linear(self.config.hidden_size, self.config.vocab_size)

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, labels=None, **kw) -> Dict[str, Tensor]:
        out = self.roberta(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids, with_pooler=False, **kw)
        seq = out["last_hidden_state"]
        B, S, d = seq.shape
        logits2d = self.lm_head(F.reshape(seq, (B * S, d)), out_dtype=np.float32)
        res = {"logits": F.reshape(logits2d, (B, S, logits2d.shape[-1])), "hidden_states": None, "attentions": None}
        if labels is not None:
            lab = labels.data if isinstance(labels, Tensor) else labels
            res["loss"] = F.cross_entropy_loss(logits2d, lab.reshape(-1))
        return res


class RoBERTaForSequenceClassification(Module):
    def __init__(self, config: Optional[RoBERTaConfig] = None, num_labels: int = 2, **kwargs):
        super().__init__()
        self.config = config or RoBERTaConfig(**kwargs)
        self.num_labels = num_labels
        self.roberta = RoBERTa(self.config)
        p = self.config.classifier_dropout if self.config.classifier_dropout is not None else self.config.hidden_dropout_prob
        self.dropout = Dropout(p)
        self.classifier = Linear(self.config.hidden_size, num_labels)

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, labels=None, **kw) -> Dict[str, Tensor]:
        out = self.roberta(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids, **kw)
        logits = self.classifier(self.dropout(out["pooler_output"]), out_dtype=np.float32)
        res = {"logits": logits, "hidden_states": None, "attentions": None}
        if labels is not None:
            lab = labels.data if isinstance(labels, Tensor) else labels
            res["loss"] = F.cross_entropy_loss(logits, lab.reshape(-1))
        return res


ROBERTA_CONFIGS = {
    "base": dict(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0, eos_token_id=2),
    "large": dict(vocab_size=50265, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                  max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0, eos_token_id=2),
}


def roberta_base(**kw):
    return RoBERTa(RoBERTaConfig(**{**ROBERTA_CONFIGS["base"], **kw}))


def roberta_large(**kw):
    return RoBERTa(RoBERTaConfig(**{**ROBERTA_CONFIGS["large"], **kw}))
