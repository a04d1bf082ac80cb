"""The translation encoder-decoder of BASELINE config 1 (examples/translation/model_v2.py:30-165): Embedding * sqrt(d)
+ sinusoidal positional encoding -> n x TransformerBlock ; decoder: n x TransformerDecoderBlock (memory ignored, as
upstream) -> Linear(d -> tgt vocab).  Built from nn.TransformerBlock / nn.TransformerDecoderBlock / nn.Linear of this
package, i.e. the reference's own layer definitions (placeholder MultiHeadAttention included)."""
from __future__ import annotations

import numpy as np

from ..array import B200Array
from ..core import Module, Tensor
from ..nn import Embedding, Linear, ModuleList, TransformerBlock, TransformerDecoderBlock


class PositionalEncoding(Module):
    """Fixed sinusoidal table added to the embeddings (examples/translation/model_v2.py:14-45)."""

    def __init__(self, d_model: int, max_len: int = 5000):
        super().__init__()
        pe = np.zeros((max_len, d_model), dtype=np.float32)
        pos = np.arange(0, max_len, dtype=np.float32)[:, None]
        div = np.exp(np.arange(0, d_model, 2).astype(np.float32) * -(np.log(10000.0) / d_model))
        pe[:, 0::2] = np.sin(pos * div)
        pe[:, 1::2] = np.cos(pos * div)
        self.pe = pe
        self._dev_cache = {}

    def forward(self, x: Tensor) -> Tensor:
        S = x.shape[1]
        if isinstance(x.data, B200Array):
            t = self._dev_cache.get(S)
            if t is None:
                t = Tensor(self.pe[None, :S], device=x.device)
                self._dev_cache[S] = t
            return x + t
        return x + Tensor(self.pe[None, :S])


class TranslationTransformer(Module):
    def __init__(self, src_vocab_size: int, tgt_vocab_size: int, d_model: int = 128, n_heads: int = 4, n_layers: int = 3, d_ff: int = 256,
                 max_seq_len: int = 100, dropout: float = 0.1):
        super().__init__()
        self.d_model = d_model
        self.src_embedding = Embedding(src_vocab_size, d_model)
        self.tgt_embedding = Embedding(tgt_vocab_size, d_model)
        self.pos_encoding = PositionalEncoding(d_model, max_seq_len)
        self.encoder_layers = ModuleList([TransformerBlock(d_model, n_heads, d_ff, dropout) for _ in range(n_layers)])
        self.decoder_layers = ModuleList([TransformerDecoderBlock(d_model, n_heads, d_ff, dropout) for _ in range(n_layers)])
        self.output_projection = Linear(d_model, tgt_vocab_size)
        self.scale = float(np.sqrt(d_model))

    def encode(self, src, src_mask=None) -> Tensor:
        x = self.pos_encoding(self.src_embedding(src) * self.scale)
        for layer in self.encoder_layers:
            x = layer(x, mask=src_mask)
        return x

    def decode(self, tgt, memory, tgt_mask=None, memory_mask=None) -> Tensor:
        x = self.pos_encoding(self.tgt_embedding(tgt) * self.scale)
        for layer in self.decoder_layers:
            x = layer(x, memory, tgt_mask=tgt_mask, memory_mask=memory_mask)
        return x

    def forward(self, src, tgt, src_mask=None, tgt_mask=None) -> Tensor:
        memory = self.encode(src, src_mask)
        return self.output_projection(self.decode(tgt, memory, tgt_mask, src_mask), out_dtype=np.float32)
