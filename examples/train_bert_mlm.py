"""BERT masked-LM training on the b200 backend: BASELINE.json's BERT-base workload (`BERTForMaskedLM`, the loss over all
tokens as models/language/bert.py:523-529 computes it, AdamW lr 2e-5 / weight decay 0.01 as in
examples/training/bert_training.py:44-45,274-278) with the loop of that script (:308-360).  Synthetic token batches:
15 % of the positions are replaced by the mask id 4, the labels are the original ids.

    python examples/train_bert_mlm.py --device cuda --hidden-size 768 --num-hidden-layers 12 --num-attention-heads 12 --intermediate-size 3072 --max-seq-length 512 --batch-size 32 --vocab-size 30522   # BERT-base
    python examples/train_bert_mlm.py --device cpu --hidden-size 32 --num-hidden-layers 2 ...   # host path (tests)
"""
import os
import sys
from dataclasses import dataclass

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import Trainer, run  # noqa: E402


@dataclass
class TrainingConfig:
    """Field names and defaults of examples/training/bert_training.py:32-58 (the classification-only `num_classes` dropped)."""
    vocab_size: int = 30000
    hidden_size: int = 512
    num_hidden_layers: int = 6
    num_attention_heads: int = 8
    intermediate_size: int = 2048
    max_seq_length: int = 128
    batch_size: int = 16
    learning_rate: float = 2e-5
    weight_decay: float = 0.01
    num_epochs: int = 3
    warmup_steps: int = 1000
    max_grad_norm: float = 1.0
    train_size: int = 5000
    val_size: int = 1000
    enable_optimizations: bool = True
    save_checkpoints: bool = True
    checkpoint_dir: str = "checkpoints/bert"


MASK_ID = 4


def synthetic_mlm(n, length, vocab, seed):
    """(masked ids, attention mask, labels): periodic sequences (a token determines the one two places on), 15 % masked,
    trailing padding (id 0, attention 0) of random length."""
    rng = np.random.default_rng(seed)
    ids = np.empty((n, length), np.int64)
    ids[:, :2] = rng.integers(5, vocab, (n, 2))
    for t in range(2, length):
        ids[:, t] = (ids[:, t - 2] * 5 + 1) % (vocab - 5) + 5
    att = np.ones((n, length), np.float32)
    for r in range(n):
        pad = int(rng.integers(0, max(1, length // 4)))
        if pad:
            ids[r, length - pad:] = 0
            att[r, length - pad:] = 0.0
    masked = np.where((rng.random((n, length)) < 0.15) & (att > 0), MASK_ID, ids)
    return masked, att, ids


class BERTTrainer(Trainer):
    name = "bert"

    def build_model(self):
        from nfb200.models import BERTConfig, BERTForMaskedLM
        c = self.config
        return BERTForMaskedLM(BERTConfig(vocab_size=c.vocab_size, hidden_size=c.hidden_size, num_hidden_layers=c.num_hidden_layers,
                                          num_attention_heads=c.num_attention_heads, intermediate_size=c.intermediate_size,
                                          max_position_embeddings=c.max_seq_length))

    def make_data(self, seed):
        c = self.config
        return synthetic_mlm(c.train_size, c.max_seq_length, c.vocab_size, seed), synthetic_mlm(c.val_size, c.max_seq_length, c.vocab_size, seed + 1)

    def forward(self, batch):
        from nfb200.core import Tensor
        masked, att, labels = batch
        out = self.net(Tensor(masked), attention_mask=Tensor(att), labels=Tensor(labels))
        return out["loss"], out["logits"], labels


def main(argv=None):
    return run(BERTTrainer, TrainingConfig, __doc__, argv)


if __name__ == "__main__":
    main()
