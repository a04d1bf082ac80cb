"""GPT-2 causal-LM training on the b200 backend: the workload of the reference's examples/training/gpt2_training.py
(config fields :33-61, loop order :417-486, per-epoch checkpoint record :528-547) on this package's mirror of the same API.
Data are synthetic token sequences (the reference's toy tokenizer / story generator is not part of the path; BASELINE.json
asks for synthetic batches).

    python examples/train_gpt2.py --device cuda                              # one B200, bf16 tensor-core path
    python -m nfb200.distributed.launcher --nproc-per-node 8 examples/train_gpt2.py --device cuda
    python examples/train_gpt2.py --device cpu --n-embd 32 --n-layer 2 ...   # host path (what tests/test_examples.py runs)

Differences from upstream, deliberately: the loss is computed inside the model's connected graph (upstream re-wraps the
logits in a fresh Tensor, which cuts the gradient: PARITY.md #1); clipping is the global-norm form of the reference's
translation trainer (examples/translation/train_conversational.py:38-54), which stays on the device."""
import os
import sys
from dataclasses import dataclass

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import Trainer, run  # noqa: E402


@dataclass
class GPT2TrainingConfig:
    """Field names and defaults of examples/training/gpt2_training.py:33-61."""
    vocab_size: int = 10000
    n_embd: int = 256
    n_layer: int = 4
    n_head: int = 4
    n_ctx: int = 128
    batch_size: int = 8
    learning_rate: float = 1e-4
    weight_decay: float = 0.1
    num_epochs: int = 5
    warmup_steps: int = 500
    max_grad_norm: float = 1.0
    train_size: int = 2000
    val_size: int = 400
    generate_every: int = 100
    max_generate_length: int = 50
    enable_optimizations: bool = True
    save_checkpoints: bool = True
    checkpoint_dir: str = "checkpoints/gpt2"


def synthetic_sequences(n, n_ctx, vocab, seed):
    """Token sequences with learnable structure: each token is a fixed function of the previous one plus noise."""
    rng = np.random.default_rng(seed)
    seq = np.empty((n, n_ctx + 1), np.int64)
    seq[:, 0] = rng.integers(4, vocab, n)
    for t in range(1, n_ctx + 1):
        nxt = (seq[:, t - 1] * 7 + 3) % (vocab - 4) + 4
        seq[:, t] = np.where(rng.random(n) < 0.9, nxt, rng.integers(4, vocab, n))
    return seq


class GPT2Trainer(Trainer):
    name = "gpt2"

    def build_model(self):
        from nfb200.models import GPT2
        c = self.config
        return GPT2({"vocab_size": c.vocab_size, "n_embd": c.n_embd, "n_layer": c.n_layer, "n_head": c.n_head, "n_ctx": c.n_ctx,
                     "n_positions": c.n_ctx})

    def make_data(self, seed):
        c = self.config
        return (synthetic_sequences(c.train_size, c.n_ctx, c.vocab_size, seed),), (synthetic_sequences(c.val_size, c.n_ctx, c.vocab_size, seed + 1),)

    def forward(self, batch):
        from nfb200.core import Tensor
        seq, = batch
        ids, targets = np.ascontiguousarray(seq[:, :-1]), np.ascontiguousarray(seq[:, 1:])
        out = self.net(Tensor(ids), labels=Tensor(targets))
        return out["loss"], out["logits"], targets


def main(argv=None):
    return run(GPT2Trainer, GPT2TrainingConfig, __doc__, argv)


if __name__ == "__main__":
    main()
