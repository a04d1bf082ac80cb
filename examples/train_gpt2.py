"""GPT-2 causal-LM training loop on the b200 backend: the workload of the reference's examples/training/gpt2_training.py
(config fields :33-61, loop order :417-486, per-epoch checkpoint record :528-547) written against this package's mirror of
the same API.  Data are synthetic token sequences (the reference's toy tokenizer / story generator is not part of the
path; BASELINE.json asks for synthetic batches).

    python examples/train_gpt2.py --device cuda                              # one B200, bf16 tensor-core path
    python -m nfb200.distributed.launcher --nproc-per-node 8 examples/train_gpt2.py --device cuda
    python examples/train_gpt2.py --device cpu --n-embd 32 --n-layer 2 ...   # host path (what tests/test_examples.py runs)

Loop order, as upstream: forward -> cross entropy on the flattened logits -> backward -> gradient clipping at
`max_grad_norm` -> AdamW step -> zero_grad -> loss / perplexity / accuracy; validation and one checkpoint per epoch.
Differences, deliberately: the loss is computed inside the model's connected graph (upstream re-wraps the logits in a
fresh Tensor, which cuts the gradient: PARITY.md #1); clipping is the global-norm form of the reference's translation
trainer (examples/translation/train_conversational.py:38-54), which stays on the device; under the launcher the model is
wrapped in DistributedDataParallel and every rank draws its shard with DistributedSampler."""
import argparse
import json
import os
import sys
import time
from dataclasses import asdict, dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "neural-network-from-scratch_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


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
        noise = rng.integers(4, vocab, n)
        seq[:, t] = np.where(rng.random(n) < 0.9, nxt, noise)
    return seq


class GPT2Trainer:
    def __init__(self, config: GPT2TrainingConfig, device: str = "cpu", precision: str = "fp32", seed: int = 0):
        from nfb200 import distributed as dist
        from nfb200 import kernels as K
        from nfb200.core import set_default_device
        from nfb200.models import GPT2
        from nfb200.optim import AdamW
        self.config, self.device, self.step = config, device, 0
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl" if device != "cpu" else "gloo")
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if device != "cpu":
            device = f"cuda:{os.environ.get('LOCAL_RANK', '0')}"
            K.set_precision(precision)
        set_default_device(device)
        np.random.seed(seed)                                   # the reference's models draw from the global NumPy RNG
        self.model = GPT2({"vocab_size": config.vocab_size, "n_embd": config.n_embd, "n_layer": config.n_layer, "n_head": config.n_head,
                           "n_ctx": config.n_ctx, "n_positions": config.n_ctx})
        self.ddp = dist.DistributedDataParallel(self.model) if self.world > 1 else None
        self.optimizer = AdamW(self.model.parameters(), lr=config.learning_rate, weight_decay=config.weight_decay)
        self.train_seq = synthetic_sequences(config.train_size, config.n_ctx, config.vocab_size, seed + 1)
        self.val_seq = synthetic_sequences(config.val_size, config.n_ctx, config.vocab_size, seed + 2)
        self.metrics = {"train_losses": [], "val_losses": [], "val_perplexities": []}

    # ------------------------------------------------------------------ one step
    def forward_pass(self, ids, targets):
        from nfb200.core import Tensor
        net = self.ddp if self.ddp is not None else self.model
        out = net(Tensor(ids), labels=Tensor(targets))
        return out["loss"], out["logits"]

    def _metrics(self, loss, logits, targets):
        lv = float(loss.item())
        pred = np.asarray(logits.numpy()).reshape(-1, self.config.vocab_size).argmax(1)
        return {"loss": lv, "perplexity": float(np.exp(min(lv, 50.0))), "accuracy": float(np.mean(pred == targets.reshape(-1)))}

    def train_step(self, ids, targets):
        from nfb200.optim import clip_grad_norm
        loss, logits = self.forward_pass(ids, targets)
        loss.backward()
        if self.ddp is not None:
            self.ddp.sync_gradients()
        clip_grad_norm(self.model.parameters(), self.config.max_grad_norm, optimizer=self.optimizer)
        self.optimizer.step()
        self.optimizer.zero_grad()
        self.step += 1
        return loss, logits

    # ------------------------------------------------------------------ epochs
    def _shard(self, n, epoch, shuffle):
        if self.world > 1:
            s = self.dist.DistributedSampler(n, num_replicas=self.world, rank=self.rank, shuffle=shuffle, seed=0)
            s.set_epoch(epoch)
            return np.fromiter(iter(s), dtype=np.int64)
        return np.random.RandomState(epoch).permutation(n) if shuffle else np.arange(n)

    def train_epoch(self, epoch: int):
        c = self.config
        order = self._shard(len(self.train_seq), epoch, True)
        tot = {"loss": 0.0, "perplexity": 0.0, "accuracy": 0.0}
        nb, t0 = 0, time.time()
        for lo in range(0, len(order) - c.batch_size + 1, c.batch_size):
            batch = self.train_seq[order[lo:lo + c.batch_size]]
            ids, targets = batch[:, :-1], batch[:, 1:]
            loss, logits = self.train_step(np.ascontiguousarray(ids), np.ascontiguousarray(targets))
            m = self._metrics(loss, logits, targets)
            for k in tot:
                tot[k] += m[k]
            nb += 1
            if self.rank == 0 and nb % 10 == 1:
                print(f"  Batch {nb}: Loss = {m['loss']:.4f}, PPL = {m['perplexity']:.2f}, Acc = {m['accuracy']:.4f}", flush=True)
        out = {k: v / max(nb, 1) for k, v in tot.items()}
        out["time"] = time.time() - t0
        return out

    def validate(self):
        from nfb200.core import no_grad
        c = self.config
        tot = {"loss": 0.0, "perplexity": 0.0, "accuracy": 0.0}
        nb, t0 = 0, time.time()
        with no_grad():
            for lo in range(0, len(self.val_seq) - c.batch_size + 1, c.batch_size):
                batch = self.val_seq[lo:lo + c.batch_size]
                ids, targets = np.ascontiguousarray(batch[:, :-1]), np.ascontiguousarray(batch[:, 1:])
                loss, logits = self.forward_pass(ids, targets)
                m = self._metrics(loss, logits, targets)
                for k in tot:
                    tot[k] += m[k]
                nb += 1
        out = {k: v / max(nb, 1) for k, v in tot.items()}
        out["time"] = time.time() - t0
        return out

    def save_checkpoint(self, epoch: int, metrics):
        """The reference's record (gpt2_training.py:528-547) plus the weights / optimizer state (nfb200/checkpoint.py)."""
        if not self.config.save_checkpoints:
            return None
        from nfb200.checkpoint import save_checkpoint
        return save_checkpoint(os.path.join(self.config.checkpoint_dir, f"gpt2_epoch_{epoch + 1}"), self.model, self.optimizer,
                               epoch=epoch, step=self.step, config=asdict(self.config), metrics=metrics)

    def train(self):
        best = float("inf")
        for epoch in range(self.config.num_epochs):
            if self.rank == 0:
                print(f"\nEpoch {epoch + 1}/{self.config.num_epochs}", flush=True)
            tr = self.train_epoch(epoch)
            va = self.validate()
            self.metrics["train_losses"].append(tr["loss"])
            self.metrics["val_losses"].append(va["loss"])
            self.metrics["val_perplexities"].append(va["perplexity"])
            best = min(best, va["perplexity"])
            if self.rank == 0:
                print(f"  Training: Loss = {tr['loss']:.4f}, PPL = {tr['perplexity']:.2f}, Acc = {tr['accuracy']:.4f}, Time = {tr['time']:.2f}s")
                print(f"  Validation: Loss = {va['loss']:.4f}, PPL = {va['perplexity']:.2f}, Acc = {va['accuracy']:.4f}", flush=True)
            self.save_checkpoint(epoch, {"train": tr, "val": va})
        return {"best_val_perplexity": best, **self.metrics}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--device", default="cuda", choices=["cpu", "cuda"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--out", default=None, help="write the final metrics as JSON here")
    for f, v in asdict(GPT2TrainingConfig()).items():
        if isinstance(v, bool):
            ap.add_argument("--" + f.replace("_", "-"), type=lambda s: s.lower() in ("1", "true", "yes"), default=v)
        else:
            ap.add_argument("--" + f.replace("_", "-"), type=type(v), default=v)
    a = ap.parse_args(argv)
    cfg = GPT2TrainingConfig(**{f: getattr(a, f) for f in asdict(GPT2TrainingConfig())})
    trainer = GPT2Trainer(cfg, device=a.device, precision=a.precision if a.device != "cpu" else "fp32")
    result = trainer.train()
    if a.out and trainer.rank == 0:
        json.dump(result, open(a.out, "w"), indent=2)
    if trainer.dist.is_initialized():
        trainer.dist.destroy_process_group()
    return result


if __name__ == "__main__":
    main()
