"""Shared training loop of the example scripts: the loop order of the reference's trainers (examples/training/
gpt2_training.py:417-486, bert_training.py:308-360, vit_training.py): forward -> loss -> backward -> gradient clipping at
`max_grad_norm` -> AdamW step -> zero_grad -> metrics; validation and one checkpoint record per epoch.  Device-agnostic:
`--device cpu` runs the host path, `--device cuda` the b200 backend (bf16 tensor-core mode by default); under
`python -m nfb200.distributed.launcher` (or torchrun) the model is wrapped in DistributedDataParallel and every rank
draws its shard with DistributedSampler."""
import argparse
import json
import os
import sys
import time
from dataclasses import asdict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "neural-network-from-scratch_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


class Trainer:
    """Subclasses provide `name`, `build_model()`, `make_data(seed) -> (train, val)` (tuples of aligned arrays), and
    `forward(batch) -> (loss Tensor, logits Tensor, targets ndarray)`."""
    name = "model"

    def __init__(self, config, device: str = "cpu", precision: str = "fp32", seed: int = 0):
        from nfb200 import distributed as dist
        from nfb200 import kernels as K
        from nfb200.core import set_default_device
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
        self.model = self.build_model()
        self.net = dist.DistributedDataParallel(self.model) if self.world > 1 else self.model
        self.optimizer = AdamW(self.model.parameters(), lr=config.learning_rate, weight_decay=config.weight_decay)
        self.train_data, self.val_data = self.make_data(seed + 1)
        self.metrics = {"train_losses": [], "val_losses": [], "val_accuracies": []}

    # ------------------------------------------------------------------ to be provided
    def build_model(self):
        raise NotImplementedError

    def make_data(self, seed):
        raise NotImplementedError

    def forward(self, batch):
        raise NotImplementedError

    # ------------------------------------------------------------------ one step
    def _metrics(self, loss, logits, targets):
        lv = float(loss.item())
        pred = np.asarray(logits.numpy())
        pred = pred.reshape(-1, pred.shape[-1]).argmax(1)
        return {"loss": lv, "perplexity": float(np.exp(min(lv, 50.0))), "accuracy": float(np.mean(pred == np.asarray(targets).reshape(-1)))}

    def train_step(self, batch):
        from nfb200.optim import clip_grad_norm
        loss, logits, targets = self.forward(batch)
        loss.backward()
        if self.net is not self.model:
            self.net.sync_gradients()
        clip_grad_norm(self.model.parameters(), self.config.max_grad_norm, optimizer=self.optimizer)
        self.optimizer.step()
        self.optimizer.zero_grad()
        self.step += 1
        return loss, logits, targets

    # ------------------------------------------------------------------ epochs
    def _shard(self, n, epoch, shuffle):
        if self.world > 1:
            s = self.dist.DistributedSampler(n, num_replicas=self.world, rank=self.rank, shuffle=shuffle, seed=0)
            s.set_epoch(epoch)
            return np.fromiter(iter(s), dtype=np.int64)
        return np.random.RandomState(epoch).permutation(n) if shuffle else np.arange(n)

    def _run(self, data, order, train):
        bs = self.config.batch_size
        tot = {"loss": 0.0, "perplexity": 0.0, "accuracy": 0.0}
        nb, t0 = 0, time.time()
        for lo in range(0, len(order) - bs + 1, bs):
            batch = tuple(np.ascontiguousarray(a[order[lo:lo + bs]]) for a in data)
            loss, logits, targets = self.train_step(batch) if train else self.forward(batch)
            m = self._metrics(loss, logits, targets)
            for k in tot:
                tot[k] += m[k]
            nb += 1
            if train and self.rank == 0 and nb % 10 == 1:
                print(f"  Batch {nb}: Loss = {m['loss']:.4f}, Acc = {m['accuracy']:.4f}", flush=True)
        out = {k: v / max(nb, 1) for k, v in tot.items()}
        out["time"] = time.time() - t0
        return out

    def train_epoch(self, epoch: int):
        return self._run(self.train_data, self._shard(len(self.train_data[0]), epoch, True), True)

    def validate(self):
        from nfb200.core import no_grad
        with no_grad():
            return self._run(self.val_data, np.arange(len(self.val_data[0])), False)

    def save_checkpoint(self, epoch: int, metrics):
        """The reference's per-epoch record (gpt2_training.py:528-547) plus weights / optimizer state (nfb200/checkpoint.py)."""
        if not self.config.save_checkpoints:
            return None
        from nfb200.checkpoint import save_checkpoint
        return save_checkpoint(os.path.join(self.config.checkpoint_dir, f"{self.name}_epoch_{epoch + 1}"), self.model, self.optimizer,
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
            self.metrics["val_accuracies"].append(va["accuracy"])
            best = min(best, va["loss"])
            if self.rank == 0:
                print(f"  Training: Loss = {tr['loss']:.4f}, Acc = {tr['accuracy']:.4f}, Time = {tr['time']:.2f}s")
                print(f"  Validation: Loss = {va['loss']:.4f}, Acc = {va['accuracy']:.4f}", flush=True)
            self.save_checkpoint(epoch, {"train": tr, "val": va})
        return {"best_val_loss": best, "best_val_perplexity": float(np.exp(min(best, 50.0))), **self.metrics}


def run(trainer_cls, config_cls, doc, argv=None):
    """Command line shared by the scripts: one flag per config field plus --device / --precision / --out."""
    ap = argparse.ArgumentParser(description=doc.split("\n")[0])
    ap.add_argument("--device", default="cuda", choices=["cpu", "cuda"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--out", default=None, help="write the final metrics as JSON here")
    defaults = asdict(config_cls())
    for f, v in defaults.items():
        if isinstance(v, bool):
            ap.add_argument("--" + f.replace("_", "-"), type=lambda s: s.lower() in ("1", "true", "yes"), default=v)
        else:
            ap.add_argument("--" + f.replace("_", "-"), type=type(v), default=v)
    a = ap.parse_args(argv)
    cfg = config_cls(**{f: getattr(a, f) for f in defaults})
    trainer = trainer_cls(cfg, device=a.device, precision=a.precision if a.device != "cpu" else "fp32")
    result = trainer.train()
    if a.out and trainer.rank == 0:
        json.dump(result, open(a.out, "w"), indent=2)
    if trainer.dist.is_initialized():
        trainer.dist.destroy_process_group()
    return result
