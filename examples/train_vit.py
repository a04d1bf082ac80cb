"""Vision Transformer classification training on the b200 backend: BASELINE.json's ViT workload (synthetic image batches,
AdamW lr 1e-3 / weight decay 0.1) with the config fields of the reference's examples/training/vit_training.py:33-69 and
the loop of that script.  Images are synthetic: a class-dependent low-frequency pattern plus N(0, 1) noise, so there is
something to learn; labels uniform over the classes.

    python examples/train_vit.py --device cuda --image-size 224 --patch-size 16 --num-classes 1000 --d-model 768 --depth 12 --num-heads 12 --batch-size 64
    python examples/train_vit.py --device cpu --image-size 16 --patch-size 4 --d-model 32 --depth 2 --num-heads 2 ...   # host path (tests)
"""
import os
import sys
from dataclasses import dataclass

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import Trainer, run  # noqa: E402


@dataclass
class ViTTrainingConfig:
    """Field names and defaults of examples/training/vit_training.py:33-69 (the augmentation switches, which act on the
    host-side data generator upstream, are accepted and unused: the batches here are synthetic)."""
    image_size: int = 32
    patch_size: int = 4
    num_classes: int = 10
    d_model: int = 384
    depth: int = 6
    num_heads: int = 6
    mlp_ratio: float = 4.0
    dropout: float = 0.1
    attention_dropout: float = 0.1
    batch_size: int = 16
    learning_rate: float = 1e-3
    weight_decay: float = 0.1
    num_epochs: int = 8
    warmup_epochs: int = 2
    max_grad_norm: float = 1.0
    train_size: int = 1500
    val_size: int = 300
    test_size: int = 150
    use_augmentation: bool = True
    horizontal_flip: bool = True
    random_crop: bool = True
    color_jitter: bool = True
    enable_optimizations: bool = True
    save_checkpoints: bool = True
    checkpoint_dir: str = "checkpoints/vit"


def synthetic_images(n, size, classes, seed):
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, classes, n)
    yy, xx = np.meshgrid(np.linspace(0, 1, size, dtype=np.float32), np.linspace(0, 1, size, dtype=np.float32), indexing="ij")
    proto = np.random.default_rng(12345)                      # the class patterns are the same for every split
    fy, fx, ph = proto.uniform(0.5, 3.0, (classes, 3)), proto.uniform(0.5, 3.0, (classes, 3)), proto.uniform(0, 6.28, (classes, 3))
    pattern = np.sin(2 * np.pi * (fy[:, :, None, None] * yy + fx[:, :, None, None] * xx) + ph[:, :, None, None]).astype(np.float32)
    img = pattern[labels] * 1.5 + rng.standard_normal((n, 3, size, size)).astype(np.float32)
    return img.astype(np.float32), labels.astype(np.int64)


class ViTTrainer(Trainer):
    name = "vit"

    def build_model(self):
        from nfb200.models import VisionTransformer
        c = self.config
        return VisionTransformer(img_size=c.image_size, patch_size=c.patch_size, num_classes=c.num_classes, embed_dim=c.d_model, depth=c.depth,
                                 num_heads=c.num_heads, mlp_ratio=c.mlp_ratio)

    def make_data(self, seed):
        c = self.config
        return synthetic_images(c.train_size, c.image_size, c.num_classes, seed), synthetic_images(c.val_size, c.image_size, c.num_classes, seed + 1)

    def forward(self, batch):
        from nfb200.core import Tensor
        from nfb200.functional import cross_entropy_loss
        img, labels = batch
        logits = self.net(Tensor(img))
        return cross_entropy_loss(logits, Tensor(labels)), logits, labels


def main(argv=None):
    return run(ViTTrainer, ViTTrainingConfig, __doc__, argv)


if __name__ == "__main__":
    main()
