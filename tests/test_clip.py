"""CLIP against a seeded forward of the unmodified reference (tests/golden/clip.npz, made by
oracle/gen_golden.py::gen_clip): host path here (always), device fp32 / bf16 paths under -m gpu."""
import os

import numpy as np
import pytest

import nfb200  # noqa: F401
from nfb200.core import Tensor
from nfb200.models import CLIP

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "clip.npz"))
TINY = dict(embed_dim=24, image_resolution=16, vision_layers=2, vision_width=64, vision_patch_size=8, context_length=9, vocab_size=50,
            transformer_width=32, transformer_heads=4, transformer_layers=2)


def _model(device=None):
    np.random.seed(41)
    m = CLIP(**TINY)
    return m.to(device) if device else m


def _check(out, get, tol):
    for k in ("image_embeds", "text_embeds", "logits_per_image"):
        np.testing.assert_allclose(get(out[k]).reshape(G[k].shape), G[k], rtol=tol, atol=tol, err_msg=k)
    np.testing.assert_allclose(float(get(out["loss"]).reshape(-1)[0]), float(G["loss"].reshape(-1)[0]), rtol=tol)


def test_clip_forward_matches_reference_host():
    m = _model()
    np.testing.assert_array_equal(np.asarray(m.transformer.positional_embedding.data)[:2], G["pos_emb_sample"])   # RNG call order
    out = m(Tensor(G["img"]), Tensor(G["txt"]))
    _check(out, lambda t: np.asarray(t.data, dtype=np.float64), 3e-6)
    np.testing.assert_allclose(np.asarray(out["logits_per_text"].data), np.asarray(out["logits_per_image"].data).T)
    only_img = m(image=Tensor(G["img"]))
    assert set(only_img) == {"image_embeds"}
    np.testing.assert_allclose(np.linalg.norm(np.asarray(only_img["image_embeds"].data), axis=-1), 1.0, rtol=1e-6)
    names = [n for n, _ in m.named_parameters()]
    assert {"logit_scale", "visual.proj.weight", "transformer.positional_embedding", "transformer.transformer.1.attn.qkv_proj.weight",
            "transformer.text_projection.weight", "visual.transformer.cls_token"} <= set(names)


def test_clip_is_one_autograd_graph_host():
    """The reference severs the graph after every step; here the contrastive loss reaches both towers and the temperature.
    Central differences on the log-temperature and on one text-projection weight."""
    m = _model()
    img, txt = Tensor(G["img"]), Tensor(G["txt"])
    m(img, txt)["loss"].backward()
    assert all(p.grad is not None for _, p in m.named_parameters())

    def loss():
        return float(np.asarray(m(img, txt)["loss"].data, dtype=np.float64).reshape(-1)[0])
    for p, idx, eps in ((m.logit_scale, (0,), 1e-2), (m.transformer.text_projection.weight, (3, 5), 1e-2), (m.visual.proj.bias, (2,), 1e-2)):
        base = np.asarray(p.data).copy()
        hi, lo = base.copy(), base.copy()
        hi[idx] += eps
        lo[idx] -= eps
        p.data = hi
        fp = loss()
        p.data = lo
        fm = loss()
        p.data = base
        fd = (fp - fm) / (2 * eps)
        assert abs(np.asarray(p.grad)[idx] - fd) <= 3e-2 * abs(fd) + 2e-5, (idx, np.asarray(p.grad)[idx], fd)


@pytest.mark.gpu
def test_clip_device_fp32_matches_reference(be):
    m = _model("cuda")
    out = m(Tensor(G["img"], device="cuda"), Tensor(G["txt"], device="cuda"))
    _check(out, lambda t: t.data.get().astype(np.float64), 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_clip_real_head_dim_train_steps_device_vs_host(be, precision):
    """Both towers at head_dim 64 (flash kernels: plain for the image tower, causal for the text tower): three AdamW steps on the
    device against the same steps on the host path."""
    from nfb200 import kernels as K
    from nfb200.core import set_default_device
    from nfb200.optim import AdamW
    rng = np.random.default_rng(8)
    img = rng.standard_normal((4, 3, 32, 32)).astype(np.float32)
    txt = rng.integers(0, 80, (4, 12))
    cfg = dict(embed_dim=32, image_resolution=32, vision_layers=2, vision_width=128, vision_patch_size=8, context_length=12, vocab_size=80,
               transformer_width=128, transformer_heads=2, transformer_layers=2, temperature_init=0.5)   # mild temperature: logits = 2 cos
    res = {}
    for dev in ("cpu", "cuda"):
        set_default_device(dev)
        K.set_precision(precision if dev == "cuda" else "fp32")
        try:
            np.random.seed(6)
            m = CLIP(**cfg)
            opt = AdamW(m.parameters(), lr=1e-4, moment_dtype="float64")   # small steps: a chaotic trajectory would amplify rounding
            losses = []
            for _ in range(3):
                opt.zero_grad()
                l = m(Tensor(img, device=dev), Tensor(txt, device=dev))["loss"]
                l.backward()
                opt.step()
                losses.append(float(np.asarray(l.data.get() if dev == "cuda" else l.data).reshape(-1)[0]))
            res[dev] = losses
        finally:
            K.set_precision("fp32")
            set_default_device("cpu")
    np.testing.assert_allclose(res["cuda"], res["cpu"], rtol=1e-4 if precision == "fp32" else 1e-2, atol=1e-3 if precision == "fp32" else 5e-3)
    assert res["cpu"][-1] < res["cpu"][0]
