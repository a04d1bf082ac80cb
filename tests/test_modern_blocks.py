"""Zoo blocks on the hot-path kernels: grouped-query / multi-query attention and the positional modules,
pinned to forwards of the UNMODIFIED reference (tests/golden/modern.npz, oracle/gen_golden.py::gen_modern).  CPU part: the
host path + true gradients by finite differences; the device paths are in tests/test_gpu_modern_blocks.py."""
import os

import numpy as np
import pytest

import nfb200  # noqa: F401
from nfb200 import functional as F
from nfb200.core import Tensor
from nfb200.nn import (GroupedQueryAttention, LayerError, LearnedPositionalEmbedding, MultiQueryAttention, RotaryPositionalEmbedding,
                       SinusoidalPositionalEncoding, create_rope, create_sinusoidal_pe)

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "modern.npz"))
GQA_CASES = {"gqa": dict(d_model=64, n_heads=4, n_kv_heads=2), "gqa_bias_keymask": dict(d_model=96, n_heads=6, n_kv_heads=3, bias=True),
             "gqa_full_mask": dict(d_model=64, n_heads=4, n_kv_heads=1), "mha": dict(d_model=32, n_heads=2)}


def load_gqa(tag, device=None):
    m = GroupedQueryAttention(**GQA_CASES[tag])
    for n in ("W_q", "W_k", "W_v", "W_o", "b_q", "b_k", "b_v", "b_o"):
        if f"{tag}_{n}" in G:
            getattr(m, n).data = G[f"{tag}_{n}"].copy()
    if device:
        m.to(device)
    mask = G[f"{tag}_mask"] if f"{tag}_mask" in G else None
    return m, G[f"{tag}_x"], mask, G[f"{tag}_out"]


@pytest.mark.parametrize("tag", list(GQA_CASES))
def test_gqa_forward_matches_reference_host(tag):
    m, x, mask, want = load_gqa(tag)
    out = m(Tensor(x), None if mask is None else Tensor(mask))
    np.testing.assert_allclose(np.asarray(out.data), want, rtol=1e-5, atol=2e-6)


def test_gqa_init_order_and_api():
    np.random.seed(7)
    m = GroupedQueryAttention(**GQA_CASES["gqa"])
    for n in ("W_q", "W_k", "W_v", "W_o"):                                   # same RNG call order and scaling as the reference
        np.testing.assert_array_equal(np.asarray(getattr(m, n).data), G[f"gqa_{n}"])
    assert m.n_rep == 2 and m.rope is not None and m.b_q is None
    np.testing.assert_array_equal(GroupedQueryAttention(d_model=32, n_heads=8, n_kv_heads=2).repeat_kv(G["repeat_in"]), G["repeat_out"])
    out, cache = m(Tensor(G["gqa_x"]), use_cache=True)
    assert cache[0].shape == (2, 4, 9, 16) and cache[1].shape == (2, 4, 9, 16)
    with pytest.raises(LayerError):
        GroupedQueryAttention(d_model=65, n_heads=4)
    with pytest.raises(LayerError):
        GroupedQueryAttention(d_model=64, n_heads=4, n_kv_heads=3)
    with pytest.raises(LayerError):
        m(Tensor(np.zeros((1, 3, 32), dtype=np.float32)))
    np.random.seed(8)
    mq = MultiQueryAttention(d_model=64, n_heads=4)
    assert mq.gqa.n_kv_heads == 1 and mq.gqa.n_rep == 4
    np.testing.assert_array_equal(np.asarray(mq.gqa.W_k.data), G["mqa_W_k"])
    np.testing.assert_allclose(np.asarray(mq(Tensor(G["mqa_x"])).data), G["mqa_out"], rtol=1e-5, atol=2e-6)


def test_gqa_true_gradients_host():
    """The reference's GQA backward is a placeholder; ours is the real gradient: check against central differences (fp64)."""
    rng = np.random.default_rng(0)
    np.random.seed(3)
    m = GroupedQueryAttention(d_model=16, n_heads=4, n_kv_heads=2, bias=True)
    x = rng.standard_normal((2, 5, 16)).astype(np.float32)
    up = (rng.standard_normal((2, 5, 16)) * 1e-2).astype(np.float32)     # small: the engine clips every node's gradient to +-10
    xt = Tensor(x, requires_grad=True)
    out = m(xt)
    out.backward(up)

    def loss(xv=None, wk=None):
        saved = m.W_k.data
        if wk is not None:
            m.W_k.data = wk
        o = np.asarray(m(Tensor(x if xv is None else xv)).data, dtype=np.float64)
        m.W_k.data = saved
        return float((o * up).sum())
    eps = 1e-2
    for idx in [(0, 0, 0), (1, 3, 7), (0, 4, 15)]:
        xp, xm = x.copy(), x.copy()
        xp[idx] += eps
        xm[idx] -= eps
        fd = (loss(xv=xp) - loss(xv=xm)) / (2 * eps)
        assert abs(np.asarray(xt.grad)[idx] - fd) <= 2e-2 * abs(fd) + 2e-5, (idx, np.asarray(xt.grad)[idx], fd)
    wk = np.asarray(m.W_k.data)
    for idx in [(0, 0), (7, 3), (15, 7)]:
        wp, wm = wk.copy(), wk.copy()
        wp[idx] += eps
        wm[idx] -= eps
        fd = (loss(wk=wp) - loss(wk=wm)) / (2 * eps)
        assert abs(np.asarray(m.W_k.grad)[idx] - fd) <= 2e-2 * abs(fd) + 2e-5, (idx, np.asarray(m.W_k.grad)[idx], fd)


def test_repeat_kv_functional_host():
    rng = np.random.default_rng(1)
    x = Tensor(rng.standard_normal((2, 2, 3, 4)).astype(np.float32), requires_grad=True)
    y = F.repeat_kv(x, 3)
    assert y.shape == (2, 6, 3, 4)
    np.testing.assert_array_equal(np.asarray(y.data)[:, 4], np.asarray(x.data)[:, 1])     # head h <- head h // n_rep
    up = rng.standard_normal(y.shape).astype(np.float32)
    y.backward(up)
    np.testing.assert_allclose(np.asarray(x.grad), np.clip(up, -10, 10).reshape(2, 2, 3, 3, 4).sum(axis=2), rtol=1e-6)
    assert F.repeat_kv(x, 1) is x


def test_positional_forward_matches_reference_host():
    x = G["pos_x"]
    np.testing.assert_allclose(np.asarray(SinusoidalPositionalEncoding(16, max_len=64)(Tensor(x), start_pos=3).data), G["sin_out"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(np.asarray(SinusoidalPositionalEncoding(16, max_len=32, base=500.0)(Tensor(x)).data), G["sin_out_base"], rtol=1e-6, atol=1e-7)
    rope = RotaryPositionalEmbedding(16, max_seq_len=32)
    rq, rk = rope(Tensor(G["rope_q"]), Tensor(G["rope_k"]), start_pos=5)
    np.testing.assert_allclose(np.asarray(rq.data), G["rope_q_out"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(np.asarray(rk.data), G["rope_k_out"], rtol=1e-6, atol=1e-7)
    r3, _ = RotaryPositionalEmbedding(16, max_seq_len=8, base=100.0)(Tensor(x), Tensor(x), start_pos=4)
    np.testing.assert_allclose(np.asarray(r3.data), G["rope3_out"], rtol=1e-6, atol=1e-7)
    np.random.seed(9)
    lpe = LearnedPositionalEmbedding(12, 16)
    np.testing.assert_array_equal(np.asarray(lpe.embedding.data), G["lpe_table"])
    xt = Tensor(x, requires_grad=True)
    out = lpe(xt, start_pos=2)
    np.testing.assert_allclose(np.asarray(out.data), G["lpe_out"], rtol=1e-6, atol=1e-7)
    out.backward(G["lpe_up"])
    np.testing.assert_allclose(np.asarray(xt.grad), G["lpe_dx"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(np.asarray(lpe.embedding.grad), G["lpe_dtable"], rtol=1e-6, atol=1e-7)
    assert isinstance(create_rope(8), RotaryPositionalEmbedding) and isinstance(create_sinusoidal_pe(8), SinusoidalPositionalEncoding)


def test_rope_gradient_is_the_transpose_rotation_host():
    rng = np.random.default_rng(2)
    rope = RotaryPositionalEmbedding(8, max_seq_len=16)
    q = Tensor(rng.standard_normal((2, 3, 5, 8)).astype(np.float32), requires_grad=True)
    k = Tensor(rng.standard_normal((2, 3, 5, 8)).astype(np.float32), requires_grad=True)
    rq, rk = rope(q, k, start_pos=1)
    up = rng.standard_normal(rq.shape).astype(np.float32)
    rq.backward(up)
    # <R q, up> = <q, R^T up>: the gradient must be the inverse rotation of the upstream gradient, and rotations keep norms
    np.testing.assert_allclose(np.asarray(rope.apply_rope(np.asarray(q.grad), 5, 1)), np.clip(up, -10, 10), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(np.asarray(rq.data)), np.linalg.norm(np.asarray(q.data)), rtol=1e-6)


def test_positional_validation_errors():
    for make in (lambda: SinusoidalPositionalEncoding(15), lambda: SinusoidalPositionalEncoding(16, max_len=0),
                 lambda: RotaryPositionalEmbedding(7), lambda: RotaryPositionalEmbedding(8, max_seq_len=0),
                 lambda: LearnedPositionalEmbedding(0, 8), lambda: LearnedPositionalEmbedding(8, 0)):
        with pytest.raises(LayerError):
            make()
    x = Tensor(np.zeros((1, 5, 8), dtype=np.float32))
    with pytest.raises(LayerError):
        SinusoidalPositionalEncoding(8, max_len=4)(x)
    with pytest.raises(LayerError):
        SinusoidalPositionalEncoding(16)(x)
    with pytest.raises(LayerError):
        LearnedPositionalEmbedding(4, 8)(x)
    with pytest.raises(LayerError):
        RotaryPositionalEmbedding(8)(x, Tensor(np.zeros((1, 4, 8), dtype=np.float32)))
    with pytest.raises(LayerError):
        RotaryPositionalEmbedding(16)(x, x)
