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


# ----------------------------------------------------------------------------------------------------------------------
# gated feed-forward blocks (nn/modern_activations.py) and differential attention (nn/differential_attention.py)
DIFF_CASES = {"diff": dict(d_model=64, n_heads=4), "diff_bias_keymask": dict(d_model=48, n_heads=6, bias=True, lambda_init=0.8),
              "diff_full_mask": dict(d_model=32, n_heads=2, lambda_init=0.3)}


def load_swiglu(device=None):
    from nfb200.nn import SwiGLU
    ff = SwiGLU(24, 40)
    for n in ("W_gate", "W_up", "W_down", "b_gate", "b_up", "b_down"):
        getattr(ff, n).data = G[f"swiglu_{n}"].copy()
    return ff.to(device) if device else ff


def load_diff(tag, device=None):
    from nfb200.nn import DifferentialAttention
    m = DifferentialAttention(**DIFF_CASES[tag])
    for n in ("lambda_param", "W_q1", "W_k1", "W_v1", "W_q2", "W_k2", "W_v2", "W_o", "b_q1", "b_k1", "b_v1", "b_q2", "b_k2", "b_v2", "b_o"):
        if f"{tag}_{n}" in G:
            getattr(m, n).data = G[f"{tag}_{n}"].copy()
    if device:
        m.to(device)
    return m, G[f"{tag}_x"], (G[f"{tag}_mask"] if f"{tag}_mask" in G else None), G[f"{tag}_out"]


def test_swiglu_forward_and_gradients_match_reference_host():
    from nfb200.nn import GeGLU, SwiGLU, Swish
    np.random.seed(21)
    fresh = SwiGLU(24, 40)
    np.testing.assert_array_equal(np.asarray(fresh.W_up.data), G["swiglu_W_up"])          # init order: W_gate, W_up, W_down
    assert SwiGLU(30).hidden_dim == int(G["swiglu_default_hidden"])
    ff = load_swiglu()
    xt = Tensor(G["swiglu_x"], requires_grad=True)
    out = ff(xt)
    np.testing.assert_allclose(np.asarray(out.data), G["swiglu_out"], rtol=1e-5, atol=2e-6)
    out.backward(G["swiglu_up"])
    np.testing.assert_allclose(np.asarray(xt.grad), G["swiglu_dx"], rtol=1e-4, atol=2e-6)
    for n in ("W_gate", "W_up", "W_down", "b_gate", "b_up", "b_down"):
        np.testing.assert_allclose(np.asarray(getattr(ff, n).grad), G[f"swiglu_d{n}"], rtol=1e-4, atol=2e-6, err_msg=n)
    np.random.seed(22)
    ge = GeGLU(24, 40, bias=False)
    np.testing.assert_array_equal(np.asarray(ge.W_down.data), G["geglu_W_down"])
    np.testing.assert_allclose(np.asarray(ge(Tensor(G["swiglu_x"])).data), G["geglu_out"], rtol=1e-5, atol=2e-6)
    xt = Tensor(G["swiglu_x"], requires_grad=True)
    sw = Swish()(xt)
    np.testing.assert_allclose(np.asarray(sw.data), G["swish_out"], rtol=1e-6, atol=1e-7)
    sw.backward(G["swiglu_up"])
    np.testing.assert_allclose(np.asarray(xt.grad), G["swish_dx"], rtol=1e-5, atol=1e-7)
    with pytest.raises(LayerError):
        SwiGLU(0)
    with pytest.raises(LayerError):
        ff(Tensor(np.zeros((2, 23), dtype=np.float32)))


@pytest.mark.parametrize("tag", list(DIFF_CASES))
def test_differential_attention_forward_matches_reference_host(tag):
    m, x, mask, want = load_diff(tag)
    out, (a1, a2) = m(Tensor(x), None if mask is None else Tensor(mask), return_attention_weights=True)
    np.testing.assert_allclose(np.asarray(out.data), want, rtol=1e-5, atol=3e-6)
    np.testing.assert_allclose(a1, G[f"{tag}_a1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(a2, G[f"{tag}_a2"], rtol=1e-5, atol=1e-6)
    st = m.get_attention_stats()
    assert abs(st["lambda_mean"] - float(np.mean(G[f"{tag}_lambda_param"]))) < 1e-6


def test_differential_block_and_model_match_reference_host():
    """No weights stored for these two: they also pin the RNG call order of every sub-module constructor."""
    from nfb200.nn import DifferentialAttention, DifferentialTransformer, DifferentialTransformerBlock
    np.random.seed(24)
    blk = DifferentialTransformerBlock(d_model=32, n_heads=4, d_ff=48)
    np.testing.assert_allclose(np.asarray(blk(Tensor(G["diffblock_x"])).data), G["diffblock_out"], rtol=1e-5, atol=3e-6)
    np.random.seed(25)
    net = DifferentialTransformer(n_layers=2, d_model=32, n_heads=4, d_ff=48, vocab_size=50, max_seq_len=16)
    np.testing.assert_allclose(np.asarray(net(Tensor(G["diffnet_ids"])).data), G["diffnet_logits"], rtol=1e-5, atol=3e-6)
    assert set(net.get_all_lambda_stats()) == {"layer_0", "layer_1", "global"}
    names = [n for n, _ in net.named_parameters()]
    assert "block_0.attention.lambda_param" in names and "block_1.ffn.W_down" in names and "output_proj.weight" in names
    with pytest.raises(LayerError):
        DifferentialAttention(d_model=30, n_heads=4)
    with pytest.raises(LayerError):
        DifferentialAttention(d_model=30, n_heads=3)


def test_differential_attention_true_gradients_host():
    """Finite differences on x, lambda and one projection of each group (the reference's backward is a placeholder)."""
    from nfb200.nn import DifferentialAttention
    rng = np.random.default_rng(4)
    np.random.seed(5)
    m = DifferentialAttention(d_model=16, n_heads=4, bias=True, lambda_init=0.4)
    x = rng.standard_normal((2, 5, 16)).astype(np.float32)
    up = (rng.standard_normal((2, 5, 16)) * 1e-2).astype(np.float32)
    xt = Tensor(x, requires_grad=True)
    m(xt).backward(up)

    def loss():
        return float((np.asarray(m(Tensor(x)).data, dtype=np.float64) * up).sum())
    eps = 1e-2
    for p, idxs in ((m.lambda_param, [(0,), (1,)]), (m.W_k2, [(0, 0), (9, 5)]), (m.W_q1, [(3, 2)]), (m.W_v1, [(7, 7)]), (m.b_q2, [(4,)])):
        base = np.asarray(p.data).copy()
        for idx in idxs:
            hi, lo = base.copy(), base.copy()
            hi[idx] += eps
            lo[idx] -= eps
            p.data = hi
            fp = loss()
            p.data = lo
            fm = loss()
            p.data = base
            fd = (fp - fm) / (2 * eps)
            assert abs(np.asarray(p.grad)[idx] - fd) <= 3e-2 * abs(fd) + 2e-5, (p.name, idx, np.asarray(p.grad)[idx], fd)


def test_transformer_encoder_stack_host():
    """nn/transformer.py TransformerEncoder: seeded forward equals the (unfused) composition of its own blocks, and parameters()
    uses the reference's layer{i}_{attn,ffn1,ffn2,norm1,norm2}_* / final_norm_* names (checked against the reference when
    generating: oracle/gen_golden.py is not needed for a pure composition)."""
    from nfb200.nn import SelfAttention, TransformerEncoder
    np.random.seed(3)
    enc = TransformerEncoder(16, num_layers=2, num_heads=4, d_ff=32)
    x = Tensor(np.random.default_rng(0).standard_normal((2, 5, 16)).astype(np.float32), requires_grad=True)
    y = enc(x)
    h = x
    for blk in enc.layers:
        h = blk(h)
    np.testing.assert_allclose(np.asarray(y.data), np.asarray(enc.norm(h).data), rtol=1e-6)
    names = list(enc.parameters().keys())
    assert names[0] == "layer0_attn_query_proj.weight" and "layer1_ffn2_bias" in names and names[-1].startswith("final_norm_") and len(names) == 34
    y.sum().backward()
    assert all(p.grad is not None for p in enc.parameters())
    assert SelfAttention(8)(x) is x
