"""Device paths of the zoo blocks: grouped-query attention through the flash kernels (K/V heads read in
place, dK/dV summed over each group inside the kernel), the interleaved-pair RoPE kernel, the positional modules and
the GroupedQueryAttention module in fp32 (1e-5 vs the reference forwards) and bf16 (1e-2) modes."""
import numpy as np
import pytest

from oracle import np_ops as O
from test_gpu_attention import _bf16, _close, _to, attn_impl  # noqa: F401  (fixture)
from test_modern_blocks import G, GQA_CASES, load_gqa

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,Hkv,S", [(2, 4, 2, 128), (1, 6, 1, 197), (2, 12, 4, 512), (1, 8, 8, 130), (1, 4, 2, 1024), (2, 3, 1, 70)])
@pytest.mark.parametrize("mode", ["none", "causal", "keymask", "causal+keymask"])
def test_flash_gqa_forward_backward(be, attn_impl, B, H, Hkv, S, mode):
    """Oracle: repeat K/V to H heads, run the reference attention core, sum dK/dV over each group (what np.repeat's
    gradient is)."""
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(B * 1000 + S + H)
    D, n_rep = 64, H // Hkv
    q = _bf16(rng.standard_normal((B, H, S, D)).astype(np.float32))
    k = _bf16(rng.standard_normal((B, Hkv, S, D)).astype(np.float32))
    v = _bf16(rng.standard_normal((B, Hkv, S, D)).astype(np.float32))
    go = _bf16(rng.standard_normal((B, H, S, D)).astype(np.float32))
    scale = 1.0 / np.sqrt(D)
    causal = "causal" in mode
    mask = None
    if "keymask" in mode:
        keep = (rng.random((B, S)) > 0.2).astype(np.float32)
        keep[:, 0] = 1.0
        mask = ((1.0 - keep) * -10000.0)[:, None, None, :].astype(np.float32)
    kr, vr = np.repeat(k, n_rep, axis=1).astype(np.float64), np.repeat(v, n_rep, axis=1).astype(np.float64)
    want, cache = O.attention_fwd(q.astype(np.float64), kr, vr, scale, "causal" if causal else "none", mask)
    wdq, wdk, wdv = O.attention_bwd(go.astype(np.float64), cache)
    wdk = wdk.reshape(B, Hkv, n_rep, S, D).sum(axis=2)
    wdv = wdv.reshape(B, Hkv, n_rep, S, D).sum(axis=2)
    qd, kd, vd = _to(be, q, be.bfloat16), _to(be, k, be.bfloat16), _to(be, v, be.bfloat16)
    md = None if mask is None else _to(be, mask)
    o, lse = Kn.attention_forward(qd, kd, vd, scale, causal=causal, mask=md)
    assert o.shape == (B, H, S, D) and lse.shape == (B, H, S)
    _close(o.get(), want)
    dq, dk, dv = Kn.attention_backward(_to(be, go, be.bfloat16), qd, kd, vd, o, lse, scale, causal=causal, mask=md)
    assert dq.shape == (B, H, S, D) and dk.shape == (B, Hkv, S, D) and dv.shape == (B, Hkv, S, D)
    _close(dv.get(), wdv)
    _close(dq.get(), wdq)
    _close(dk.get(), wdk)


def test_flash_gqa_rejects_bad_groups(be):
    from nfb200 import B200Error
    from nfb200 import kernels as Kn
    q = be.from_numpy(np.zeros((1, 5, 64, 64), np.float32), be.bfloat16)
    k = be.from_numpy(np.zeros((1, 2, 64, 64), np.float32), be.bfloat16)
    with pytest.raises(B200Error):
        Kn.attention_forward(q, k, k, 0.125)


@pytest.mark.parametrize("shape", [(2, 3, 7, 16), (2, 7, 16), (1, 4, 130, 64), (2, 2, 33, 6)])
@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_rope_interleaved_kernel(be, shape, dt):
    from nfb200.core import Tensor
    from nfb200.nn import RotaryPositionalEmbedding
    rng = np.random.default_rng(sum(shape))
    D, S = shape[-1], shape[-2]
    rope = RotaryPositionalEmbedding(D, max_seq_len=256)
    x = rng.standard_normal(shape).astype(np.float32)
    up = rng.standard_normal(shape).astype(np.float32)
    if dt == "bf16":
        x, up = _bf16(x), _bf16(up)
    want = rope.apply_rope(x, S, 3)
    want_g = rope.apply_rope(up, S, 3, inverse=True)
    xd = be.from_numpy(x, be.bfloat16 if dt == "bf16" else None)
    qt = Tensor(xd, requires_grad=True)
    kt = Tensor(xd.copy(), requires_grad=True)
    rq, rk = rope(qt, kt, start_pos=3)
    tol = 1e-2 if dt == "bf16" else 1e-5
    np.testing.assert_allclose(rq.data.get().astype(np.float32), want, rtol=tol, atol=tol)
    np.testing.assert_allclose(rk.data.get().astype(np.float32), want, rtol=tol, atol=tol)
    np.testing.assert_array_equal(xd.get(), be.from_numpy(x, be.bfloat16 if dt == "bf16" else None).get())   # input untouched
    g = rope.apply_rope(be.from_numpy(up, be.bfloat16 if dt == "bf16" else None), S, 3, inverse=True)
    np.testing.assert_allclose(g.get().astype(np.float32), want_g, rtol=tol, atol=tol)
    if dt == "f32" and len(shape) == 4 and shape[-1] == 16:
        r2, _ = RotaryPositionalEmbedding(16, max_seq_len=32)(Tensor(be.from_numpy(G["rope_q"])), Tensor(be.from_numpy(G["rope_k"])), start_pos=5)
        np.testing.assert_allclose(r2.data.get(), G["rope_q_out"], rtol=1e-5, atol=1e-6)


def test_positional_modules_device(be):
    from nfb200.core import Tensor
    from nfb200.nn import LearnedPositionalEmbedding, SinusoidalPositionalEncoding
    x = G["pos_x"]
    out = SinusoidalPositionalEncoding(16, max_len=64)(Tensor(x, device="cuda"), start_pos=3)
    np.testing.assert_allclose(out.data.get(), G["sin_out"], rtol=1e-6, atol=1e-6)
    np.random.seed(9)
    lpe = LearnedPositionalEmbedding(12, 16).to("cuda")
    xt = Tensor(x, device="cuda", requires_grad=True)
    out = lpe(xt, start_pos=2)
    np.testing.assert_allclose(out.data.get(), G["lpe_out"], rtol=1e-6, atol=1e-6)
    out.backward(be.from_numpy(G["lpe_up"]))
    np.testing.assert_allclose(xt.grad.get(), G["lpe_dx"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(lpe.embedding.grad.get(), G["lpe_dtable"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", list(GQA_CASES))
def test_gqa_module_fp32_matches_reference_device(be, tag):
    from nfb200.core import Tensor
    m, x, mask, want = load_gqa(tag, "cuda")
    out = m(Tensor(x, device="cuda"), None if mask is None else Tensor(mask, device="cuda"))
    np.testing.assert_allclose(out.data.get(), want, rtol=2e-5, atol=5e-6)


@pytest.mark.parametrize("n_kv", [1, 2, 8])
@pytest.mark.parametrize("masked", [False, True])
def test_gqa_module_bf16_flash_vs_fp32(be, n_kv, masked):
    """head_dim 64: the bf16 path runs the tcgen05 GEMMs + grouped flash kernels; compare output and every gradient with
    the fp32 composition (explicit repeat_kv) of the same module: max error <= 2e-2 of the tensor's scale, relative L2 <= 2e-2."""
    from nfb200 import kernels as K
    from nfb200.core import Tensor
    from nfb200.nn import GroupedQueryAttention
    rng = np.random.default_rng(5 + n_kv)
    B, S, d, H = 2, 200, 512, 8
    x = rng.standard_normal((B, S, d)).astype(np.float32)
    up = (rng.standard_normal((B, S, d)) * 1e-2).astype(np.float32)
    mask = None
    if masked:
        mask = np.where(rng.random((B, 1, 1, S)) < 0.25, -1e4, 0.0).astype(np.float32)
        mask[..., 0] = 0.0
    res = {}
    try:
        for prec in ("fp32", "bf16"):
            K.set_precision(prec)
            np.random.seed(11)
            m = GroupedQueryAttention(d, H, n_kv, bias=True).to("cuda")
            xt = Tensor(x, device="cuda", requires_grad=True)
            out = m(xt, None if mask is None else Tensor(mask, device="cuda"))
            out.backward(be.from_numpy(up))
            res[prec] = {"out": out.data.get().astype(np.float32), "dx": xt.grad.get().astype(np.float32),
                         **{n: p.grad.get().astype(np.float32) for n, p in m.named_parameters()}}
    finally:
        K.set_precision("fp32")
    assert set(res["bf16"]) >= {"out", "dx", "W_q", "W_k", "W_v", "W_o", "b_q", "b_k", "b_v", "b_o"}
    for name, want in res["fp32"].items():
        got = res["bf16"][name]
        scale = float(np.abs(want).max())
        if name == "b_k":
            # softmax is invariant to a per-head constant added to every key: the true gradient of b_k is exactly 0 (fp32 run:
            # ~1e-7); the bf16 run leaves rounding noise, bounded against the size of the neighbouring bias gradients
            assert scale <= 1e-5
            assert float(np.abs(got).max()) <= 2e-2 * float(np.abs(res["fp32"]["b_q"]).max() + np.abs(res["fp32"]["b_v"]).max())
            continue
        assert float(np.abs(got - want).max()) <= 2e-2 * scale, (name, float(np.abs(got - want).max()), scale)
        l2 = float(np.linalg.norm(got - want) / (np.linalg.norm(want) + 1e-30))
        assert l2 <= 2e-2, (name, l2)   # a chain of four bf16 kernels (each held to 1e-2 in isolation); measured 0.3-1.0e-2


# ----------------------------------------------------------------------------------------------------------------------
def test_swiglu_module_fp32_matches_reference_device(be):
    from nfb200.core import Tensor
    from test_modern_blocks import load_swiglu
    ff = load_swiglu("cuda")
    xt = Tensor(G["swiglu_x"], device="cuda", requires_grad=True)
    out = ff(xt)
    np.testing.assert_allclose(out.data.get(), G["swiglu_out"], rtol=2e-5, atol=5e-6)
    out.backward(be.from_numpy(G["swiglu_up"]))
    np.testing.assert_allclose(xt.grad.get(), G["swiglu_dx"], rtol=1e-4, atol=5e-6)
    for n in ("W_gate", "W_up", "W_down", "b_gate", "b_up", "b_down"):     # gradients reach each parameter THROUGH the concatenation
        np.testing.assert_allclose(getattr(ff, n).grad.get(), G[f"swiglu_d{n}"], rtol=1e-4, atol=5e-6, err_msg=n)


@pytest.mark.parametrize("tag", ["diff", "diff_bias_keymask", "diff_full_mask"])
def test_differential_attention_fp32_matches_reference_device(be, tag):
    from nfb200.core import Tensor
    from test_modern_blocks import load_diff
    m, x, mask, want = load_diff(tag, "cuda")
    out, (a1, a2) = m(Tensor(x, device="cuda"), None if mask is None else Tensor(mask, device="cuda"), return_attention_weights=True)
    np.testing.assert_allclose(out.data.get(), want, rtol=2e-5, atol=5e-6)
    np.testing.assert_allclose(a1, G[f"{tag}_a1"], rtol=1e-4, atol=1e-6)


def test_differential_model_fp32_matches_reference_device(be):
    from nfb200.core import Tensor
    from nfb200.nn import DifferentialTransformer
    np.random.seed(25)
    net = DifferentialTransformer(n_layers=2, d_model=32, n_heads=4, d_ff=48, vocab_size=50, max_seq_len=16).to("cuda")
    out = net(Tensor(G["diffnet_ids"], device="cuda"))
    np.testing.assert_allclose(out.data.get(), G["diffnet_logits"], rtol=5e-5, atol=1e-5)


@pytest.mark.parametrize("overlap", [False, True], ids=["one_stream", "wgrad_side_stream"])
def test_differential_block_bf16_flash_vs_fp32(be, overlap):
    """head_dim 64: one concatenated-projection GEMM -> packed flash attention for both maps -> lambda combination -> SwiGLU FFN
    with its concatenated [up|gate] GEMM; output and EVERY parameter gradient (through the differentiable concatenations, with
    the weight-gradient GEMMs on the side stream in the second variant) against the fp32 composition."""
    from nfb200 import kernels as K
    from nfb200.core import Tensor
    from nfb200.nn import DifferentialTransformerBlock
    rng = np.random.default_rng(17)
    B, S, d, H = 2, 160, 256, 4
    x = rng.standard_normal((B, S, d)).astype(np.float32)
    up = (rng.standard_normal((B, S, d)) * 1e-2).astype(np.float32)
    res = {}
    try:
        for prec in ("fp32", "bf16"):
            K.set_precision(prec)
            K.set_wgrad_overlap(bool(overlap) and prec == "bf16")
            np.random.seed(13)
            blk = DifferentialTransformerBlock(d_model=d, n_heads=H, d_ff=384, lambda_init=0.6).to("cuda")
            xt = Tensor(x, device="cuda", requires_grad=True)
            out = blk(xt)
            out.backward(be.from_numpy(up))
            res[prec] = {"out": out.data.get().astype(np.float32), "dx": xt.grad.get().astype(np.float32),
                         **{n: p.grad.get().astype(np.float32) for n, p in blk.named_parameters()}}
    finally:
        K.set_wgrad_overlap(False)
        K.set_precision("fp32")
    assert {"attention.lambda_param", "attention.W_k2", "ffn.W_gate", "ffn.b_up", "norm1.weight"} <= set(res["bf16"])
    for name, want in res["fp32"].items():
        got = res["bf16"][name]
        scale = float(np.abs(want).max())
        assert scale > 0, name
        assert float(np.abs(got - want).max()) <= 3e-2 * scale, (name, float(np.abs(got - want).max()), scale)
        l2 = float(np.linalg.norm(got - want) / (np.linalg.norm(want) + 1e-30))
        assert l2 <= 2.5e-2, (name, l2)   # whole-block chain of bf16 kernels (each 1e-2 in isolation)
