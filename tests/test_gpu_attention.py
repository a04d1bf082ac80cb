"""GPU parity of RoPE and the fused flash attention forward/backward against the oracle
restatement of the reference's attention cores (GPT-2 causal -1e4 replace, BERT additive key mask, ViT plain).
bf16 tensor-core path: stated tolerance 1e-2 relative/absolute on activations and gradients (north star)."""
import numpy as np
import pytest

from oracle import np_ops as O

pytestmark = pytest.mark.gpu


def _bf16(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(a))


def _to(be, a, dtype=None):
    return be.from_numpy(np.asarray(a), dtype)


def _close(got, want, tol=1e-2):
    scale = max(1e-6, float(np.abs(want).max()))
    err = float(np.abs(got - want).max())
    assert err <= tol * scale + 1e-3 * tol, f"max err {err} vs scale {scale}"


@pytest.fixture(params=[3, 1, 0], ids=["tcgen05", "tcgen05_fwd_only", "mma_sync"])
def attn_impl(request):
    """3 = tcgen05/TMEM/TMA forward and backward (csrc/attention_tc.cu, the default), 1 = tcgen05 forward with the
    mma.sync backward, 0 = the mma.sync flash kernels only."""
    from nfb200 import _lib
    _lib.call("nfb_flash_attn_set_tc", request.param)
    yield request.param
    _lib.call("nfb_flash_attn_set_tc", 3)


@pytest.mark.parametrize("B,H,S", [(2, 3, 128), (1, 2, 197), (2, 12, 512), (1, 1, 64), (1, 2, 70), (1, 2, 1024), (2, 2, 300)])
@pytest.mark.parametrize("mode", ["none", "causal", "keymask", "causal+keymask"])
def test_flash_forward_backward(be, attn_impl, B, H, S, mode):
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(B * 1000 + S)
    D = 64
    q = _bf16(rng.standard_normal((B, H, S, D)).astype(np.float32))
    k = _bf16(rng.standard_normal((B, H, S, D)).astype(np.float32))
    v = _bf16(rng.standard_normal((B, H, S, D)).astype(np.float32))
    go = _bf16(rng.standard_normal((B, H, S, D)).astype(np.float32))
    scale = 1.0 / np.sqrt(D)
    causal = "causal" in mode
    mask = None
    if "keymask" in mode:
        keep = (rng.random((B, S)) > 0.2).astype(np.float32)
        keep[:, 0] = 1.0
        mask = ((1.0 - keep) * -10000.0)[:, None, None, :].astype(np.float32)  # bert.py:456-457
    want, cache = O.attention_fwd(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64), scale, "causal" if causal else "none", mask)
    wdq, wdk, wdv = O.attention_bwd(go.astype(np.float64), cache)
    dq_, dk_, dv_ = _to(be, q, be.bfloat16), _to(be, k, be.bfloat16), _to(be, v, be.bfloat16)
    o, lse = Kn.attention_forward(dq_, dk_, dv_, scale, causal=causal, mask=None if mask is None else _to(be, mask))
    assert o.shape == (B, H, S, D)
    _close(o.get(), want)
    # LSE = log sum exp of the masked, scaled scores
    s = np.matmul(q.astype(np.float64), np.swapaxes(k.astype(np.float64), -1, -2)) * scale
    if causal:
        s = np.where(np.tril(np.ones((S, S))) == 0, -1e4, s)
    if mask is not None:
        s = s + mask
    want_lse = np.log(np.exp(s - s.max(-1, keepdims=True)).sum(-1)) + s.max(-1)
    np.testing.assert_allclose(lse.get(), want_lse, rtol=1e-3, atol=2e-3)
    dq, dk, dv = Kn.attention_backward(_to(be, go, be.bfloat16), dq_, dk_, dv_, o, lse, scale, causal=causal,
                                       mask=None if mask is None else _to(be, mask))
    _close(dv.get(), wdv)
    _close(dq.get(), wdq)
    _close(dk.get(), wdk)


def test_flash_packed_qkv_layout_and_rope(be, attn_impl):
    """The fused qkv projection output (B,S,3,H,D) is consumed in place; RoPE rotates q,k in place; the packed
    dqkv gradient buffer comes back ready for the dgrad GEMM (gpt2.py:228-262)."""
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(7)
    B, S, H, D = 2, 256, 4, 64
    qkv = _bf16(rng.standard_normal((B, S, 3, H, D)).astype(np.float32))
    go = _bf16(rng.standard_normal((B, S, H * D)).astype(np.float32))
    cos, sin = O.rope_tables(D, S)
    q = qkv[:, :, 0].transpose(0, 2, 1, 3)
    k = qkv[:, :, 1].transpose(0, 2, 1, 3)
    v = qkv[:, :, 2].transpose(0, 2, 1, 3)
    qr, kr = O.rope_fwd(q, cos, sin), O.rope_fwd(k, cos, sin)
    scale = 1.0 / np.sqrt(D)
    want, cache = O.attention_fwd(_bf16(qr).astype(np.float64), _bf16(kr).astype(np.float64), v.astype(np.float64), scale, "causal")
    gobh = go.reshape(B, S, H, D).transpose(0, 2, 1, 3)
    wdq, wdk, wdv = O.attention_bwd(gobh.astype(np.float64), cache)
    wdq, wdk = O.rope_bwd(wdq, cos, sin), O.rope_bwd(wdk, cos, sin)

    dev = _to(be, qkv, be.bfloat16)
    Kn.rope_packed_qk_inplace(dev, H)
    got_rot = dev.get()
    np.testing.assert_allclose(got_rot[:, :, 0].transpose(0, 2, 1, 3), qr, rtol=1e-2, atol=1e-2)
    assert np.array_equal(got_rot[:, :, 2], qkv[:, :, 2])  # v untouched
    qd = dev[:, :, 0].transpose(0, 2, 1, 3)
    kd = dev[:, :, 1].transpose(0, 2, 1, 3)
    vd = dev[:, :, 2].transpose(0, 2, 1, 3)
    o, lse = Kn.attention_forward(qd, kd, vd, scale, causal=True)
    o_flat = o.transpose(0, 2, 1, 3).reshape(B, S, H * D)  # free view: stored (B,S,H,D)
    assert o_flat.is_contiguous
    _close(o_flat.get(), want.transpose(0, 2, 1, 3).reshape(B, S, H * D))
    dgo = _to(be, go, be.bfloat16).reshape(B, S, H, D).transpose(0, 2, 1, 3)
    dq, dk, dv, packed = Kn.attention_backward(dgo, qd, kd, vd, o, lse, scale, causal=True, packed=True)
    assert packed.shape == (B, S, 3, H, D) and packed.is_contiguous
    Kn.rope_packed_qk_inplace(packed, H, inverse=True)
    got = packed.get()
    _close(got[:, :, 0].transpose(0, 2, 1, 3), wdq)
    _close(got[:, :, 1].transpose(0, 2, 1, 3), wdk)
    _close(got[:, :, 2].transpose(0, 2, 1, 3), wdv)


def test_rope_fp32_exact_tables(be):
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(8)
    x = rng.standard_normal((2, 3, 100, 64)).astype(np.float32)
    cos, sin = O.rope_tables(64, 100)
    d = _to(be, x.copy())
    Kn.rope_inplace(d)
    np.testing.assert_allclose(d.get(), O.rope_fwd(x, cos, sin), rtol=1e-6, atol=1e-6)
    Kn.rope_inplace(d, inverse=True)
    np.testing.assert_allclose(d.get(), x, rtol=1e-5, atol=1e-5)  # rotation is orthogonal


def test_backend_flash_attention_entry_point_and_unfused_fp32(be):
    """CuPy-precedent entry point backend.flash_attention(q,k,v,scale,block_size) (cuda_backend.py:354-375) and
    the fp32 unfused path used for the 1e-5 contract."""
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(9)
    q, k, v = (rng.standard_normal((2, 2, 96, 64)).astype(np.float32) for _ in range(3))
    want, _ = O.attention_fwd(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64), 0.125)
    got = be.flash_attention(_to(be, q), _to(be, k), _to(be, v), 0.125, 64)
    _close(got.get(), want, 2e-2)
    o32, _ = Kn.attention_reference_unfused(_to(be, q), _to(be, k), _to(be, v), 0.125, causal=True)
    want_c, _ = O.attention_fwd(q, k, v, np.float32(0.125), "causal")
    np.testing.assert_allclose(o32.get(), want_c, rtol=1e-5, atol=2e-6)


def test_linear_forward_backward_both_precisions(be):
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(10)
    x = rng.standard_normal((4, 50, 256)).astype(np.float32)
    W = (rng.standard_normal((384, 256)) * 0.05).astype(np.float32)
    b = rng.standard_normal(384).astype(np.float32)
    g = rng.standard_normal((4, 50, 384)).astype(np.float32)
    for act in (None, "gelu", "relu"):
        want, cache = O.linear_fwd(x.astype(np.float64), W.astype(np.float64), b.astype(np.float64), act)
        wdx, wdW, wdb = O.linear_bwd(g.astype(np.float64), cache)
        y, ctx = Kn.linear_forward(_to(be, x), _to(be, W), _to(be, b), activation=act, precision="fp32")
        np.testing.assert_allclose(y.get(), want, rtol=1e-5, atol=1e-5)
        dx, dW, db = Kn.linear_backward(_to(be, g), ctx)
        np.testing.assert_allclose(dx.get(), wdx, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(dW.get(), wdW, rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(db.get(), wdb, rtol=1e-5, atol=1e-4)
        y, ctx = Kn.linear_forward(_to(be, x), _to(be, W), _to(be, b), activation=act, precision="bf16")
        _close(y.get(), want, 2e-2)
        dx, dW, db = Kn.linear_backward(_to(be, g), ctx)
        if act == "relu":
            # the ReLU mask is discontinuous: pre-activations within bf16 rounding of 0 may flip, so bound the
            # MEAN error instead of the max (a handful of elements differ by O(|g W|))
            for got, w in ((dx, wdx), (dW, wdW), (db, wdb)):
                assert np.abs(got.get() - w).mean() <= 2e-2 * np.abs(w).max()
            continue
        _close(dx.get(), wdx, 2e-2)
        _close(dW.get(), wdW, 2e-2)
        _close(db.get(), wdb, 2e-2)
