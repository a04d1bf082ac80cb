"""GPU parity of the fused bandwidth-bound kernels against the CPU oracle (oracle/np_ops.py):
LayerNorm/RMSNorm fwd+bwd (a7, a8), GELU/SiLU/SwiGLU/ReLU/tanh/sigmoid fwd+bwd (a4, a11, a12), softmax (a5),
cross-entropy (a10), embedding gather + bit-exact scatter-add (a9), Adam/AdamW/SGD (a13), grad clip."""
import numpy as np
import pytest

from oracle import np_ops as O

pytestmark = pytest.mark.gpu


def _to(be, a, dtype=None):
    return be.from_numpy(np.asarray(a), dtype)


def _bf16_round(a):
    """Round-to-nearest-even fp32 -> bf16 -> fp32 on the host (for tolerance-free comparison of bf16 I/O)."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(a))


@pytest.mark.parametrize("rows,cols", [(64, 768), (37, 128), (5, 1024), (16, 256), (9, 100), (3, 1500)])
@pytest.mark.parametrize("rms", [False, True])
def test_norm_fwd_bwd_fp32(be, rows, cols, rms):
    rng = np.random.default_rng(rows * cols)
    x = (rng.standard_normal((rows, cols)) * 2 + 0.3).astype(np.float32)
    w = rng.standard_normal(cols).astype(np.float32)
    b = rng.standard_normal(cols).astype(np.float32)
    g = rng.standard_normal((rows, cols)).astype(np.float32)
    eps = 1e-5
    if rms:
        want, cache = O.rmsnorm_fwd(x, w, eps)
        wdx, wdw = O.rmsnorm_bwd(g, cache)
        wdb = None
    else:
        want, cache = O.layernorm_fwd(x, w, b, eps)
        wdx, wdw, wdb = O.layernorm_bwd(g, cache)
    out = be.norm_forward(_to(be, x), _to(be, w), None if rms else _to(be, b), eps, rms=rms)
    y, mean, rstd = out[:3]
    np.testing.assert_allclose(y.get(), want, rtol=1e-5, atol=2e-6)
    dx, dw, db = be.norm_backward(_to(be, g), _to(be, x), _to(be, w), mean, rstd, rms=rms)
    np.testing.assert_allclose(dx.get(), wdx, rtol=1e-4, atol=3e-6)
    np.testing.assert_allclose(dw.get(), wdw, rtol=1e-5, atol=1e-4)
    if not rms:
        np.testing.assert_allclose(db.get(), wdb, rtol=1e-5, atol=1e-4)


def test_layernorm_eps_1e12_and_stats(be):
    """BERT uses eps=1e-12 (bert.py:59-107); tests/test_mathematical_correctness.py:166-190 checks mean~0, var~1."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((128, 768)).astype(np.float32) * 5 + 2
    y = be.layernorm(_to(be, x), _to(be, np.ones(768, np.float32)), _to(be, np.zeros(768, np.float32)), 1e-12).get()
    assert np.abs(y.mean(-1)).max() < 1e-5 and np.abs(y.var(-1) - 1).max() < 1e-4
    np.testing.assert_allclose(y, O.layernorm_fwd(x, np.ones(768, np.float32), np.zeros(768, np.float32), 1e-12)[0], rtol=1e-5, atol=3e-6)


def test_norm_bf16_out_residual_fusion(be):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((96, 768)).astype(np.float32)
    r = rng.standard_normal((96, 768)).astype(np.float32)
    w = rng.standard_normal(768).astype(np.float32)
    y, _, rstd, s = be.norm_forward(_to(be, x), _to(be, w), None, 1e-5, rms=True, out_dtype=be.bfloat16, residual=_to(be, r))
    assert np.array_equal(s.get(), x + r)
    want = O.rmsnorm_fwd(x + r, w, 1e-5)[0]
    np.testing.assert_allclose(y.get(), want, rtol=1e-2, atol=1e-2)  # bf16 output: stated 1e-2 tolerance
    g = rng.standard_normal((96, 768)).astype(np.float32)
    dres = rng.standard_normal((96, 768)).astype(np.float32)
    dx, dw, _ = be.norm_backward(_to(be, g, be.bfloat16), s, _to(be, w), None, rstd, rms=True, dresidual=_to(be, dres))
    wdx, wdw = O.rmsnorm_bwd(_bf16_round(g), O.rmsnorm_fwd(x + r, w, 1e-5)[1])
    np.testing.assert_allclose(dx.get(), wdx + dres, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dw.get(), wdw, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("op,fwd,bwd,src_is_out", [
    ("relu", O.relu_fwd, O.relu_bwd, False), ("tanh", O.tanh_fwd, O.tanh_bwd, True), ("sigmoid", O.sigmoid_fwd, O.sigmoid_bwd, True),
    ("gelu_erf", O.gelu_erf_fwd, O.gelu_erf_bwd, False), ("gelu_tanh", O.gelu_tanh_fwd, O.gelu_tanh_bwd, False),
    ("silu", O.silu_fwd, O.silu_bwd, False)])
def test_activations_fwd_bwd(be, op, fwd, bwd, src_is_out):
    from nfb200 import array as A
    rng = np.random.default_rng(2)
    x = (rng.standard_normal((257, 513)) * 3).astype(np.float32)
    x.flat[:6] = [0.0, -0.0, 1000.0, -1000.0, 30.0, -30.0]  # stability at extremes (test_mathematical_correctness.py:315-332)
    g = rng.standard_normal(x.shape).astype(np.float32)
    y = A.unary(op, _to(be, x)).get()
    want = fwd(x.astype(np.float64)).astype(np.float32) if op != "relu" else fwd(x)
    np.testing.assert_allclose(y, want, rtol=1e-5, atol=1e-6)
    assert np.all(np.isfinite(y))
    src = want if src_is_out else x
    dx = be.unary_backward(op, _to(be, src), _to(be, g)).get()
    wdx = bwd(g.astype(np.float64), src.astype(np.float64)).astype(np.float32)
    np.testing.assert_allclose(dx, wdx, rtol=2e-5, atol=2e-6)


def test_gelu_exact_vs_erf_reference_values(be):
    """tests/test_mathematical_correctness.py:77-106 known values: gelu(0)=0, gelu(1)=0.8413447, gelu(-1)=-0.1586553."""
    y = be.gelu(_to(be, np.array([0.0, 1.0, -1.0, 2.0], np.float32))).get()
    np.testing.assert_allclose(y, [0.0, 0.8413447460685429, -0.15865525393145707, 1.9544997361036416], rtol=1e-6, atol=1e-7)
    yt = be.gelu(_to(be, np.array([0.5, 1.0, 2.0], np.float32)), approximate=True).get()
    assert np.abs(yt - be.gelu(_to(be, np.array([0.5, 1.0, 2.0], np.float32))).get()).max() < 1e-3


@pytest.mark.parametrize("shape,axis", [((4, 12, 64, 64), -1), ((8, 512), -1), ((3, 7, 33), 1), ((16, 50257), -1), ((5, 1, 2049), -1)])
def test_softmax_fwd_bwd(be, shape, axis):
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(shape) * 4).astype(np.float32)
    g = rng.standard_normal(shape).astype(np.float32)
    p = be.softmax(_to(be, x), axis=axis)
    want = O.softmax_fwd(x, axis)
    np.testing.assert_allclose(p.get(), want, rtol=1e-5, atol=1e-8)
    dx = be.softmax_backward(p, _to(be, g), axis=axis).get()
    np.testing.assert_allclose(dx, O.softmax_bwd(g, want, axis), rtol=1e-4, atol=2e-7)


def test_softmax_extremes_and_causal_constant(be):
    x = np.array([[1000.0, -1000.0, 0.0], [-1e4, -1e4, 3.0], [-1e4, -1e4, -1e4]], np.float32)
    p = be.softmax(_to(be, x)).get()
    np.testing.assert_allclose(p, O.softmax_fwd(x), rtol=1e-6, atol=1e-12)
    assert np.all(np.isfinite(p))


@pytest.mark.parametrize("rows,V", [(64, 1000), (16, 50257), (7, 30522), (5, 17), (3, 70001)])
@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_cross_entropy(be, rows, V, reduction):
    """Value + gradient (softmax - onehot)/B at rtol 1e-6-ish (tests/test_mathematical_correctness.py:286-313)."""
    rng = np.random.default_rng(rows + V)
    z = (rng.standard_normal((rows, V)) * 3).astype(np.float32)
    t = rng.integers(0, V, rows)
    wl, cache = O.cross_entropy_fwd(z, t, reduction)
    wd = O.cross_entropy_bwd(np.float32(1.0), cache)
    loss, dl = be.cross_entropy(_to(be, z), _to(be, t), reduction)
    np.testing.assert_allclose(loss.get(), wl, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dl.get(), wd, rtol=1e-5, atol=1e-9)
    loss32, _ = be.cross_entropy(_to(be, z), _to(be, t.astype(np.int32)), reduction, need_grad=False)
    np.testing.assert_allclose(loss32.get(), wl, rtol=1e-5, atol=1e-6)


def test_cross_entropy_bf16_grad_and_onehot_targets(be):
    rng = np.random.default_rng(5)
    z = (rng.standard_normal((32, 4096)) * 2).astype(np.float32)
    t = rng.integers(0, 4096, 32)
    onehot = np.eye(4096, dtype=np.float32)[t]
    loss, dl = be.cross_entropy(_to(be, z), _to(be, onehot), "mean", grad_dtype=be.bfloat16)
    wl, cache = O.cross_entropy_fwd(z, onehot, "mean")
    np.testing.assert_allclose(loss.get(), wl, rtol=1e-5)
    np.testing.assert_allclose(dl.get(), O.cross_entropy_bwd(np.float32(1), cache), rtol=1e-2, atol=1e-6)


def _bf16_round(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(a))


@pytest.mark.parametrize("rows,V", [(20, 50257), (300, 30522), (9, 8192), (5, 53248), (1000, 8195)])
@pytest.mark.parametrize("mode", ["f32_reg", "bf16_bulk", "bf16_reg"])
def test_cross_entropy_wide_rows_lm_head_layout(be, rows, V, mode):
    """The LM-head layout: (T, V) logits with the row pitch padded to 8 elements (garbage in the pad), fp32 through the
    register-resident kernel, bf16 through the bulk-copy shared-memory kernel and the register kernel; fp32 softmax math
    either way (functional/loss.py:15-112).  bf16 logits are compared against the oracle ON the bf16-rounded logits."""
    from nfb200 import _lib
    from nfb200.array import B200Array
    rng = np.random.default_rng(rows + V)
    z = (rng.standard_normal((rows, V)) * 3).astype(np.float32)
    t = rng.integers(0, V, rows)
    t[0], t[-1] = V - 1, 0
    ld = (V + 7) // 8 * 8
    zp = np.full((rows, ld), np.nan, np.float32)   # NaN pad: a kernel that reads the pad poisons max / sum
    zp[:, :V] = z
    if mode == "f32_reg":
        dev = _to(be, zp)[:, :V]
        zin = z
    else:
        dev = be.astype(_to(be, zp), be.bfloat16)[:, :V]
        zin = _bf16_round(z)
    wl, cache = O.cross_entropy_fwd(zin, t, "mean")
    wd = O.cross_entropy_bwd(np.float32(1.0), cache)
    _lib.call("nfb_cross_entropy_set_bulk", 0 if mode == "bf16_reg" else 1)
    try:
        loss, dl = be.cross_entropy(dev, _to(be, t), "mean", grad_dtype=np.float32)
        np.testing.assert_allclose(loss.get(), wl, rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(dl.get(), wd, rtol=1e-5, atol=1e-9)
        loss2, dl2 = be.cross_entropy(dev, _to(be, t), "mean", grad_dtype=be.bfloat16)
        np.testing.assert_allclose(loss2.get(), wl, rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(dl2.get(), wd, rtol=1e-2, atol=1e-7)
        loss3, _ = be.cross_entropy(dev, _to(be, t.astype(np.int32)), "sum", need_grad=False)
        np.testing.assert_allclose(loss3.get(), O.cross_entropy_fwd(zin, t, "sum")[0], rtol=2e-6)
    finally:
        _lib.call("nfb_cross_entropy_set_bulk", 1)


def test_embedding_gather_known_and_bit_exact(be):
    """tests/test_layers.py:95-104 (identity gather) and tests/test_nn_embedding_95_coverage.py:133-181."""
    W = np.array([[1, 2], [3, 4], [5, 6]], np.float32)
    out = be.embedding_forward(_to(be, W), _to(be, np.array([[0, 1, 2]]))).get()
    assert np.array_equal(out, W[None])
    rng = np.random.default_rng(6)
    W = rng.standard_normal((50257, 768)).astype(np.float32)
    idx = rng.integers(0, 50257, (8, 512))
    got = be.embedding_forward(_to(be, W), _to(be, idx)).get()
    assert np.array_equal(got, O.embedding_fwd(W, idx))
    got32 = be.embedding_forward(_to(be, W), _to(be, idx.astype(np.int32))).get()
    assert np.array_equal(got32, got)


def test_embedding_scatter_add_known_answer(be):
    """tests/test_layers.py:116-129: idx [[0,0,1]], upstream ones -> dW[0,0]==2.0, dW[1,0]==1.0 exactly."""
    dW = be.embedding_backward(_to(be, np.ones((1, 3, 2), np.float32)), _to(be, np.array([[0, 0, 1]])), 3).get()
    assert dW[0, 0] == 2.0 and dW[1, 0] == 1.0 and dW[2, 0] == 0.0


@pytest.mark.parametrize("n,d,V,hot", [(4096, 768, 50257, False), (4096, 768, 50257, True), (640, 128, 1000, True), (513, 100, 37, True), (16384, 768, 2, True)])
def test_embedding_scatter_add_bit_exact(be, n, d, V, hot):
    """Bit-exact against the reference's sequential loop order (nn/embedding.py:52-55), with heavy duplicates."""
    rng = np.random.default_rng(n + d)
    idx = rng.integers(0, min(V, 50) if hot else V, n)
    g = rng.standard_normal((n, d)).astype(np.float32)
    want = O.embedding_bwd_fast(g, idx, V)
    if n <= 640:
        assert np.array_equal(want, O.embedding_bwd(g, idx, V))  # the vectorised oracle equals the literal loop
    got = be.embedding_backward(_to(be, g), _to(be, idx), V).get()
    assert np.array_equal(got, want)
    # accumulate into an existing gradient (tied lm_head/wte weights): existing + segment_sum, in that order
    base = rng.standard_normal((V, d)).astype(np.float32)
    acc = be.embedding_backward(_to(be, g), _to(be, idx.astype(np.int32)), V, out=_to(be, base.copy()), accumulate=True).get()
    assert np.array_equal(acc, base + want)


def test_index_out_of_range_is_detected(be):
    import ctypes as C
    from nfb200 import _lib
    W = _to(be, np.zeros((4, 8), np.float32))
    be.embedding_forward(W, _to(be, np.array([0, 9])))
    flag = C.c_int(0)
    _lib.call("nfb_index_error_check", C.byref(flag))
    assert flag.value == 1
    be.embedding_forward(W, _to(be, np.array([0, 3, -1])))
    _lib.call("nfb_index_error_check", C.byref(flag))
    assert flag.value == 0


def test_swiglu(be):
    rng = np.random.default_rng(7)
    gate = rng.standard_normal((64, 1536)).astype(np.float32)
    up = rng.standard_normal((64, 1536)).astype(np.float32)
    d = rng.standard_normal((64, 1536)).astype(np.float32)
    out = be.swiglu_forward(_to(be, gate), _to(be, up)).get()
    np.testing.assert_allclose(out, O.silu_fwd(gate) * up, rtol=1e-5, atol=1e-6)
    dg, du = be.swiglu_backward(_to(be, gate), _to(be, up), _to(be, d))
    np.testing.assert_allclose(du.get(), d * O.silu_fwd(gate), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dg.get(), O.silu_bwd(d * up, gate), rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("variant", ["adam", "adamw"])
@pytest.mark.parametrize("moments", ["f64", "f32"])
def test_adam_family_trajectory(be, variant, moments):
    """10 optimizer steps against optim/adam.py / optim/adamw.py semantics."""
    import ctypes as C
    from nfb200 import _lib
    from nfb200.array import B200Array
    rng = np.random.default_rng(8)
    n = 100003
    p0 = rng.standard_normal(n).astype(np.float32)
    p_ref = p0.copy()
    st = O.AdamState(n)
    p = _to(be, p0)
    mdt = np.float64 if moments == "f64" else np.float32
    m, v = be.zeros((n,), mdt), be.zeros((n,), mdt)
    hyper = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-5, weight_decay=0.0) if variant == "adam" else dict(lr=5e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.1)
    for step in range(1, 11):
        g = (rng.standard_normal(n) * (10.0 ** rng.integers(-4, 1))).astype(np.float32)
        p_ref = (O.adam_step if variant == "adam" else O.adamw_step)(p_ref, g, st, **hyper)
        gd = _to(be, g)
        _lib.call("nfb_adam_step", 0 if variant == "adam" else 1, _lib.F64 if moments == "f64" else _lib.F32, p.ptr, gd.ptr, m.ptr, v.ptr, None, n,
                  hyper["lr"], hyper["betas"][0], hyper["betas"][1], hyper["eps"], hyper["weight_decay"], step, 0, None, None)
        if moments == "f64":
            np.testing.assert_allclose(p.get(), p_ref, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(p.get(), p_ref, rtol=1e-5, atol=1e-6)
    if moments == "f64":
        np.testing.assert_allclose(m.get(), st.exp_avg, rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(v.get(), st.exp_avg_sq, rtol=1e-12, atol=1e-18)


def test_adam_nonfinite_scrub_and_update_clip(be):
    from nfb200 import _lib
    p0 = np.array([1.0, 2.0, 3.0, 4.0], np.float32)
    g = np.array([np.nan, np.inf, 1e30, 1.0], np.float32)
    st = O.AdamState(4)
    with np.errstate(all="ignore"):
        want = O.adam_step(p0.copy(), g, st, lr=1.0, eps=1e-5)
    p = _to(be, p0)
    m, v = be.zeros((4,), np.float64), be.zeros((4,), np.float64)
    _lib.call("nfb_adam_step", 0, _lib.F64, p.ptr, _to(be, g).ptr, m.ptr, v.ptr, None, 4, 1.0, 0.9, 0.999, 1e-5, 0.0, 1, 0, None, None)
    np.testing.assert_allclose(p.get(), want, rtol=1e-6)


def test_sgd_momentum(be):
    from nfb200 import _lib
    rng = np.random.default_rng(9)
    n = 4099
    p_ref = rng.standard_normal(n).astype(np.float32)
    p = _to(be, p_ref)
    buf_ref, buf = None, be.zeros((n,))
    for step in range(5):
        g = rng.standard_normal(n).astype(np.float32)
        p_ref, buf_ref = O.sgd_step(p_ref, g, buf_ref, 0.1, momentum=0.9, weight_decay=1e-4, nesterov=True)
        _lib.call("nfb_sgd_step", p.ptr, _to(be, g).ptr, buf.ptr, None, n, 0.1, 0.9, 0.0, 1e-4, 1, 1 if step == 0 else 0, 0)
    np.testing.assert_allclose(p.get(), p_ref, rtol=1e-5, atol=1e-6)


def test_global_norm_clip(be):
    from nfb200 import _lib
    rng = np.random.default_rng(10)
    gs = [rng.standard_normal(s).astype(np.float32) * 3 for s in ((1000, 33), (517,), (64, 64))]
    want, norm = O.clip_grad_norm([g.copy() for g in gs], 1.0)
    total = be.zeros((1,), np.float64)
    coef, nrm = be.zeros((1,)), be.zeros((1,))
    dev = [_to(be, g) for g in gs]
    for i, d in enumerate(dev):
        _lib.call("nfb_sumsq", d.ptr, d.size, total.ptr, 1 if i else 0)
    _lib.call("nfb_clip_coef", total.ptr, 1.0, 1e-6, coef.ptr, nrm.ptr)
    for d in dev:
        _lib.call("nfb_scale_by_device_scalar", d.ptr, d.size, coef.ptr)
    np.testing.assert_allclose(nrm.get()[0], norm, rtol=1e-6)
    for d, w in zip(dev, want):
        np.testing.assert_allclose(d.get(), w, rtol=1e-6, atol=1e-9)
