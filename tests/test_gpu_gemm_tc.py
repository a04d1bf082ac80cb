"""GPU parity of the tcgen05/TMEM bf16 GEMM against a float64 NumPy contraction of the
SAME bf16-rounded operands (so the only error left is fp32 accumulation order): all four operand
layouts (forward x.W^T, dgrad dY.W, wgrad dY^T.X), ragged M/N/K, every tile width, epilogues, split-K."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bf16(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(a))


def _to(be, a, dtype=None):
    return be.from_numpy(np.asarray(a), dtype)


def _ref(a, b, ta, tb):
    A = _bf16(a).astype(np.float64)
    B = _bf16(b).astype(np.float64)
    return (A.T if ta else A) @ (B.T if tb else B)


LAYOUTS = [(False, True), (False, False), (True, False), (True, True)]


@pytest.fixture(params=[0, 1, 2], ids=["auto", "cta1", "cta_pair"])
def cta_group(request):
    """0 = the library's own choice, 1 = single-CTA tiles, 2 = cta_group::2 SM-pair tiles (256 x BLOCK_N)."""
    from nfb200 import _lib
    _lib.call("nfb_gemm_bf16_force_cta_group", request.param)
    yield request.param
    _lib.call("nfb_gemm_bf16_force_cta_group", 0)


@pytest.mark.parametrize("ta,tb", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 128, 128), (512, 768, 768), (300, 200, 136), (129, 72, 72), (64, 40, 8), (4096, 2304, 768)])
def test_gemm_layouts_and_ragged_shapes(be, cta_group, ta, tb, M, N, K):
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    got = Kn.gemm_bf16(_to(be, a), _to(be, b), ta, tb).get()
    want = _ref(a, b, ta, tb)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-4 * np.sqrt(K / 64))


@pytest.mark.parametrize("bn", [64, 128, 192, 256])
def test_every_tile_width(be, cta_group, bn):
    from nfb200 import _lib, kernels as Kn
    rng = np.random.default_rng(bn)
    a = rng.standard_normal((384, 320)).astype(np.float32)
    b = rng.standard_normal((1000, 320)).astype(np.float32)
    _lib.call("nfb_gemm_bf16_force_tile", bn)
    try:
        got = Kn.gemm_bf16(_to(be, a), _to(be, b), False, True).get()
        got2 = Kn.gemm_bf16(_to(be, a.T.copy()), _to(be, b.T.copy()), True, False).get()
    finally:
        _lib.call("nfb_gemm_bf16_force_tile", 0)
    want = _ref(a, b, False, True)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=5e-4)
    np.testing.assert_allclose(got2, want, rtol=2e-5, atol=5e-4)


def test_many_tiles_persistent_schedule(be, cta_group):
    """More tiles than SMs: exercises smem-ring wrap-around and the double-buffered TMEM accumulator."""
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(1)
    a = rng.standard_normal((2048, 192)).astype(np.float32)
    b = rng.standard_normal((5000, 192)).astype(np.float32)
    got = Kn.gemm_bf16(_to(be, a), _to(be, b), False, True).get()
    np.testing.assert_allclose(got, _ref(a, b, False, True), rtol=2e-5, atol=5e-4)


def test_unaligned_output_pitch_logits_shape(be, cta_group):
    """N = 50257-style odd row pitch: coalesced scalar stores, no alignment requirement on C."""
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(2)
    x = rng.standard_normal((256, 128)).astype(np.float32)
    w = rng.standard_normal((5027, 128)).astype(np.float32)
    got = Kn.gemm_bf16(_to(be, x), _to(be, w), False, True).get()
    np.testing.assert_allclose(got, _ref(x, w, False, True), rtol=2e-5, atol=4e-4)


def test_epilogues_bias_gelu_relu_residual_bf16_out(be, cta_group):
    from nfb200 import kernels as Kn
    from oracle import np_ops as O
    rng = np.random.default_rng(3)
    x = rng.standard_normal((200, 256)).astype(np.float32)
    w = (rng.standard_normal((384, 256)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(384).astype(np.float32)
    res = rng.standard_normal((200, 384)).astype(np.float32)
    z = _ref(x, w, False, True) + bias
    y = Kn.gemm_bf16(_to(be, x), _to(be, w), False, True, bias=_to(be, bias)).get()
    np.testing.assert_allclose(y, z, rtol=2e-5, atol=3e-4)
    yg, aux = Kn.gemm_bf16(_to(be, x), _to(be, w), False, True, bias=_to(be, bias), epilogue="gelu", aux=True)
    np.testing.assert_allclose(yg.get(), O.gelu_tanh_fwd(z), rtol=1e-4, atol=3e-4)
    np.testing.assert_allclose(aux.get(), z, rtol=1e-2, atol=1e-2)  # bf16 pre-activation
    yr = Kn.gemm_bf16(_to(be, x), _to(be, w), False, True, bias=_to(be, bias), epilogue="relu", residual=_to(be, res)).get()
    np.testing.assert_allclose(yr, np.maximum(z, 0) + res, rtol=2e-5, atol=3e-4)
    yb = Kn.gemm_bf16(_to(be, x), _to(be, w), False, True, bias=_to(be, bias), out_dtype=be.bfloat16)
    assert yb.dtype is be.bfloat16
    np.testing.assert_allclose(yb.get(), z, rtol=1e-2, atol=1e-2)
    yc = Kn.gemm_bf16(_to(be, x), _to(be, w), False, True, bias=_to(be, bias), clip=0.5).get()
    np.testing.assert_allclose(yc, np.clip(z, -0.5, 0.5), rtol=2e-5, atol=3e-4)


@pytest.mark.parametrize("out_bf16", [False, True])
@pytest.mark.parametrize("M,N,K", [(512, 768, 256), (300, 200, 136), (4096, 2304, 768)])
def test_register_epilogue_without_tma_store(be, cta_group, M, N, K, out_bf16):
    """The TMA-store epilogue is the default for 16-byte addressable C; this pins the register (LDS + STG) epilogue."""
    from nfb200 import _lib, kernels as Kn
    rng = np.random.default_rng(M + N)
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((N, K)).astype(np.float32)
    _lib.call("nfb_gemm_bf16_set_tma_store", 0)
    try:
        got = Kn.gemm_bf16(_to(be, a), _to(be, b), False, True, out_dtype=be.bfloat16 if out_bf16 else np.float32).get()
    finally:
        _lib.call("nfb_gemm_bf16_set_tma_store", 1)
    if out_bf16:
        np.testing.assert_allclose(got, _ref(a, b, False, True), rtol=1e-2, atol=1e-2 * np.sqrt(K / 64))
    else:
        np.testing.assert_allclose(got, _ref(a, b, False, True), rtol=2e-5, atol=2e-4 * np.sqrt(K / 64))


@pytest.mark.parametrize("split_k", [0, 1, 3, 8])
def test_wgrad_split_k_accumulate(be, cta_group, split_k):
    """dW (+)= dY^T X with K = tokens: split-K partial sums land with red.global.add on a pre-initialised C."""
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(4)
    T, dout, din = 4096, 256, 384
    dy = (rng.standard_normal((T, dout)) * 0.1).astype(np.float32)
    x = rng.standard_normal((T, din)).astype(np.float32)
    base = rng.standard_normal((dout, din)).astype(np.float32)
    out = _to(be, base.copy())
    Kn.gemm_bf16(_to(be, dy), _to(be, x), True, False, out=out, accumulate=True, split_k=split_k)
    want = base + _ref(dy, x, True, False)
    np.testing.assert_allclose(out.get(), want, rtol=3e-5, atol=2e-3)


@pytest.mark.parametrize("raster", [1, 2], ids=["m_fastest", "n_fastest"])
def test_tile_order_for_l2_reuse(be, cta_group, raster):
    """Both tile orders of the persistent scheduler (m fastest / n fastest; the library picks by an L2-reuse estimate, which
    only flips for operands larger than L2, e.g. the LM-head weight gradient) on shapes with several ragged m AND n tiles:
    a forward, and a split-K weight gradient accumulating into C."""
    from nfb200 import _lib, kernels as Kn
    rng = np.random.default_rng(6)
    a = rng.standard_normal((1100, 200)).astype(np.float32)
    b = rng.standard_normal((900, 200)).astype(np.float32)
    T, dout, din = 2048, 1100, 600
    dy = (rng.standard_normal((T, dout)) * 0.1).astype(np.float32)
    x = rng.standard_normal((T, din)).astype(np.float32)
    base = rng.standard_normal((dout, din)).astype(np.float32)
    out = _to(be, base.copy())
    _lib.call("nfb_gemm_bf16_force_raster", raster)
    try:
        got = Kn.gemm_bf16(_to(be, a), _to(be, b), False, True).get()
        Kn.gemm_bf16(_to(be, dy), _to(be, x), True, False, out=out, accumulate=True, split_k=3)
    finally:
        _lib.call("nfb_gemm_bf16_force_raster", 0)
    np.testing.assert_allclose(got, _ref(a, b, False, True), rtol=2e-5, atol=5e-4)
    np.testing.assert_allclose(out.get(), base + _ref(dy, x, True, False), rtol=3e-5, atol=2e-3)


@pytest.mark.parametrize("ta,tb", LAYOUTS)
@pytest.mark.parametrize("out_dtype", ["f32", "bf16"])
def test_internal_split_k_forced(be, cta_group, ta, tb, out_dtype):
    """Non-accumulating launches cut along K: every k-split stores its partial product into its own fp32 slice and a finishing
    pass sums the slices in a fixed order, clips and casts (taken automatically only for under-filled long-K launches such
    as the LM-head input gradient; forced here on small ragged shapes)."""
    from nfb200 import _lib, kernels as Kn
    rng = np.random.default_rng(8)
    M, N, K = 300, 200, 2000
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    odt = be.bfloat16 if out_dtype == "bf16" else np.float32
    a_d, b_d = Kn.to_bf16(_to(be, a)), Kn.to_bf16(_to(be, b))
    n0 = _lib.launch_count()
    _lib.call("nfb_gemm_bf16_set_auto_split", 2)
    try:
        got = Kn.gemm_bf16(a_d, b_d, ta, tb, out_dtype=odt, clip=30.0)
    finally:
        _lib.call("nfb_gemm_bf16_set_auto_split", 1)
    launches = _lib.launch_count() - n0
    got = got.get()
    want = np.clip(_ref(a, b, ta, tb), -30.0, 30.0)
    assert np.abs(want).max() == 30.0                                  # the clip is active
    if out_dtype == "bf16":
        np.testing.assert_allclose(got, want, rtol=1e-2, atol=1e-2 * np.sqrt(K / 64))
    else:
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-4 * np.sqrt(K / 64))
    assert launches >= 2, launches                                      # GEMM + finishing pass (+ a pitch-aligning operand copy)


def test_internal_split_k_lm_head_dgrad(be):
    """The shape the cost model splits on its own: dX (4096, 768) = dlogits (4096, 50257) . wte (50257, 768).  Same result as
    the single launch (fp32 accumulation either way, one rounding to bf16) and two launches where there was one."""
    from nfb200 import _lib, kernels as Kn
    rng = np.random.default_rng(9)
    T, V, d = 4096, 50257, 768
    ldv = (V + 7) // 8 * 8
    dl = Kn.to_bf16(_to(be, (rng.standard_normal((T, ldv)) * 1e-2).astype(np.float32)))[:, :V]
    w = Kn.to_bf16(_to(be, (rng.standard_normal((V, d)) * 0.02).astype(np.float32)))
    _lib.call("nfb_gemm_bf16_set_auto_split", 0)
    _lib.call("nfb_gemm_bf16_set_stream_k", 0)     # (the stream-K schedule supersedes the internal split when it is on)
    try:
        n0 = _lib.launch_count()
        one = Kn.gemm_bf16(dl, w, False, False, out_dtype=be.bfloat16, clip=10.0)
        n1 = _lib.launch_count()
        _lib.call("nfb_gemm_bf16_set_auto_split", 1)
        split = Kn.gemm_bf16(dl, w, False, False, out_dtype=be.bfloat16, clip=10.0)
        n2 = _lib.launch_count()
        again = Kn.gemm_bf16(dl, w, False, False, out_dtype=be.bfloat16, clip=10.0).get()
    finally:
        _lib.call("nfb_gemm_bf16_set_auto_split", 1)
        _lib.call("nfb_gemm_bf16_set_stream_k", 1)
    assert n1 - n0 == 1 and n2 - n1 == 2, (n1 - n0, n2 - n1)
    a, b = one.get(), split.get()
    assert np.isfinite(b).all()
    assert np.array_equal(again, b)                                      # slices summed in a fixed order: run-to-run bit-identical
    # the same shape under the stream-K schedule: one launch, bit-identical run to run
    _lib.call("nfb_gemm_bf16_set_stream_k", 2)
    try:
        n3 = _lib.launch_count()
        sk1 = Kn.gemm_bf16(dl, w, False, False, out_dtype=be.bfloat16, clip=10.0)
        n4 = _lib.launch_count()
        sk1 = sk1.get()
        sk2 = Kn.gemm_bf16(dl, w, False, False, out_dtype=be.bfloat16, clip=10.0).get()
    finally:
        _lib.call("nfb_gemm_bf16_set_stream_k", 1)
    assert n4 - n3 == 1
    assert np.array_equal(sk1, sk2)
    np.testing.assert_allclose(sk1, a, rtol=8e-3, atol=8e-3 * float(np.abs(a).max()))
    np.testing.assert_allclose(b, a, rtol=8e-3, atol=8e-3 * float(np.abs(a).max()))   # one bf16 ulp of the larger values
    # against fp64 on a slice of rows
    ref = dl[:64].get().astype(np.float64) @ w.get().astype(np.float64)
    np.testing.assert_allclose(b[:64], ref, rtol=1e-2, atol=1e-2 * float(np.abs(ref).max()))


@pytest.fixture
def stream_k_forced():
    """Every eligible launch takes the stream-K schedule (equal (tile, k-block) ranges per cluster + in-kernel fix-up)."""
    from nfb200 import _lib
    _lib.call("nfb_gemm_bf16_set_stream_k", 2)
    yield
    _lib.call("nfb_gemm_bf16_set_stream_k", 1)


@pytest.mark.parametrize("ta,tb", LAYOUTS)
@pytest.mark.parametrize("M,N,K", [(256, 128, 128), (512, 768, 768), (300, 200, 136), (129, 72, 72), (4096, 768, 3072), (1000, 520, 4100), (2304, 768, 4096),
                                   (4096, 3072, 256), (8192, 2100, 136)])
def test_stream_k_layouts_and_ragged_shapes(be, cta_group, stream_k_forced, ta, tb, M, N, K):
    """Tiles cut by range boundaries are finished by the cluster holding their last k-block, which adds the partial tiles of
    the earlier clusters in cluster order: same values as the one-tile-per-cluster schedule up to fp32 summation order, and
    bit-identical from run to run."""
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    ad, bd = _to(be, a), _to(be, b)
    got = Kn.gemm_bf16(ad, bd, ta, tb).get()
    want = _ref(a, b, ta, tb)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-4 * np.sqrt(K / 64))
    assert np.array_equal(Kn.gemm_bf16(ad, bd, ta, tb).get(), got)


@pytest.mark.parametrize("bn", [64, 128, 192, 256])
def test_stream_k_every_tile_width_and_epilogue(be, cta_group, stream_k_forced, bn):
    """Owner-side epilogues under stream-K: bias + GELU with the bf16 pre-activation output (register epilogue), residual,
    clip and a bf16 TMA-store output; the accumulating (weight-gradient) form adds the finished tile once."""
    from nfb200 import _lib, kernels as Kn
    from oracle import np_ops as O
    rng = np.random.default_rng(bn)
    M, N, K = 520, 600, 1100
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    res = rng.standard_normal((M, N)).astype(np.float32)
    z = _ref(x, w, False, True) + bias
    _lib.call("nfb_gemm_bf16_force_tile", bn)
    try:
        xd, wd, bd = _to(be, x), _to(be, w), _to(be, bias)
        yg, aux = Kn.gemm_bf16(xd, wd, False, True, bias=bd, epilogue="gelu", aux=True)
        yr = Kn.gemm_bf16(xd, wd, False, True, bias=bd, residual=_to(be, res), clip=3.0).get()
        yb = Kn.gemm_bf16(xd, wd, False, True, bias=bd, out_dtype=be.bfloat16).get()
        base = rng.standard_normal((K, N)).astype(np.float32)
        acc = _to(be, base.copy())
        Kn.gemm_bf16(xd, _to(be, res), True, False, out=acc, accumulate=True)   # (K, N) += x^T res, contraction over M
    finally:
        _lib.call("nfb_gemm_bf16_force_tile", 0)
    np.testing.assert_allclose(yg.get(), O.gelu_tanh_fwd(z), rtol=1e-4, atol=4e-4)
    np.testing.assert_allclose(aux.get(), z, rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(yr, np.clip(z + res, -3.0, 3.0), rtol=2e-5, atol=6e-4)
    np.testing.assert_allclose(yb, z, rtol=1e-2, atol=2e-2)
    np.testing.assert_allclose(acc.get(), base + _ref(x, res, True, False), rtol=3e-5, atol=2e-3)


@pytest.mark.parametrize("out_dtype", ["f32", "bf16"])
def test_grouped_launch_weight_gradients(be, cta_group, out_dtype):
    """nfb_gemm_bf16_grouped: the four weight gradients of a Transformer block (ragged variants included) in one stream-K
    launch; every output matches its own single launch, the accumulating form adds onto the existing values, and two runs
    are bit-identical (each element receives exactly one store / reduce-add)."""
    from nfb200 import _lib, kernels as Kn
    rng = np.random.default_rng(11)
    T = 2048
    shapes = [(768, 1536), (520, 304), (3072, 768), (200, 72)]
    probs, wants, bases = [], [], []
    for (m, n) in shapes:
        dz = (rng.standard_normal((T, m)) * 0.1).astype(np.float32)
        x = rng.standard_normal((T, n)).astype(np.float32)
        probs.append((Kn.to_bf16(_to(be, dz)), Kn.to_bf16(_to(be, x))))
        wants.append(_ref(dz, x, True, False))
        bases.append(rng.standard_normal((m, n)).astype(np.float32))
    acc = out_dtype == "f32"
    odt = np.float32 if acc else be.bfloat16

    def run():
        if acc:
            outs = [_to(be, b.copy()) for b in bases]
        else:
            from nfb200.array import B200Array
            outs = [B200Array.empty(b.shape, odt) for b in bases]
        n0 = _lib.launch_count()
        Kn.gemm_bf16_grouped([(a, b, o) for (a, b), o in zip(probs, outs)], True, False, accumulate=acc)
        assert _lib.launch_count() - n0 == 1
        return [o.get() for o in outs]

    got = run()
    for g, w, b in zip(got, wants, bases):
        if acc:
            np.testing.assert_allclose(g, b + w, rtol=3e-5, atol=2e-3)
        else:
            np.testing.assert_allclose(g, w, rtol=1e-2, atol=1e-2 * np.sqrt(T / 64))
    for g, g2 in zip(got, run()):
        assert np.array_equal(g, g2)


@pytest.mark.parametrize("bn", [0, 64, 128, 192, 256])
def test_fused_rotary_epilogue(be, cta_group, bn):
    """qkv projection with the rotary embedding of q and k (models/language/gpt2.py:78-95) in the GEMM epilogue: output columns
    [0, 2d) -- heads of 64 columns -- are rotated by the position row % S BEFORE the rounding to bf16; the v columns are not.
    Against the fp64 projection + the oracle's rope_fwd, at the bf16 contract; and within one bf16 rounding of the unfused
    route (projection rounded to bf16, then k_rope in place)."""
    from nfb200 import _lib, kernels as Kn
    from oracle import np_ops as O
    rng = np.random.default_rng(21 + bn)
    Bn, S, H, D, K = 3, 100, 4, 64, 200
    d = H * D
    x = rng.standard_normal((Bn * S, K)).astype(np.float32)
    w = (rng.standard_normal((3 * d, K)) * 0.08).astype(np.float32)
    bias = rng.standard_normal(3 * d).astype(np.float32)
    z = (_ref(x, w, False, True) + bias).reshape(Bn, S, 3, H, D)
    cos, sin = O.rope_tables(D, S)
    want = z.copy()
    for part in (0, 1):
        want[:, :, part] = O.rope_fwd(z[:, :, part].transpose(0, 2, 1, 3), cos, sin).transpose(0, 2, 1, 3)
    _lib.call("nfb_gemm_bf16_force_tile", bn)
    try:
        fused = Kn.gemm_bf16(_to(be, x), _to(be, w), False, True, bias=_to(be, bias), out_dtype=be.bfloat16, rope=(S, 2 * d, D))
        plain = Kn.gemm_bf16(_to(be, x), _to(be, w), False, True, bias=_to(be, bias), out_dtype=be.bfloat16)
    finally:
        _lib.call("nfb_gemm_bf16_force_tile", 0)
    got = fused.get().reshape(Bn, S, 3, H, D)
    assert np.array_equal(got[:, :, 2], plain.get().reshape(Bn, S, 3, H, D)[:, :, 2])        # v: exactly the plain bf16 projection
    Kn.rope_packed_qk_inplace(plain.reshape(Bn, S, 3, H, D), H)
    np.testing.assert_allclose(got, want, rtol=1e-2, atol=1e-2 * float(np.abs(want).max()) / 4)
    unfused = plain.get().reshape(Bn, S, 3, H, D)
    np.testing.assert_allclose(got, unfused, rtol=2e-2, atol=2e-2 * float(np.abs(want).max()) / 4)
    exact = _bf16(want)                                                                      # one rounding of the rotated fp32 value
    assert np.mean(got != exact) < 0.02, np.mean(got != exact)                                # (fp32 summation-order flips only)


def test_gemm_f32_wrapper(be):
    from nfb200 import kernels as Kn
    rng = np.random.default_rng(5)
    x = rng.standard_normal((300, 200)).astype(np.float32)
    w = rng.standard_normal((150, 200)).astype(np.float32)
    b = rng.standard_normal(150).astype(np.float32)
    got = Kn.gemm_f32(_to(be, x), _to(be, w), False, True, bias=_to(be, b)).get()
    np.testing.assert_allclose(got, x.astype(np.float64) @ w.T.astype(np.float64) + b, rtol=1e-5, atol=2e-5)
    got = Kn.gemm_f32(_to(be, x), _to(be, x), True, False).get()
    np.testing.assert_allclose(got, x.T.astype(np.float64) @ x.astype(np.float64), rtol=1e-5, atol=3e-5)


@pytest.mark.parametrize("M,N,K,ta,tb", [(512, 768, 256, False, True), (1024, 2304, 768, False, False), (768, 768, 4096, True, False),
                                         (4096, 768, 3072, False, True), (256, 264, 1032, True, True), (256, 512, 9216, False, True),
                                         (512, 256, 8200, True, False)])
def test_fp32_contraction_through_the_bf16x3_split(be, M, N, K, ta, tb):
    """fp32-mode contractions on the tensor cores: both operands split three ways into bf16 (24 significand bits), six products
    as ONE bf16 GEMM over a six-fold K with fp32 accumulation (nfb_split_bf16x3 + nfb_gemm_bf16).  Held to the fp32 contract:
    <= 1e-5 of the row scale against the float64 product (functional/arithmetic.py:441-575; the FFMA kernel is held to the same),
    with bias and with accumulation into an existing C, and bit-identical run to run.  K > 4096 runs in chunks (the pipe's
    truncating accumulation error grows with K) whose results meet in round-to-nearest fp32 adds."""
    from nfb200 import kernels as Kn
    from nfb200.array import B200Array
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    a *= np.exp(rng.uniform(-6, 6, a.shape)).astype(np.float32)          # 5 decades of dynamic range inside every row
    bias = rng.standard_normal(N).astype(np.float32)
    c0 = rng.standard_normal((M, N)).astype(np.float32)
    A64 = (a.T if ta else a).astype(np.float64)
    B64 = (b.T if tb else b).astype(np.float64)
    want = A64 @ B64
    scale = np.abs(A64) @ np.abs(B64)                                     # the natural error scale of a dot product
    da, db_, dbias = B200Array.from_numpy(a), B200Array.from_numpy(b), B200Array.from_numpy(bias)
    Kn.set_fp32_tensor_core(True)
    try:
        got = Kn.gemm_f32(da, db_, ta, tb).get().astype(np.float64)
        assert np.max(np.abs(got - want) / scale) < 2e-6                   # the tensor pipe's own fp32 accumulation error; far inside 1e-5
        got2 = Kn.gemm_f32(da, db_, ta, tb).get().astype(np.float64)
        assert np.array_equal(got, got2)
        gb = Kn.gemm_f32(da, db_, ta, tb, bias=dbias).get().astype(np.float64)
        assert np.max(np.abs(gb - (want + bias)) / (scale + np.abs(bias))) < 2e-6
        out = B200Array.from_numpy(c0)
        Kn.gemm_f32(da, db_, ta, tb, out=out, beta=1.0)
        assert np.max(np.abs(out.get() - (want + c0)) / (scale + np.abs(c0))) < 2e-6
        # against the FFMA kernel: both inside the contract, so they agree to it
        Kn.set_fp32_tensor_core(False)
        ref = Kn.gemm_f32(da, db_, ta, tb).get().astype(np.float64)
        assert np.max(np.abs(ref - want) / scale) < 2e-6
    finally:
        Kn.set_fp32_tensor_core(True)
