"""GPU parity at the edges of the op surface: zero-size operands, single
rows / single columns, NaN ordering in argmax / max, and buffers beyond 2^31 elements (64-bit indexing in the
elementwise, reduction and gather kernels).  NumPy on the same inputs is the oracle; integer / index results are
bit-exact, one-IEEE-op results too, reductions within the fp32 contract (1e-5 relative)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _to(be, a):
    return be.from_numpy(np.asarray(a))


def test_zero_size_elementwise_and_views(be):
    a = np.zeros((0, 5), np.float32)
    b = np.ones((5,), np.float32)
    x = _to(be, a)
    for op in ("add", "subtract", "multiply", "divide", "maximum"):
        got = be.to_numpy(getattr(be, op)(x, _to(be, b)))
        want = getattr(np, op)(a, b)
        assert got.shape == want.shape == (0, 5) and got.dtype == want.dtype
    assert be.to_numpy(be.exp(x)).shape == (0, 5)
    assert be.to_numpy(be.transpose(x, (1, 0))).shape == (5, 0)
    assert be.to_numpy(be.reshape(x, (5, 0, 1))).shape == (5, 0, 1)
    assert be.to_numpy(be.astype(x, np.int64)).dtype == np.int64
    assert be.to_numpy(be.where(be.greater(x, 0), x, 0.0)).shape == (0, 5)
    full = _to(be, np.arange(10, dtype=np.float32).reshape(2, 5))
    cat = be.to_numpy(be.concatenate([x, full, x], axis=0))
    assert np.array_equal(cat, np.arange(10, dtype=np.float32).reshape(2, 5))
    assert be.to_numpy(full[1:1]).shape == (0, 5)
    full[1:1] = 7.0                                                       # assignment to an empty slice is a no-op
    assert np.array_equal(be.to_numpy(full), np.arange(10, dtype=np.float32).reshape(2, 5))


def test_zero_size_reductions(be):
    a = np.zeros((0, 5), np.float32)
    x = _to(be, a)
    assert be.to_numpy(be.sum(x)) == 0.0
    assert np.array_equal(be.to_numpy(be.sum(x, axis=0)), np.zeros(5, np.float32))
    assert be.to_numpy(be.sum(x, axis=1)).shape == (0,)
    with np.errstate(all="ignore"):
        assert np.isnan(be.to_numpy(be.mean(x)))
    with pytest.raises(ValueError):
        be.max(x)
    with pytest.raises(ValueError):
        be.argmax(x)
    assert be.to_numpy(be.max(x, axis=1)).shape == (0,)                   # reducing a non-empty axis of an empty array is fine


def test_zero_size_matmul(be):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((0, 4)).astype(np.float32)
    w = rng.standard_normal((4, 3)).astype(np.float32)
    assert be.to_numpy(be.matmul(_to(be, a), _to(be, w))).shape == (0, 3)
    # inner dimension 0: an empty sum is 0 (np.matmul((3,0),(0,4)) = zeros)
    got = be.to_numpy(be.matmul(_to(be, np.zeros((3, 0), np.float32)), _to(be, np.zeros((0, 4), np.float32))))
    assert np.array_equal(got, np.zeros((3, 4), np.float32))
    got = be.to_numpy(be.matmul(_to(be, np.zeros((2, 0, 4), np.float32)), _to(be, w)))
    assert got.shape == (2, 0, 3)


def test_zero_rows_through_fused_ops(be):
    d, V = 64, 97
    x0 = _to(be, np.zeros((0, d), np.float32))
    g = _to(be, np.ones(d, np.float32))
    y, mean, rstd = be.norm_forward(x0, g, g, 1e-5, rms=False)
    assert be.to_numpy(y).shape == (0, d)
    assert be.to_numpy(be.softmax(x0, axis=-1)).shape == (0, d)
    W = _to(be, np.arange(V * d, dtype=np.float32).reshape(V, d))
    idx = _to(be, np.zeros((2, 0), np.int64))
    out = be.embedding_forward(W, idx)
    assert be.to_numpy(out).shape == (2, 0, d)
    dW = be.embedding_backward(_to(be, np.zeros((2, 0, d), np.float32)), idx, V)
    assert np.array_equal(be.to_numpy(dW), np.zeros((V, d), np.float32))


def test_single_row_single_column(be):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((1, 768)).astype(np.float32)
    g = rng.standard_normal(768).astype(np.float32)
    b = rng.standard_normal(768).astype(np.float32)
    y, mean, rstd = be.norm_forward(_to(be, x), _to(be, g), _to(be, b), 1e-5, rms=False)
    mu = x.mean(-1, keepdims=True)
    want = g * (x - mu) / np.sqrt(x.var(-1, keepdims=True) + 1e-5) + b
    np.testing.assert_allclose(be.to_numpy(y), want, rtol=1e-5, atol=1e-5)
    # softmax over an axis of length 1 is exactly 1
    s = be.to_numpy(be.softmax(_to(be, rng.standard_normal((7, 1)).astype(np.float32)), axis=-1))
    assert np.array_equal(s, np.ones((7, 1), np.float32))
    # a (1, 1) contraction and a dot of two vectors
    assert be.to_numpy(be.matmul(_to(be, np.float32([[3.0]])), _to(be, np.float32([[4.0]])))) == 12.0
    u, v = rng.standard_normal(1000).astype(np.float32), rng.standard_normal(1000).astype(np.float32)
    np.testing.assert_allclose(be.to_numpy(be.dot(_to(be, u), _to(be, v))), np.dot(u.astype(np.float64), v.astype(np.float64)), rtol=1e-5, atol=1e-4)


def test_cross_entropy_single_row_and_extreme_logits(be):
    """functional/loss.py:15-112: softmax of max-shifted logits, loss = -log(p + 1e-8): a one-hot-certain wrong answer gives
    -log(1e-8), never inf."""
    V = 50257
    logits = np.full((1, V), -30.0, np.float32)
    logits[0, 5] = 30.0
    for tgt, want in ((5, -np.log(1.0 + 1e-8)), (6, -np.log(np.exp(np.float32(-60.0)) + 1e-8))):
        loss, dlog = be.cross_entropy(_to(be, logits), _to(be, np.array([tgt], np.int64)), "mean", True)
        got = float(np.asarray(be.to_numpy(loss)).reshape(-1)[0])
        assert np.isfinite(got)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
        d = be.to_numpy(dlog)
        assert d.shape == (1, V) and np.isfinite(d).all()
        np.testing.assert_allclose(d.sum(), 0.0, atol=1e-5)                # softmax - onehot sums to zero


def test_argmax_nan_and_ties(be):
    """np.argmax: NaN compares as the maximum, first occurrence wins; ties -> first index (bit-exact index contract)."""
    a = np.array([[1.0, np.nan, 3.0, np.nan], [2.0, 2.0, 1.0, 2.0], [-np.inf, -np.inf, -np.inf, -np.inf]], np.float32)
    x = _to(be, a)
    assert np.array_equal(be.to_numpy(be.argmax(x, axis=1)), np.argmax(a, axis=1))
    assert np.array_equal(be.to_numpy(be.argmin(x, axis=1)), np.argmin(a, axis=1))
    assert be.to_numpy(be.argmax(x)) == np.argmax(a)
    long = np.zeros(1 << 20, np.float32)
    long[[12345, 700000]] = 5.0
    assert be.to_numpy(be.argmax(_to(be, long))) == 12345
    with np.errstate(all="ignore"):
        got = be.to_numpy(be.max(x, axis=1))
        want = np.max(a, axis=1)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])


def test_beyond_2_31_elements(be):
    """64-bit indexing: an elementwise op, a full reduction, a strided view and a row gather on buffers of more than 2^31
    elements (8.6 GB each; sized for the 180 GB of a B200)."""
    n = (1 << 31) + 1024
    a = be.full((n,), 1.0, dtype=np.float32)
    b = be.full((n,), 2.0, dtype=np.float32)
    a[n - 3:] = 5.0                                                       # writes through a view that starts past 2^31
    c = be.add(a, b)
    assert np.array_equal(be.to_numpy(c[n - 5:]), np.float32([3, 3, 7, 7, 7]))
    assert np.array_equal(be.to_numpy(c[(1 << 31) - 2:(1 << 31) + 2]), np.float32([3, 3, 3, 3]))
    total = float(be.to_numpy(be.sum(c)))
    np.testing.assert_allclose(total, 3.0 * n + 12.0, rtol=1e-5)
    del a, b
    # row gather from a table whose last rows sit beyond 2^31 elements
    d = 1024
    rows = n // d
    table = be.reshape(c[:rows * d], (rows, d))
    idx = np.array([0, rows - 1, rows - 2, 1 << 20], np.int64)
    got = be.to_numpy(be.embedding_forward(table, _to(be, idx)))
    assert got.shape == (4, d) and np.array_equal(got[:, 0], np.float32([3, 3, 3, 3]))
    tail = be.to_numpy(table[rows - 1])                                   # the last row ends with the three 7s written above
    assert np.array_equal(tail, np.concatenate([np.full(d - 3, 3.0, np.float32), np.full(3, 7.0, np.float32)]))
    # strided (every-other-row) view keeps 64-bit offsets: rows-1 = 2^21 is even, so the view ends with the last row
    ev = be.to_numpy(be.sum(table[::2][rows // 2 - 2:], axis=1))
    assert np.array_equal(ev, np.float32([3.0 * d, 3.0 * d, 3.0 * d + 12.0]))


@pytest.mark.parametrize("S,Skv,causal", [(1, 1, False), (1, 1, True), (20, 17, False), (130, 300, False), (300, 1, False),
                                          (1, 513, False), (17, 20, True), (257, 129, True)])
def test_attention_tiny_and_cross_shapes(be, S, Skv, causal):
    """Query and key ranges that differ (the translation decoder's cross-attention, examples/translation/model_v2.py) and
    one-token sequences, against the oracle's unfused softmax(q k^T) v; bf16 contract 1e-2."""
    from oracle import np_ops as O
    from nfb200 import kernels as Kn
    from nfb200.array import B200Array
    rng = np.random.default_rng(S * 1000 + Skv)
    B, H, D = 2, 3, 64

    def bf(x):
        return be.to_numpy(be.astype(be.astype(_to(be, x.astype(np.float32)), be.bfloat16), np.float32))
    q, k, v, go = (bf(rng.standard_normal(s)) for s in ((B, H, S, D), (B, H, Skv, D), (B, H, Skv, D), (B, H, S, D)))
    scale = 1.0 / np.sqrt(D)
    want, cache = O.attention_fwd(q.astype(np.float64), k.astype(np.float64), v.astype(np.float64), scale, "causal" if causal else "none", None)
    wdq, wdk, wdv = O.attention_bwd(go.astype(np.float64), cache)
    dq_, dk_, dv_ = (be.astype(_to(be, x), be.bfloat16) for x in (q, k, v))
    o, lse = Kn.attention_forward(dq_, dk_, dv_, scale, causal=causal)
    assert isinstance(o, B200Array) and o.shape == (B, H, S, D)

    def close(got, want_, tol=1e-2):
        sc = max(1e-6, float(np.abs(want_).max()))
        assert float(np.abs(got - want_).max()) <= tol * sc + 1e-3 * tol
    close(o.get(), want)
    dq, dk, dv = Kn.attention_backward(be.astype(_to(be, go), be.bfloat16), dq_, dk_, dv_, o, lse, scale, causal=causal)
    assert dq.shape == (B, H, S, D) and dk.shape == (B, H, Skv, D) and dv.shape == (B, H, Skv, D)
    close(dv.get(), wdv)
    close(dq.get(), wdq)
    close(dk.get(), wdk)
