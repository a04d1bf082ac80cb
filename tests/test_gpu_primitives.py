"""GPU parity: Backend primitive ops vs NumPy on the same
seeded inputs.  Mirrors the structure of the reference's tests/test_backends.py:46-284 and
tests/test_backend_accuracy.py:76-184 (known answers + cross-backend tolerances)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6  # north star: fp32 path within 1e-5 relative


def _to(be, a):
    return be.from_numpy(np.asarray(a))


def test_known_answers_2x2(be):
    """tests/test_backends.py:99-147 of the reference."""
    a = be.array([[1, 2], [3, 4]], dtype=np.float32)
    b = be.array([[5, 6], [7, 8]], dtype=np.float32)
    assert np.array_equal(be.to_numpy(be.add(a, b)), [[6, 8], [10, 12]])
    assert np.array_equal(be.to_numpy(be.subtract(a, b)), [[-4, -4], [-4, -4]])
    assert np.array_equal(be.to_numpy(be.multiply(a, b)), [[5, 12], [21, 32]])
    np.testing.assert_allclose(be.to_numpy(be.divide(a, b)), np.array([[1, 2], [3, 4]], np.float32) / np.array([[5, 6], [7, 8]], np.float32), rtol=1e-7)
    assert np.array_equal(be.to_numpy(be.power(a, 2)), [[1, 4], [9, 16]])
    assert np.array_equal(be.to_numpy(be.matmul(a, b)), [[19, 22], [43, 50]])


def test_reductions_known(be):
    """tests/test_backends.py:149-181."""
    a = be.array([[1, 2, 3], [4, 5, 6]], dtype=np.float32)
    assert be.to_numpy(be.sum(a)) == 21
    assert np.array_equal(be.to_numpy(be.sum(a, axis=0)), [5, 7, 9])
    assert np.array_equal(be.to_numpy(be.sum(a, axis=1)), [6, 15])
    assert be.to_numpy(be.mean(a)) == 3.5
    assert be.to_numpy(be.max(a)) == 6 and be.to_numpy(be.min(a)) == 1
    assert be.to_numpy(be.argmax(a)) == 5 and be.to_numpy(be.argmin(a)) == 0
    assert np.array_equal(be.to_numpy(be.argmax(a, axis=1)), [2, 2])


@pytest.mark.parametrize("shape_a,shape_b", [((64, 128), (64, 128)), ((8, 16, 768), (768,)), ((8, 1, 32), (1, 5, 32)),
                                             ((4, 12, 33, 33), (4, 1, 1, 33)), ((7,), ()), ((3, 5), (3, 1))])
@pytest.mark.parametrize("op", ["add", "subtract", "multiply", "divide", "maximum", "minimum"])
def test_binary_broadcast_bit_exact(be, op, shape_a, shape_b):
    rng = np.random.default_rng(0)
    a = rng.standard_normal(shape_a).astype(np.float32)
    b = (rng.standard_normal(shape_b).astype(np.float32) + 3.0)
    got = be.to_numpy(getattr(be, op)(_to(be, a), _to(be, b)))
    want = getattr(np, op)(a, b)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want)  # single IEEE op per element: bit exact


@pytest.mark.parametrize("dtype", [np.float64, np.int32, np.int64])
def test_binary_other_dtypes(be, dtype):
    rng = np.random.default_rng(1)
    a = (rng.standard_normal((5, 17)) * 10).astype(dtype)
    b = (rng.standard_normal((17,)) * 10).astype(dtype) + 1
    for op in ("add", "subtract", "multiply"):
        assert np.array_equal(be.to_numpy(getattr(be, op)(_to(be, a), _to(be, b))), getattr(np, op)(a, b))
    for op in ("less", "greater_equal", "equal", "not_equal"):
        got = be.to_numpy(getattr(be, op)(_to(be, a), _to(be, b)))
        assert got.dtype == np.bool_ and np.array_equal(got, getattr(np, op)(a, b))


def test_scalar_operands_and_operators(be):
    rng = np.random.default_rng(2)
    a = rng.standard_normal((33, 65)).astype(np.float32)
    x = _to(be, a)
    for fn in (lambda v: v + 2.5, lambda v: 2.5 - v, lambda v: v * 0.1, lambda v: 1.0 / (v * v + 1.0), lambda v: -v,
               lambda v: v ** 2, lambda v: (v > 0.0), lambda v: abs(v)):
        got, want = fn(x), fn(a)
        assert np.array_equal(got.get(), want), fn
    i = _to(be, np.arange(12, dtype=np.int64).reshape(3, 4))
    assert np.array_equal((i * 2 + 1).get(), np.arange(12).reshape(3, 4) * 2 + 1)
    assert (i / 2).get().dtype == np.float64


@pytest.mark.parametrize("name,fn,tol", [
    ("exp", np.exp, 2e-7), ("log", None, 2e-7), ("sqrt", None, 0), ("abs", np.abs, 0), ("sign", np.sign, 0), ("tanh", np.tanh, 4e-7)])
def test_unary(be, name, fn, tol):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((1000, 37)).astype(np.float32) * 3
    if name in ("log", "sqrt"):
        a = np.abs(a) + 0.1
        fn = getattr(np, name)
    got = be.to_numpy(getattr(be, name)(_to(be, a)))
    if tol == 0:
        assert np.array_equal(got, fn(a))
    else:
        np.testing.assert_allclose(got, fn(a), rtol=tol * 10, atol=tol)


def test_clip_where_astype_concat_stack_split(be):
    rng = np.random.default_rng(4)
    a = rng.standard_normal((6, 10)).astype(np.float32)
    x = _to(be, a)
    assert np.array_equal(be.to_numpy(be.clip(x, -0.5, 0.5)), np.clip(a, -0.5, 0.5))
    assert np.array_equal(be.to_numpy(be.where(be.greater(x, 0), x, be.zeros(a.shape))), np.where(a > 0, a, 0))
    assert np.array_equal(be.to_numpy(be.where(be.greater(x, 0), x, -1e4)), np.where(a > 0, a, np.float32(-1e4)))
    assert np.array_equal(be.to_numpy(be.astype(x, np.int32)), a.astype(np.int32))
    assert np.array_equal(be.to_numpy(be.concatenate([x, x], axis=1)), np.concatenate([a, a], axis=1))
    assert np.array_equal(be.to_numpy(be.stack([x, x, x], axis=0)), np.stack([a, a, a]))
    parts = be.split(_to(be, np.array([1, 2, 3, 4, 5, 6])), 3)
    assert len(parts) == 3 and np.array_equal(be.to_numpy(parts[0]), [1, 2])
    assert np.array_equal(be.to_numpy(be.arange(0, 10, 2)), np.arange(0, 10, 2, dtype=np.float32))


def test_views_transpose_reshape_index(be):
    rng = np.random.default_rng(5)
    a = rng.standard_normal((4, 6, 8, 10)).astype(np.float32)
    x = _to(be, a)
    assert np.array_equal(x.transpose(0, 2, 1, 3).get(), a.transpose(0, 2, 1, 3))
    assert np.array_equal(x.transpose(0, 1, 3, 2).reshape(4, 6, 80).get(), a.transpose(0, 1, 3, 2).reshape(4, 6, 80))
    assert np.array_equal(x[:, 0].get(), a[:, 0])
    assert np.array_equal(x[1:3, ::2, -1].get(), a[1:3, ::2, -1])
    assert np.array_equal(x[:, None, :, 3:7, 0].get(), a[:, None, :, 3:7, 0])
    big = rng.standard_normal((300, 500)).astype(np.float32)
    assert np.array_equal(_to(be, big).T.contiguous().get(), big.T)
    y = _to(be, a.copy())
    y[0, :, 2] = 7.0
    y[1] = x[2]
    b = a.copy()
    b[0, :, 2] = 7.0
    b[1] = a[2]
    assert np.array_equal(y.get(), b)
    idx = np.array([[3, 0], [1, 3]])
    assert np.array_equal(x[_to(be, idx)].get(), a[idx])


@pytest.mark.parametrize("axis", [None, 0, 1, 2, (0, 1), (1, 2), (0, 2), -1])
@pytest.mark.parametrize("op", ["sum", "mean", "max", "min"])
def test_reduce_axes(be, op, axis):
    """tests/test_backends.py:334-363 tolerance: rtol 1e-4 / atol 1e-6; we hold 1e-5."""
    rng = np.random.default_rng(6)
    a = rng.standard_normal((10, 20, 30)).astype(np.float32)
    for keep in (False, True):
        got = be.to_numpy(getattr(be, op)(_to(be, a), axis=axis, keepdims=keep))
        want = getattr(np, op)(a, axis=axis, keepdims=keep)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=2e-5 if op in ("sum",) else ATOL)


def test_reduce_large_and_columns(be):
    rng = np.random.default_rng(7)
    a = rng.standard_normal((4096, 768)).astype(np.float32)
    np.testing.assert_allclose(be.to_numpy(be.sum(_to(be, a), axis=0)), a.sum(axis=0, dtype=np.float64), rtol=1e-5, atol=2e-4)
    np.testing.assert_allclose(be.to_numpy(be.sum(_to(be, a))), a.sum(dtype=np.float64), rtol=1e-5, atol=1e-2)
    np.testing.assert_allclose(be.to_numpy(be.mean(_to(be, a), axis=-1)), a.mean(axis=-1), rtol=1e-5, atol=1e-6)
    ints = rng.integers(0, 50257, (8, 512))
    assert np.array_equal(be.to_numpy(be.argmax(_to(be, ints), axis=1)), np.argmax(ints, axis=1))
    assert int(be.to_numpy(be.sum(_to(be, ints)))) == int(ints.sum())


def test_argmax_ties_first_occurrence(be):
    a = np.array([[1, 5, 5, 2], [7, 7, 7, 7], [0, -1, -1, 0]], np.float32)
    assert np.array_equal(be.to_numpy(be.argmax(_to(be, a), axis=1)), np.argmax(a, axis=1))
    assert np.array_equal(be.to_numpy(be.argmin(_to(be, a), axis=1)), np.argmin(a, axis=1))
    assert np.array_equal(be.to_numpy(be.argmax(_to(be, a), axis=0)), np.argmax(a, axis=0))


@pytest.mark.parametrize("shape_a,shape_b", [((64, 64), (64, 64)), ((10, 3, 4), (10, 4, 5)), ((128, 768), (768, 300)),
                                             ((2, 12, 50, 64), (2, 12, 64, 50)), ((5, 7, 33), (33, 9)), ((17,), (17, 4)), ((6, 17), (17,))])
def test_matmul_fp32(be, shape_a, shape_b):
    """tests/test_backends.py:300-332 (rtol 1e-3) and test_backend_accuracy.py:76-108 (rtol 1e-5 / atol 1e-7)."""
    rng = np.random.default_rng(8)
    a = rng.standard_normal(shape_a).astype(np.float32)
    b = rng.standard_normal(shape_b).astype(np.float32)
    got = be.to_numpy(be.matmul(_to(be, a), _to(be, b)))
    want = np.matmul(a.astype(np.float64), b.astype(np.float64))
    assert got.shape == want.shape
    scale = np.sqrt(a.shape[-1])
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=2e-6 * scale)


def test_matmul_transposed_views(be):
    rng = np.random.default_rng(9)
    x = rng.standard_normal((37, 96)).astype(np.float32)
    w = rng.standard_normal((50, 96)).astype(np.float32)
    got = (_to(be, x) @ _to(be, w).T).get()
    np.testing.assert_allclose(got, x @ w.T, rtol=RTOL, atol=1e-5)
    got2 = (_to(be, x).T @ _to(be, x)).get()
    np.testing.assert_allclose(got2, x.T @ x, rtol=RTOL, atol=2e-5)


def test_einsum(be):
    rng = np.random.default_rng(10)
    a = rng.standard_normal((2, 3, 5, 4)).astype(np.float32)
    b = rng.standard_normal((2, 3, 6, 4)).astype(np.float32)
    np.testing.assert_allclose(be.to_numpy(be.einsum("bhqd,bhkd->bhqk", _to(be, a), _to(be, b))), np.einsum("bhqd,bhkd->bhqk", a, b), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(be.to_numpy(be.einsum("ij->ji", _to(be, a[0, 0]))), a[0, 0].T)


def test_numpy_protocols(be):
    """reference layers call np.* directly on tensor.data."""
    rng = np.random.default_rng(11)
    a = rng.standard_normal((4, 9, 16)).astype(np.float32)
    x = _to(be, a)
    np.testing.assert_allclose(np.mean(x, axis=-1, keepdims=True).get(), np.mean(a, axis=-1, keepdims=True), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(np.var(x, axis=-1, keepdims=True).get(), np.var(a, axis=-1, keepdims=True), rtol=1e-5, atol=1e-7)
    assert np.array_equal(np.sqrt(np.abs(x)).get(), np.sqrt(np.abs(a)))
    assert np.array_equal(np.maximum(0.0, x).get(), np.maximum(0.0, a))
    assert np.array_equal(np.transpose(x, (0, 2, 1)).get(), np.transpose(a, (0, 2, 1)))
    assert np.array_equal(np.concatenate((-x[..., 8:], x[..., :8]), axis=-1).get(), np.concatenate((-a[..., 8:], a[..., :8]), axis=-1))
    assert np.array_equal(np.where(x > 0, x, np.float32(-1e4)).get(), np.where(a > 0, a, np.float32(-1e4)))
    assert bool(np.all(np.isfinite(x)))
    with pytest.raises(TypeError):
        np.asarray(x)  # no implicit device->host copies
    with pytest.raises(NotImplementedError):
        np.fft.fft(x)  # unsupported functions raise; there is no CPU fallback


def test_device_contract(be):
    """tests/test_backends.py:448-474."""
    arr = be.array([1, 2, 3, 4])
    arr2 = be.to_device(arr, "cuda")
    assert "cuda" in be.device_of(arr2)
    with pytest.raises(ValueError):
        be.to_device(arr, "cpu")
    assert be.is_array(arr) and be.shape(arr) == (4,) and be.size(arr) == 4
    z = be.zeros((2, 3))
    assert be.dtype(z) == np.float32 and np.array_equal(be.to_numpy(z), np.zeros((2, 3), np.float32))
    assert np.array_equal(be.to_numpy(be.full((2,), 7.5)), np.full((2,), 7.5, np.float32))
    np.random.seed(0)
    r = be.to_numpy(be.random_normal((3, 3)))
    np.random.seed(0)
    assert np.array_equal(r, np.random.normal(0, 1, (3, 3)).astype(np.float32))
