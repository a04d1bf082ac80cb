"""The known-answer tests the REFERENCE's own suite holds for the hot path, restated against this package's
public API with the same inputs and expected values -- on the host path (`device="cpu"`, CPU suite) and on the b200
backend (`device="cuda"`, GPU suite).  Each test cites the reference test it restates.  Tolerances: the reference asserts in
float64 terms; tensors here are float32 like the reference's (core dtype default), so value checks use the fp32 contract
(1e-5 relative, north star) where the reference's own tolerance is tighter than fp32 can hold; index / count results are
exact."""
import math

import numpy as np
import pytest

DEVICES = [pytest.param("cpu", id="host"), pytest.param("cuda", id="b200", marks=pytest.mark.gpu)]


@pytest.fixture(params=DEVICES)
def dev(request):
    if request.param == "cuda":
        import nfb200
        assert nfb200.get_backend("b200").is_available        # fail loudly, never fall back
    return request.param


def _np(x):
    return np.asarray(x.get() if hasattr(x, "get") else x)


def test_gelu_exact_and_tanh_values(dev):
    """tests/test_mathematical_correctness.py:77-105: exact GELU = 0.5 x (1 + erf(x / sqrt 2)); tanh form within 1e-3."""
    from scipy.special import erf
    from nfb200.core import Tensor
    from nfb200.functional import gelu
    for xv in (-3.0, -1.0, -0.5, 0.0, 0.5, 1.0, 3.0):
        x = Tensor([xv], requires_grad=True, device=dev)
        expected = 0.5 * xv * (1 + erf(xv / math.sqrt(2)))
        np.testing.assert_allclose(_np(gelu(x, approximate=False).data), [expected], rtol=1e-5, atol=1e-7)
        assert abs(float(_np(gelu(x, approximate=True).data)[0]) - expected) < 1e-3


def test_gelu_gradient_vs_numeric(dev):
    """tests/test_mathematical_correctness.py:107-131: analytic gradient vs central differences of the erf form (rtol 1e-3)."""
    from scipy.special import erf
    from nfb200.core import Tensor
    from nfb200.functional import gelu
    f = lambda v: 0.5 * v * (1 + erf(v / math.sqrt(2)))     # noqa: E731
    for xv in (-2.0, -0.5, 0.0, 0.5, 2.0):
        x = Tensor([xv], requires_grad=True, device=dev)
        gelu(x, approximate=False).backward()
        num = (f(xv + 1e-5) - f(xv - 1e-5)) / 2e-5
        np.testing.assert_allclose(_np(x.grad), [num], rtol=1e-3, atol=1e-5)


def test_layernorm_statistics_and_gradient_flow(dev):
    """tests/test_mathematical_correctness.py:166-190: rows of LayerNorm output have mean 0 (atol 1e-6), variance 1 (rtol 1e-5
    up to eps); x, weight and bias all receive gradients."""
    from nfb200.core import Tensor
    from nfb200.nn import LayerNorm
    rng = np.random.default_rng(0)
    x = Tensor(rng.standard_normal((4, 8)).astype(np.float32), requires_grad=True, device=dev)
    ln = LayerNorm(8, eps=1e-5)
    if dev == "cuda":
        ln.to(dev)
    y = ln(x)
    yd = _np(y.data).astype(np.float64)
    np.testing.assert_allclose(yd.mean(-1), 0.0, atol=1e-6)
    np.testing.assert_allclose(yd.var(-1), 1.0, rtol=1e-4)     # 1 / (1 + eps / var) in fp32
    (y * y).sum().backward()
    assert x.grad is not None and ln.weight.grad is not None and ln.bias.grad is not None


def test_add_mul_matmul_gradients(dev):
    """tests/test_mathematical_correctness.py:247-284 and tests/test_tensor.py:139-157 (dz/dx = 4, dz/dy = 2)."""
    from nfb200.core import Tensor
    from nfb200.functional import add, matmul, mul
    a = Tensor([2.0, 3.0], requires_grad=True, device=dev)
    b = Tensor([4.0, 5.0], requires_grad=True, device=dev)
    add(a, b).backward(np.array([1.0, 1.0], np.float32))
    assert np.array_equal(_np(a.grad), [1.0, 1.0]) and np.array_equal(_np(b.grad), [1.0, 1.0])
    a.zero_grad()
    b.zero_grad()
    mul(a, b).backward(np.array([1.0, 1.0], np.float32))
    assert np.array_equal(_np(a.grad), [4.0, 5.0]) and np.array_equal(_np(b.grad), [2.0, 3.0])
    x = Tensor([2.0], requires_grad=True, device=dev)
    y = Tensor([3.0], requires_grad=True, device=dev)
    add(mul(x, y), x).backward(np.array([1.0], np.float32))
    assert np.array_equal(_np(x.grad), [4.0]) and np.array_equal(_np(y.grad), [2.0])
    rng = np.random.default_rng(1)
    A = rng.standard_normal((3, 4)).astype(np.float32)
    B = rng.standard_normal((4, 5)).astype(np.float32)
    G = rng.standard_normal((3, 5)).astype(np.float32)
    At, Bt = Tensor(A, requires_grad=True, device=dev), Tensor(B, requires_grad=True, device=dev)
    matmul(At, Bt).backward(G)
    np.testing.assert_allclose(_np(At.grad), G @ B.T, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_np(Bt.grad), A.T @ G, rtol=1e-5, atol=1e-6)


def test_incoming_gradient_clip(dev):
    """tests/test_tensor.py:208-214: a 1e10 upstream gradient arrives clipped to |g| <= 10."""
    from nfb200.core import Tensor
    x = Tensor([1000.0], requires_grad=True, device=dev)
    x.backward(np.array([1e10], np.float32))
    assert abs(float(_np(x.grad)[0])) <= 10.0


def test_cross_entropy_value_and_gradient(dev):
    """tests/test_mathematical_correctness.py:286-313: loss = -mean log softmax[target]; grad = (softmax - onehot) / B."""
    from nfb200.core import Tensor
    from nfb200.functional import cross_entropy_loss
    rng = np.random.default_rng(2)
    z = rng.standard_normal((4, 5)).astype(np.float32)
    t = np.array([0, 1, 2, 3])
    logits = Tensor(z, requires_grad=True, device=dev)
    loss = cross_entropy_loss(logits, Tensor(t, device=dev))
    p = np.exp(z.astype(np.float64))
    p /= p.sum(1, keepdims=True)
    np.testing.assert_allclose(float(_np(loss.data)), -np.mean(np.log(p[np.arange(4), t])), rtol=1e-5)
    loss.backward()
    want = p.copy()
    want[np.arange(4), t] -= 1
    np.testing.assert_allclose(_np(logits.grad), want / 4, rtol=1e-4, atol=1e-6)


def test_sigmoid_softmax_extreme_values(dev):
    """tests/test_mathematical_correctness.py:315-332: finite, in range and normalised for |x| up to 1000."""
    from nfb200.core import Tensor
    from nfb200.functional import sigmoid, softmax
    for val in (-1000.0, -100.0, -10.0, 0.0, 10.0, 100.0, 1000.0):
        with np.errstate(over="ignore"):
            s = _np(sigmoid(Tensor([val], device=dev)).data)
        assert np.isfinite(s).all() and 0 <= s[0] <= 1
        sm = _np(softmax(Tensor([[val, val - 1]], device=dev)).data)
        assert np.isfinite(sm).all() and abs(sm.sum() - 1.0) < 1e-6


def test_embedding_lookup_and_repeated_index_gradient(dev):
    """tests/test_layers.py:95-104 (W = [[1,2],[3,4],[5,6]], idx [[0,1,2]] -> W itself) and :116-129 (idx [[0,0,1]], upstream
    ones -> dW[0,0] == 2.0, dW[1,0] == 1.0, EXACT equality; the indices are a raw ndarray, not a Tensor)."""
    from nfb200.nn import Embedding
    W = np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]], np.float32)
    layer = Embedding(3, 2)
    if dev == "cuda":
        layer.to(dev)
        import nfb200
        layer.weight.data = nfb200.get_backend("b200").from_numpy(W)
    else:
        layer.weight.data = W.copy()
    out = layer(np.array([[0, 1, 2]]))
    assert np.array_equal(_np(out.data), W[None])
    layer.weight.zero_grad()
    emb = layer(np.array([[0, 0, 1]]))
    emb.backward(np.ones((1, 3, 2), np.float32))
    g = _np(layer.weight.grad)
    assert g[0, 0] == 2.0 and g[1, 0] == 1.0 and g[2, 0] == 0.0


def test_adam_first_steps(dev):
    """tests/test_adam_optimizer.py:54-120: a unit gradient moves every parameter down; exp_avg is kept; the larger gradient
    gets the smaller effective step; the step counter feeds the bias correction."""
    from nfb200.core import Parameter
    from nfb200.optim import Adam

    def param(a):
        return Parameter(np.asarray(a, np.float32), device=dev)

    def grad(a):
        return np.asarray(a, np.float32)
    p = param(np.zeros((5, 5)))
    p.grad = grad(np.ones((5, 5)))
    Adam([p], lr=0.1).step()
    assert np.all(_np(p.data) < 0)
    p = param(np.zeros((3, 3)))
    opt = Adam([p], lr=0.1, beta1=0.9)
    for _ in range(5):
        p.grad = grad(np.ones((3, 3)))
        opt.step()
    st = opt.get_state_dict()["state"]
    st = st[list(st.keys())[0]]
    assert st["step"] == 5 and not np.allclose(_np(st["exp_avg"]), 0.0)
    p = param(np.zeros((2, 2)))
    p.grad = grad([[1.0, 0.1], [0.1, 1.0]])
    Adam([p], lr=0.1, beta2=0.999).step()
    d = _np(p.data)
    assert abs(d[0, 0]) < abs(d[0, 1]) * 2
