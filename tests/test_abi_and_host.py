"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/nfb200.h declares, the ctypes
table matches the header, the product path fails LOUDLY without a GPU (no CPU fallback), and the host-side mirror of
the reference's backend registry / Device / DType / Module contracts behaves as pinned upstream."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nfb200.h")


def _declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nfb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from nfb200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 80
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"libnfb200.so is missing {missing}"


def test_ctypes_table_matches_header():
    from nfb200 import _lib
    bound = set(_lib._SIGNATURES) | set(_lib._SPECIAL_RESTYPE)
    declared = set(_declared())
    assert declared <= bound, f"header symbols without a ctypes signature: {sorted(declared - bound)}"
    assert bound <= declared, f"ctypes signatures missing from the header: {sorted(bound - declared)}"
    # argument counts agree with the header prototypes
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in _lib._SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^)]*)\)", txt)
        n = 0 if m.group(1).strip() in ("", "void") else len(m.group(1).split(","))
        assert n == len(args), f"{name}: header has {n} parameters, ctypes table has {len(args)}"


def _plan(trans_b, M, N, K, accumulate=0, epilogue=0, split_k=0, sms=148):
    from nfb200 import _lib
    out = (ctypes.c_int * 6)()
    _lib.call("nfb_gemm_bf16_plan", int(trans_b), M, N, K, accumulate, epilogue, split_k, sms, out)
    return dict(zip(("bn", "cg", "sk", "n_fast", "internal_split", "stream_k"), out))


def test_gemm_launch_plans_for_the_gpt2_small_step():
    """The host-side launch decisions of csrc/gemm_tc.cu (tile width, SM pairing, split-K, tile order, internal split-K)
    for the GEMM shapes of the bench configuration (T = 4096 tokens, d = 768, V = 50257) on 148 SMs.  Pure host logic:
    guards the measured choices (profiles/r01_gemm_shapes.txt) against accidental changes."""
    T, d, V = 4096, 768, 50257
    # forward x.W^T (B stored (N,K): transB=1): SM pairs on 256-row tiles; N = 768 takes 192-wide tiles (64 pairs of 74)
    assert _plan(1, T, 2304, d) == dict(bn=256, cg=2, sk=1, n_fast=0, internal_split=0, stream_k=0)
    assert _plan(1, T, d, d)["bn"] == 192 and _plan(1, T, d, 1536)["bn"] == 192
    # layer input gradients dY.W (MN-major B: 192-wide pair tiles are not available) stay one launch
    for N, K in ((1536, d), (d, 3072), (d, 2304), (d, d)):
        pl = _plan(0, T, N, K)
        assert pl["cg"] == 2 and pl["bn"] == 256 and pl["internal_split"] == 0, (N, K, pl)
    # LM head: logits and their input gradient; the latter fills 48 of 74 SM pairs -> cut 3 ways along K = 50257
    assert _plan(1, T, V, d)["internal_split"] == 0 and _plan(1, T, V, d)["n_fast"] == 0
    assert _plan(0, T, d, V) == dict(bn=256, cg=2, sk=3, n_fast=0, internal_split=3, stream_k=0)
    assert _plan(0, T, d, V, epilogue=1)["internal_split"] == 0          # a fused activation cannot be split
    # weight gradients accumulate (split-K by reduce-add); the LM-head one reads dlogits^T (412 MB > L2) once: n fastest
    lm = _plan(0, V, d, T, accumulate=1)
    assert lm["n_fast"] == 1 and lm["cg"] == 2 and lm["internal_split"] == 0
    for M, N in ((d, 1536), (3072, d), (d, d), (2304, d)):
        pl = _plan(0, M, N, T, accumulate=1)
        assert pl["sk"] > 1 and pl["n_fast"] == 0 and pl["internal_split"] == 0, (M, N, pl)
    assert _plan(0, d, d, T, accumulate=1, split_k=5)["sk"] == 5        # an explicit factor is honoured
    # a single-row-block problem cannot pair SMs
    assert _plan(1, 128, 512, 256)["cg"] == 1
    # single launches keep the classic schedule (the stream-K fix-up costs what the balance wins at these sizes, see
    # profiles/r02_gemm_streamk.txt); only the test hook and grouped launches (nfb_gemm_bf16_grouped) take stream-K
    from nfb200 import _lib
    assert all(_plan(tb, M, N, K)["stream_k"] == 0 for tb, M, N, K in ((0, T, 1536, d), (1, T, 3072, d), (0, T, d, 3072)))
    _lib.call("nfb_gemm_bf16_set_stream_k", 2)
    try:
        assert _plan(0, T, 1536, d)["stream_k"] == 1 and _plan(0, d, d, T, accumulate=1)["stream_k"] == 1
        assert _plan(0, d, d, T, accumulate=1, split_k=5)["stream_k"] == 0   # an explicit split-K factor keeps the classic path
    finally:
        _lib.call("nfb_gemm_bf16_set_stream_k", 1)


def test_no_cpu_fallback_without_gpu():
    import nfb200
    from nfb200 import _lib
    if _lib.device_available():
        pytest.skip("a GPU is present")
    assert "b200" not in nfb200.available_backends()
    with pytest.raises(RuntimeError):          # same contract as an unavailable reference backend (backend.py:348-349)
        nfb200.get_backend("b200")
    with pytest.raises(RuntimeError):
        nfb200.get_backend("cuda")
    from nfb200.core import Tensor
    with pytest.raises(RuntimeError):          # a cuda Tensor does NOT silently fall back to NumPy (PARITY.md)
        Tensor([1.0, 2.0], device="cuda")
    from nfb200.array import B200Array
    with pytest.raises(_lib.B200Error):
        B200Array.empty((4,))


def test_registry_contract():
    """tests/test_backends.py:24-43 of the reference."""
    import nfb200
    assert "numpy" in nfb200.available_backends()
    with pytest.raises(ValueError, match="Unknown backend"):
        nfb200.get_backend("invalid_backend")
    a, b = nfb200.get_backend("numpy"), nfb200.get_backend("numpy")
    assert a is not b and a.name == "numpy"    # a NEW instance per call (backend.py:334-351)
    nfb200.set_backend("numpy")
    assert nfb200.current_backend().name == "numpy"
    from nfb200.plugin import ABSTRACT_MEMBERS, Backend
    assert len(ABSTRACT_MEMBERS) == 54         # the 54 abstract members of backends/backend.py:9-322 (SURVEY says 56: a miscount)
    with pytest.raises(TypeError):
        Backend()
    from nfb200.b200_backend import B200Backend
    B200Backend()                              # implements all of them (instantiation would fail otherwise)
    assert nfb200.get_backend_for_device("cuda:3") == "cuda" and nfb200.get_backend_for_device("cpu") == "numpy"


def test_numpy_backend_known_answers():
    """tests/test_backends.py:99-181 on the host backend."""
    import nfb200
    be = nfb200.get_backend("numpy")
    a, b = be.array([[1, 2], [3, 4]], dtype=np.float32), be.array([[5, 6], [7, 8]], dtype=np.float32)
    assert np.array_equal(be.matmul(a, b), [[19, 22], [43, 50]])
    assert be.zeros((2, 2)).dtype == np.float32 and be.random_normal((3,)).dtype == np.float32
    with pytest.raises(ValueError):
        be.to_device(a, "cuda")
    assert be.device_of(a) == "cpu"


def test_device_and_dtype_contracts():
    """tests/test_device_complete.py:22-215 and tests/test_core_dtype_comprehensive.py:20-135 of the reference."""
    from nfb200.core import Device, DeviceType, DType, get_default_device, get_default_dtype, set_default_device, set_default_dtype
    assert [d.value for d in DeviceType] == ["cpu", "cuda", "mps"]
    assert str(Device.cuda(2)) == "cuda:2" and repr(Device.cuda(1)) == "Device('cuda', 1)" and Device.cpu().index is None
    assert Device.from_string("cuda:2") == Device.cuda(2) and Device.cuda(0).is_gpu and Device.cpu().is_cpu
    with pytest.raises(ValueError):
        Device.from_string("gpu")
    with pytest.raises(ValueError):
        Device(DeviceType.CUDA, -1)
    with pytest.raises(TypeError):
        set_default_device(3)
    assert get_default_device() == Device.cpu()
    assert len(list(DType)) == 5 and str(DType.FLOAT32) == "float32" and repr(DType.INT64) == "DType.INT64"
    assert DType.from_numpy("f4") is DType.FLOAT32 and DType.from_numpy(np.int64) is DType.INT64
    assert DType.FLOAT32 == np.float32 and DType.FLOAT64.bytes_per_element == 8 and DType.INT32.is_integer
    with pytest.raises(ValueError):
        DType.from_numpy(np.float16)
    with pytest.raises(TypeError):
        set_default_dtype("float32")
    assert get_default_dtype() is DType.FLOAT32


def test_tensor_autograd_contract_on_host():
    """tests/test_tensor.py:48-62,139-157,208-214,235-244 and tests/test_core_tensor_coverage_boost.py:97-144."""
    from nfb200.core import Parameter, Tensor, is_grad_enabled, no_grad
    x = Tensor([[1.0, 2.0]], requires_grad=True)
    y = Tensor([[3.0, 4.0]], requires_grad=True)
    z = x * y + x
    z.backward(np.ones((1, 2), np.float32))
    assert np.array_equal(x.grad, [[4.0, 5.0]]) and np.array_equal(y.grad, [[1.0, 2.0]])
    z2 = x * y
    z2.backward(np.ones((1, 2), np.float32))            # gradients accumulate across backward calls
    assert np.array_equal(x.grad, [[7.0, 9.0]])
    x.zero_grad()
    assert x.grad is None
    big = Tensor([1.0], requires_grad=True)
    (big * 1000.0).backward(np.array([100.0], np.float32))
    assert np.all(np.abs(big.grad) <= 10.0 * 1000.0) and np.all(np.isfinite(big.grad))
    w = Tensor([1.0], requires_grad=True)
    w.backward(np.array([1e6], np.float32))             # +-10 clip on the incoming gradient (test_tensor.py:208-214)
    assert w.grad[0] == 10.0
    frozen = Tensor([1.0])
    frozen.backward(np.array([1.0]))
    assert frozen.grad is None
    with pytest.raises(TypeError):
        Tensor("string")
    with pytest.raises(TypeError):
        Tensor([1.0], requires_grad="yes")
    assert Tensor(np.arange(3)).dtype.name == "INT64" and Tensor([1, 2]).dtype.name == "FLOAT32"
    with no_grad():
        assert not is_grad_enabled()
        assert (x + y).grad_fn is None
    p = Parameter(np.ones((2, 2), np.float32))
    assert p.requires_grad and "Parameter" in repr(p)
    # broadcasting reduces gradients to operand shapes (test_tensor.py:226-233)
    a, b = Tensor(np.ones((4, 3), np.float32), requires_grad=True), Tensor(np.ones((3,), np.float32), requires_grad=True)
    (a + b).backward(np.ones((4, 3), np.float32))
    assert b.grad.shape == (3,) and np.array_equal(b.grad, [4, 4, 4])


def test_module_parameter_view_and_state_dict():
    """tests/test_core_base_comprehensive.py:69-81,178-219,282-305 and the dict-style uses in nn/transformer.py:101-114."""
    from nfb200 import nn
    np.random.seed(0)
    lin = nn.Linear(4, 3)
    assert lin.weight.shape == (3, 4) and nn.StandardLinear(4, 3).weight.shape == (4, 3)   # (out,in) vs (in,out)
    ps = lin.parameters()
    assert len(list(ps)) == 2 and "weight" in ps and ps["weight"] is lin.weight and dict(ps.items())["bias"] is lin.bias
    blk = nn.TransformerBlock(8, 2, 16)
    names = [n for n, _ in blk.named_parameters()]
    assert "self_attn.query_proj.weight" in names and len(names) == len(set(names)) == 16
    sd = blk.state_dict()
    blk2 = nn.TransformerBlock(8, 2, 16)
    blk2.load_state_dict(sd)
    assert all(np.array_equal(sd[k], v) for k, v in blk2.state_dict().items())
    assert blk.training and not blk.eval().training and blk.train().self_attn.training
    with pytest.raises(ValueError):
        lin(__import__("nfb200").core.Tensor(np.ones((2, 5), np.float32)))


def test_gradients_by_finite_differences_on_host():
    """The reference's own method (tests/test_mathematical_correctness.py:34-59): every differentiable op of the hot path."""
    from nfb200 import functional as F
    from nfb200.core import Tensor, set_grad_clip
    rng = np.random.default_rng(0)
    set_grad_clip(None)
    try:
        def check(fn, *shapes, tol=2e-3):
            xs = [rng.standard_normal(s) for s in shapes]
            ts = [Tensor(x.astype(np.float64), requires_grad=True, dtype=np.float64) for x in xs]
            out = fn(*ts)
            w = rng.standard_normal(out.shape)
            out.backward(w)
            for i, x in enumerate(xs):
                num = np.zeros_like(x)
                it = np.nditer(x, flags=["multi_index"])
                for _ in it:
                    idx = it.multi_index
                    for sgn in (1, -1):
                        xp = [a.copy() for a in xs]
                        xp[i][idx] += sgn * 1e-5
                        val = float((fn(*[Tensor(a, dtype=np.float64) for a in xp]).data * w).sum())
                        num[idx] += sgn * val / 2e-5
                np.testing.assert_allclose(ts[i].grad, num, rtol=tol, atol=1e-6)

        check(lambda a, b: F.matmul(a, b), (3, 4), (4, 2))
        check(lambda a, b: F.div(F.mul(a, b), F.add(F.mul(b, b), 1.0)), (2, 3), (3,))
        check(lambda a: F.softmax(a, axis=-1), (2, 5))
        check(lambda a: F.gelu(a), (7,))
        check(lambda a: F.gelu(a, approximate=True), (7,))
        check(lambda a: F.silu(a), (7,))
        check(lambda a: F.tanh(F.sigmoid(a)), (6,))
        check(lambda a, w, b: F.layer_norm(a, w, b, 1e-5), (3, 8), (8,), (8,))
        check(lambda a, w: F.rms_norm(a, w, 1e-6), (3, 8), (8,))
        check(lambda a, w, b: F.linear(a, w, b, "gelu"), (2, 3, 4), (5, 4), (5,))
        check(lambda a, w, r: F.linear(a, w, None, None, r, None, "in_out"), (3, 4), (4, 5), (3, 5))
        check(lambda q: F.rope(q), (1, 2, 5, 8))
        check(lambda q, k, v: F.scaled_dot_product_attention(q, k, v, 0.5, causal=True), (1, 2, 4, 8), (1, 2, 4, 8), (1, 2, 4, 8))
        check(lambda x: F.packed_qkv_attention(x, 2, causal=True, use_rope=True), (1, 4, 48))
        check(lambda x: F.swiglu_packed(x), (3, 8))
        check(lambda z: F.cross_entropy_loss(z, np.array([1, 0, 3])), (3, 4))
        check(lambda a: F.mean(F.reshape(F.transpose(a, (1, 0, 2)), (4, 6)), axis=1), (2, 4, 3))
    finally:
        set_grad_clip(10.0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/neural_arch/backends"), reason="needs the reference checkout (build container only)")
def test_drop_in_registration_into_the_real_reference_registry():
    """INTEGRATION.md step 1 against the UNMODIFIED reference package: `nfb200.install_into_reference()` registers a class that
    IS a `neural_arch.backends.Backend` (the reference's own ABC) with every abstract member implemented, under the names
    "b200" and "cuda"; without a GPU the reference's registry answers as it does for any unavailable backend
    (`RuntimeError`, backends/backend.py:348-349) and never lists it as available -- no silent fallback.  Runs in a child
    process so the reference package does not leak into this one."""
    import subprocess
    import sys
    code = r'''
import sys, json
sys.path[:0] = [%r, %r]
import neural_arch.backends.backend as rb          # the reference's registry, through the namespace shim of oracle/compat
import nfb200
assert nfb200.install_into_reference() is True
cls = rb._BACKENDS["b200"]
assert rb._BACKENDS["cuda"] is cls and issubclass(cls, rb.Backend)
res = {"abstract_left": sorted(getattr(cls, "__abstractmethods__", ())), "n_abstract": len(rb.Backend.__abstractmethods__)}
be = cls()                                           # instantiation fails if any abstract member were missing
res["name"], res["available"] = be.name, bool(be.is_available)
try:
    rb.get_backend("b200")
    res["get_backend"] = "returned"
except RuntimeError as e:
    res["get_backend"] = "RuntimeError"
try:
    rb.get_backend("no-such-backend")
except ValueError:
    res["unknown"] = "ValueError"
res["listed"] = "b200" in rb.available_backends()
print(json.dumps(res))
''' % (os.path.join(ROOT, "oracle", "compat"), os.path.join(ROOT, "neural-network-from-scratch_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["abstract_left"] == [] and res["n_abstract"] >= 50
    assert res["name"] == "b200" and res["unknown"] == "ValueError"
    if not res["available"]:                                             # this container: no GPU
        assert res["get_backend"] == "RuntimeError" and res["listed"] is False
