"""Child process of tests/test_gpu_reference_dropin.py: imports the UNMODIFIED reference package (oracle/_ref/neural_arch, staged
by oracle/make_ref.py; its missing `core` comes from oracle/compat) and runs the reference's OWN layer / functional / optimizer
code on Device.cuda tensors -- i.e. on B200Array through the reference's NumPy-protocol call sites -- next to the same code on
Device.cpu (the reference's NumPy backend).  Prints one JSON object: {check: {"max_rel": .., "ok": ..} | {"error": ..}}."""
import json
import os
import sys
import traceback

os.environ.setdefault("NUMBA_DISABLE_JIT", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "neural_arch")
os.environ["NEURAL_ARCH_REFERENCE"] = REF
sys.path[:0] = [os.path.join(ROOT, "oracle", "compat"), ROOT, os.path.join(ROOT, "neural-network-from-scratch_b200")]

import numpy as np  # noqa: E402

import neural_arch  # noqa: E402  (the shim: core from nfb200.core, everything else from the reference tree)
import nfb200  # noqa: E402
from neural_arch.core import Parameter, Tensor, set_default_device  # noqa: E402
from neural_arch.optimization_config import configure  # noqa: E402

RESULTS = {}


def host(a):
    a = a.data if isinstance(a, Tensor) else a
    return np.asarray(a.get() if hasattr(a, "get") else a)


def rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30))


def check(name, tol=1e-5):
    def deco(fn):
        try:
            pairs = fn()
            errs = [rel(host(g), host(w)) for g, w in pairs]
            worst = max(errs)
            RESULTS[name] = {"max_rel": worst, "ok": bool(worst <= tol), "n": len(pairs), "each": [float(f"{e:.3e}") for e in errs]}
        except Exception:
            RESULTS[name] = {"error": traceback.format_exc()[-1500:]}
        return fn
    return deco


def on(device, build):
    """Run `build()` (which creates layers / tensors) with `device` as the default device; same NumPy seed both times."""
    set_default_device(device)
    try:
        np.random.seed(7)
        return build(device)
    finally:
        set_default_device("cpu")


configure(enable_jit=False, enable_fusion=False)   # the reference's plain NumPy Linear path
# the reference's optimizers call np.asarray(...) on what they computed (optim/adam.py:245): that can only be served by a host
# copy, which B200Array refuses unless the caller opts in; the number of such copies is reported
from nfb200.array import implicit_host_conversions, set_implicit_host_conversion  # noqa: E402
set_implicit_host_conversion(True)
rng = np.random.default_rng(0)
X2 = rng.standard_normal((6, 16)).astype(np.float32)
U2 = rng.standard_normal((6, 8)).astype(np.float32)
X3 = (rng.standard_normal((3, 5, 16)) * 2 + 0.5).astype(np.float32)
U3 = rng.standard_normal((3, 5, 16)).astype(np.float32)


@check("registry: install_into_reference")
def _registry():
    import neural_arch.backends as RB
    from neural_arch.backends import backend as ref_backend
    assert nfb200.install_into_reference()
    names = RB.available_backends()
    assert "b200" in names and "cuda" in names, names
    be = RB.get_backend("b200")
    assert isinstance(be, ref_backend.Backend) and be.is_available and be.name in ("b200", "cuda")
    a = be.from_numpy(np.array([[1, 2], [3, 4]], np.float32))
    b = be.from_numpy(np.array([[5, 6], [7, 8]], np.float32))
    return [(be.to_numpy(be.matmul(a, b)), np.array([[19, 22], [43, 50]], np.float32))]     # the KAT of tests/test_backends.py:133-147


@check("nn.Linear (OptimizedLinear, reference nn/optimized.py) fwd+bwd")
def _linear():
    from neural_arch.nn import Linear

    def build(dev):
        lin = Linear(16, 8)
        t = Tensor(X2, requires_grad=True, device=dev)
        out = lin(t)
        out.backward(Tensor(U2, device=dev).data)
        return [out, t.grad, lin.weight.grad, lin.bias.grad]
    return list(zip(on("cuda", build), on("cpu", build)))


@check("nn.linear.Linear (in,out) (reference nn/linear.py:13-207) fwd+bwd")
def _std_linear():
    from neural_arch.nn.linear import Linear as StandardLinear

    def build(dev):
        lin = StandardLinear(16, 8)
        t = Tensor(X2, requires_grad=True, device=dev)
        out = lin(t)
        out.backward(Tensor(U2, device=dev).data)
        return [out, t.grad, lin.weight.grad, lin.bias.grad]
    return list(zip(on("cuda", build), on("cpu", build)))


@check("nn.LayerNorm (reference nn/normalization.py:86-221) fwd+bwd")
def _layernorm():
    from neural_arch.nn import LayerNorm

    def build(dev):
        ln = LayerNorm(16, eps=1e-5)
        t = Tensor(X3, requires_grad=True, device=dev)
        out = ln(t)
        out.backward(Tensor(U3, device=dev).data)
        return [out, t.grad, ln.weight.grad, ln.bias.grad]
    return list(zip(on("cuda", build), on("cpu", build)))


@check("nn.RMSNorm (reference nn/normalization.py:228-359) fwd+bwd")
def _rmsnorm():
    from neural_arch.nn import RMSNorm

    def build(dev):
        rn = RMSNorm(16, eps=1e-8)
        t = Tensor(X3, requires_grad=True, device=dev)
        out = rn(t)
        out.backward(Tensor(U3, device=dev).data)
        return [out, t.grad, rn.weight.grad]
    return list(zip(on("cuda", build), on("cpu", build)))


@check("nn.Embedding (reference nn/embedding.py:27-63) gather + ordered scatter-add", tol=0.0)
def _embedding():
    from neural_arch.nn import Embedding
    idx = np.array([[0, 0, 1, 12], [5, 1, 0, 5]])
    ue = rng.standard_normal((2, 4, 8)).astype(np.float32)

    def build(dev):
        emb = Embedding(13, 8)
        out = emb(Tensor(idx, device=dev))
        out.backward(Tensor(ue, device=dev).data)
        if hasattr(out, "_backward"):
            out._backward()
        return [out, emb.weight.grad]
    return list(zip(on("cuda", build), on("cpu", build)))


@check("functional.gelu / softmax / relu / sigmoid / tanh (reference functional/activation.py) fwd+bwd", tol=2e-5)
def _activations():
    import neural_arch.functional as RF
    x = (rng.standard_normal((5, 9)) * 2.5).astype(np.float32)
    u = rng.standard_normal((5, 9)).astype(np.float32)
    fns = [RF.gelu, lambda t: RF.gelu(t, approximate=True), lambda t: RF.softmax(t, axis=-1), RF.relu, RF.sigmoid, RF.tanh]

    def build(dev):
        outs = []
        for fn in fns:
            t = Tensor(x, requires_grad=True, device=dev)
            out = fn(t)
            out.backward(Tensor(u, device=dev).data)
            outs += [out, t.grad]
        return outs
    got, want = on("cuda", build), on("cpu", build)
    # A defect of the reference itself, reproduced faithfully: for a backend that offers `gelu` and whose arrays have a CuPy-style
    # `.get()`, functional/activation.py:336-340 feeds erf with the RAW input (`np_erf_input = x.backend_data.get()`) instead
    # of x / sqrt(2) (computed two lines above, then ignored).  The device result must equal what that code says.
    from scipy.special import erf
    want[1] = u * (0.5 * (1.0 + erf(x)) + x * np.exp(-0.5 * x ** 2) / np.sqrt(2.0 * np.pi))
    return list(zip(got, want))


@check("functional.add / mul / matmul (reference functional/arithmetic.py) fwd+bwd")
def _arith():
    import neural_arch.functional as RF
    a = rng.standard_normal((4, 6)).astype(np.float32)
    b = (rng.standard_normal((6,)) + 3).astype(np.float32)
    m2 = rng.standard_normal((6, 5)).astype(np.float32)

    def build(dev):
        outs = []
        for fn in (RF.add, RF.sub, RF.mul, RF.div):
            ta, tb = Tensor(a, requires_grad=True, device=dev), Tensor(b, requires_grad=True, device=dev)
            out = fn(ta, tb)
            out.backward(Tensor(np.ones((4, 6), np.float32), device=dev).data)
            outs += [out, ta.grad, tb.grad]
        ta, tb = Tensor(a, requires_grad=True, device=dev), Tensor(m2, requires_grad=True, device=dev)
        out = RF.matmul(ta, tb)
        out.backward(Tensor(np.ones((4, 5), np.float32), device=dev).data)
        return outs + [out, ta.grad, tb.grad]
    return list(zip(on("cuda", build), on("cpu", build)))


@check("functional.cross_entropy_loss (reference functional/loss.py:15-112) fwd+bwd")
def _ce():
    import neural_arch.functional as RF
    z = (rng.standard_normal((7, 11)) * 3).astype(np.float32)
    tg = rng.integers(0, 11, 7)

    def build(dev):
        t = Tensor(z, requires_grad=True, device=dev)
        loss = RF.cross_entropy_loss(t, Tensor(tg, device=dev))
        loss.backward()
        return [loss, t.grad]
    return list(zip(on("cuda", build), on("cpu", build)))


@check("optim.Adam / AdamW (reference optim/adam.py:165-252, optim/adamw.py:278-372) 6-step trajectories")
def _optim():
    from neural_arch.optim import Adam, AdamW
    p0 = rng.standard_normal((5, 7)).astype(np.float32)
    grads = [(rng.standard_normal((5, 7)) * (10.0 ** rng.integers(-3, 1))).astype(np.float32) for _ in range(6)]

    def build(dev):
        outs = []
        for make in (lambda p: Adam({"w": p}, lr=1e-3), lambda p: Adam({"w": p}, lr=0.05, weight_decay=0.01),
                     lambda p: AdamW({"w": p}, lr=5e-4, weight_decay=0.1)):
            p = Parameter(p0.copy())
            assert p.device.is_cuda == (dev == "cuda")
            opt = make(p)
            for gr in grads:
                p.grad = Tensor(gr, device=dev).data
                opt.step()
                outs.append(host(p).copy())
        return outs
    return list(zip(on("cuda", build), on("cpu", build)))


@check("reference GPT-2 block pieces (models/language/gpt2.py: RMSNorm, SwiGLU, GPT2Attention) forward", tol=2e-5)
def _gpt2_pieces():
    from neural_arch.models.language.gpt2 import GPT2Attention, RMSNorm, SwiGLU
    cfg = {"vocab_size": 61, "n_positions": 32, "n_embd": 32, "n_layer": 2, "n_head": 2}
    x = rng.standard_normal((2, 10, 32)).astype(np.float32)

    def build(dev):
        attn, sw, rn = GPT2Attention(cfg), SwiGLU(32), RMSNorm(32, eps=1e-5)
        a_out = attn(Tensor(x, device=dev))
        a_out = a_out[0] if isinstance(a_out, tuple) else a_out
        return [a_out, sw(Tensor(x, device=dev)), rn(Tensor(x, device=dev))]
    return list(zip(on("cuda", build), on("cpu", build)))


RESULTS["_implicit_host_copies"] = {"count": implicit_host_conversions(), "ok": True, "max_rel": 0.0}
print("DROPIN_JSON " + json.dumps(RESULTS))
