"""B200Backend: the from-scratch B200 (sm_100a) implementation of the reference's Backend plugin.

Registered as ``"b200"`` and as ``"cuda"`` (the reference maps ``Device.cuda(i)`` to the backend
named "cuda", backends/utils.py:72-90, and its CuPy backend only registers when CuPy imports).
Every method launches hand-written CUDA kernels from libnfb200.so through ctypes; arrays are
:class:`nfb200.array.B200Array` objects living in HBM.  No CuPy, no Triton, no PyTorch, no CPU
fallback: if the library or a B200 is missing the backend reports ``is_available == False`` (so
``get_backend("b200")`` raises RuntimeError exactly like an unavailable reference backend) and any
direct use raises :class:`B200Error`.

Beyond the 54 abstract members it provides the optional hooks the reference's functional layer
probes with ``hasattr`` (relu/maximum/softmax/sigmoid/tanh/gelu, functional/activation.py:28-283),
the CuPy-precedent fused entry points (``fused_linear_gelu``, ``layernorm``, ``flash_attention``,
``benchmark_kernel``, ``has_custom_kernels``; backends/cuda_backend.py:303-386) and the fused
training kernels the nn/optim layers of this repo call (norm/linear/attention/cross-entropy/
embedding forward+backward, Adam/AdamW).
"""
from __future__ import annotations

import numbers

import numpy as np

from . import _lib
from . import array as A
from ._lib import call
from .array import B200Array, bfloat16, dtype_code
from . import kernels as K
from .plugin import Backend, register_backend


def _arr(x, dtype=None) -> B200Array:
    if isinstance(x, B200Array):
        return x
    return B200Array.from_numpy(np.asarray(x), dtype)


class B200Backend(Backend):
    float32, float64, int32, int64, bool = np.float32, np.float64, np.int32, np.int64, np.bool_
    bfloat16 = bfloat16

    # ------------------------------------------------------------------ identity
    @property
    def name(self) -> str:
        return "b200"

    @property
    def is_available(self) -> bool:
        return _lib.lib_available() and _lib.device_available()

    available = is_available

    @property
    def supports_gradients(self) -> bool:
        return True

    @property
    def has_custom_kernels(self) -> bool:
        return self.is_available

    def _dev(self) -> str:
        return f"cuda:{_lib.ensure_init()}"

    # ------------------------------------------------------------------ creation (float32 default)
    def array(self, data, dtype=None):
        if isinstance(data, B200Array):
            return data.astype(dtype) if dtype is not None else data.copy()
        return B200Array.from_numpy(np.array(data, dtype=None if dtype is bfloat16 else dtype), dtype if dtype is bfloat16 else None)

    def zeros(self, shape, dtype=None):
        return B200Array.full(shape, 0, dtype or np.float32)

    def ones(self, shape, dtype=None):
        return B200Array.full(shape, 1, dtype or np.float32)

    def full(self, shape, fill_value, dtype=None):
        return B200Array.full(shape, fill_value, dtype or np.float32)

    def empty(self, shape, dtype=None):
        return B200Array.empty(shape, dtype or np.float32)

    def arange(self, start, stop, step=1.0, dtype=None):
        return A.arange(start, stop, step, dtype or np.float32)

    # Weight init in the reference draws from the GLOBAL NumPy RNG (np.random.seed before model
    # construction); to stay bit-identical to the NumPy path the numbers are drawn on
    # the host with the same calls and copied once.  Initialisation is not on the hot path.
    def random_normal(self, shape, mean=0.0, std=1.0, dtype=None):
        return B200Array.from_numpy(np.random.normal(mean, std, shape).astype(np.float32))

    def random_uniform(self, shape, low=0.0, high=1.0, dtype=None):
        return B200Array.from_numpy(np.random.uniform(low, high, shape).astype(np.float32))

    # ------------------------------------------------------------------ shape
    def reshape(self, x, shape):
        return _arr(x).reshape(shape)

    def flatten(self, x):
        return _arr(x).flatten()

    def transpose(self, x, axes=None):
        return _arr(x).transpose(axes)

    def squeeze(self, x, axis=None):
        return _arr(x).squeeze(axis)

    def expand_dims(self, x, axis):
        return _arr(x).expand_dims(axis)

    # ------------------------------------------------------------------ math
    def _bin(self, op, x, y):
        if not isinstance(x, B200Array) and not isinstance(y, B200Array):
            x = _arr(x)
        return A.binary(op, x, y)

    def add(self, x, y):
        return self._bin("add", x, y)

    def subtract(self, x, y):
        return self._bin("sub", x, y)

    def multiply(self, x, y):
        return self._bin("mul", x, y)

    mul = multiply

    def divide(self, x, y):
        return self._bin("div", x, y)

    def power(self, x, y):
        return self._bin("pow", x, y)

    def maximum(self, x, y):
        return self._bin("max", x, y)

    def minimum(self, x, y):
        return self._bin("min", x, y)

    def matmul(self, x, y):
        return A.matmul(_arr(x), _arr(y))

    def dot(self, x, y):
        return A.dot(x if isinstance(x, numbers.Number) else _arr(x), y if isinstance(y, numbers.Number) else _arr(y))

    # ------------------------------------------------------------------ reductions
    def sum(self, x, axis=None, keepdims=False):
        return A.reduce("sum", _arr(x), axis, keepdims)

    def mean(self, x, axis=None, keepdims=False):
        return A.reduce("mean", _arr(x), axis, keepdims)

    def max(self, x, axis=None, keepdims=False):
        return A.reduce("max", _arr(x), axis, keepdims)

    def min(self, x, axis=None, keepdims=False):
        return A.reduce("min", _arr(x), axis, keepdims)

    def argmax(self, x, axis=None):
        return A.argreduce(True, _arr(x), axis)

    def argmin(self, x, axis=None):
        return A.argreduce(False, _arr(x), axis)

    # ------------------------------------------------------------------ unary
    def exp(self, x):
        return A.unary("exp", _arr(x))

    def log(self, x):
        return A.unary("log", _arr(x))

    def sqrt(self, x):
        return A.unary("sqrt", _arr(x))

    def abs(self, x):
        return A.unary("abs", _arr(x))

    def sign(self, x):
        return A.unary("sign", _arr(x))

    def clip(self, x, min_val, max_val):
        return _arr(x).clip(min_val, max_val)

    def tanh(self, x):
        return A.unary("tanh", _arr(x))

    def sigmoid(self, x):
        return A.unary("sigmoid", _arr(x))

    def relu(self, x):
        return A.unary("relu", _arr(x))

    def silu(self, x):
        return A.unary("silu", _arr(x))

    def gelu(self, x, approximate: bool = False):
        """Exact erf GELU by default (functional/activation.py:402-466); tanh form when approximate."""
        return A.unary("gelu_tanh" if approximate else "gelu_erf", _arr(x))

    def softmax(self, x, axis=-1):
        """max-subtracted softmax with the reference's 1e-8 floor on the denominator (activation.py:65-107)."""
        x = _arr(x)
        if x.dtype != np.dtype(np.float32):
            x = x.astype(np.float32)
        ax = axis + x.ndim if axis < 0 else axis
        moved = ax != x.ndim - 1
        xs = x.swapaxes(ax, -1) if moved else x
        xc = xs.contiguous()
        out = B200Array.empty(xc.shape, np.float32)
        if out.size:
            call("nfb_softmax_fwd", xc.ptr, out.ptr, xc.size // xc.shape[-1], xc.shape[-1])
        return out.swapaxes(ax, -1).contiguous() if moved else out

    def softmax_backward(self, p, g, axis=-1):
        p, g = _arr(p), _arr(g)
        ax = axis + p.ndim if axis < 0 else axis
        moved = ax != p.ndim - 1
        pc = (p.swapaxes(ax, -1) if moved else p).contiguous()
        gc = (g.swapaxes(ax, -1) if moved else g).contiguous()
        out = B200Array.empty(pc.shape, np.float32)
        if out.size:
            call("nfb_softmax_bwd", pc.ptr, gc.ptr, out.ptr, pc.size // pc.shape[-1], pc.shape[-1])
        return out.swapaxes(ax, -1).contiguous() if moved else out

    def unary_backward(self, op, src, grad):
        """grad * f'(.)  for op in relu/tanh/sigmoid/gelu_erf/gelu_tanh/silu (src = x, or y for tanh/sigmoid)."""
        return A.unary_bwd(op, _arr(src), _arr(grad))

    # ------------------------------------------------------------------ comparisons
    def equal(self, x, y):
        return self._bin("eq", x, y)

    def not_equal(self, x, y):
        return self._bin("ne", x, y)

    def less(self, x, y):
        return self._bin("lt", x, y)

    def less_equal(self, x, y):
        return self._bin("le", x, y)

    def greater(self, x, y):
        return self._bin("gt", x, y)

    def greater_equal(self, x, y):
        return self._bin("ge", x, y)

    # ------------------------------------------------------------------ manipulation
    def concatenate(self, arrays, axis=0):
        return A.concatenate([_arr(a) for a in arrays], axis)

    def stack(self, arrays, axis=0):
        return A.stack([_arr(a) for a in arrays], axis)

    def split(self, x, indices_or_sections, axis=0):
        return A.split(_arr(x), indices_or_sections, axis)

    # ------------------------------------------------------------------ conversion / device
    def astype(self, x, dtype):
        return _arr(x).astype(dtype)

    def to_numpy(self, x):
        if isinstance(x, B200Array):
            return x.get()
        return np.asarray(x)

    def from_numpy(self, x, dtype=None):
        return B200Array.from_numpy(x, dtype)

    def to_device(self, x, device):
        d = str(device)
        if not (d.startswith("cuda") or d.startswith("b200")):
            raise ValueError(f"B200 backend only supports cuda devices, got {device}")
        if ":" in d:
            want = int(d.split(":")[1])
            have = _lib.ensure_init()
            if want != have:
                raise ValueError(f"this process is bound to cuda:{have} (one process per GPU); cannot move to {device}")
        return _arr(x)

    def device_of(self, x):
        return self._dev()

    def is_array(self, x):
        return isinstance(x, B200Array)

    def shape(self, x):
        return tuple(x.shape)

    def size(self, x):
        return int(x.size)

    def dtype(self, x):
        return x.dtype

    # ------------------------------------------------------------------ advanced
    def where(self, condition, x, y):
        return A.where(condition, x, y)

    def einsum(self, equation, *operands):
        """One- and two-operand einsum lowered to permutes + the batched GEMM kernel."""
        eq = equation.replace(" ", "")
        if "..." in eq:
            raise NotImplementedError("b200 einsum: ellipsis is not supported")
        lhs, rhs = eq.split("->") if "->" in eq else (eq, None)
        terms = lhs.split(",")
        ops = [_arr(o) for o in operands]
        if len(terms) != len(ops):
            raise ValueError("einsum: operand count does not match the equation")
        if rhs is None:
            counts = {}
            for t in terms:
                for c in t:
                    counts[c] = counts.get(c, 0) + 1
            rhs = "".join(sorted(c for c, n in counts.items() if n == 1))
        for t, o in zip(terms, ops):
            if len(t) != o.ndim:
                raise ValueError(f"einsum: term '{t}' does not match an operand of rank {o.ndim}")
            if len(set(t)) != len(t):
                raise NotImplementedError("b200 einsum: repeated index inside one operand (diagonal) is not supported")
        if len(ops) == 1:
            t, x = terms[0], ops[0]
            drop = tuple(i for i, c in enumerate(t) if c not in rhs)
            if drop:
                x = A.reduce("sum", x, drop, False)
                t = "".join(c for c in t if c in rhs)
            return x.transpose([t.index(c) for c in rhs]).contiguous() if t != rhs else x
        if len(ops) != 2:
            raise NotImplementedError("b200 einsum: more than two operands")
        ta, tb = terms
        a, b = ops
        # sum out indices private to one operand and absent from the output
        for which in (0, 1):
            t, o = (ta, a) if which == 0 else (tb, b)
            other = tb if which == 0 else ta
            drop = tuple(i for i, c in enumerate(t) if c not in rhs and c not in other)
            if drop:
                o = A.reduce("sum", o, drop, False)
                t = "".join(c for i, c in enumerate(t) if i not in drop)
            if which == 0:
                ta, a = t, o
            else:
                tb, b = t, o
        batch = [c for c in ta if c in tb and c in rhs]
        contr = [c for c in ta if c in tb and c not in rhs]
        a_only = [c for c in ta if c not in tb]
        b_only = [c for c in tb if c not in ta]
        dims = {c: a.shape[ta.index(c)] for c in ta}
        dims.update({c: b.shape[tb.index(c)] for c in tb})
        ap = a.transpose([ta.index(c) for c in batch + a_only + contr])
        bp = b.transpose([tb.index(c) for c in batch + contr + b_only])
        nb = A._prod([dims[c] for c in batch])
        m = A._prod([dims[c] for c in a_only])
        k = A._prod([dims[c] for c in contr])
        n = A._prod([dims[c] for c in b_only])
        r = A.matmul(ap.contiguous().reshape(nb, m, k), bp.contiguous().reshape(nb, k, n))
        r = r.reshape([dims[c] for c in batch + a_only + b_only])
        cur = batch + a_only + b_only
        return r.transpose([cur.index(c) for c in rhs]).contiguous() if cur != list(rhs) else r

    def unique(self, x, return_counts=False):
        raise NotImplementedError("b200: unique() has no device kernel yet (it is not on the training path); "
                                  "copy to the host explicitly with to_numpy() if you need it")

    # ------------------------------------------------------------------ CuPy-precedent fused entry points
    def layernorm(self, x, weight, bias, eps: float = 1e-5):
        """(x-mean)/sqrt(var+eps)*weight+bias over the last dim (cuda_backend.py:334-352)."""
        y, _, _ = self.norm_forward(x, weight, bias, eps, rms=False)
        return y

    def fused_linear_gelu(self, x, weight, bias):
        """gelu_tanh(x @ weight.T + bias), weight (out,in) (cuda_backend.py:321-332)."""
        from . import kernels
        return kernels.linear_forward(_arr(x), _arr(weight), None if bias is None else _arr(bias), activation="gelu")[0]

    def flash_attention(self, q, k, v, scale, block_size: int = 64, causal: bool = False, mask=None):
        """softmax(q k^T * scale) v for (B,H,S,D) operands (cuda_backend.py:354-375), online softmax."""
        from . import kernels
        return kernels.attention_forward(_arr(q), _arr(k), _arr(v), float(scale), causal=causal, mask=mask)[0]

    def benchmark_kernel(self, fn, *args, num_runs: int = 100) -> float:
        """Mean device time in ms of ``fn(*args)`` measured with CUDA events on the compute stream."""
        from .runtime import time_ms
        return time_ms(lambda: fn(*args), iters=num_runs, warmup=3)

    # ------------------------------------------------------------------ fused training kernels
    def norm_forward(self, x, weight, bias, eps, rms=False, out_dtype=None, residual=None):
        """Returns (y, mean|None, rstd[, x+residual]).  nn/normalization.py:101-145 / :300-320."""
        x = _arr(x).contiguous()
        cols = x.shape[-1]
        rows = x.size // max(cols, 1)
        ydt = out_dtype or x.dtype
        y = B200Array.empty(x.shape, ydt)
        mean = None if rms else B200Array.empty((rows,), np.float32)
        rstd = B200Array.empty((rows,), np.float32)
        res = None if residual is None else _arr(residual).contiguous()
        summed = None if res is None else B200Array.empty(x.shape, x.dtype)
        w = None if weight is None else _arr(weight).contiguous()
        b = None if bias is None else _arr(bias).contiguous()
        call("nfb_norm_fwd", 1 if rms else 0, dtype_code(x.dtype), dtype_code(ydt), x.ptr, res.ptr if res is not None else None,
             summed.ptr if summed is not None else None, w.ptr if w is not None else None, b.ptr if b is not None else None,
             y.ptr, mean.ptr if mean is not None else None, rstd.ptr, rows, cols, float(eps))
        return (y, mean, rstd) if res is None else (y, mean, rstd, summed)

    def norm_backward(self, dy, x, weight, mean, rstd, rms=False, need_wgrad=True, need_bgrad=True, dresidual=None,
                      dweight_out=None, dbias_out=None, accumulate=False, clip=0.0, bf16_twin=False):
        """Returns (dx, dweight|None, dbias|None).  nn/normalization.py:147-200 / :300-341."""
        x = _arr(x).contiguous()
        dy = _arr(dy).contiguous()
        cols = x.shape[-1]
        rows = x.size // max(cols, 1)
        dx = B200Array.empty(x.shape, x.dtype)
        dw = dweight_out if dweight_out is not None else (B200Array.empty((cols,), np.float32) if need_wgrad and weight is not None else None)
        db = dbias_out if dbias_out is not None else (B200Array.empty((cols,), np.float32) if need_bgrad and not rms else None)
        w = None if weight is None else _arr(weight).contiguous()
        dres = None if dresidual is None else _arr(dresidual).contiguous()
        twin = B200Array.empty(x.shape, bfloat16) if (bf16_twin and x.dtype == np.dtype(np.float32)) else None
        call("nfb_norm_bwd", 1 if rms else 0, dtype_code(x.dtype), dtype_code(dy.dtype), dy.ptr, x.ptr, w.ptr if w is not None else None,
             mean.ptr if mean is not None else None, rstd.ptr, dres.ptr if dres is not None else None, dx.ptr,
             dw.ptr if dw is not None else None, db.ptr if db is not None else None, 1 if accumulate else 0, rows, cols,
             float(clip or 0.0), twin.ptr if twin is not None else None)
        if dw is not None or db is not None:
            K.mark_side_stream_dirty()   # with an auxiliary stream set, the parameter reduction ran there
        if clip:
            dx.clipped = True
        dx.twin = twin
        return dx, dw, db

    def cross_entropy(self, logits, targets, reduction="mean", need_grad=True, grad_dtype=None, ld=None):
        """Fused softmax + NLL (+ gradient).  Returns (loss, dlogits|None) where dlogits already carries the
        1/B factor of reduction='mean' (functional/loss.py:30-105)."""
        logits = _arr(logits)
        if logits.ndim != 2:
            raise ValueError("cross_entropy expects 2-D logits (batch, classes)")
        rows, V = logits.shape
        # bf16 logits (the bf16 tensor-core LM head) are consumed as they are when the wide-row kernel applies
        ldv0 = logits.strides[0] if rows > 1 else max(V, 1)
        keep_bf16 = (logits.dtype is bfloat16 and logits.strides[-1] == 1 and ldv0 >= V and ldv0 % 4 == 0 and 8192 <= V <= 13 * 4096)
        if logits.dtype != np.dtype(np.float32) and not keep_bf16:
            logits = logits.astype(np.float32)
        if logits.strides[-1] != 1 or (logits.strides[0] < V):
            logits = logits.contiguous()
        ldv = logits.strides[0] if rows > 1 else max(V, 1)
        t = _arr(targets)
        if t.ndim != 1:
            t = A.argreduce(True, t, 1)
        if t.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
            t = t.astype(np.int64)
        t = t.contiguous()
        if t.shape[0] != rows:
            raise ValueError(f"cross_entropy: {rows} logits rows but {t.shape[0]} targets")
        per_row = B200Array.empty((rows,), np.float32)
        dl = None
        dld = (V + 7) // 8 * 8  # 16-byte row pitch: the gradient feeds TMA-loaded GEMM operands without a repack
        if need_grad:
            dl = B200Array.empty((rows, dld), grad_dtype or np.float32)[:, :V]
        scale = 1.0 / rows if reduction == "mean" else 1.0
        call("nfb_cross_entropy_ex", dtype_code(logits.dtype), logits.ptr, t.ptr, 1 if t.dtype == np.dtype(np.int64) else 0, per_row.ptr,
             dl.ptr if dl is not None else None, dtype_code(dl.dtype) if dl is not None else 0, rows, V, ldv, dld, float(scale))
        if reduction == "mean":
            loss = A.reduce("mean", per_row, None, False)
        elif reduction == "sum":
            loss = A.reduce("sum", per_row, None, False)
        elif reduction == "none":
            loss = per_row
        else:
            raise ValueError(f"Unknown reduction: {reduction}")
        return loss, dl

    def embedding_forward(self, weight, indices, out_dtype=None):
        idx = _arr(indices)
        if idx.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
            idx = idx.astype(np.int32)  # the reference casts with astype(np.int32) (nn/embedding.py:34)
        return A.take_rows(_arr(weight), idx, out_dtype)

    def embedding_backward(self, grad, indices, vocab_size, out=None, accumulate=False):
        """Deterministic, NumPy-order scatter-add (nn/embedding.py:44-57).  Returns dW (vocab, d) fp32."""
        g = _arr(grad).contiguous()
        idx = _arr(indices)
        if idx.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
            idx = idx.astype(np.int32)
        idx = idx.contiguous()
        d = g.shape[-1]
        n = idx.size
        if out is None:
            out = B200Array.full((vocab_size, d), 0, np.float32)
            accumulate = False
        call("nfb_scatter_add_rows", dtype_code(idx.dtype), dtype_code(g.dtype), g.ptr, idx.ptr, out.ptr, n, d, vocab_size,
             1 if accumulate else 0)
        return out

    def colsum(self, x, out=None, accumulate=False):
        """out[c] (+)= sum over rows of a 2-D fp32/bf16 array (bias gradients), fp32 result."""
        x = _arr(x)
        if x.ndim != 2 or x.strides[1] != 1:
            x = x.reshape(-1, x.shape[-1]).contiguous()
        R, Cc = x.shape
        if out is None:
            out = B200Array.empty((Cc,), np.float32)
            accumulate = False
        call("nfb_colsum", dtype_code(x.dtype), x.ptr, R, Cc, x.strides[0] if R > 1 else Cc, out.ptr, 1 if accumulate else 0)
        return out

    def swiglu_forward(self, gate, up):
        g, u = _arr(gate), _arr(up)
        ld = g.strides[-2] if g.ndim >= 2 else g.shape[-1]
        Cc = g.shape[-1]
        rows = g.size // max(Cc, 1)
        if not (g.strides[-1] == 1 and u.strides[-1] == 1 and u.strides[-2:] == g.strides[-2:]) or g.ndim > 2 and not g.is_contiguous and ld == Cc:
            g, u, ld = g.contiguous(), u.contiguous(), Cc
        out = B200Array.empty(g.shape, g.dtype)
        call("nfb_swiglu_fwd", dtype_code(g.dtype), g.ptr, u.ptr, out.ptr, rows, Cc, ld)
        return out

    def swiglu_backward(self, gate, up, dout):
        g, u, d = _arr(gate).contiguous(), _arr(up).contiguous(), _arr(dout).contiguous()
        Cc = g.shape[-1]
        rows = g.size // max(Cc, 1)
        dg, du = B200Array.empty(g.shape, g.dtype), B200Array.empty(g.shape, g.dtype)
        call("nfb_swiglu_bwd", dtype_code(g.dtype), g.ptr, u.ptr, d.ptr, dg.ptr, du.ptr, rows, Cc, Cc)
        return dg, du


def make_reference_subclass(ref_backend_cls):
    """Build a class that is BOTH the reference's ``Backend`` (so ``isinstance`` checks upstream pass)
    and our implementation.  Used by plugin.install_into_reference()."""
    ns = {k: v for k, v in B200Backend.__dict__.items() if not k.startswith("__") and k != "_abc_impl"}
    return type("B200Backend", (ref_backend_cls,), ns)


register_backend("b200", B200Backend)
register_backend("cuda", B200Backend)
