"""CPU ORACLE (test infrastructure, NOT product code) -- NumPy restatement of the reference's per-op math.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product path (the ``nfb200`` package + libnfb200.so) never does.

Each function follows the cited lines of the reference (paths relative to
``/root/reference/src/neural_arch``) closely enough that ``oracle/validate_against_reference.py``
can check it against the reference's own modules, imported unmodified over the reconstructed
``neural_arch.core`` in ``oracle/compat``.  Pinning status is recorded in
``oracle/PINNED.md``.

Everything is forward + backward as *pure functions* (no autograd objects): ``*_fwd`` returns
``(out, cache)``, ``*_bwd(grad_out, cache)`` returns input gradients, composed by ``np_model.py`` in a
connected graph (decision 2).
"""
from __future__ import annotations

import math

import numpy as np

try:  # functional/activation.py:306-308, 403-405 prefers SciPy's erf
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

GRAD_CLIP = 10.0


def clip_grad(g, enabled=True):
    """GradientFunction.apply: nan_to_num then clip to [-10, 10] (tests/test_tensor.py:208-214)."""
    if not enabled:
        return g
    if not np.all(np.isfinite(g)):
        g = np.nan_to_num(g, nan=0.0, posinf=GRAD_CLIP, neginf=-GRAD_CLIP)
    return np.clip(g, -GRAD_CLIP, GRAD_CLIP)


def reduce_gradient(grad, shape):
    """functional/utils.py:33-96 -- sum a broadcast gradient back to the operand's shape."""
    grad = np.asarray(grad)
    shape = tuple(shape)
    if grad.shape == shape:
        return grad
    while grad.ndim > len(shape):
        grad = grad.sum(axis=0)
    for i, s in enumerate(shape):
        if s == 1 and grad.shape[i] != 1:
            grad = grad.sum(axis=i, keepdims=True)
    return grad.reshape(shape)


# ------------------------------------------------------------------ arithmetic (functional/arithmetic.py:14-438)
def add_bwd(g, a_shape, b_shape):
    return reduce_gradient(g, a_shape), reduce_gradient(g, b_shape)


def sub_bwd(g, a_shape, b_shape):
    return reduce_gradient(g, a_shape), reduce_gradient(-g, b_shape)


def mul_bwd(g, a, b):
    return reduce_gradient(g * b, a.shape), reduce_gradient(g * a, b.shape)


def div_bwd(g, a, b):
    return reduce_gradient(g / b, a.shape), reduce_gradient(-g * a / (b * b), b.shape)


def matmul_fwd(a, b):
    """functional/arithmetic.py:441-483 (np.matmul through numpy_backend.py:134-135)."""
    return np.matmul(a, b)


def matmul_bwd(g, a, b):
    """functional/arithmetic.py:488-570: dA = G . swap(B), dB = swap(A) . G, broadcast-reduced."""
    ga = np.matmul(g, np.swapaxes(b, -1, -2))
    gb = np.matmul(np.swapaxes(a, -1, -2), g)
    return reduce_gradient(ga, a.shape), reduce_gradient(gb, b.shape)


# ------------------------------------------------------------------ activations (functional/activation.py)
def relu_fwd(x):
    return np.maximum(0, x)  # :14-61


def relu_bwd(g, x):
    return g * (x > 0)


def sigmoid_fwd(x):
    """:151-214 -- the overflow-safe split form."""
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    e = np.exp(x[~pos])
    out[~pos] = e / (1.0 + e)
    return out


def sigmoid_bwd(g, y):
    return g * y * (1.0 - y)


def tanh_fwd(x):
    return np.tanh(x)  # :217-260


def tanh_bwd(g, y):
    return g * (1.0 - y * y)


def gelu_erf_fwd(x):
    """:402-466 exact: 0.5 x (1 + erf(x / sqrt 2))."""
    return (0.5 * x * (1.0 + _erf(x / np.sqrt(2.0)))).astype(x.dtype)


def gelu_erf_bwd(g, x):
    """:447-462 : 0.5 (1 + erf) + x exp(-x^2/2) / sqrt(2 pi)."""
    cdf = 0.5 * (1.0 + _erf(x / np.sqrt(2.0)))
    pdf = np.exp(-0.5 * x * x) / np.sqrt(2.0 * np.pi)
    return (g * (cdf + x * pdf)).astype(x.dtype)


def gelu_tanh_fwd(x):
    """:374-399 ; optimization/fusion.py:55-70 -- the form fused Linear+GELU uses (BERT/ViT MLP)."""
    c = np.sqrt(2.0 / np.pi)
    return (0.5 * x * (1.0 + np.tanh(c * (x + 0.044715 * x ** 3)))).astype(x.dtype)


def gelu_tanh_bwd(g, x):
    """nn/optimized.py:194-205."""
    c = np.sqrt(2.0 / np.pi)
    t = np.tanh(c * (x + 0.044715 * x ** 3))
    dinner = c * (1.0 + 3.0 * 0.044715 * x * x)
    return (g * (0.5 * (1.0 + t) + 0.5 * x * (1.0 - t * t) * dinner)).astype(x.dtype)


def silu_fwd(x):
    """models/language/gpt2.py:127-130 : x / (1 + exp(-x))."""
    return x * (1.0 / (1.0 + np.exp(-x)))


def silu_bwd(g, x):
    s = 1.0 / (1.0 + np.exp(-x))
    return g * (s * (1.0 + x * (1.0 - s)))


def softmax_fwd(x, axis=-1):
    """:65-107 : exp(x - max) / max(sum, 1e-8)."""
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))
    s = np.maximum(np.sum(e, axis=axis, keepdims=True), 1e-8)
    return e / s


def softmax_bwd(g, p, axis=-1):
    """:122-140 : p (g - sum(g p))."""
    return p * (g - np.sum(g * p, axis=axis, keepdims=True))


# ------------------------------------------------------------------ normalisation (nn/normalization.py)
def layernorm_fwd(x, gamma, beta, eps):
    """:101-117."""
    mean = np.mean(x, axis=-1, keepdims=True)
    var = np.var(x, axis=-1, keepdims=True, ddof=0)
    std = np.sqrt(var + eps)
    xhat = (x - mean) / std
    out = xhat
    if gamma is not None:
        out = gamma * xhat
        if beta is not None:
            out = out + beta
    return out, (x, mean, var, std, xhat, gamma, eps)


def layernorm_bwd(g, cache):
    """:147-200, term by term."""
    x, mean, var, std, xhat, gamma, eps = cache
    N = x.shape[-1]
    gn = g * gamma if gamma is not None else g
    gvar = np.sum(gn * (x - mean) * (-0.5) * (var + eps) ** (-1.5), axis=-1, keepdims=True)
    gmean = np.sum(gn * (-1.0 / std), axis=-1, keepdims=True) + gvar * np.sum(-2.0 * (x - mean), axis=-1, keepdims=True) / N
    dx = gn / std + gvar * 2.0 * (x - mean) / N + gmean / N
    lead = tuple(range(g.ndim - 1))
    dgamma = np.sum(g * xhat, axis=lead)
    dbeta = np.sum(g, axis=lead)
    return dx, dgamma, dbeta


def rmsnorm_fwd(x, weight, eps):
    """nn/normalization.py:300-312 and models/language/gpt2.py:106-116 (same forward)."""
    ms = np.mean(x ** 2, axis=-1, keepdims=True)
    rms = np.sqrt(ms + eps)
    normalized = x / rms
    out = weight * normalized if weight is not None else normalized
    return out, (x, ms, rms, normalized, weight, eps)


def rmsnorm_bwd(g, cache):
    """nn/normalization.py:322-341."""
    x, ms, rms, normalized, weight, eps = cache
    N = x.shape[-1]
    gn = g * weight if weight is not None else g
    rms_inv = 1.0 / rms
    grms = np.sum(gn * normalized * (-rms_inv), axis=-1, keepdims=True)
    gms = grms * 0.5 * (ms + eps) ** (-0.5)
    dx = gn * rms_inv + gms * 2.0 * x / N
    dw = np.sum(g * normalized, axis=tuple(range(g.ndim - 1)))
    return dx, dw


# ------------------------------------------------------------------ embedding (nn/embedding.py:27-63)
def embedding_fwd(weight, indices):
    flat = np.asarray(indices).flatten().astype(np.int32)
    return weight[flat].reshape(list(np.shape(indices)) + [weight.shape[1]])


def embedding_bwd(grad, indices, vocab_size):
    """The sequential loop of :52-55 (order matters for bit-exactness)."""
    d = grad.shape[-1]
    flat_idx = np.asarray(indices).flatten()
    flat_grad = grad.reshape(-1, d)
    wg = np.zeros((vocab_size, d), dtype=grad.dtype)
    for i, idx in enumerate(flat_idx):
        wg[idx] += flat_grad[i]
    return wg


def embedding_bwd_fast(grad, indices, vocab_size):
    """Same result as embedding_bwd, vectorised: np.add.at applies duplicates in index order."""
    d = grad.shape[-1]
    wg = np.zeros((vocab_size, d), dtype=grad.dtype)
    np.add.at(wg, np.asarray(indices).flatten(), grad.reshape(-1, d))
    return wg


# ------------------------------------------------------------------ loss (functional/loss.py:15-112)
def cross_entropy_fwd(logits, targets, reduction="mean"):
    probs = softmax_fwd(logits)
    t = np.asarray(targets)
    t = t.astype(int) if t.ndim == 1 else np.argmax(t, axis=1)
    B = logits.shape[0]
    tp = probs[np.arange(B), t]
    loss = -np.log(tp + 1e-8)
    if reduction == "mean":
        loss = np.mean(loss)
    elif reduction == "sum":
        loss = np.sum(loss)
    elif reduction != "none":
        raise ValueError(f"Unknown reduction: {reduction}")
    return loss, (probs, t, reduction)


def cross_entropy_bwd(g, cache):
    probs, t, reduction = cache
    B = probs.shape[0]
    d = probs.copy()
    d[np.arange(B), t] -= 1.0
    if reduction == "mean":
        d = d / B
    return d * g


# ------------------------------------------------------------------ linear (nn/optimized.py, nn/linear.py)
def linear_fwd(x, W, b, activation=None, weight_layout="out_in"):
    """OptimizedLinear: x @ W.T + b with W (out,in) (nn/optimized.py:299-317, _jit_forward :231-250);
    StandardLinear: x @ W + b with W (in,out) (nn/linear.py:202-205). Optional fused tanh-GELU / ReLU."""
    z = np.matmul(x, W.T if weight_layout == "out_in" else W)
    if b is not None:
        z = z + b
    if activation is None:
        y = z
    elif activation.lower() == "gelu":
        y = gelu_tanh_fwd(z)
    elif activation.lower() == "relu":
        y = np.maximum(0.0, z)
    else:
        y = z
    return y, (x, W, b, z, activation, weight_layout)


def linear_bwd(g, cache):
    """dX = dZ W ; dW = dZ^T X ; db = sum rows (nn/optimized.py:261-291 generalised to (...,in) inputs by
    flattening the leading dims -- SURVEY 3.2 notes the reference's own closure is 2-D only)."""
    x, W, b, z, activation, layout = cache
    if activation is None:
        dz = g
    elif activation.lower() == "gelu":
        dz = gelu_tanh_bwd(g, z)
    elif activation.lower() == "relu":
        dz = g * (z > 0).astype(z.dtype)
    else:
        dz = g
    dz2 = dz.reshape(-1, dz.shape[-1])
    x2 = x.reshape(-1, x.shape[-1])
    if layout == "out_in":
        dx = np.matmul(dz, W)
        dW = np.matmul(dz2.T, x2)
    else:
        dx = np.matmul(dz, W.T)
        dW = np.matmul(x2.T, dz2)
    db = dz2.sum(axis=0) if b is not None else None
    return dx, dW, db


# ------------------------------------------------------------------ attention cores
def rope_tables(head_dim, seq_len, base=10000):
    """models/language/gpt2.py:25-54."""
    inv_freq = 1.0 / (base ** (np.arange(0, head_dim, 2).astype(np.float32) / head_dim))
    t = np.arange(seq_len, dtype=np.float32)
    freqs = np.outer(t, inv_freq)
    emb = np.concatenate((freqs, freqs), axis=-1)
    return np.cos(emb)[None, None, :, :], np.sin(emb)[None, None, :, :]


def rotate_half(x):
    """gpt2.py:78-83."""
    h = x.shape[-1] // 2
    return np.concatenate((-x[..., h:], x[..., :h]), axis=-1)


def rope_fwd(x, cos, sin):
    """gpt2.py:86-95."""
    return x * cos + rotate_half(x) * sin


def rope_bwd(g, cos, sin):
    """Transpose of the rotation: d/dx [x cos + R(x) sin] = g cos + R^T(g sin), R^T(y) = [y2, -y1]."""
    gs = g * sin
    h = g.shape[-1] // 2
    return g * cos + np.concatenate((gs[..., h:], -gs[..., :h]), axis=-1)


def attention_fwd(q, k, v, scale, mode="none", mask=None):
    """(B,H,S,D) attention exactly as the model files spell it:
       gpt2  (models/language/gpt2.py:191-215): s = q k^T / sqrt(hd); s = where(tril == 0, -1e4, s); (+mask)
       bert  (models/language/bert.py:186-213): s = q k^T * (1/sqrt hd) + mask, mask = (1 - m)[:,None,None,:] * -1e4
       vit   (models/vision/vision_transformer.py:117-126): s = (q k^T) * hd^-0.5
    then softmax(axis=-1) and @ v."""
    s = np.matmul(q, np.swapaxes(k, -1, -2)) * scale
    if mode == "causal":
        S, Skv = s.shape[-2], s.shape[-1]
        tril = np.tril(np.ones((S, Skv)))
        s = np.where(tril == 0, -1e4, s)
    if mask is not None:
        s = s + mask
    p = softmax_fwd(s.astype(q.dtype), axis=-1)
    o = np.matmul(p, v)
    return o, (q, k, v, p, scale, mode)


def attention_bwd(go, cache):
    """Chain rule through matmul / softmax / mask-replace (masked positions get zero gradient)."""
    q, k, v, p, scale, mode = cache
    dv = np.matmul(np.swapaxes(p, -1, -2), go)
    dp = np.matmul(go, np.swapaxes(v, -1, -2))
    ds = softmax_bwd(dp, p, axis=-1)
    if mode == "causal":
        S, Skv = ds.shape[-2], ds.shape[-1]
        ds = np.where(np.tril(np.ones((S, Skv))) == 0, 0.0, ds)
    ds = ds * scale
    dq = np.matmul(ds, k)
    dk = np.matmul(np.swapaxes(ds, -1, -2), q)
    return dq.astype(q.dtype), dk.astype(q.dtype), dv.astype(q.dtype)


# ------------------------------------------------------------------ optimizers
class AdamState:
    def __init__(self, shape):
        self.step = 0
        self.exp_avg = np.zeros(shape, dtype=np.float64)
        self.exp_avg_sq = np.zeros(shape, dtype=np.float64)


def adam_step(p, g, st, lr=1e-3, betas=(0.9, 0.999), eps=1e-5, weight_decay=0.0, maximize=False):
    """optim/adam.py:165-252 (amsgrad off)."""
    b1, b2 = betas
    if maximize:
        g = -g
    if weight_decay != 0:
        g = g + weight_decay * p
    st.step += 1
    st.exp_avg = b1 * st.exp_avg + (1 - b1) * g.astype(np.float64)
    st.exp_avg_sq = b2 * st.exp_avg_sq + (1 - b2) * (g.astype(np.float64) ** 2)
    bc1, bc2 = 1 - b1 ** st.step, 1 - b2 ** st.step
    mhat, vhat = st.exp_avg / bc1, st.exp_avg_sq / bc2
    denom = np.sqrt(vhat + eps)
    eff = lr * min(5.0, 0.1 / lr) if (lr <= 0.02 and lr > 0) else lr
    upd = (eff * mhat / denom).astype(p.dtype)
    upd = np.clip(upd, -10.0, 10.0)
    newp = np.asarray(p - upd, dtype=p.dtype)
    if not np.all(np.isfinite(newp)):
        newp = np.nan_to_num(newp, nan=0.0, posinf=1.0, neginf=-1.0)
    return newp


def adamw_step(p, g, st, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, maximize=False):
    """optim/adamw.py:278-372 (amsgrad off), double bias-correction included."""
    b1, b2 = betas
    if maximize:
        g = -g
    st.step += 1
    st.exp_avg = b1 * st.exp_avg + (1 - b1) * g.astype(np.float64)
    st.exp_avg_sq = b2 * st.exp_avg_sq + (1 - b2) * (g.astype(np.float64) ** 2)
    bc1, bc2 = 1 - b1 ** st.step, 1 - b2 ** st.step
    mhat, vhat = st.exp_avg / bc1, st.exp_avg_sq / bc2
    denom = np.sqrt(vhat + eps)
    step_size = lr / bc1
    upd = (step_size * mhat / denom).astype(p.dtype)
    upd = np.clip(upd, -10.0, 10.0)
    if weight_decay != 0:
        newp = p * (1 - lr * weight_decay) - upd
    else:
        newp = p - upd
    if not np.all(np.isfinite(newp)):
        newp = np.nan_to_num(newp, nan=0.0, posinf=1.0, neginf=-1.0)
    return newp.astype(p.dtype)


def sgd_step(p, g, buf, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, maximize=False):
    """optim/sgd.py:87-115.  Returns (new_p, new_buf)."""
    if maximize:
        g = -g
    if weight_decay != 0:
        g = g + weight_decay * p
    if momentum != 0:
        buf = g.copy() if buf is None else momentum * buf + (1 - dampening) * g
        g = g + momentum * buf if nesterov else buf
    return (p - lr * g).astype(p.dtype), buf


def clip_grad_norm(grads, max_norm):
    """examples/translation/train_conversational.py:38-54: global L2 norm, scale by max_norm/(norm+1e-6) if above."""
    total = 0.0
    for g in grads:
        total += float(np.sum(g.astype(np.float64) ** 2))
    norm = math.sqrt(total)
    coef = max_norm / (norm + 1e-6)
    if coef < 1:
        grads = [g * np.float32(coef) for g in grads]
    return grads, norm


# ------------------------------------------------------------------ bf16 emulation (oracle of the tensor-core path)
# The reference has no 16-bit dtype: the B200 backend's bf16 mode is an internal compute mode whose
# rounding points are restated here so that "inherent bf16 noise" and "kernel error" can be told apart: the same fp32
# formulas as above, with operands / stored activations rounded to bfloat16 (round-to-nearest-even) exactly where the
# CUDA path stores bf16 (neural-network-from-scratch_b200/csrc: GEMM operands and bf16 outputs, RoPE in place, P and dS
# inside the flash kernels, bf16 logits / dlogits).
def bf16_round(a):
    """fp32 -> nearest bfloat16 (ties to even), returned as fp32 values."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = r.astype(np.uint32).view(np.float32).reshape(a.shape)
    return np.where(np.isfinite(a), out, a)


def attention_fwd_bf16(q, k, v, scale, mode="none", mask=None):
    """csrc/attention_tc.cu k_flash_fwd_tc: S = q k^T (bf16 operands, fp32 accumulate); softmax statistics from the
    UNROUNDED exponentials; P rounded to bf16 for the P.V product; O = (P_bf16 V) / l rounded to bf16; LSE in fp32."""
    s = np.matmul(q.astype(np.float32), np.swapaxes(k, -1, -2).astype(np.float32)) * np.float32(scale)
    if mode == "causal":
        S, Skv = s.shape[-2], s.shape[-1]
        s = np.where(np.tril(np.ones((S, Skv))) == 0, np.float32(-1e4), s)
    if mask is not None:
        s = s + mask
    s = s.astype(np.float32)
    m = s.max(axis=-1, keepdims=True)
    e = np.exp(s - m)
    l = np.maximum(e.sum(axis=-1, keepdims=True), np.float32(1e-8))
    o = bf16_round(np.matmul(bf16_round(e), v.astype(np.float32)) / l)
    lse = m + np.log(l)
    return o, (q, k, v, o, lse, scale, mode, mask)


def attention_bwd_bf16(go, cache, key_block=128, dq_f32_blocks=4):
    """csrc/attention_tc.cu k_flash_bwd_tc: P = exp(s - LSE) and dS = P_bf16 o (dP - delta) * scale are rounded to bf16
    before the dV / dK / dQ products; dQ is the sum of one bf16 partial per 128-key block (added in bf16 for <= 4 blocks,
    in fp32 beyond); dK, dV leave as bf16."""
    q, k, v, o, lse, scale, mode, mask = cache
    q32, k32, v32, go = q.astype(np.float32), k.astype(np.float32), v.astype(np.float32), go.astype(np.float32)
    s = np.matmul(q32, np.swapaxes(k32, -1, -2)) * np.float32(scale)
    S, Skv = s.shape[-2], s.shape[-1]
    keep = None
    if mode == "causal":
        keep = np.tril(np.ones((S, Skv))) != 0
        s = np.where(keep, s, np.float32(-1e4))
    if mask is not None:
        s = s + mask
    p = bf16_round(np.exp(s.astype(np.float32) - lse))
    dv = bf16_round(np.matmul(np.swapaxes(p, -1, -2), go))
    dp = np.matmul(go, np.swapaxes(v32, -1, -2))
    delta = np.sum(go * o.astype(np.float32), axis=-1, keepdims=True)
    ds = bf16_round(p * (dp - delta) * np.float32(scale))
    if keep is not None:
        ds = np.where(keep, ds, np.float32(0.0))
    dk = bf16_round(np.matmul(np.swapaxes(ds, -1, -2), q32))
    nblk = (Skv + key_block - 1) // key_block
    dq = np.zeros(q32.shape, np.float32)
    for j in range(nblk):
        part = np.matmul(ds[..., j * key_block:(j + 1) * key_block], k32[..., j * key_block:(j + 1) * key_block, :])
        if nblk <= dq_f32_blocks:
            dq = bf16_round(dq + bf16_round(part))
        else:
            dq = dq + part
    return bf16_round(dq), dk, dv
