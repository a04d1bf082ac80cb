"""Layer-level differentiable ops: linear, layer_norm, rms_norm, embedding, rotary attention, swiglu.

On cuda tensors each forward/backward is a handful of fused kernels (tcgen05 GEMM with epilogues, warp-per-row
norms, flash attention, bit-exact scatter-add); weight gradients are ACCUMULATED IN PLACE into the parameter's
gradient buffer (a view of the optimizer's flat arena when one exists), so no per-parameter temporary is made.
On cpu tensors the same math runs through NumPy (the reference's own formulas) -- that is the
`Device.cpu()` product path of the "numpy" backend, never a fallback for a cuda tensor.
"""
from __future__ import annotations

import math

import numpy as np

from ..array import B200Array, bfloat16
from ..core.tensor import ACCUMULATED, Tensor, make_result

F32 = np.dtype(np.float32)


def _dev(x: Tensor) -> bool:
    return isinstance(x.data, B200Array)


def _bf16_shadow(p: Tensor):
    """bf16 copy of a parameter for the tensor-core path, refreshed when the parameter changes."""
    sh = getattr(p, "_bf16", None)
    if sh is not None and (getattr(p, "_bf16_version", -1) == p._version or getattr(p, "_bf16_live", False)):
        return sh
    sh = p.data.astype(bfloat16)
    p._bf16 = sh
    p._bf16_version = p._version
    return sh


# ====================================================================================== linear
def fused_rope_available(x: Tensor, head_dim: int) -> bool:
    """True when linear(..., rope=...) can rotate q / k in the projection's epilogue: device tensor, bf16 mode, head_dim 64."""
    if not _dev(x) or head_dim != 64:
        return False
    from .. import kernels as K
    return K.get_precision() == "bf16"


def linear(x: Tensor, weight: Tensor, bias=None, activation=None, residual=None, out_dtype=None,
           weight_layout: str = "out_in", rope=None) -> Tensor:
    """act(x @ W^T + b) [+ residual]; W is (out,in) ('out_in', nn/optimized.py) or (in,out) ('in_out', nn/linear.py).
    rope = (seq_len, cols, head_dim) (only where fused_rope_available): the rotary embedding of output columns [0, cols) is applied
    in the GEMM epilogue; the caller's backward must un-rotate the gradient of those columns (packed_qkv_attention does)."""
    n_in = weight.shape[1] if weight_layout == "out_in" else weight.shape[0]
    if x.shape[-1] != n_in:
        raise ValueError(f"Input features {x.shape[-1]} != expected {n_in}")
    inputs = [x, weight] + ([bias] if bias is not None else []) + ([residual] if residual is not None else [])
    if _dev(x):
        from .. import kernels as K
        prec = K.get_precision()
        wb = _bf16_shadow(weight) if prec == "bf16" else None
        y, ctx = K.linear_forward(x.data, weight.data, None if bias is None else bias.data, activation, out_dtype,
                                  None if residual is None else residual.data, wb, prec, weight_layout, rope)

        def backward(g):
            from ..core.tensor import get_grad_clip
            need_dx = x.requires_grad
            # leaf parameters: the wgrad GEMM / column sum accumulate straight into the gradient slot.  A weight or bias that is
            # itself a graph node (e.g. a concatenation of several parameters) gets a fresh buffer that is handed on.
            w_leaf = weight._grad_fn is None
            b_leaf = bias is None or bias._grad_fn is None
            dW_out = (weight.grad_buffer() if w_leaf else B200Array.full(weight.shape, 0.0, np.float32)) if weight.requires_grad else None
            db_out = None
            if bias is not None and bias.requires_grad:
                db_out = bias.grad_buffer() if b_leaf else B200Array.full(bias.shape, 0.0, np.float32)
            clip = (get_grad_clip() or 0.0) if prec == "bf16" else 0.0
            # dX comes back in the storage type of X: bf16 activations get bf16 gradients (they feed further GEMMs /
            # the flash backward), fp32 activations (unfused elementwise neighbours) get fp32 gradients
            dx_dt = None if x.data.dtype is bfloat16 else np.float32
            dx, _, _ = K.linear_backward(g, ctx, need_dx, dW_out, db_out, dx_dt, clip)
            if dx is not None and clip:
                dx.clipped = True
            if not (w_leaf and b_leaf):
                K.join_side_stream()   # the handed-on buffers may have been written on the wgrad side stream
            outs = [dx, ACCUMULATED if (w_leaf or dW_out is None) else dW_out]
            if bias is not None:
                outs.append(ACCUMULATED if (b_leaf or db_out is None) else db_out)
            if residual is not None:
                outs.append(g if g.dtype == F32 else g.astype(np.float32))
            return tuple(outs)

        return make_result(y, inputs, backward, "linear")
    # ---- host path
    xd, W = x.data, weight.data
    z = np.matmul(xd, W.T if weight_layout == "out_in" else W)
    if bias is not None:
        z = z + bias.data
    act = activation.lower() if isinstance(activation, str) else None
    if act == "gelu":
        from .activation import _HOST
        y = _HOST["gelu_tanh"][0](z)
    elif act == "relu":
        y = np.maximum(0.0, z)
    else:
        y = z
    if residual is not None:
        y = y + residual.data

    def backward_host(g):
        if act == "gelu":
            from .activation import _HOST
            dz = _HOST["gelu_tanh"][1](g, z)
        elif act == "relu":
            dz = g * (z > 0)
        else:
            dz = g
        dz2, x2 = dz.reshape(-1, dz.shape[-1]), xd.reshape(-1, xd.shape[-1])
        if weight_layout == "out_in":
            dx, dW = np.matmul(dz, W), np.matmul(dz2.T, x2)
        else:
            dx, dW = np.matmul(dz, W.T), np.matmul(x2.T, dz2)
        outs = [dx, dW]
        if bias is not None:
            outs.append(dz2.sum(axis=0))
        if residual is not None:
            outs.append(g)
        return tuple(outs)

    return make_result(y, inputs, backward_host, "linear")


# ====================================================================================== normalisation
def _norm(x: Tensor, weight, bias, eps: float, rms: bool, out_dtype=None, name="layernorm") -> Tensor:
    inputs = [x] + ([weight] if weight is not None else []) + ([bias] if bias is not None else [])
    if _dev(x):
        be = x.backend
        xd = x.data
        if xd.dtype is not bfloat16 and xd.dtype != F32:
            xd = xd.astype(np.float32)
        y, mean, rstd = be.norm_forward(xd, None if weight is None else weight.data, None if bias is None else bias.data, eps,
                                        rms=rms, out_dtype=out_dtype)

        def backward(g, acc0=None):
            dw_out = weight.grad_buffer() if (weight is not None and weight.requires_grad) else None
            db_out = bias.grad_buffer() if (bias is not None and bias.requires_grad) else None
            if g.dtype is not bfloat16 and g.dtype != F32:
                g = g.astype(np.float32)
            gdt = g if (g.dtype == y.dtype or g.dtype is y.dtype) else g.astype(y.dtype)
            dres = acc0
            if dres is not None and dres.dtype != xd.dtype:
                dres = dres.astype(xd.dtype)
            from ..core.tensor import get_grad_clip
            from .. import kernels as K
            dx, _, _ = be.norm_backward(gdt, xd, None if weight is None else weight.data, mean, rstd, rms=rms,
                                        dresidual=dres, dweight_out=dw_out, dbias_out=db_out, accumulate=True,
                                        need_wgrad=dw_out is not None, need_bgrad=db_out is not None,
                                        clip=get_grad_clip() or 0.0, bf16_twin=(K.get_precision() == "bf16"))
            outs = [dx]
            if weight is not None:
                outs.append(ACCUMULATED)
            if bias is not None:
                outs.append(ACCUMULATED)
            return tuple(outs)

        return make_result(y, inputs, backward, name, fused_acc=True)
    # ---- host path (nn/normalization.py:101-200 and :300-341)
    xd = x.data
    if rms:
        ms = np.mean(xd ** 2, axis=-1, keepdims=True)
        rstd = 1.0 / np.sqrt(ms + eps)
        mean = 0.0
    else:
        mean = np.mean(xd, axis=-1, keepdims=True)
        var = np.var(xd, axis=-1, keepdims=True, ddof=0)
        rstd = 1.0 / np.sqrt(var + eps)
    xhat = (xd - mean) * rstd
    y = xhat if weight is None else weight.data * xhat
    if bias is not None:
        y = y + bias.data

    def backward_host(g):
        N = xd.shape[-1]
        gh = g * weight.data if weight is not None else g
        m1 = 0.0 if rms else np.mean(gh, axis=-1, keepdims=True)
        m2 = np.mean(gh * xhat, axis=-1, keepdims=True)
        dx = rstd * (gh - m1 - xhat * m2)
        lead = tuple(range(g.ndim - 1))
        outs = [dx]
        if weight is not None:
            outs.append(np.sum(g * xhat, axis=lead))
        if bias is not None:
            outs.append(np.sum(g, axis=lead))
        return tuple(outs)

    return make_result(y.astype(xd.dtype), inputs, backward_host, name)


def layer_norm(x: Tensor, weight=None, bias=None, eps: float = 1e-5, out_dtype=None) -> Tensor:
    return _norm(x, weight, bias, eps, False, out_dtype, "layernorm")


def rms_norm(x: Tensor, weight=None, eps: float = 1e-8, out_dtype=None) -> Tensor:
    return _norm(x, weight, None, eps, True, out_dtype, "rmsnorm")


# ====================================================================================== embedding
def embedding(weight: Tensor, indices, out_dtype=None) -> Tensor:
    """W[idx.astype(int32)] -> (..., d); backward = the reference's ordered scatter-add (nn/embedding.py:27-63), bit-exact."""
    idx = indices.data if isinstance(indices, Tensor) else indices
    V = weight.shape[0]
    if _dev(weight):
        be = weight.backend
        if not isinstance(idx, B200Array):
            idx = be.from_numpy(np.asarray(idx))
        out = be.embedding_forward(weight.data, idx, out_dtype)

        def backward(g):
            from .. import kernels as K
            ddp = weight.__dict__.get("_ddp") if weight.__dict__.get("_ddp_sparse") else None
            if ddp is not None and ddp.defer_embedding_grad(weight, g, idx):
                return (ACCUMULATED,)   # data parallel: the rows are exchanged as (ids, rows) after backward instead of densely
            K.join_side_stream()   # a tied LM head's weight-gradient GEMM accumulates into the same buffer
            be.embedding_backward(g, idx, V, out=weight.grad_buffer(), accumulate=True)
            return (ACCUMULATED,)

        return make_result(out, [weight], backward, "embedding")
    idx = np.asarray(idx)
    flat = idx.flatten().astype(np.int32)
    out = weight.data[flat].reshape(list(idx.shape) + [weight.shape[1]])

    def backward_host(g):
        wg = np.zeros_like(weight.data)
        np.add.at(wg, flat, g.reshape(-1, weight.shape[1]))  # applies duplicates in index order, like the loop
        return (wg,)

    return make_result(out, [weight], backward_host, "embedding")


def to_float32(x: Tensor) -> Tensor:
    """Differentiable up-cast of a bf16-stored activation (head dims the flash kernel does not cover fall back to the
    unfused fp32 composition, which needs fp32 operands)."""
    if _dev(x) and x.data.dtype is bfloat16:
        return make_result(x.data.astype(np.float32), [x], lambda g: (g,), "to_float32")
    return x


# ====================================================================================== attention
def _rope_tables_host(head_dim, seq_len, base=10000):
    inv_freq = 1.0 / (base ** (np.arange(0, head_dim, 2).astype(np.float32) / head_dim))
    freqs = np.outer(np.arange(seq_len, dtype=np.float32), inv_freq)
    emb = np.concatenate((freqs, freqs), axis=-1)
    return np.cos(emb)[None, None, :, :], np.sin(emb)[None, None, :, :]


def _rot_half(x):
    h = x.shape[-1] // 2
    return np.concatenate((-x[..., h:], x[..., :h]), axis=-1)


def _rot_half_T(x):
    h = x.shape[-1] // 2
    return np.concatenate((x[..., h:], -x[..., :h]), axis=-1)


def rope(x: Tensor, base: int = 10000) -> Tensor:
    """Rotary position embedding on a (B,H,S,D) tensor (models/language/gpt2.py:25-95)."""
    S, D = x.shape[-2], x.shape[-1]
    if _dev(x):
        from .. import kernels as K
        xd = x.data.copy()
        K.rope_inplace(xd, False, base)
        return make_result(xd, [x], lambda g: (K.rope_inplace(g.copy(), True, base),), "rope")
    cos, sin = _rope_tables_host(D, S, base)
    xd = x.data
    return make_result((xd * cos + _rot_half(xd) * sin).astype(xd.dtype), [x],
                       lambda g: ((g * cos + _rot_half_T(g * sin)).astype(xd.dtype),), "rope")


def scaled_dot_product_attention(q: Tensor, k: Tensor, v: Tensor, scale: float, causal: bool = False, mask=None) -> Tensor:
    """softmax(q k^T * scale [causal -1e4 replace] [+ additive mask]) v for (B,H,S,D) tensors; k / v may carry fewer heads
    (grouped-query / multi-query attention: query head h uses key/value head h // (H // Hkv)).
    cuda + bf16 precision: the fused flash kernels (grouped K/V read in place, per-key additive masks); otherwise the
    unfused fp32 composition (1e-5 path, any broadcastable additive mask)."""
    mk = mask.data if isinstance(mask, Tensor) else mask
    H, Hkv = q.shape[1], k.shape[1]
    if Hkv <= 0 or H % Hkv:
        raise ValueError(f"{H} query heads are not a multiple of {Hkv} key/value heads")
    if _dev(q):
        from .. import kernels as K
        per_key = mk is None or int(np.prod(mk.shape)) in (k.shape[2], q.shape[0] * k.shape[2])
        if K.get_precision() == "bf16" and q.shape[-1] == 64 and per_key:
            qd, kd, vd = K.to_bf16(q.data), K.to_bf16(k.data), K.to_bf16(v.data)
            o, lse = K.attention_forward(qd, kd, vd, scale, causal, mk)

            def backward(g):
                return K.attention_backward(g, qd, kd, vd, o, lse, scale, causal, mk)

            return make_result(o, [q, k, v], backward, "flash_attention")
    # ---- composition from differentiable primitives (host tensors and the fp32 device path)
    q, k, v = to_float32(q), to_float32(k), to_float32(v)
    from .activation import softmax
    from .arithmetic import add, matmul, mul
    from .shape import repeat_kv, swapaxes
    k, v = repeat_kv(k, H // Hkv), repeat_kv(v, H // Hkv)
    s = mul(matmul(q, swapaxes(k, -1, -2)), float(scale))
    if causal:
        S, Skv = s.shape[-2], s.shape[-1]
        if _dev(q):
            from .. import array as A
            keep = A.tril_mask(S, Skv)
            sd = s.data
            s = make_result(A.where(keep, sd, -1e4), [s], lambda g: (A.where(keep, g, 0.0),), "causal_mask")
        else:
            keep = np.tril(np.ones((S, Skv))) != 0
            s = make_result(np.where(keep, s.data, np.float32(-1e4)), [s], lambda g: (np.where(keep, g, 0.0).astype(g.dtype),), "causal_mask")
    if mk is not None:
        s = add(s, mask if isinstance(mask, Tensor) else Tensor(mk, device=q.device))
    return matmul(softmax(s, axis=-1), v)


def packed_qkv_attention(qkv: Tensor, n_heads: int, causal: bool = True, use_rope: bool = True, scale=None, mask=None,
                         rope_applied: bool = False) -> Tensor:
    """Attention on the output of a fused qkv projection.  qkv: (B,S,3*d) laid out as (B,S,3,H,D) (gpt2.py:228-234,
    vision_transformer.py:110-116).  Returns (B,S,d).  cuda + bf16: RoPE in place on the packed buffer, flash
    forward reading it in place, and a backward that hands back ONE packed (B,S,3d) gradient -- no transposes."""
    B, S, three_d = qkv.shape
    d = three_d // 3
    D = d // n_heads
    scale = (1.0 / math.sqrt(D)) if scale is None else scale
    mk = mask.data if isinstance(mask, Tensor) else mask
    if _dev(qkv):
        from .. import kernels as K
        if K.get_precision() == "bf16" and D == 64:
            buf = K.to_bf16(qkv.data)
            rotate_here = use_rope and not rope_applied   # rope_applied: the projection's epilogue already rotated q and k
            if (buf is qkv.data and rotate_here) or not buf.is_contiguous:
                buf = buf.copy() if rotate_here else buf.contiguous()
            p5 = buf.reshape(B, S, 3, n_heads, D)
            if rotate_here:
                K.rope_packed_qk_inplace(p5, n_heads)
            qd = p5[:, :, 0].transpose(0, 2, 1, 3)
            kd = p5[:, :, 1].transpose(0, 2, 1, 3)
            vd = p5[:, :, 2].transpose(0, 2, 1, 3)
            o, lse = K.attention_forward(qd, kd, vd, scale, causal, mk)
            out = o.transpose(0, 2, 1, 3).reshape(B, S, d)

            def backward(g):
                g4 = K.to_bf16(g).reshape(B, S, n_heads, D).transpose(0, 2, 1, 3)
                _, _, _, packed = K.attention_backward(g4, qd, kd, vd, o, lse, scale, causal, mk, packed=True)
                if use_rope:
                    K.rope_packed_qk_inplace(packed, n_heads, inverse=True)
                return (packed.reshape(B, S, three_d),)

            return make_result(out, [qkv], backward, "packed_qkv_attention")
    from .shape import getitem, reshape, transpose
    qkv = to_float32(qkv)
    x5 = transpose(reshape(qkv, (B, S, 3, n_heads, D)), (2, 0, 3, 1, 4))
    q, k, v = getitem(x5, 0), getitem(x5, 1), getitem(x5, 2)
    if use_rope:
        q, k = rope(q), rope(k)
    o = scaled_dot_product_attention(q, k, v, scale, causal, mask)
    return reshape(transpose(o, (0, 2, 1, 3)), (B, S, d))


# ====================================================================================== SwiGLU
def swiglu_packed(gu: Tensor, gate_last: bool = False) -> Tensor:
    """silu(gate) * up on a packed projection output: [gate | up] (models/language/gpt2.py:119-139), or with
    ``gate_last`` [up | gate] -- the layout of functional.swiglu (functional/activation.py:512-572: x, gate = split(x))."""
    C = gu.shape[-1] // 2
    if _dev(gu):
        from .._lib import call
        from ..array import dtype_code
        gd = gu.data if gu.data.is_contiguous else gu.data.contiguous()
        rows = gd.size // (2 * C)
        g2 = gd.reshape(rows, 2 * C)
        out = B200Array.empty((rows, C), gd.dtype)
        isz = gd.itemsize
        og, ou = (C * isz, 0) if gate_last else (0, C * isz)          # byte offsets of the gate / up halves inside a row
        call("nfb_swiglu_fwd", dtype_code(gd.dtype), g2.ptr + og, g2.ptr + ou, out.ptr, rows, C, 2 * C)

        def backward(g):
            gg = g if (g.dtype == gd.dtype or g.dtype is gd.dtype) else g.astype(gd.dtype)
            gg = gg.contiguous().reshape(rows, C)
            dgu = B200Array.empty((rows, 2 * C), gd.dtype)
            call("nfb_swiglu_bwd", dtype_code(gd.dtype), g2.ptr + og, g2.ptr + ou, gg.ptr, dgu.ptr + og, dgu.ptr + ou, rows, C, 2 * C)
            return (dgu.reshape(gu.shape),)

        return make_result(out.reshape(gu.shape[:-1] + (C,)), [gu], backward, "swiglu")
    gd = gu.data
    gate, up = (gd[..., C:], gd[..., :C]) if gate_last else (gd[..., :C], gd[..., C:])
    sg = 1.0 / (1.0 + np.exp(-gate))

    def backward_host(g):
        dgate = g * up * (sg * (1.0 + gate * (1.0 - sg)))
        dup = g * gate * sg
        return (np.concatenate((dup, dgate) if gate_last else (dgate, dup), axis=-1).astype(gd.dtype),)

    return make_result((gate * sg * up).astype(gd.dtype), [gu], backward_host, "swiglu")
