"""Layer-level fused kernels (host wrappers): tensor-core GEMM with epilogues, Linear fwd/bwd,
RoPE, fused flash attention fwd/bwd.  These are what nn.Linear / attention blocks / the model zoo of
this repo call; each is a handful of launches of hand-written sm_100a kernels.

Precision modes (the reference has no 16-bit Tensor dtype, so bf16 is an internal
compute mode of the backend):
  * "fp32": SIMT FFMA contraction (csrc/gemm_simt.cu)          -- 1e-5 relative contract
  * "bf16": tcgen05/TMEM GEMM on bf16 operands, fp32 accumulate  -- 1e-2 contract, throughput path
"""
from __future__ import annotations


import os

import numpy as np

from . import _lib
from ._lib import B200Error, call
from .array import B200Array, bfloat16, dtype_code

_PRECISION = "fp32"
F32 = np.dtype(np.float32)
# (per-launch CUDA-event timing of every kernel class: _lib.PROFILE; bench.py)
GEMM_SPAN = None     # set to {"buf": device uint64 array of 2*cap words armed to all-ones, "cap": cap, "meta": []}: every tcgen05 GEMM
#                      launch i then stamps %globaltimer into words 2i (min after griddepcontrol.wait) and 2i+1 (~max at CTA exit)


def _arm_span(flops, shape):
    sp = GEMM_SPAN
    if sp is None or len(sp["meta"]) >= sp["cap"]:
        return
    call("nfb_gemm_bf16_set_span", sp["buf"].ptr + 16 * len(sp["meta"]))
    sp["meta"].append((flops, shape))


def set_precision(mode: str):
    """Select the contraction path used by Linear / attention: 'fp32' (default) or 'bf16'."""
    global _PRECISION
    if mode not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


def to_bf16(x: B200Array) -> B200Array:
    return x if x.dtype is bfloat16 else x.astype(bfloat16)


def _rowmajor_2d(x: B200Array):
    """-> (array, ld) for a 2-D array whose last dim has unit stride (copy otherwise)."""
    if x.ndim != 2:
        raise ValueError("expected a 2-D operand")
    if x.strides[1] == 1 and x.strides[0] >= x.shape[1] and (x.strides[0] % 8 == 0) and (x.ptr % 16 == 0):
        return x, x.strides[0]
    if x.shape[0] == 1 and x.strides[1] == 1 and x.shape[1] % 8 == 0 and x.ptr % 16 == 0:
        return x, x.shape[1]
    if x.shape[1] % 8 != 0:
        # pad the leading dimension to a multiple of 8 elements (TMA needs 16-byte row pitch)
        ld = (x.shape[1] + 7) // 8 * 8
        buf = B200Array.empty((x.shape[0], ld), x.dtype)
        buf[:, :x.shape[1]] = x
        return buf[:, :x.shape[1]], ld
    xc = x.contiguous()
    return xc, xc.shape[1]


def gemm_bf16(a: B200Array, b: B200Array, trans_a=False, trans_b=False, out=None, out_dtype=np.float32, bias=None,
              residual=None, epilogue=None, accumulate=False, split_k=0, clip=0.0, aux=False, ldc=None, rope=None):
    """out[M,N] (+)= epi(op(a) @ op(b) + bias) + residual on the tcgen05 path.

    a: (M,K) or, with trans_a, (K,M);  b: (K,N) or, with trans_b, (N,K).  bf16 operands (fp32 inputs are cast).
    Returns out (and the bf16 pre-activation when aux=True).
    """
    a, b = to_bf16(a), to_bf16(b)
    a2, lda = _rowmajor_2d(a)
    b2, ldb = _rowmajor_2d(b)
    M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
    K2, N = (b.shape[1], b.shape[0]) if trans_b else b.shape
    if K != K2:
        raise ValueError(f"gemm_bf16: inner dimensions differ ({K} vs {K2})")
    if out is None:
        if accumulate:
            raise ValueError("gemm_bf16: accumulate=True needs an existing `out`")
        odt = bfloat16 if (out_dtype is bfloat16 or out_dtype == "bfloat16") else F32
        if ldc is None and N % 8 != 0 and N >= 1024:
            ldc = (N + 7) // 8 * 8  # wide ragged outputs (logits): keep rows 16-byte aligned for the consumers
        if ldc is None or ldc == N:
            out = B200Array.empty((M, N), odt)
        else:
            out = B200Array.empty((M, ldc), odt)[:, :N]
    if out.shape != (M, N) or out.strides[1] != 1:
        raise ValueError(f"gemm_bf16: out has shape {out.shape} strides {out.strides}, expected ({M},{N}) row-major")
    ldc_v = out.strides[0] if M > 1 else max(N, out.strides[0])
    res_ptr, ldr = None, 0
    if residual is not None:
        if residual.dtype != F32 or residual.shape != (M, N) or residual.strides[1] != 1:
            raise ValueError("gemm_bf16: residual must be an fp32 (M,N) row-major array")
        res_ptr, ldr = residual.ptr, residual.strides[0]
    bias_ptr = None
    if bias is not None:
        if bias.dtype != F32 or bias.size != N:
            raise ValueError("gemm_bf16: bias must be fp32 of length N")
        bias = bias.contiguous()
        bias_ptr = bias.ptr
    aux_arr = B200Array.empty((M, N), bfloat16) if aux else None
    epi = {None: 0, "none": 0, "gelu": 1, "relu": 2}[epilogue.lower() if isinstance(epilogue, str) else epilogue]
    if rope is not None:
        # (seq_len, cols, head_dim): rotate output columns [0, cols) -- heads of 64 -- by the position row % seq_len in the epilogue
        S_r, cols_r, D_r = rope
        if D_r != 64 or out.dtype is not bfloat16:
            raise ValueError("gemm_bf16: the fused rotary epilogue needs head_dim 64 and a bf16 output")
        cos_t, sin_t = rope_tables(D_r, S_r)
        call("nfb_gemm_bf16_set_rope", cos_t.ptr, sin_t.ptr, int(S_r), int(cols_r))
    _arm_span(2.0 * M * N * K, (M, N, K, bool(trans_a), bool(trans_b)))
    call("nfb_gemm_bf16", 1 if trans_a else 0, 1 if trans_b else 0, M, N, K, a2.ptr, lda, b2.ptr, ldb, out.ptr,
         dtype_code(out.dtype), ldc_v, bias_ptr, res_ptr, ldr, aux_arr.ptr if aux_arr is not None else None, N, epi,
         1 if accumulate else 0, int(split_k), float(clip))
    return (out, aux_arr) if aux else out


def gemm_bf16_grouped(problems, trans_a=False, trans_b=False, accumulate=False):
    """Up to 4 independent contractions out_g (+)= op(a_g) @ op(b_g) (same layouts, fp32 or bf16 outputs of one type) in ONE
    persistent stream-K launch of the tcgen05 kernel.  problems: [(a, b, out), ...] with `out` preallocated."""
    import ctypes as C
    n = len(problems)
    if n == 0:
        return
    Ms, Ns, Ks, As, ldas, Bs, ldbs, Cs, ldcs = [], [], [], [], [], [], [], [], []
    keep = []
    odt = None
    flops = 0.0
    for a, b, out in problems:
        a, b = to_bf16(a), to_bf16(b)
        a2, lda = _rowmajor_2d(a)
        b2, ldb = _rowmajor_2d(b)
        M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
        K2, N = (b.shape[1], b.shape[0]) if trans_b else b.shape
        if K != K2:
            raise ValueError(f"gemm_bf16_grouped: inner dimensions differ ({K} vs {K2})")
        if out.shape != (M, N) or out.strides[1] != 1:
            raise ValueError(f"gemm_bf16_grouped: out has shape {out.shape} strides {out.strides}, expected ({M},{N}) row-major")
        if odt is None:
            odt = out.dtype
        elif out.dtype is not odt and out.dtype != odt:
            raise ValueError("gemm_bf16_grouped: outputs must share one dtype")
        keep.append((a2, b2))
        Ms.append(M); Ns.append(N); Ks.append(K)
        As.append(a2.ptr); ldas.append(lda); Bs.append(b2.ptr); ldbs.append(ldb)
        Cs.append(out.ptr); ldcs.append(out.strides[0] if M > 1 else max(N, out.strides[0]))
        flops += 2.0 * M * N * K
    ia, vpa, la = (C.c_int * n), (C.c_void_p * n), (C.c_int64 * n)
    _arm_span(flops, ("grouped", tuple(zip(Ms, Ns, Ks)), bool(trans_a), bool(trans_b)))
    call("nfb_gemm_bf16_grouped", n, 1 if trans_a else 0, 1 if trans_b else 0, ia(*Ms), ia(*Ns), ia(*Ks), vpa(*As), la(*ldas), vpa(*Bs), la(*ldbs),
         vpa(*Cs), la(*ldcs), dtype_code(odt), 1 if accumulate else 0)


# fp32 contractions on the tensor cores (three-way bf16 split, six-fold K: csrc/gemm_simt.cu nfb_split_bf16x3).  Same 1e-5
# contract as the FFMA kernel (O(2^-24) per product term, fp32 accumulation), ~3x its throughput on GPT-2-small layer shapes.
_F32_TC = os.environ.get("NFB200_FP32_TENSOR_CORE", "1") != "0"
_F32_TC_MIN_MNK = 1 << 24      # below this the two split passes cost more than the FFMA kernel


def set_fp32_tensor_core(on: bool):
    """fp32-mode 2-D contractions through the bf16x3 split on tcgen05 (True, default) or the FFMA kernel (False)."""
    global _F32_TC
    _F32_TC = bool(on)


_F32_TC_KCHUNK = 4096          # the pipe's truncating fp32 accumulation grows with K (3e-7 of |A||B| at 768, 1e-6 at 8192): longer
                               # contractions go in chunks whose results meet in round-to-nearest fp32 adds


def _gemm_f32_split(a2, lda, b2, ldb, trans_a, trans_b, M, N, K, out, bias, accumulate):
    # A' : (M, 6K) [K along the columns] or, with trans_a, (6K, M);  B' : (6K, N) or, with trans_b, (N, 6K)
    for k0 in range(0, K, _F32_TC_KCHUNK):
        kc = min(_F32_TC_KCHUNK, K - k0)
        ac = a2[k0:k0 + kc, :] if trans_a else a2[:, k0:k0 + kc]
        bc = b2[:, k0:k0 + kc] if trans_b else b2[k0:k0 + kc, :]
        ra, ca = ac.shape
        rb, cb = bc.shape
        ap = B200Array.empty((6 * ra, ca) if trans_a else (ra, 6 * ca), bfloat16)
        bp = B200Array.empty((rb, 6 * cb) if trans_b else (6 * rb, cb), bfloat16)
        call("nfb_split_bf16x3", ac.ptr, ra, ca, lda, 0, 0 if trans_a else 1, ap.ptr)
        call("nfb_split_bf16x3", bc.ptr, rb, cb, ldb, 1, 1 if trans_b else 0, bp.ptr)
        # split_k = 1: one pass over K' per output tile, no racing reduce-adds -> run-to-run deterministic like the FFMA kernel
        gemm_bf16(ap, bp, trans_a, trans_b, out=out, out_dtype=np.float32, bias=bias if k0 == 0 else None,
                  accumulate=accumulate or k0 > 0, split_k=1)
    return out


def gemm_f32(a: B200Array, b: B200Array, trans_a=False, trans_b=False, out=None, bias=None, beta=0.0):
    """fp32 contraction with the same calling convention as gemm_bf16 (2-D operands): the bf16x3 split on the tensor cores for
    large aligned problems, the FFMA kernel otherwise."""
    def prep(x):
        if x.dtype != F32:
            x = x.astype(np.float32)
        if x.strides[1] == 1 and x.strides[0] >= x.shape[1]:
            return x, x.strides[0]
        xc = x.contiguous()
        return xc, xc.shape[1]
    a2, lda = prep(a)
    b2, ldb = prep(b)
    M, K = (a.shape[1], a.shape[0]) if trans_a else a.shape
    K2, N = (b.shape[1], b.shape[0]) if trans_b else b.shape
    if K != K2:
        raise ValueError(f"gemm_f32: inner dimensions differ ({K} vs {K2})")
    if out is None:
        out = B200Array.empty((M, N), np.float32)
        beta = 0.0
    bias_c = bias.contiguous() if bias is not None else None
    if (_F32_TC and beta in (0.0, 1.0) and float(M) * N * K >= _F32_TC_MIN_MNK and M % 8 == 0 and N % 8 == 0 and K % 8 == 0
            and lda % 4 == 0 and ldb % 4 == 0 and a2.ptr % 16 == 0 and b2.ptr % 16 == 0 and out.dtype == F32 and out.strides[1] == 1
            and (bias_c is None or bias_c.dtype == F32)):
        return _gemm_f32_split(a2, lda, b2, ldb, trans_a, trans_b, M, N, K, out, bias_c, beta == 1.0)
    call("nfb_gemm_f32", 1 if trans_a else 0, 1 if trans_b else 0, M, N, K, 1.0, a2.ptr, lda, 0, b2.ptr, ldb, 0, float(beta),
         out.ptr, out.strides[0] if M > 1 else N, 0, 1, bias_c.ptr if bias_c is not None else None)
    return out


# ======================================================================================
# Linear (nn/optimized.py:149-317, nn/linear.py:202-205): y = act(x @ W^T + b) [+ residual]
# ======================================================================================
def linear_forward(x: B200Array, W: B200Array, b=None, activation=None, out_dtype=None, residual=None, w_bf16=None,
                   precision=None, weight_layout="out_in", rope=None):
    """x (..., in); W (out,in) [or (in,out) with weight_layout='in_out'].  Returns (y, ctx) where ctx is what
    linear_backward needs.  precision None -> the module-level mode."""
    precision = precision or _PRECISION
    lead = x.shape[:-1]
    k_in = x.shape[-1]
    x2 = x.reshape(-1, k_in)
    n_out = W.shape[0] if weight_layout == "out_in" else W.shape[1]
    tb = weight_layout == "out_in"
    act = activation.lower() if isinstance(activation, str) else None
    if act not in (None, "gelu", "relu"):
        act = None
    res2 = None if residual is None else residual.reshape(-1, n_out)
    if precision == "bf16":
        xb = to_bf16(x2)
        wb = w_bf16 if w_bf16 is not None else to_bf16(W)
        # the residual stream is fp32: a GEMM that folds the skip connection into its epilogue writes fp32
        odt = out_dtype if out_dtype is not None else (F32 if (x.dtype == F32 or residual is not None) else bfloat16)
        if act == "gelu":
            y, z = gemm_bf16(xb, wb, False, tb, out_dtype=odt, bias=b, residual=res2, epilogue="gelu", aux=True)
        else:
            y = gemm_bf16(xb, wb, False, tb, out_dtype=odt, bias=b, residual=res2, epilogue=act, rope=rope)
            z = y if act == "relu" else None  # relu'(z) == (y > 0)
        ctx = (xb, W, wb, z, act, weight_layout, b is not None, "bf16")
    else:
        xf = x2 if x2.dtype == F32 else x2.astype(np.float32)
        z = gemm_f32(xf, W, False, tb, bias=b)
        if act == "gelu":
            from . import array as A
            y = A.unary("gelu_tanh", z)
        elif act == "relu":
            from . import array as A
            y = A.unary("relu", z)
        else:
            y = z
        if res2 is not None:
            y = y + res2
        ctx = (xf, W, None, z if act else None, act, weight_layout, b is not None, "fp32")
    return y.reshape(lead + (n_out,)), ctx


# ---- weight-gradient GEMMs on a side stream ------------------------------------------------------------------------
# dW (+)= dZ^T X depends on nothing downstream until the optimizer, while the layer GEMMs of GPT-2-small-sized models fill
# only 0.65-2 waves of the 148 SMs.  With overlap on, every wgrad GEMM of the bf16 path is issued on a second stream
# (a fork/join branch of the captured CUDA graph): its CTAs fill the SMs the dgrad / norm / attention kernels of the main
# chain leave idle.  Operands are kept referenced until the join so the stream-ordered allocator cannot recycle them.
_WGRAD_OVERLAP = False
_side = None
_side_dirty = False
_held = []


# ---- grouped weight gradients ------------------------------------------------------------------------------------------
# dW_g += dZ_g^T X_g of consecutive Linear layers are independent of each other and of everything before the optimizer, so
# they are QUEUED (with the bias column sums that go with them) and run four at a time as one grouped stream-K launch
# (nfb_gemm_bf16_grouped): the four weight gradients of a GPT-2-small block are 90 tiles x 64 k-blocks, which one launch
# cuts into 74 equal ranges, where four separate launches each split K ~8 ways and pushed 8x their output through L2
# reduce-adds.  The queue is flushed when it holds WGRAD_GROUP contractions, before anything else is issued on the side
# stream, before a DDP bucket is handed to NCCL and at the end of backward() (join_side_stream).
WGRAD_GROUP = 4
GEMM_WGRAD_DEFER = True      # False: every weight gradient is launched where linear_backward computes it
_WGRAD_BIG_FLOP = 1.0e11     # LM-head-sized problems keep their own launch (tile order chosen for L2 reuse)
_pending_wgrad = []          # (a, b, out): out += a^T b
_pending_colsum = []         # (dz, db_out)


def set_wgrad_group(n: int):
    """How many queued weight-gradient contractions go into one grouped launch (1 = launch each on its own)."""
    global WGRAD_GROUP
    flush_wgrad()
    WGRAD_GROUP = max(1, min(4, int(n)))


def _issue_wgrad(items, colsums):
    """Every weight gradient is DETERMINISTIC: the grouped stream-K launch adds each finished tile to the gradient buffer
    exactly once (partial tiles meet in a workspace in a fixed order), and the LM-head-sized ones run with split_k = 1
    (one reduce-add per element).  No launch splits K over racing reduce-adds."""
    big = [it for it in items if 2.0 * it[0].shape[1] * it[1].shape[1] * it[0].shape[0] >= _WGRAD_BIG_FLOP]
    small = [it for it in items if not any(it is b for b in big)]
    if small:
        gemm_bf16_grouped(small, True, False, accumulate=True)
    for a, b, out in big:
        gemm_bf16(a, b, True, False, out=out, accumulate=True, split_k=1)
    for dz, db_out in colsums:
        _colsum(dz, db_out)


def flush_wgrad():
    """Launch the queued weight-gradient contractions / bias sums (side stream when the overlap is on)."""
    global _pending_wgrad, _pending_colsum
    if not _pending_wgrad and not _pending_colsum:
        return
    items, colsums = _pending_wgrad, _pending_colsum
    _pending_wgrad, _pending_colsum = [], []
    if _WGRAD_OVERLAP:
        keep = [x for it in items for x in it] + [x for it in colsums for x in it]
        run_on_side_stream(lambda: _issue_wgrad(items, colsums), *keep, _flush=False)
    else:
        _issue_wgrad(items, colsums)


def queue_wgrad(a: B200Array, b: B200Array, out: B200Array, colsum=None):
    """out += a^T b (a (T,M), b (T,N) bf16), deferred; colsum = (dz, db_out) rides along."""
    _pending_wgrad.append((a, b, out))
    if colsum is not None:
        _pending_colsum.append(colsum)
    if len(_pending_wgrad) >= WGRAD_GROUP:
        flush_wgrad()


def set_wgrad_overlap(on: bool):
    """Weight / bias / norm-parameter gradients (everything only the optimizer consumes) on a side stream."""
    global _WGRAD_OVERLAP, _side
    join_side_stream()
    _WGRAD_OVERLAP = bool(on)
    if _WGRAD_OVERLAP and _side is None:
        from .runtime import Stream
        _side = Stream()
    # the C side moves the norm-backward parameter reduction to this stream (csrc/norm.cu)
    call("nfb_set_aux_stream", _side.handle if _WGRAD_OVERLAP else None)


def mark_side_stream_dirty():
    """The library issued optimizer-only work on the auxiliary stream: the next join must wait for it."""
    global _side_dirty
    if _WGRAD_OVERLAP:
        _side_dirty = True


def join_side_stream():
    """Make the compute stream wait for every weight-gradient GEMM issued on the side stream (no-op when none is pending).
    Called at the end of backward(), before a DDP bucket is all-reduced and before the embedding scatter-add."""
    global _side_dirty
    flush_wgrad()
    if _side_dirty:
        call("nfb_norm_flush")   # queued norm-parameter reductions (csrc/norm.cu) go out before the fence
        from .runtime import Event, compute_stream_wait
        compute_stream_wait(Event(timing=False).record(_side))
        _held.clear()
        _side_dirty = False


def side_stream_event():
    """An event recorded on the side stream behind everything issued there so far, or None when nothing is pending.
    Lets ANOTHER stream (the DDP communication stream) wait for the weight gradients without stalling the compute stream."""
    flush_wgrad()
    if not _side_dirty:
        return None
    call("nfb_norm_flush")
    from .runtime import Event
    return Event(timing=False).record(_side)


def run_on_side_stream(fn, *keep, _flush=True):
    """Issue `fn`'s kernels on the side stream, ordered after everything enqueued on the compute stream so far; `keep`
    is held until the next join so the stream-ordered allocator cannot recycle those buffers."""
    global _side, _side_dirty
    from .runtime import Event, Stream, current_stream_handle
    if _flush:
        flush_wgrad()   # queued weight gradients precede whatever consumes them on this stream
        call("nfb_norm_flush")   # ... and so do the queued norm-parameter reductions (csrc/norm.cu)
    if _side is None:
        _side = Stream()
    _side.wait_event(Event(timing=False).record())   # operands are final on the compute stream
    main = current_stream_handle()
    call("nfb_set_stream", _side.handle)
    try:
        fn()
    finally:
        call("nfb_set_stream", main)
    _held.append(keep)
    _side_dirty = True


def linear_backward(dy: B200Array, ctx, need_dx=True, dW_out=None, db_out=None, dx_dtype=None, clip=0.0):
    """Returns (dx, dW, db).  dW / db are ACCUMULATED into dW_out / db_out when given (gradient arena views),
    otherwise freshly allocated.  dX = dZ W ; dW = dZ^T X ; db = column sums of dZ."""
    from . import array as A
    x2, W, wb, z, act, layout, has_bias, precision = ctx
    lead = dy.shape[:-1]
    n_out = dy.shape[-1]
    twin = getattr(dy, "twin", None)
    dz = dy.reshape(-1, n_out)
    out_in = layout == "out_in"
    if precision == "bf16":
        dz = twin.reshape(-1, n_out) if (twin is not None and dy.dtype != bfloat16) else to_bf16(dz)
        if act == "gelu":
            dz = A.unary_bwd("gelu_tanh", z.reshape(dz.shape), dz)
        elif act == "relu":
            dz = A.unary_bwd("relu", to_bf16(z.reshape(dz.shape)), dz)
        dx = None
        if need_dx:
            odt = dx_dtype if dx_dtype is not None else bfloat16
            # dX (T,in) = dZ (T,out) . W(out,in): B operand is W stored (K=out, N=in) -> trans_b=False
            dx = gemm_bf16(dz, wb, False, not out_in, out_dtype=odt, clip=clip)
        side = _WGRAD_OVERLAP and dW_out is not None
        if dW_out is not None and WGRAD_GROUP > 1 and GEMM_WGRAD_DEFER:
            # gradient-arena target: nothing reads it before the optimizer / the bucket all-reduce -> queue for a grouped launch
            pair = (dz, x2) if out_in else (x2, dz)
            queue_wgrad(pair[0], pair[1], dW_out, (dz, db_out) if (has_bias and db_out is not None) else None)
            db = db_out if has_bias else None
            if has_bias and db_out is None:
                db = _colsum(dz, None)
            return (dx.reshape(lead + (x2.shape[1],)) if dx is not None else None), dW_out, db
        if dW_out is None:
            dW_out = B200Array.full(W.shape, 0.0, np.float32)
        if out_in:   # dW (out,in) = dZ^T (out,T) . X (T,in)
            wg = lambda: _issue_wgrad([(dz, x2, dW_out)], [])
        else:        # dW (in,out) = X^T . dZ
            wg = lambda: _issue_wgrad([(x2, dz, dW_out)], [])
        db = None
        if side and has_bias and db_out is not None:
            # the bias gradient (column sums of dZ) feeds only the optimizer too: same side stream
            def wg_and_bias(wg=wg):
                wg()
                _colsum(dz, db_out)
            run_on_side_stream(wg_and_bias, dz, x2, dW_out, db_out)
            db = db_out
        elif side:
            run_on_side_stream(wg, dz, x2, dW_out)
        else:
            wg()
        if has_bias and db is None:
            db = _colsum(dz, db_out)
        return (dx.reshape(lead + (x2.shape[1],)) if dx is not None else None), dW_out, db
    # ---- fp32 path
    if dz.dtype != F32:
        dz = dz.astype(np.float32)
    dz = dz.contiguous()
    if act == "gelu":
        dz = A.unary_bwd("gelu_tanh", z.reshape(dz.shape), dz)
    elif act == "relu":
        dz = A.unary_bwd("relu", z.reshape(dz.shape), dz)
    dx = None
    if need_dx:
        dx = gemm_f32(dz, W, False, not out_in)
        if clip > 0:
            dx = dx.clip(-clip, clip)
        dx = dx.reshape(lead + (x2.shape[1],))
    if out_in:
        dW = gemm_f32(dz, x2, True, False, out=dW_out, beta=1.0 if dW_out is not None else 0.0)
    else:
        dW = gemm_f32(x2, dz, True, False, out=dW_out, beta=1.0 if dW_out is not None else 0.0)
    db = None
    if has_bias:
        db = _colsum(dz, db_out)
    return dx, dW, db


def _colsum(dz: B200Array, out=None) -> B200Array:
    """Bias gradient: column sums of dZ accumulated straight into the gradient buffer (one 2-stage kernel)."""
    R, Cc = dz.shape
    acc = out is not None
    if out is None:
        out = B200Array.empty((Cc,), np.float32)
    call("nfb_colsum", dtype_code(dz.dtype), dz.ptr, R, Cc, dz.strides[0] if R > 1 else Cc, out.ptr, 1 if acc else 0)
    return out


# ======================================================================================
# RoPE + fused attention
# ======================================================================================
_ROPE_CACHE = {}


def rope_tables(head_dim: int, seq_len: int, base: int = 10000):
    """cos/sin tables (S, D/2) built with the reference's own NumPy expressions (gpt2.py:25-54) so angles are
    bit-identical to the NumPy path; cached on device."""
    key = (head_dim, seq_len, base)
    hit = _ROPE_CACHE.get(key)
    if hit is None:
        inv_freq = 1.0 / (base ** (np.arange(0, head_dim, 2).astype(np.float32) / head_dim))
        t = np.arange(seq_len, dtype=np.float32)
        freqs = np.outer(t, inv_freq)
        hit = (B200Array.from_numpy(np.cos(freqs).astype(np.float32)), B200Array.from_numpy(np.sin(freqs).astype(np.float32)))
        _ROPE_CACHE[key] = hit
    return hit


def _bhsd_strides(x: B200Array):
    """Logical (B,H,S,D) array -> (sb, ss, sh) element strides; requires unit stride along D."""
    if x.ndim != 4 or x.strides[3] != 1:
        raise ValueError("attention operands must be 4-D (B,H,S,D) with a contiguous last dimension")
    return x.strides[0], x.strides[2], x.strides[1]


def rope_interleaved_inplace(x: B200Array, cos_t: B200Array, sin_t: B200Array, inverse=False):
    """Interleaved-pair rotary embedding (nn/positional.py:212-270) in place on a logical (B,H,S,D) or (B,S,D) array;
    cos_t / sin_t: device fp32 (S, D/2) tables of the positions of these S rows."""
    if x.ndim == 3:
        Bn, S, D = x.shape
        H, sb, ss, sh = 1, x.strides[0], x.strides[1], 0
        if x.strides[2] != 1:
            raise ValueError("rope: the last dimension must be contiguous")
    else:
        Bn, H, S, D = x.shape
        sb, ss, sh = _bhsd_strides(x)
    call("nfb_rope_interleaved", dtype_code(x.dtype), x.ptr, cos_t.ptr, sin_t.ptr, Bn, S, H, D, sb, ss, sh, 1 if inverse else 0)
    return x


def rope_inplace(x: B200Array, inverse=False, base=10000):
    """Rotate a logical (B,H,S,D) array (any strides) in place."""
    Bn, H, S, D = x.shape
    cos_t, sin_t = rope_tables(D, S, base)
    sb, ss, sh = _bhsd_strides(x)
    call("nfb_rope", dtype_code(x.dtype), x.ptr, cos_t.ptr, sin_t.ptr, 1, 0, Bn, S, H, D, sb, ss, sh, 1 if inverse else 0)
    return x


def rope_packed_qk_inplace(qkv: B200Array, n_heads: int, inverse=False, base=10000):
    """qkv: packed (B,S,3,H,D) [or (T=B*S,3*H*D) with B,S given by shape]: rotates the q and k parts in place."""
    Bn, S, three, H, D = qkv.shape
    cos_t, sin_t = rope_tables(D, S, base)
    call("nfb_rope", dtype_code(qkv.dtype), qkv.ptr, cos_t.ptr, sin_t.ptr, 2, H * D, Bn, S, H, D, S * 3 * H * D, 3 * H * D, D,
         1 if inverse else 0)
    return qkv


def _key_mask(mask, Bn, Skv):
    if mask is None:
        return None
    m = mask if isinstance(mask, B200Array) else B200Array.from_numpy(np.asarray(mask, np.float32))
    if m.dtype != F32:
        m = m.astype(np.float32)
    if m.size == Bn * Skv:
        return m.reshape(Bn, Skv).contiguous()
    if m.size == Skv:
        return m.reshape(1, Skv).broadcast_to((Bn, Skv)).contiguous()
    raise ValueError(f"attention mask must be an additive per-key mask broadcastable to (B,1,1,S_kv); got shape {m.shape}")


def attention_forward(q: B200Array, k: B200Array, v: B200Array, scale: float, causal=False, mask=None, need_lse=True):
    """Fused softmax(q k^T * scale [masks]) v.  q,k,v logical (B,H,S,D), bf16 (fp32 inputs are cast), D == 64.
    k / v may have fewer heads than q (grouped-query / multi-query attention, nn/modern_attention.py:26-358): query head h
    then reads key/value head h // (H // Hkv) in place.
    Returns (o, lse): o logical (B,H,S,D) stored as (B,S,H,D) so o.transpose(0,2,1,3).reshape(B,S,H*D) is free."""
    q, k, v = to_bf16(q), to_bf16(k), to_bf16(v)
    Bn, H, S, D = q.shape
    Hkv, Skv = k.shape[1], k.shape[2]
    o_store = B200Array.empty((Bn, S, H, D), bfloat16)
    lse = B200Array.empty((Bn, H, S), np.float32) if need_lse else None
    km = _key_mask(mask, Bn, Skv)
    qs, ks_, vs = _bhsd_strides(q), _bhsd_strides(k), _bhsd_strides(v)
    if ks_ != vs:
        v = v.contiguous()
        k = k.contiguous()
        ks_ = _bhsd_strides(k)
    mode = (1 if causal else 0) | (2 if km is not None else 0)
    call("nfb_flash_attn_fwd_gqa", _lib.BF16, q.ptr, k.ptr, v.ptr, o_store.ptr, lse.ptr if lse is not None else None, Bn, H, Hkv, S, D,
         qs[0], qs[1], qs[2], ks_[0], ks_[1], ks_[2], S * H * D, H * D, D, Skv, float(scale), mode, km.ptr if km is not None else None)
    return o_store.transpose(0, 2, 1, 3), lse


def attention_backward(dout: B200Array, q, k, v, o, lse, scale, causal=False, mask=None, packed=False):
    """Returns (dq, dk, dv) as logical (B,H,S,D) bf16 arrays (dk / dv with k's head count: grouped-query gradients are
    summed over each group inside the kernel).  With packed=True the three live in ONE (B,S,3,H,D) buffer (returned
    as a 4th value) laid out like a fused qkv projection output."""
    q, k, v = to_bf16(q), to_bf16(k), to_bf16(v)
    Bn, H, S, D = q.shape
    Hkv, Skv = k.shape[1], k.shape[2]
    dout = to_bf16(dout)
    os_ = _bhsd_strides(o)
    if _bhsd_strides(dout) != os_:
        # bring dout into o's storage layout (B,S,H,D)
        tmp = B200Array.empty((Bn, S, H, D), bfloat16).transpose(0, 2, 1, 3)
        if _bhsd_strides(tmp) != os_:
            o = o.transpose(0, 2, 1, 3).contiguous().transpose(0, 2, 1, 3)
            os_ = _bhsd_strides(o)
        tmp._assign(dout)
        dout = tmp
    qs, ks_ = _bhsd_strides(q), _bhsd_strides(k)
    if _bhsd_strides(v) != ks_:
        k, v = k.contiguous(), v.contiguous()
        ks_ = _bhsd_strides(k)
    km = _key_mask(mask, Bn, Skv)
    mode = (1 if causal else 0) | (2 if km is not None else 0)
    if packed:
        if S != Skv or H != Hkv:
            raise ValueError("packed dqkv needs S == S_kv and as many key/value heads as query heads")
        buf = B200Array.empty((Bn, S, 3, H, D), bfloat16)
        dq = buf[:, :, 0].transpose(0, 2, 1, 3)
        dk = buf[:, :, 1].transpose(0, 2, 1, 3)
        dv = buf[:, :, 2].transpose(0, 2, 1, 3)
    else:
        buf = None
        dq = B200Array.empty((Bn, S, H, D), bfloat16).transpose(0, 2, 1, 3)
        dk = B200Array.empty((Bn, Skv, Hkv, D), bfloat16).transpose(0, 2, 1, 3)
        dv = B200Array.empty((Bn, Skv, Hkv, D), bfloat16).transpose(0, 2, 1, 3)
    # the kernels address dq with q's strides and dk/dv with k's strides: make operand layouts agree
    if _bhsd_strides(dq) != qs:
        q = _relayout_like(q, dq)
        qs = _bhsd_strides(q)
    if _bhsd_strides(dk) != ks_:
        k = _relayout_like(k, dk)
        v = _relayout_like(v, dv)
        ks_ = _bhsd_strides(k)
    call("nfb_flash_attn_bwd_gqa", _lib.BF16, q.ptr, k.ptr, v.ptr, o.ptr, dout.ptr, lse.ptr, dq.ptr, dk.ptr, dv.ptr, Bn, H, Hkv, S, D,
         qs[0], qs[1], qs[2], ks_[0], ks_[1], ks_[2], os_[0], os_[1], os_[2], Skv, float(scale), mode, km.ptr if km is not None else None)
    return (dq, dk, dv, buf) if packed else (dq, dk, dv)


def _relayout_like(x: B200Array, like: B200Array) -> B200Array:
    """Copy logical (B,H,S,D) x into a fresh buffer with the same stride pattern as `like`."""
    Bn, H, S, D = x.shape
    sb, ss, sh = _bhsd_strides(like)
    if (sb, ss, sh) == (S * 3 * H * D, 3 * H * D, D):
        buf = B200Array.empty((Bn, S, 3, H, D), bfloat16)
        out = buf[:, :, 0].transpose(0, 2, 1, 3)
    else:
        out = B200Array.empty((Bn, S, H, D), bfloat16).transpose(0, 2, 1, 3)
        if _bhsd_strides(out) != (sb, ss, sh):
            raise B200Error("attention: unsupported gradient layout")
    out._assign(x)
    return out


def attention_reference_unfused(q, k, v, scale, causal=False, mask=None):
    """fp32 unfused attention (batched SIMT GEMM + softmax kernel + GEMM): the always-correct 1e-5 path; it
    materialises the (B,H,S,S) score tensor like the reference does.  Returns (o, p)."""
    from . import array as A
    from .b200_backend import B200Backend
    be = B200Backend()
    s = A.matmul(q, k.transpose(0, 1, 3, 2)) * float(scale)
    if causal:
        S, Skv = s.shape[-2], s.shape[-1]
        s = A.where(A.tril_mask(S, Skv), s, -1e4)
    if mask is not None:
        s = s + mask
    p = be.softmax(s, axis=-1)
    return A.matmul(p, v), p
