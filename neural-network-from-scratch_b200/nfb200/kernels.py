"""Layer-level fused kernels (host wrappers): tensor-core GEMM with epilogues, Linear fwd/bwd,
RoPE, fused flash attention fwd/bwd.  These are what nn.Linear / attention blocks / the model zoo of
this repo call; each is a handful of launches of hand-written sm_100a kernels.

Precision modes (the reference has no 16-bit Tensor dtype, so bf16 is an internal
compute mode of the backend):
  * "fp32": SIMT FFMA contraction (csrc/gemm_simt.cu)          -- 1e-5 relative contract
  * "bf16": tcgen05/TMEM GEMM on bf16 operands, fp32 accumulate  -- 1e-2 contract, throughput path
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from ._lib import B200Error, call
from .array import B200Array, bfloat16, dtype_code

_PRECISION = "fp32"
F32 = np.dtype(np.float32)


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
              residual=None, epilogue=None, accumulate=False, split_k=0, clip=0.0, aux=False, ldc=None):
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
    call("nfb_gemm_bf16", 1 if trans_a else 0, 1 if trans_b else 0, M, N, K, a2.ptr, lda, b2.ptr, ldb, out.ptr,
         dtype_code(out.dtype), ldc_v, bias_ptr, res_ptr, ldr, aux_arr.ptr if aux_arr is not None else None, N, epi,
         1 if accumulate else 0, int(split_k), float(clip))
    return (out, aux_arr) if aux else out


def gemm_f32(a: B200Array, b: B200Array, trans_a=False, trans_b=False, out=None, bias=None, beta=0.0):
    """fp32 SIMT contraction with the same calling convention as gemm_bf16 (2-D operands)."""
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
    call("nfb_gemm_f32", 1 if trans_a else 0, 1 if trans_b else 0, M, N, K, 1.0, a2.ptr, lda, 0, b2.ptr, ldb, 0, float(beta),
         out.ptr, out.strides[0] if M > 1 else N, 0, 1, bias_c.ptr if bias_c is not None else None)
    return out
