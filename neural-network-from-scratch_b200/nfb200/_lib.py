"""ctypes binding of libnfb200.so (the C-ABI declared in include/nfb200.h).

This is the only place the host code touches native code.  There is no CPU fallback: every
wrapper raises :class:`B200Error` when the shared library is missing, was not built, or when a
call returns a non-zero status (the message comes from ``nfb_last_error``).  The reference's CuPy
backend swallowed kernel errors and fell back to NumPy (backends/cuda_backend.py:305-314); we
deliberately do not.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NFB200_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libnfb200.so"))


class B200Error(RuntimeError):
    """A libnfb200 call failed (or the library / a B200 is not present)."""


vp, i32, i64, f32, f64, sz, u64p = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t, C.POINTER(C.c_ulonglong)
i64p = C.POINTER(C.c_int64)

# name -> argtypes  (all return int status unless listed in _SPECIAL_RESTYPE)
_SIGNATURES = {
    # runtime
    "nfb_device_count": [C.POINTER(C.c_int)],
    "nfb_init": [i32],
    "nfb_device_name": [C.c_char_p, i32],
    "nfb_malloc": [sz, C.POINTER(vp)],
    "nfb_free": [vp],
    "nfb_empty_cache": [],
    "nfb_mem_stats": [u64p],
    "nfb_host_alloc": [sz, C.POINTER(vp)],
    "nfb_host_free": [vp],
    "nfb_memcpy_h2d": [vp, vp, sz],
    "nfb_memcpy_h2d_async": [vp, vp, sz],
    "nfb_memcpy_d2h": [vp, vp, sz],
    "nfb_memcpy_d2h_async": [vp, vp, sz],
    "nfb_memcpy_d2d": [vp, vp, sz],
    "nfb_memset": [vp, i32, sz],
    "nfb_stream_create": [C.POINTER(vp)],
    "nfb_stream_destroy": [vp],
    "nfb_set_stream": [vp],
    "nfb_stream_sync": [vp],
    "nfb_device_sync": [],
    "nfb_event_create": [C.POINTER(vp)],
    "nfb_event_create_notiming": [C.POINTER(vp)],
    "nfb_event_destroy": [vp],
    "nfb_event_record": [vp, vp],
    "nfb_event_record_external": [vp, vp],
    "nfb_event_sync": [vp],
    "nfb_event_elapsed_ms": [vp, vp, C.POINTER(C.c_float)],
    "nfb_stream_wait_event": [vp, vp],
    "nfb_graph_begin_capture": [],
    "nfb_graph_end_capture": [C.POINTER(vp)],
    "nfb_graph_launch": [vp],
    "nfb_graph_destroy": [vp],
    "nfb_flush_l2": [],
    # elementwise
    "nfb_binary": [i32, i32, vp, vp, vp, i32, i64p, i64p, i64p],
    "nfb_binary_scalar": [i32, i32, vp, f64, vp, i64, i32],
    "nfb_unary": [i32, i32, vp, vp, i64, f64, f64],
    "nfb_unary_bwd": [i32, i32, vp, vp, vp, i64],
    "nfb_cast": [i32, i32, vp, vp, i64],
    "nfb_where": [i32, vp, i64p, vp, i64p, vp, i64p, vp, i32, i64p],
    "nfb_copy_strided": [i32, vp, vp, i32, i64p, i64p, i64p],
    "nfb_fill": [i32, vp, i64, f64],
    "nfb_arange": [i32, vp, i64, f64, f64],
    "nfb_swiglu_fwd": [i32, vp, vp, vp, i64, i64, i64],
    "nfb_swiglu_bwd": [i32, vp, vp, vp, vp, vp, i64, i64, i64],
    "nfb_add3": [vp, vp, vp, vp, i64],
    # reductions
    "nfb_reduce": [i32, i32, vp, vp, i64, i64, i64],
    "nfb_argreduce": [i32, i32, vp, vp, i64, i64, i64],
    "nfb_colsum": [i32, vp, i64, i32, i64, vp, i32],
    # norms / softmax / loss
    "nfb_norm_fwd": [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, f32],
    "nfb_norm_bwd": [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i64, i32, f32, vp],
    "nfb_softmax_fwd": [vp, vp, i64, i32],
    "nfb_softmax_bwd": [vp, vp, vp, i64, i32],
    "nfb_cross_entropy": [vp, vp, i32, vp, vp, i32, i64, i32, i64, i64, f32],
    "nfb_cross_entropy_set_bulk": [i32],
    "nfb_cross_entropy_ex": [i32, vp, vp, i32, vp, vp, i32, i64, i32, i64, i64, f32],
    # embedding
    "nfb_gather_rows": [i32, i32, i32, vp, vp, vp, i64, i32, i64],
    "nfb_scatter_add_rows": [i32, i32, vp, vp, vp, i64, i32, i64, i32],
    "nfb_index_error_check": [C.POINTER(C.c_int)],
    # optimizers
    "nfb_adam_step": [i32, i32, vp, vp, vp, vp, vp, i64, f64, f64, f64, f64, f64, i32, i32, vp, vp],
    "nfb_adam_step_ex": [i32, i32, vp, vp, vp, vp, vp, i64, f64, f64, f64, f64, f64, i32, i32, vp, vp, vp],
    "nfb_lion_step": [i32, vp, vp, vp, vp, i64, f64, f64, f64, f64, i32, vp, vp],
    "nfb_grad_unscale": [vp, vp, i32, i32, f32, f32, vp, vp],
    "nfb_add_i32": [vp, i32],
    "nfb_sgd_step": [vp, vp, vp, vp, i64, f64, f64, f64, f64, i32, i32, i32],
    "nfb_sgd_step_ex": [vp, vp, vp, vp, i64, f64, f64, f64, f64, i32, i32, i32, vp],
    "nfb_sumsq": [vp, i64, vp, i32],
    "nfb_clip_coef": [vp, f32, f32, vp, vp],
    "nfb_scale_by_device_scalar": [vp, i64, vp],
    # contractions
    "nfb_gemm_f32": [i32, i32, i32, i32, i32, f32, vp, i64, i64, vp, i64, i64, f32, vp, i64, i64, i32, vp],
    "nfb_gemm_f64": [i32, i32, i32, i32, i32, vp, i64, i64, vp, i64, i64, vp, i64, i64, i32],
    "nfb_split_bf16x3": [vp, i64, i64, i64, i32, i32, vp],
    "nfb_gemm_bf16": [i32, i32, i32, i32, i32, vp, i64, vp, i64, vp, i32, i64, vp, vp, i64, vp, i64, i32, i32, i32, f32],
    "nfb_gemm_bf16_force_tile": [i32],
    "nfb_gemm_bf16_force_cta_group": [i32],
    "nfb_gemm_bf16_force_raster": [i32],
    "nfb_gemm_bf16_set_auto_split": [i32],
    "nfb_gemm_bf16_plan": [i32, i32, i32, i32, i32, i32, i32, i32, C.POINTER(C.c_int)],
    "nfb_gemm_bf16_grouped": [i32, i32, i32, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(vp), i64p, C.POINTER(vp), i64p,
                              C.POINTER(vp), i64p, i32, i32],
    "nfb_gemm_bf16_set_rope": [vp, vp, i32, i32],
    "nfb_gemm_bf16_set_stream_k": [i32],
    "nfb_gemm_bf16_set_sm_limit": [i32],
    "nfb_gemm_bf16_set_span": [vp],
    "nfb_set_pdl": [i32],
    "nfb_set_aux_stream": [vp],
    "nfb_norm_flush": [],
    "nfb_gemm_bf16_set_tma_store": [i32],
    "nfb_gemm_bf16_set_timeline": [vp],
    # attention
    "nfb_rope": [i32, vp, vp, vp, i32, i64, i32, i32, i32, i32, i64, i64, i64, i32],
    "nfb_flash_attn_set_tc": [i32],
    "nfb_flash_attn_set_poly": [i32, i32],
    "nfb_flash_attn_set_timeline": [vp],
    "nfb_flash_attn_fwd": [i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, i64, i32, f32, i32, vp],
    "nfb_flash_attn_bwd": [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, i64,
                           i32, f32, i32, vp],
    "nfb_flash_attn_set_dq_f32_blocks": [i32],
    "nfb_rope_interleaved": [i32, vp, vp, vp, i32, i32, i32, i32, i64, i64, i64, i32],
    "nfb_flash_attn_fwd_gqa": [i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, i64, i32, f32, i32, vp],
    "nfb_flash_attn_bwd_gqa": [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, i64,
                               i32, f32, i32, vp],
    # collectives
    "nfb_comm_load": [C.c_char_p],
    "nfb_comm_unique_id": [C.c_char_p, i32],
    "nfb_comm_init": [i32, i32, C.c_char_p, i32, C.POINTER(vp)],
    "nfb_comm_destroy": [vp],
    "nfb_comm_allreduce": [vp, vp, vp, i64, i32, i32, vp],
    "nfb_comm_broadcast": [vp, vp, i64, i32, i32, vp],
    "nfb_comm_allgather": [vp, vp, vp, i64, i32, vp],
    "nfb_comm_reduce_scatter": [vp, vp, vp, i64, i32, i32, vp],
}
_SPECIAL_RESTYPE = {
    "nfb_last_error": C.c_char_p,
    "nfb_version": C.c_char_p,
    "nfb_launch_count": C.c_ulonglong,
    "nfb_get_stream": vp,
    "nfb_current_device": C.c_int,
    "nfb_get_sm_count": C.c_int,
}

# dtype codes (csrc/common.cuh NfbDType)
F32, F64, I32, I64, BOOL, BF16 = 0, 1, 2, 3, 4, 5

_lib = None
_lib_error = None
_lock = threading.Lock()
_initialised_device = None


def load():
    """Load the shared library (once).  Raises B200Error when it is absent."""
    global _lib, _lib_error
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            _lib_error = f"libnfb200.so not found at {LIB_PATH}; run `python -c 'import __graft_entry__ as g; g.build()'` (make -C csrc)"
            raise B200Error(_lib_error)
        try:
            lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        except OSError as e:  # pragma: no cover
            _lib_error = f"cannot load {LIB_PATH}: {e}"
            raise B200Error(_lib_error) from e
        for name, args in _SIGNATURES.items():
            fn = getattr(lib, name, None)
            if fn is None:
                continue  # checked by tests/test_abi.py against include/nfb200.h
            fn.argtypes = args
            fn.restype = C.c_int
        for name, res in _SPECIAL_RESTYPE.items():
            fn = getattr(lib, name, None)
            if fn is not None:
                fn.restype = res
                fn.argtypes = []
        _lib = lib
        return lib


def lib_available() -> bool:
    try:
        load()
        return True
    except B200Error:
        return False


def last_error() -> str:
    return load().nfb_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = ""):
    if rc != 0:
        raise B200Error(f"{what + ': ' if what else ''}{last_error()}")


# Measurement hook (bench.py): with PROFILE set to a list, every call of an entry point named in PROFILE_NAMES is bracketed by
# a pair of timing events recorded on the compute stream (external event-record nodes when the step is being captured into a
# CUDA graph) and appended as (name, args, start_event, end_event).  Off (None) in production: one attribute load per call.
PROFILE = None
PROFILE_NAMES = frozenset()


def call(name: str, *args):
    """Call an entry point and raise on a non-zero status."""
    lib = load()
    fn = getattr(lib, name, None)
    if fn is None:
        raise B200Error(f"libnfb200.so does not export {name} (stale build?)")
    prof = PROFILE
    if prof is not None and name in PROFILE_NAMES:
        from .runtime import Event
        e0 = Event().record(external=True)
        rc = fn(*args)
        prof.append((name, args, e0, Event().record(external=True)))
    else:
        rc = fn(*args)
    if rc != 0:
        raise B200Error(f"{name}: {last_error()}")


def device_count() -> int:
    lib = load()
    n = C.c_int(0)
    rc = lib.nfb_device_count(C.byref(n))
    if rc != 0:
        return 0
    return n.value


def device_available() -> bool:
    try:
        return device_count() > 0
    except B200Error:
        return False


def init(device=None) -> int:
    """Bind this process to one B200 (default: LOCAL_RANK or 0).  Idempotent."""
    global _initialised_device
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if _initialised_device is None else _initialised_device
    if _initialised_device == device:
        return device
    call("nfb_init", int(device))
    _initialised_device = device
    return device


def ensure_init() -> int:
    if _initialised_device is None:
        return init()
    return _initialised_device


def launch_count() -> int:
    return int(load().nfb_launch_count())


def sm_count() -> int:
    ensure_init()
    return int(load().nfb_get_sm_count())


def i64_array(values):
    arr = (C.c_int64 * max(1, len(values)))(*[int(v) for v in values])
    return arr
