"""Thin host wrappers over the runtime part of the C-ABI: events, streams, CUDA graphs, timing."""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import call


class Event:
    def __init__(self, timing: bool = True):
        _lib.ensure_init()
        h = C.c_void_p()
        call("nfb_event_create" if timing else "nfb_event_create_notiming", C.byref(h))
        self.handle = h.value

    def record(self, stream=None, external=False):
        """external=True: when the stream is being captured, record as an external event node (readable after replays)."""
        call("nfb_event_record_external" if external else "nfb_event_record", self.handle, stream.handle if isinstance(stream, Stream) else stream)
        return self

    def synchronize(self):
        call("nfb_event_sync", self.handle)

    def elapsed_ms(self, later: "Event") -> float:
        ms = C.c_float(0)
        call("nfb_event_elapsed_ms", self.handle, later.handle, C.byref(ms))
        return float(ms.value)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                _lib.load().nfb_event_destroy(C.c_void_p(h))
            except Exception:
                pass


class Stream:
    def __init__(self):
        _lib.ensure_init()
        h = C.c_void_p()
        call("nfb_stream_create", C.byref(h))
        self.handle = h.value

    def synchronize(self):
        call("nfb_stream_sync", self.handle)

    def wait_event(self, ev: Event):
        call("nfb_stream_wait_event", self.handle, ev.handle)


def current_stream_handle():
    _lib.ensure_init()
    return _lib.load().nfb_get_stream()


def compute_stream_wait(ev: Event):
    call("nfb_stream_wait_event", None, ev.handle)


def synchronize():
    _lib.ensure_init()
    call("nfb_device_sync")


def stream_synchronize():
    _lib.ensure_init()
    call("nfb_stream_sync", None)


def flush_l2():
    call("nfb_flush_l2")


def mem_stats() -> dict:
    _lib.ensure_init()
    s = (C.c_ulonglong * 4)()
    call("nfb_mem_stats", s)
    return {"live_bytes": int(s[0]), "cached_bytes": int(s[1]), "peak_bytes": int(s[2]), "cuda_mallocs": int(s[3])}


def empty_cache():
    call("nfb_empty_cache")


_CAPTURING = 0


def is_capturing() -> bool:
    """True while a ``Graph`` is recording the compute stream (host code that must not be frozen into the graph checks it)."""
    return _CAPTURING > 0


class Graph:
    """Capture the kernels issued inside the ``with`` block on the compute stream; ``replay()`` relaunches them.

    All device buffers touched during capture must stay alive (and at the same address) for replays:
    the caching allocator hands out deterministic addresses once warmed up, and callers keep the
    step's persistent tensors referenced.
    """

    def __init__(self):
        self.exec = None

    def __enter__(self):
        _lib.ensure_init()
        global _CAPTURING
        call("nfb_graph_begin_capture")
        _CAPTURING += 1
        return self

    def __exit__(self, et, ev, tb):
        global _CAPTURING
        _CAPTURING = max(0, _CAPTURING - 1)
        h = C.c_void_p()
        if et is not None:
            try:
                call("nfb_graph_end_capture", C.byref(h))
            except Exception:
                pass
            return False
        call("nfb_graph_end_capture", C.byref(h))
        self.exec = h.value
        return False

    def replay(self):
        call("nfb_graph_launch", self.exec)

    def __del__(self):
        h, self.exec = getattr(self, "exec", None), None
        if h:
            try:
                _lib.load().nfb_graph_destroy(C.c_void_p(h))
            except Exception:
                pass


def time_ms(fn, iters: int = 20, warmup: int = 3, flush: bool = False) -> float:
    """Mean device milliseconds per call of ``fn`` (CUDA events on the launching stream)."""
    for _ in range(warmup):
        fn()
    stream_synchronize()
    total = 0.0
    if flush:
        for _ in range(iters):
            flush_l2()
            a, b = Event(), Event()
            a.record()
            fn()
            b.record()
            b.synchronize()
            total += a.elapsed_ms(b)
        return total / iters
    a, b = Event(), Event()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_ms(b) / iters
