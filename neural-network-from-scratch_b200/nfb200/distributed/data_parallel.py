"""DistributedDataParallel / DataParallel / DistributedSampler (mirror of distributed/data_parallel.py:21-388).

B200 design: one process per GPU; the module's parameters are re-homed into one flat fp32 arena with a
matching flat gradient arena; gradients are all-reduced (NCCL ncclAvg over NVLink/NVSwitch) in contiguous ~25 MB
BUCKETS of that arena -- the reference accepts bucket_size_mb=25 but all-reduces one tensor at a time,
synchronously (data_parallel.py:250-287).  A bucket is launched on a dedicated communication stream as soon as the
backward pass has produced the last gradient that lives in it (use-count tracking in the autograd engine), so the
transfer overlaps the rest of backward; finish_gradient_sync() fences the compute stream before the optimizer.
"""
from __future__ import annotations

import os
import weakref

from contextlib import contextmanager
from typing import List, Optional

import numpy as np

from ..array import B200Array
from ..core import Module, Tensor
from ..optim.arena import ALIGN, ParamArena
from . import communication as comm
from .communication import ReduceOp


_DEBUG_SKIP_ALLREDUCE = bool(int(os.environ.get("NFB_DDP_DEBUG_SKIP_ALLREDUCE", "0")))


class _Bucket:
    __slots__ = ("lo", "hi", "params", "pending", "launched", "event", "seq", "pieces", "piece_events")

    def __init__(self, lo, hi, params):
        self.lo, self.hi, self.params = lo, hi, params
        self.pending = 0
        self.launched = False
        self.event = None
        self.seq = 0   # launch order of the all-reduce within the step
        self.pieces = [(lo, hi)]     # a bucket holding ONE oversized parameter is all-reduced in bucket-sized pieces
        self.piece_events = []


class DistributedDataParallel(Module):
    def __init__(self, module: Module, device_ids=None, output_device=None, bucket_size_mb: float = 25,
                 find_unused_parameters: bool = False, broadcast_parameters: bool = True, overlap: bool = True,
                 sparse_embedding_grads: bool = False, grad_dtype: str = "float32"):
        super().__init__()
        if grad_dtype not in ("float32", "bfloat16"):
            raise ValueError("grad_dtype must be 'float32' or 'bfloat16'")
        # bfloat16: every bucket is cast to bf16 on the communication stream, all-reduced (ncclAvg) in bf16 and cast back into
        # the fp32 gradient arena -- half the bytes on the wire and half the time NCCL's channel CTAs hold SMs next to the
        # backward GEMMs; the averaged gradient carries one bf16 rounding (2^-9 relative), inside the bf16 mode's 1e-2 contract
        self.grad_dtype = grad_dtype
        self._g16 = None
        self.module = module
        self.bucket_size_mb = bucket_size_mb
        self.find_unused_parameters = find_unused_parameters
        # find_unused_parameters: ranks may run different control flow, so the order in which buckets become ready is not the
        # same everywhere -- and NCCL calls must be issued in the same order on every rank.  All buckets are then reduced
        # after backward, in bucket order (no overlap), with zeros standing in for the gradients a rank did not produce.
        self.overlap = overlap and not find_unused_parameters
        self._sync_consumed = True
        self.group = comm.get_backend() if comm.is_initialized() else None
        self.world_size = self.group.world_size if self.group else 1
        self.rank = self.group.rank if self.group else 0
        self._require_sync = True
        _LIVE_DDP.add(self)
        params = list(module.parameters())
        self._dev_params = [p for p in params if isinstance(p.data, B200Array)]
        self._host_params = [p for p in params if not isinstance(p.data, B200Array)]
        self.arena: Optional[ParamArena] = None
        self.buckets: List[_Bucket] = []
        if self._dev_params:
            from .. import kernels as K
            self.arena = ParamArena(self._dev_params, np.float32, 0, with_bf16=(K.get_precision() == "bf16"))
            self._build_buckets()
            for p in self.arena.params:
                p._ddp = self
            # Large embedding tables: the gradient of the lookup reaches at most `tokens` rows and is only final when the LAST
            # node of backward has run, so as part of a dense bucket it would put the whole table (154 MB for the tied GPT-2
            # wte, 94 MB for BERT's) on the wire AFTER backward, with nothing left to hide it behind.  Instead the table's
            # dense consumers (a tied LM head) are bucketed as usual -- ready early -- and the lookup's rows are exchanged as
            # (ids, rows) with one all-gather after backward and scatter-added, scaled by 1/world, in rank order.
            self._sparse_ids = {}
            self._pending_sparse = []
            # (opt-in: measured slower on GPT-2 small at 2 GPUs -- 6.22 vs 6.11 ms -- because the bucket-pipelined AdamW already hides
            # the table's all-reduce, and the exchange gives that pipelining up)
            if sparse_embedding_grads and self.world_size > 1 and self.group is not None and self.group.backend == "nccl":
                from ..nn.layers import Embedding
                for m in module.modules():
                    w = getattr(m, "weight", None)
                    if isinstance(m, Embedding) and w is not None and id(w) in self.arena.offsets and w.size >= int(os.environ.get("NFB200_DDP_SPARSE_MIN", 1 << 20)):
                        w._ddp_sparse = True
                        self._sparse_ids[id(w)] = w
        if broadcast_parameters and self.world_size > 1:
            self.broadcast_parameters()

    # ------------------------------------------------------------------ buckets
    def _build_buckets(self):
        """Contiguous slices of the gradient arena, filled in REVERSE registration order (backward order)."""
        cap = max(1, int(self.bucket_size_mb * (1 << 20) / 4))
        a = self.arena
        cur, cur_hi = [], None
        for p in reversed(a.params):
            o = a.offsets[id(p)]
            n = (p.size + ALIGN - 1) // ALIGN * ALIGN
            if cur and (cur_hi - o) > cap:
                self.buckets.append(_Bucket(min(a.offsets[id(q)] for q in cur), cur_hi, cur))
                cur, cur_hi = [], None
            if cur_hi is None:
                cur_hi = o + n
            cur.append(p)
        if cur:
            self.buckets.append(_Bucket(min(a.offsets[id(q)] for q in cur), cur_hi, cur))
        for b in self.buckets:   # e.g. the 154 MB tied embedding gradient: pieces let the optimizer start on piece i while i+1 is on the wire
            if b.hi - b.lo > 2 * cap:
                b.pieces = [(o, min(o + cap, b.hi)) for o in range(b.lo, b.hi, cap)]
        self._bucket_of = {}
        for b in self.buckets:
            for p in b.params:
                self._bucket_of[id(p)] = b

    def broadcast_parameters(self, src: int = 0):
        if self.arena is not None and self.group is not None and self.group.backend == "nccl":
            self.group.broadcast_buffer(self.arena.flat_p, src)
            if self.arena.flat_bf16 is not None:
                self.group.broadcast_buffer(self.arena.flat_bf16, src)
            for p in self.arena.params:
                p._version += 1
        for p in self._host_params:
            if self.group is not None:
                self.group.broadcast(p, src)

    # ------------------------------------------------------------------ forward / overlap machinery
    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def _begin_backward(self, counts):
        """Called by the autograd engine at the start of a backward pass with {id(p): (p, number of graph nodes that
        consume p)} for the parameters of this wrapper the pass will reach.  A bucket is handed to NCCL when the last
        consumer of its last reachable parameter has run."""
        if self.arena is None or not self._require_sync or self.world_size <= 1:
            return
        if any(b.launched for b in self.buckets) and not self._sync_consumed:
            raise RuntimeError("DistributedDataParallel: backward() ran again before the gradients of the previous backward were "
                               "consumed (finish_gradient_sync / step_optimizer / sync_gradients); accumulate gradients under "
                               "no_sync() and synchronise on the last micro-batch only")
        self._sync_consumed = False
        for b in self.buckets:
            b.pending = 0
            b.launched = False
        for p in self.arena.params:
            p._ddp_uses = 0
        for p, n in counts.values():
            b = self._bucket_of.get(id(p))
            if b is None:
                continue
            p._ddp_uses = n
            b.pending += 1

    def _note_use(self, p):   # kept for engines that count at graph-build time; counting now happens in _begin_backward
        pass

    # ------------------------------------------------------------------ sparse embedding gradients
    def defer_embedding_grad(self, weight, g, idx) -> bool:
        """Called by the embedding backward for a table flagged _ddp_sparse.  Stages this rank's (ids, rows / world) in
        persistent send buffers (on the compute stream); _flush_sparse() exchanges and applies them.  Returns False when the
        dense path must be used (no synchronisation this pass: no_sync, single rank)."""
        if not (self._require_sync and self.world_size > 1 and id(weight) in self._sparse_ids):
            return False
        from .._lib import call
        from ..array import dtype_code
        d = weight.shape[1]
        rows = g.reshape(-1, d)
        if rows.dtype != np.dtype(np.float32):
            rows = rows.astype(np.float32)
        rows = rows.contiguous()
        T = rows.shape[0]
        st = self.__dict__.setdefault("_sparse_bufs", {}).get(id(weight))
        if st is None or st["T"] != T:
            N = self.world_size
            st = {"T": T, "send_rows": B200Array.empty((T, d), np.float32), "send_idx": B200Array.empty((T,), np.int64),
                  "all_rows": B200Array.empty((N * T, d), np.float32), "all_idx": B200Array.empty((N * T,), np.int64)}
            self._sparse_bufs[id(weight)] = st
        call("nfb_binary_scalar", 2, dtype_code(np.float32), rows.ptr, 1.0 / self.world_size, st["send_rows"].ptr, rows.size, 0)   # rows * (1 / world)
        ids = idx if isinstance(idx, B200Array) else B200Array.from_numpy(np.asarray(idx))
        st["send_idx"]._assign(ids.reshape(-1))
        self._pending_sparse.append((weight, st))
        return True

    def _flush_sparse(self):
        """All-gather the staged (ids, rows) of every rank on the communication stream, then scatter-add them -- rank 0's rows
        first, each rank's in token order: deterministic -- into the (already averaged) gradient of the table on the
        compute stream.  Runs after every bucket has been launched; returns the ids of the parameters it touched."""
        if not self._pending_sparse:
            return set()
        from .._lib import call
        from ..array import dtype_code
        from ..runtime import Event, compute_stream_wait
        g = self.group
        pend, self._pending_sparse = self._pending_sparse, []
        g.comm_stream.wait_event(Event(timing=False).record())          # the staged send buffers are final on the compute stream
        for weight, st in pend:
            for src, dst in ((st["send_rows"], st["all_rows"]), (st["send_idx"], st["all_idx"])):
                call("nfb_comm_allgather", g.comm, src.ptr, dst.ptr, src.size, dtype_code(src.dtype), g.comm_stream.handle)
        compute_stream_wait(Event(timing=False).record(g.comm_stream))   # ... behind the bucket all-reduces issued on that stream before
        touched = set()
        be = None
        for weight, st in pend:
            be = be or weight.backend
            be.embedding_backward(st["all_rows"], st["all_idx"], weight.shape[0], out=weight._grad_slot, accumulate=True)
            weight._grad = weight._grad_slot
            weight._slot_zero = False
            touched.add(id(weight))
        return touched

    def _note_grad_done(self, p):
        """Called by the autograd engine after a consumer node of p has run its backward."""
        if not (self._require_sync and self.overlap and self.world_size > 1):
            return
        p._ddp_uses -= 1
        if p._ddp_uses == 0:
            b = self._bucket_of[id(p)]
            b.pending -= 1
            if b.pending == 0 and not b.launched:
                self._launch_bucket(b)

    def _launch_bucket(self, b: _Bucket):
        from ..runtime import Event
        g = self.group
        # parameters of this bucket that received no gradient this step still take part with zeros
        for p in b.params:
            if p._grad is None and not p._slot_zero:
                p._grad_slot.fill(0.0)
                p._slot_zero = True
        from .. import kernels as K
        K.flush_wgrad()                                 # queued (grouped) weight-gradient GEMMs of this bucket go out first
        ready = Event(timing=False).record()            # on the compute stream: gradients of this bucket are final
        g.comm_stream.wait_event(ready)
        side_ev = K.side_stream_event()                 # weight gradients computed on the side stream are part of the bucket:
        if side_ev is not None:                         # the COMMUNICATION stream waits for them, the compute stream runs on
            g.comm_stream.wait_event(side_ev)
        b.piece_events = []
        bf16 = self.grad_dtype == "bfloat16"
        if bf16 and self._g16 is None:
            from ..array import bfloat16 as _bf16
            self._g16 = B200Array.empty((max(self.arena.total, 1),), _bf16)
        for lo, hi in b.pieces:
            if not _DEBUG_SKIP_ALLREDUCE:   # measurement aid: same streams / events / optimizer pipeline, no NCCL kernel
                if bf16:
                    from .._lib import call
                    from ..runtime import current_stream_handle
                    main = current_stream_handle()
                    call("nfb_set_stream", g.comm_stream.handle)     # the two casts run on the communication stream, around the all-reduce
                    try:
                        call("nfb_cast", 0, 5, self.arena.flat_g.ptr + lo * 4, self._g16.ptr + lo * 2, hi - lo)
                        g.allreduce_buffer(self._g16, hi - lo, ReduceOp.AVERAGE, g.comm_stream, offset=lo)
                        call("nfb_cast", 5, 0, self._g16.ptr + lo * 2, self.arena.flat_g.ptr + lo * 4, hi - lo)
                    finally:
                        call("nfb_set_stream", main)
                else:
                    g.allreduce_buffer(self.arena.flat_g, hi - lo, ReduceOp.AVERAGE, g.comm_stream, offset=lo)
            b.piece_events.append(Event(timing=False).record(g.comm_stream))
        b.event = b.piece_events[-1]
        b.launched = True
        self._launch_seq = getattr(self, "_launch_seq", 0) + 1
        b.seq = self._launch_seq

    def finish_gradient_sync(self):
        """Fence: launch what has not been launched (unused parameters, overlap off) and make the compute stream
        wait for every bucket.  Call after backward(), before optimizer.step()."""
        if self.world_size == 1 or not self._require_sync:
            return
        from ..runtime import compute_stream_wait
        if self.arena is not None:
            for b in self.buckets:
                if not b.launched:
                    self._launch_bucket(b)
            self._flush_sparse()           # (its wait on the communication stream also covers every bucket launched so far)
            for b in self.buckets:
                compute_stream_wait(b.event)
            for p in self.arena.params:   # every parameter now holds the averaged gradient (zeros if unused)
                if p._grad is None:
                    p._grad = p._grad_slot
                p._slot_zero = False
            self._sync_consumed = True
        for p in self._host_params:
            if p.grad is not None:
                p.grad = self.group.all_reduce(Tensor(p.grad), ReduceOp.AVERAGE).data

    def step_optimizer(self, optimizer):
        """finish_gradient_sync() + optimizer.step() with the fused optimizer kernel pipelined behind the bucket
        all-reduces: slice i of the arena is updated as soon as bucket i has arrived, while later buckets (the big
        tied embedding gradient comes last) are still on the wire."""
        import inspect
        pipelined = "segments" in inspect.signature(optimizer.step).parameters     # Adam / AdamW; SGD and Lion take the plain path
        if self.world_size == 1 or not self._require_sync or self.arena is None or self._host_params or not pipelined or \
                getattr(optimizer, "arena", None) is None or optimizer.arena.flat_g is not self.arena.flat_g or \
                len(optimizer.arena.params) != len(self.arena.params):   # (an optimizer over a SUBSET of the module steps after the full sync)
            self.finish_gradient_sync()
            optimizer.step()
            return
        for b in self.buckets:
            if not b.launched:
                self._launch_bucket(b)
        flushed = bool(self._flush_sparse())   # the compute stream now sits behind every all-reduce issued so far
        for p in self.arena.params:   # every parameter will hold the averaged gradient (zeros if unused)
            if p._grad is None:
                p._grad = p._grad_slot
            p._slot_zero = False
        self._sync_consumed = True
        order = sorted(self.buckets, key=lambda b: getattr(b, "seq", 0))
        optimizer.step(segments=[(lo, hi, None if flushed else ev) for b in order for (lo, hi), ev in zip(b.pieces, b.piece_events)])

    def sync_gradients(self):
        """Reference API (data_parallel.py:271-287): all-reduce(AVERAGE) every gradient after backward."""
        self.finish_gradient_sync()

    @contextmanager
    def no_sync(self):
        prev, self._require_sync = self._require_sync, False
        try:
            yield
        finally:
            self._require_sync = prev

    def parameters(self, recurse: bool = True):
        return self.module.parameters(recurse)

    def named_parameters(self, prefix: str = "", recurse: bool = True, _memo=None):
        return self.module.named_parameters(prefix, recurse, _memo)

    def state_dict(self):
        return self.module.state_dict()

    def load_state_dict(self, sd, strict=True):
        return self.module.load_state_dict(sd, strict)

    def train(self, mode: bool = True):
        self.module.train(mode)
        return self

    def eval(self):
        self.module.eval()
        return self


class DataParallel(DistributedDataParallel):
    """Single-process multi-GPU DataParallel does not exist in a one-process-per-GPU design; this is DDP under the
    reference's other class name (data_parallel.py:21-183), so scripts using either spelling run."""


class DistributedSampler:
    """Rank-strided partition with padding and epoch-seeded shuffling, bit-identical to data_parallel.py:304-388."""

    def __init__(self, dataset_size: int, num_replicas: Optional[int] = None, rank: Optional[int] = None, shuffle: bool = True,
                 seed: int = 0, drop_last: bool = False):
        if num_replicas is None:
            num_replicas = comm.get_world_size() if comm.is_initialized() else 1
        if rank is None:
            rank = comm.get_rank() if comm.is_initialized() else 0
        self.dataset_size, self.num_replicas, self.rank = dataset_size, num_replicas, rank
        self.shuffle, self.seed, self.drop_last, self.epoch = shuffle, seed, drop_last, 0
        if drop_last and dataset_size % num_replicas != 0:
            self.num_samples = dataset_size // num_replicas
        else:
            self.num_samples = (dataset_size + num_replicas - 1) // num_replicas
        self.total_size = self.num_samples * num_replicas

    def __iter__(self):
        if self.shuffle:
            indices = np.random.RandomState(self.seed + self.epoch).permutation(self.dataset_size).tolist()
        else:
            indices = list(range(self.dataset_size))
        if not self.drop_last:
            pad = self.total_size - len(indices)
            if pad <= len(indices):
                indices += indices[:pad]
            else:
                indices += (indices * (pad // len(indices) + 1))[:pad]
        else:
            indices = indices[:self.total_size]
        indices = indices[self.rank:self.total_size:self.num_replicas]
        assert len(indices) == self.num_samples
        return iter(indices)

    def __len__(self):
        return self.num_samples

    def set_epoch(self, epoch: int):
        self.epoch = epoch


# ---------------------------------------------------------------------- module-level helpers (data_parallel.py:391-476)
def gather(tensor, dst: int = 0, async_op: bool = False):
    """All ranks contribute `tensor`; rank `dst` gets the list ordered by rank, the others None (data_parallel.py:391-412:
    an all_gather whose result only `dst` keeps).  Without a process group: `[tensor]`."""
    if not comm.is_initialized():
        return [tensor]
    parts = comm.all_gather(tensor)
    return parts if comm.get_rank() == dst else None


def scatter(tensor_list, src: int = 0):
    """Rank r receives `tensor_list[r]` of rank `src` (data_parallel.py:415-446 -- whose non-source branch returns the
    broadcast of a dummy; this does what its docstring says).  Only `src` needs the list; it must hold one tensor per
    rank (`ValueError` otherwise, as upstream).  Host path over the rendezvous sockets; device tensors are staged through
    the host (control-plane use: the batch itself is sharded by `DistributedSampler`, gradients go through NCCL)."""
    if not comm.is_initialized():
        return tensor_list[0] if tensor_list else Tensor(np.zeros((0,), np.float32))
    g = comm.get_backend()
    mine = None
    if g.rank == src:
        if tensor_list is None or len(tensor_list) != g.world_size:
            raise ValueError("tensor_list must contain tensors for all processes")
        mine = [comm._to_host(t) for t in tensor_list]
    parts = g.tcp.allgather(mine)[src]
    if parts is None:
        raise RuntimeError(f"scatter: rank {src} did not provide a tensor list")
    if g.rank == src:
        return tensor_list[src]
    on_device = g.backend == "nccl"
    return Tensor(parts[g.rank], device="cuda" if on_device else "cpu")


_LIVE_DDP = weakref.WeakSet()


@contextmanager
def no_sync():
    """Suspend gradient synchronisation of every live DistributedDataParallel wrapper (data_parallel.py:449-457 is a
    placeholder that only yields): backward passes inside the block accumulate local gradients, the first backward after
    it all-reduces the sum."""
    wrappers = list(_LIVE_DDP)
    prev = [w._require_sync for w in wrappers]
    for w in wrappers:
        w._require_sync = False
    try:
        yield
    finally:
        for w, pv in zip(wrappers, prev):
            w._require_sync = pv


def is_distributed_training_available() -> bool:
    """data_parallel.py:460-462."""
    return comm.is_initialized() and comm.get_world_size() > 1


def get_distributed_info():
    """data_parallel.py:465-476 (the backend entry names the transport instead of the upstream constant "distributed")."""
    if not comm.is_initialized():
        return {"available": False, "world_size": 1, "rank": 0, "backend": None}
    g = comm.get_backend()
    return {"available": True, "world_size": g.world_size, "rank": g.rank, "backend": g.backend}
