"""Process groups and collectives (mirror of neural_arch.distributed.communication:25-600).

Two real backends (the reference's NCCL path cannot rendezvous -- its unique-id exchange is a stub,
communication.py:339-347 -- and its "gloo" backend only simulates collectives arithmetically, :379-400):

  * "nccl": one NCCL communicator per process/GPU through the C-ABI (csrc/comm.cu, libnccl dlopen'ed); the unique id
    is created on rank 0 and shipped to the other ranks over the TCP rendezvous below.  Collectives run on device
    buffers (B200Array / cuda Tensors) over NVLink 5 / NVSwitch.
  * "gloo": host tensors (NumPy) exchanged over the same TCP rendezvous (gather to rank 0, reduce, broadcast):
    real multi-process collectives for the host-side logic and the CPU tests.

Rendezvous: env:// with RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT (distributed/launcher.py:145-162).
torchrun keeps its own store on MASTER_PORT, so this module listens on MASTER_PORT + NFB200_PORT_OFFSET (default 37).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import hmac
import os
import socket
import struct
import time
from enum import Enum
from typing import List, Optional

import numpy as np

from .. import _lib
from ..array import B200Array, dtype_code
from ..core.tensor import Tensor


class ReduceOp(Enum):
    SUM = "sum"
    PRODUCT = "product"
    MIN = "min"
    MAX = "max"
    AVERAGE = "average"


_NCCL_OP = {ReduceOp.SUM: 0, ReduceOp.AVERAGE: 1, ReduceOp.MAX: 2, ReduceOp.MIN: 3, ReduceOp.PRODUCT: 4}
_MAGIC = b"NFB200RV"


# ---- wire format of the rendezvous / host collectives.  A closed set of value types in a fixed binary framing: nothing
# received from a socket is ever executed or unpickled (any host that can reach the port could otherwise run code in the job).
#   N none | T / F bool | i int64 | f float64 | s utf-8 str | b bytes | a ndarray (dtype str, ndim, dims, raw C-order bytes)
#   l list | t tuple | d dict with str keys
_WIRE_DTYPES = {"float16", "float32", "float64", "int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64", "bool"}


def _encode(obj, out: bytearray):
    if obj is None:
        out += b"N"
    elif isinstance(obj, (bool, np.bool_)):
        out += b"T" if obj else b"F"
    elif isinstance(obj, (int, np.integer)):
        out += b"i" + struct.pack("!q", int(obj))
    elif isinstance(obj, (float, np.floating)):
        out += b"f" + struct.pack("!d", float(obj))
    elif isinstance(obj, str):
        raw = obj.encode("utf-8")
        out += b"s" + struct.pack("!Q", len(raw)) + raw
    elif isinstance(obj, (bytes, bytearray)):
        out += b"b" + struct.pack("!Q", len(obj)) + bytes(obj)
    elif isinstance(obj, np.ndarray):
        if obj.dtype.name not in _WIRE_DTYPES:
            raise TypeError(f"cannot send an ndarray of dtype {obj.dtype} over the rendezvous")
        name = obj.dtype.name.encode()
        raw = np.ascontiguousarray(obj).tobytes()
        out += b"a" + struct.pack("!B", len(name)) + name + struct.pack("!B", obj.ndim) + struct.pack(f"!{obj.ndim}Q", *obj.shape)
        out += struct.pack("!Q", len(raw)) + raw
    elif isinstance(obj, (list, tuple)):
        out += (b"l" if isinstance(obj, list) else b"t") + struct.pack("!Q", len(obj))
        for x in obj:
            _encode(x, out)
    elif isinstance(obj, dict):
        out += b"d" + struct.pack("!Q", len(obj))
        for k, v in obj.items():
            if not isinstance(k, str):
                raise TypeError("only str keys can be sent over the rendezvous")
            _encode(k, out)
            _encode(v, out)
    else:
        raise TypeError(f"cannot send a {type(obj).__name__} over the rendezvous (allowed: None, bool, int, float, str, bytes, ndarray, list, tuple, dict)")


def _decode(buf: memoryview, pos: int, depth: int = 0):
    if depth > 32:
        raise ValueError("rendezvous message nested too deeply")
    tag = bytes(buf[pos:pos + 1])
    pos += 1
    if tag == b"N":
        return None, pos
    if tag in (b"T", b"F"):
        return tag == b"T", pos
    if tag == b"i":
        return struct.unpack_from("!q", buf, pos)[0], pos + 8
    if tag == b"f":
        return struct.unpack_from("!d", buf, pos)[0], pos + 8
    if tag in (b"s", b"b"):
        (n,) = struct.unpack_from("!Q", buf, pos)
        pos += 8
        if n > len(buf) - pos:
            raise ValueError("rendezvous message truncated")
        raw = bytes(buf[pos:pos + n])
        return (raw.decode("utf-8") if tag == b"s" else raw), pos + n
    if tag == b"a":
        (ln,) = struct.unpack_from("!B", buf, pos)
        name = bytes(buf[pos + 1:pos + 1 + ln]).decode()
        pos += 1 + ln
        if name not in _WIRE_DTYPES:
            raise ValueError(f"rendezvous message carries an ndarray of unsupported dtype {name!r}")
        (nd,) = struct.unpack_from("!B", buf, pos)
        shape = struct.unpack_from(f"!{nd}Q", buf, pos + 1)
        pos += 1 + 8 * nd
        (n,) = struct.unpack_from("!Q", buf, pos)
        pos += 8
        dt = np.dtype(name)
        if n != int(np.prod(shape, dtype=np.int64)) * dt.itemsize or n > len(buf) - pos:
            raise ValueError("rendezvous message: ndarray payload does not match its shape")
        arr = np.frombuffer(bytes(buf[pos:pos + n]), dtype=dt).reshape(shape).copy()
        return arr, pos + n
    if tag in (b"l", b"t"):
        (n,) = struct.unpack_from("!Q", buf, pos)
        pos += 8
        if n > len(buf):
            raise ValueError("rendezvous message truncated")
        items = []
        for _ in range(n):
            x, pos = _decode(buf, pos, depth + 1)
            items.append(x)
        return (items if tag == b"l" else tuple(items)), pos
    if tag == b"d":
        (n,) = struct.unpack_from("!Q", buf, pos)
        pos += 8
        if n > len(buf):
            raise ValueError("rendezvous message truncated")
        d = {}
        for _ in range(n):
            k, pos = _decode(buf, pos, depth + 1)
            if not isinstance(k, str):
                raise ValueError("rendezvous message: dict key is not a str")
            d[k], pos = _decode(buf, pos, depth + 1)
        return d, pos
    raise ValueError(f"rendezvous message: unknown type tag {tag!r}")


def _send_msg(sock, obj):
    data = bytearray()
    _encode(obj, data)
    sock.sendall(struct.pack("!Q", len(data)) + bytes(data))


def _recv_exact(sock, n):
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(1 << 20, n - len(buf)))
        if not chunk:
            raise ConnectionError("rendezvous peer closed the connection")
        buf += chunk
    return bytes(buf)


def _recv_msg(sock):
    (n,) = struct.unpack("!Q", _recv_exact(sock, 8))
    obj, pos = _decode(memoryview(_recv_exact(sock, n)), 0)
    if pos != n:
        raise ValueError("rendezvous message has trailing bytes")
    return obj


def _job_token() -> bytes:
    """Shared secret of the job: NFB200_RDZV_TOKEN (the launcher draws a random one per job and exports it to every rank);
    under other launchers the run id they export, else the rendezvous endpoint itself (guards against cross-talk only)."""
    tok = os.environ.get("NFB200_RDZV_TOKEN") or os.environ.get("TORCHELASTIC_RUN_ID") or ""
    return (tok + "|" + os.environ.get("MASTER_ADDR", "127.0.0.1") + ":" + os.environ.get("MASTER_PORT", "29500")).encode()


def _hello_mac(rank: int, world: int) -> bytes:
    return hmac.new(_job_token(), _MAGIC + struct.pack("!II", rank, world), hashlib.sha256).digest()


class _TcpGroup:
    """Star topology over TCP: rank 0 holds one socket per peer.  Used for rendezvous and host collectives."""

    def __init__(self, rank: int, world: int, addr: str, port: int, timeout: float):
        self.rank, self.world = rank, world
        self.peers: List[Optional[socket.socket]] = [None] * world
        self.sock: Optional[socket.socket] = None
        if world == 1:
            return
        if rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr if addr not in ("localhost",) else "127.0.0.1", port))
            srv.listen(world)
            srv.settimeout(timeout)
            got = 0
            while got < world - 1:
                conn, _ = srv.accept()
                conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                try:
                    conn.settimeout(10)
                    hello = _recv_exact(conn, len(_MAGIC) + 4 + 32)
                    conn.settimeout(None)
                except (OSError, ConnectionError):
                    conn.close()
                    continue
                (r,) = struct.unpack("!I", hello[len(_MAGIC):len(_MAGIC) + 4])
                # a peer is admitted only with the magic, a rank in 1..world-1 that has not checked in yet, and the job's MAC
                if hello[:len(_MAGIC)] != _MAGIC or not (0 < r < world) or self.peers[r] is not None or \
                        not hmac.compare_digest(hello[len(_MAGIC) + 4:], _hello_mac(r, world)):
                    conn.close()
                    continue
                self.peers[r] = conn
                got += 1
            srv.close()
        else:
            deadline = time.time() + timeout
            while True:
                try:
                    s = socket.create_connection((addr, port), timeout=5)
                    s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                    s.sendall(_MAGIC + struct.pack("!I", rank) + _hello_mac(rank, world))
                    s.settimeout(timeout)
                    self.sock = s
                    break
                except OSError:
                    if time.time() > deadline:
                        raise TimeoutError(f"rank {rank}: cannot reach the rendezvous at {addr}:{port}")
                    time.sleep(0.2)

    def gather(self, obj):
        """Everyone sends; rank 0 returns the list [obj_0, ..., obj_{w-1}], the others None."""
        if self.world == 1:
            return [obj]
        if self.rank == 0:
            out = [obj] + [None] * (self.world - 1)
            for r in range(1, self.world):
                out[r] = _recv_msg(self.peers[r])
            return out
        _send_msg(self.sock, obj)
        return None

    def bcast(self, obj):
        if self.world == 1:
            return obj
        if self.rank == 0:
            for r in range(1, self.world):
                _send_msg(self.peers[r], obj)
            return obj
        return _recv_msg(self.sock)

    def allgather(self, obj):
        return self.bcast(self.gather(obj))

    def barrier(self):
        self.allgather(None)

    def close(self):
        for s in self.peers + [self.sock]:
            if s is not None:
                try:
                    s.close()
                except OSError:
                    pass


class ProcessGroup:
    def __init__(self, backend: str, rank: int, world_size: int, timeout: float = 1800):
        self.backend, self.rank, self.world_size = backend, rank, world_size
        addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(os.environ.get("MASTER_PORT", "29500")) + int(os.environ.get("NFB200_PORT_OFFSET", "37"))
        self.tcp = _TcpGroup(rank, world_size, addr, port, timeout)
        self.comm = None
        self.comm_stream = None
        if backend == "nccl":
            from ..runtime import Stream
            local = int(os.environ.get("LOCAL_RANK", str(rank)))
            _lib.init(local)
            path = os.environ.get("NFB200_NCCL_LIB", "")
            if not path:
                try:
                    import nvidia.nccl as _n  # the pip wheel that ships libnccl.so.2 in this image
                    cand = os.path.join(list(_n.__path__)[0], "lib", "libnccl.so.2")
                    path = cand if os.path.exists(cand) else ""
                except Exception:
                    path = ""
            _lib.call("nfb_comm_load", path.encode())
            uid = C.create_string_buffer(128)
            if rank == 0:
                _lib.call("nfb_comm_unique_id", uid, 128)
            raw = self.tcp.bcast(bytes(uid.raw))
            h = C.c_void_p()
            _lib.call("nfb_comm_init", rank, world_size, raw, 128, C.byref(h))
            self.comm = h.value
            self.comm_stream = Stream()
        elif backend != "gloo":
            raise ValueError(f"Unsupported backend: {backend}")

    # ------------------------------------------------------------------ device collectives (raw buffers)
    def allreduce_buffer(self, arr: B200Array, count: Optional[int] = None, op: ReduceOp = ReduceOp.SUM, stream=None, offset: int = 0):
        """In-place all-reduce of `count` elements of a contiguous device buffer starting at element `offset`."""
        n = arr.size - offset if count is None else count
        ptr = arr.ptr + offset * arr.itemsize
        _lib.call("nfb_comm_allreduce", self.comm, ptr, ptr, n, dtype_code(arr.dtype), _NCCL_OP[op],
                  stream.handle if stream is not None else None)

    def broadcast_buffer(self, arr: B200Array, src: int = 0, stream=None):
        _lib.call("nfb_comm_broadcast", self.comm, arr.ptr, arr.size, dtype_code(arr.dtype), src, stream.handle if stream is not None else None)

    # ------------------------------------------------------------------ tensor-level API
    def _host_reduce(self, arrs, op):
        a = np.stack(arrs, 0)
        if op is ReduceOp.SUM:
            return a.sum(0, dtype=arrs[0].dtype)
        if op is ReduceOp.AVERAGE:
            return (a.sum(0, dtype=np.float64) / len(arrs)).astype(arrs[0].dtype)
        if op is ReduceOp.MAX:
            return a.max(0)
        if op is ReduceOp.MIN:
            return a.min(0)
        return a.prod(0)

    def all_reduce(self, tensor, op: ReduceOp = ReduceOp.SUM):
        data = tensor.data if isinstance(tensor, Tensor) else tensor
        if isinstance(data, B200Array):
            if self.backend != "nccl":
                raise RuntimeError("device tensors need the nccl backend")
            buf = data if data.is_contiguous else data.contiguous()
            self.allreduce_buffer(buf, op=op)
            if buf is not data:
                data._assign(buf)
            return tensor
        arr = np.ascontiguousarray(data)
        parts = self.tcp.gather(arr)
        red = self.tcp.bcast(self._host_reduce(parts, op) if self.rank == 0 else None)
        if isinstance(tensor, Tensor):
            tensor.data = red
            return tensor
        data[...] = red
        return data

    def all_gather(self, tensor):
        data = tensor.data if isinstance(tensor, Tensor) else tensor
        if isinstance(data, B200Array):
            src = data.contiguous()
            out = B200Array.empty((self.world_size,) + src.shape, src.dtype)
            _lib.call("nfb_comm_allgather", self.comm, src.ptr, out.ptr, src.size, dtype_code(src.dtype), None)
            return [Tensor(out[i]) for i in range(self.world_size)]
        parts = self.tcp.allgather(np.ascontiguousarray(data))
        return [Tensor(p) for p in parts]

    def reduce_scatter(self, tensor, op: ReduceOp = ReduceOp.SUM):
        data = tensor.data if isinstance(tensor, Tensor) else tensor
        if data.shape[0] % self.world_size:
            raise ValueError(f"reduce_scatter: dim 0 ({data.shape[0]}) must be divisible by the world size ({self.world_size})")
        chunk = data.shape[0] // self.world_size
        if isinstance(data, B200Array):
            src = data.contiguous()
            out = B200Array.empty((chunk,) + src.shape[1:], src.dtype)
            _lib.call("nfb_comm_reduce_scatter", self.comm, src.ptr, out.ptr, out.size, dtype_code(src.dtype), _NCCL_OP[op], None)
            return Tensor(out)
        parts = self.tcp.gather(np.ascontiguousarray(data))
        red = self.tcp.bcast(self._host_reduce(parts, op) if self.rank == 0 else None)
        return Tensor(red[self.rank * chunk:(self.rank + 1) * chunk])

    def broadcast(self, tensor, src: int = 0):
        data = tensor.data if isinstance(tensor, Tensor) else tensor
        if isinstance(data, B200Array):
            buf = data if data.is_contiguous else data.contiguous()
            self.broadcast_buffer(buf, src)
            if buf is not data:
                data._assign(buf)
            return tensor
        parts = self.tcp.allgather(np.ascontiguousarray(data))
        if isinstance(tensor, Tensor):
            tensor.data = parts[src]
            return tensor
        data[...] = parts[src]
        return data

    def barrier(self):
        if self.backend == "nccl":
            from ..runtime import synchronize
            synchronize()
        self.tcp.barrier()

    def destroy(self):
        if self.comm is not None:
            try:
                _lib.call("nfb_comm_destroy", self.comm)
            except Exception:
                pass
            self.comm = None
        self.tcp.close()


_group: Optional[ProcessGroup] = None


def init_process_group(backend: str = "auto", init_method: str = "env://", world_size: Optional[int] = None,
                       rank: Optional[int] = None, timeout: int = 1800) -> ProcessGroup:
    global _group
    if world_size is None:
        world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if rank is None:
        rank = int(os.environ.get("RANK", "0"))
    if backend == "auto":
        backend = "nccl" if _lib.device_available() else "gloo"
    if backend not in ("nccl", "gloo"):
        raise ValueError(f"Unsupported backend: {backend}")
    if init_method.startswith("tcp://"):
        host, _, port = init_method[len("tcp://"):].partition(":")
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = host, port or "29500"
    _group = ProcessGroup(backend, rank, world_size, timeout)
    return _group


def destroy_process_group():
    global _group
    if _group is not None:
        _group.destroy()
    _group = None


def get_backend() -> ProcessGroup:
    if _group is None:
        raise RuntimeError("Process group not initialized. Call init_process_group() first.")
    return _group


def is_initialized() -> bool:
    return _group is not None


def is_available() -> bool:
    return True


def get_world_size() -> int:
    return get_backend().world_size


def get_rank() -> int:
    return get_backend().rank


def barrier(timeout: Optional[int] = None):
    get_backend().barrier()


def all_reduce(tensor, op: ReduceOp = ReduceOp.SUM):
    return get_backend().all_reduce(tensor, op)


def all_gather(tensor):
    return get_backend().all_gather(tensor)


def reduce_scatter(tensor, op: ReduceOp = ReduceOp.SUM):
    return get_backend().reduce_scatter(tensor, op)


def broadcast(tensor, src: int = 0):
    return get_backend().broadcast(tensor, src)


def _to_host(tensor):
    data = tensor.data if isinstance(tensor, Tensor) else tensor
    return data.get() if isinstance(data, B200Array) else np.ascontiguousarray(data)


def _store(tensor, host: np.ndarray):
    """Write a received host array into `tensor` (device tensors get one host->device copy) and return it."""
    if isinstance(tensor, Tensor):
        tensor.data = B200Array.from_numpy(host) if isinstance(tensor.data, B200Array) else host
        return tensor
    if isinstance(tensor, B200Array):
        tensor._assign(B200Array.from_numpy(host))
        return tensor
    tensor[...] = host
    return tensor


def send(tensor, dst: int, tag: int = 0):
    """Point-to-point send (communication.py:317-321, 585-587; the reference fakes it with a broadcast).  Here the payload
    really travels over the rendezvous sockets: a star with rank 0 in the middle, so one end of every pair must be rank 0
    (control-plane traffic such as metrics or shards of a checkpoint; gradients go through the NCCL collectives)."""
    g = get_backend()
    if dst == g.rank:
        raise ValueError("send: destination is this rank")
    if not 0 <= dst < g.world_size:
        raise ValueError(f"send: rank {dst} is out of range for world size {g.world_size}")
    if g.rank != 0 and dst != 0:
        raise NotImplementedError("send/recv between two non-zero ranks is not routed; use broadcast / all_gather")
    _send_msg(g.tcp.sock if g.rank != 0 else g.tcp.peers[dst], (int(tag), _to_host(tensor)))


def recv(tensor, src: int, tag: int = 0):
    """Point-to-point receive into `tensor` (communication.py:323-326, 590-592); returns it."""
    g = get_backend()
    if src == g.rank:
        raise ValueError("recv: source is this rank")
    if not 0 <= src < g.world_size:
        raise ValueError(f"recv: rank {src} is out of range for world size {g.world_size}")
    if g.rank != 0 and src != 0:
        raise NotImplementedError("send/recv between two non-zero ranks is not routed; use broadcast / all_gather")
    got_tag, host = _recv_msg(g.tcp.sock if g.rank != 0 else g.tcp.peers[src])
    if got_tag != int(tag):
        raise RuntimeError(f"recv: expected tag {tag} from rank {src}, got {got_tag}")
    return _store(tensor, host)


def all_reduce_max_host(value: float) -> float:
    """max over ranks of a host scalar (bench timing: the job's time is the slowest rank's)."""
    g = get_backend()
    vals = g.tcp.allgather(float(value))
    return max(vals)
