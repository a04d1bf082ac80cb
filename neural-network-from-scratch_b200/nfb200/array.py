"""B200Array: a device-resident, NumPy-protocol duck array backed by libnfb200.

The reference's Tensor/functional/nn code touches ``tensor.data`` with NumPy syntax everywhere
(``x.data + y.data``, ``np.mean(x.data, axis=-1, keepdims=True)``, ``w.data[idx]``,
``grad[rows, cols] -= 1`` ...), so the array a backend returns is the wide part of the plugin
boundary.  B200Array implements that surface (operators, ``__array_ufunc__`` / ``__array_function__``
(NEP-13/18), basic + row-gather indexing, reshape/transpose views, reductions) on HBM-resident
buffers; every operation is a CUDA kernel launch through the C-ABI.  Anything unsupported raises
-- there is no silent NumPy fallback, and implicit ``np.asarray(x)`` is refused (use ``.get()``).
"""
from __future__ import annotations

import ctypes as C
import math
import numbers

import numpy as np

from . import _lib
from ._lib import B200Error, call


class _BFloat16:
    """Marker dtype for the backend-internal bf16 storage (the reference has no 16-bit DType)."""

    name = "bfloat16"
    itemsize = 2
    kind = "V"

    def __repr__(self):
        return "bfloat16"

    def __eq__(self, other):
        return other is self or other == "bfloat16"

    def __hash__(self):
        return hash("bfloat16")


bfloat16 = _BFloat16()

_CODE = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64, np.dtype(np.int32): _lib.I32,
         np.dtype(np.int64): _lib.I64, np.dtype(np.bool_): _lib.BOOL}


# implicit device -> host conversion through np.asarray(array) / __array__: off (TypeError, like CuPy) unless a caller that
# runs UNCHANGED reference code opts in; see B200Array.__array__
_IMPLICIT_HOST = {"allow": False, "count": 0}


def set_implicit_host_conversion(allow: bool) -> int:
    """Allow / forbid np.asarray(B200Array).  Returns the number of implicit host copies made so far."""
    _IMPLICIT_HOST["allow"] = bool(allow)
    return _IMPLICIT_HOST["count"]


def implicit_host_conversions() -> int:
    return _IMPLICIT_HOST["count"]


def dtype_code(dt) -> int:
    if dt is bfloat16 or dt == "bfloat16":
        return _lib.BF16
    try:
        return _CODE[np.dtype(dt)]
    except KeyError:
        raise TypeError(f"dtype {dt} is not supported on the b200 backend (float32/float64/int32/int64/bool)") from None


def _norm_dtype(dt):
    if dt is None:
        return None
    if dt is bfloat16 or (isinstance(dt, str) and dt == "bfloat16"):
        return bfloat16
    if hasattr(dt, "numpy_dtype"):  # core.DType
        dt = dt.numpy_dtype
    d = np.dtype(dt)
    if d not in _CODE:
        raise TypeError(f"dtype {d} is not supported on the b200 backend")
    return d


def _itemsize(dt) -> int:
    return 2 if dt is bfloat16 else np.dtype(dt).itemsize


class _Buffer:
    """Owns one device allocation from the caching allocator; freed on refcount drop."""

    __slots__ = ("ptr", "nbytes", "__weakref__")

    def __init__(self, nbytes: int):
        _lib.ensure_init()
        out = C.c_void_p()
        call("nfb_malloc", max(int(nbytes), 1), C.byref(out))
        self.ptr = out.value
        self.nbytes = int(nbytes)

    def __del__(self):
        p, self.ptr = getattr(self, "ptr", None), None
        if p:
            try:
                _lib.load().nfb_free(C.c_void_p(p))
            except Exception:  # interpreter shutdown
                pass


def _contig_strides(shape):
    st, e = [], 1
    for s in reversed(shape):
        st.append(e)
        e *= max(int(s), 1) if s != 0 else 1
    return tuple(reversed(st))


def _prod(xs):
    n = 1
    for x in xs:
        n *= int(x)
    return n


BIN = {"add": 0, "sub": 1, "mul": 2, "div": 3, "pow": 4, "max": 5, "min": 6,
       "eq": 10, "ne": 11, "lt": 12, "le": 13, "gt": 14, "ge": 15}
UN = {"neg": 0, "exp": 1, "log": 2, "sqrt": 3, "abs": 4, "sign": 5, "tanh": 6, "sigmoid": 7, "relu": 8,
      "gelu_erf": 9, "gelu_tanh": 10, "silu": 11, "square": 12, "recip": 13, "clip": 14, "affine": 15,
      "pows": 16, "rsqrt": 17, "nan_to_num": 18, "sin": 19, "cos": 20, "floor": 21}
RED = {"sum": 0, "mean": 1, "max": 2, "min": 3, "prod": 4}


class B200Array:
    __slots__ = ("_buf", "_off", "shape", "strides", "dtype", "clipped", "twin", "__weakref__")
    __array_priority__ = 1000.0

    # ------------------------------------------------------------------ construction
    def __init__(self, buf: _Buffer, shape, dtype, strides=None, offset=0):
        self._buf = buf
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype
        self.strides = tuple(strides) if strides is not None else _contig_strides(self.shape)  # in ELEMENTS
        self._off = int(offset)  # in elements
        self.clipped = False  # gradient already clamped to +-clip by the kernel that produced it
        self.twin = None      # optional bf16 copy written by the producing kernel (feeds tensor-core GEMMs)

    @staticmethod
    def empty(shape, dtype=np.float32) -> "B200Array":
        if isinstance(shape, numbers.Integral):
            shape = (int(shape),)
        dtype = _norm_dtype(dtype) or np.dtype(np.float32)
        shape = tuple(int(s) for s in shape)
        if any(s < 0 for s in shape):
            raise ValueError("negative dimensions are not allowed")
        return B200Array(_Buffer(_prod(shape) * _itemsize(dtype)), shape, dtype)

    @staticmethod
    def from_numpy(a, dtype=None) -> "B200Array":
        if isinstance(a, B200Array):
            return a.astype(dtype) if dtype is not None and _norm_dtype(dtype) != a.dtype else a
        arr = np.asarray(a)
        want = _norm_dtype(dtype)
        if want is bfloat16:
            return B200Array.from_numpy(arr.astype(np.float32)).astype(bfloat16)
        if want is not None:
            arr = arr.astype(want, copy=False)
        elif arr.dtype not in _CODE:
            if arr.dtype.kind == "f":
                arr = arr.astype(np.float32)
            elif arr.dtype.kind in "iu":
                arr = arr.astype(np.int64)
            else:
                raise TypeError(f"cannot move dtype {arr.dtype} to the b200 backend")
        arr = np.ascontiguousarray(arr)
        out = B200Array.empty(arr.shape, arr.dtype)
        if arr.size:
            call("nfb_memcpy_h2d", out.ptr, arr.ctypes.data, arr.nbytes)
        return out

    @staticmethod
    def full(shape, value, dtype=np.float32) -> "B200Array":
        out = B200Array.empty(shape, dtype)
        if out.size:
            call("nfb_fill", dtype_code(out.dtype), out.ptr, out.size, float(value))
        return out

    # ------------------------------------------------------------------ basic properties
    @property
    def ptr(self) -> int:
        return self._buf.ptr + self._off * _itemsize(self.dtype)

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:
        return _prod(self.shape)

    @property
    def itemsize(self) -> int:
        return _itemsize(self.dtype)

    @property
    def nbytes(self) -> int:
        return self.size * self.itemsize

    @property
    def device(self) -> str:
        return f"cuda:{_lib.ensure_init()}"

    @property
    def T(self) -> "B200Array":
        return self.transpose()

    @property
    def is_contiguous(self) -> bool:
        e = 1
        for s, st in zip(reversed(self.shape), reversed(self.strides)):
            if s != 1 and st != e:
                return False
            e *= s
        return True

    def __len__(self):
        if not self.shape:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def __repr__(self):
        return f"B200Array(shape={self.shape}, dtype={getattr(self.dtype, 'name', self.dtype)}, device='{self.device}')"

    # ------------------------------------------------------------------ host transfer
    def get(self) -> np.ndarray:
        """Device -> host copy (CuPy-style name, probed by functional/activation.py:47-48)."""
        src = self.contiguous()
        if src.dtype is bfloat16:
            src = src.astype(np.float32)
        out = np.empty(src.shape, dtype=src.dtype)
        if out.size:
            call("nfb_memcpy_d2h", out.ctypes.data, src.ptr, out.nbytes)
        return out

    def __array__(self, dtype=None, copy=None):
        if _IMPLICIT_HOST["allow"]:
            # opt-in (set_implicit_host_conversion): the reference's own optimizers ask for NumPy arrays themselves --
            # `param.data = np.asarray(new_data, dtype=...)` (optim/adam.py:245) -- which np.asarray can only satisfy with a
            # host copy; the Tensor.data setter moves the result back into the parameter's device buffer.  Counted, so a
            # caller can see how many round trips unchanged reference code costs.
            _IMPLICIT_HOST["count"] += 1
            out = self.get()
            return out if dtype is None else out.astype(dtype, copy=False)
        raise TypeError("Implicit conversion of a B200Array to a NumPy array is not allowed; "
                        "call .get() or backend.to_numpy() to copy it to the host explicitly.")

    def item(self):
        if self.size != 1:
            raise ValueError("can only convert an array of size 1 to a Python scalar")
        return self.get().reshape(()).item()

    def tolist(self):
        return self.get().tolist()

    def __float__(self):
        return float(self.item())

    def __int__(self):
        return int(self.item())

    def __bool__(self):
        if self.size != 1:
            raise ValueError("The truth value of an array with more than one element is ambiguous.")
        return bool(self.item())

    def __index__(self):
        if self.dtype is bfloat16 or np.dtype(self.dtype).kind not in "iu" or self.size != 1:
            raise TypeError("only integer scalar arrays can be converted to a scalar index")
        return int(self.item())

    # ------------------------------------------------------------------ layout
    def contiguous(self) -> "B200Array":
        if self.is_contiguous:
            return self
        out = B200Array.empty(self.shape, self.dtype)
        if out.size:
            n = self.ndim
            call("nfb_copy_strided", self.itemsize, self.ptr, out.ptr, n, _lib.i64_array(self.shape),
                 _lib.i64_array(self.strides), _lib.i64_array(out.strides))
        return out

    def copy(self) -> "B200Array":
        if not self.is_contiguous:
            return self.contiguous()
        out = B200Array.empty(self.shape, self.dtype)
        if out.size:
            call("nfb_memcpy_d2d", out.ptr, self.ptr, self.nbytes)
        return out

    def reshape(self, *shape, order="C") -> "B200Array":
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = [int(s) for s in shape]
        if shape.count(-1) > 1:
            raise ValueError("can only specify one unknown dimension")
        if -1 in shape:
            known = _prod([s for s in shape if s != -1])
            if known == 0 or self.size % known:
                raise ValueError(f"cannot reshape array of size {self.size} into shape {tuple(shape)}")
            shape[shape.index(-1)] = self.size // known
        if _prod(shape) != self.size:
            raise ValueError(f"cannot reshape array of size {self.size} into shape {tuple(shape)}")
        if tuple(shape) == self.shape:
            return self
        if not self.is_contiguous and self.ndim >= 2 and shape and shape[-1] == self.shape[-1] and self.strides[-1] == 1:
            # row-padded arrays (e.g. logits with a 16-byte-aligned pitch): regrouping the leading dims stays a view
            ld = None
            e, ok = 1, True
            for s, st in zip(reversed(self.shape[:-1]), reversed(self.strides[:-1])):
                if s == 1:
                    continue
                if ld is None:
                    ld = st
                if st != ld * e:
                    ok = False
                    break
                e *= s
            if ok and ld is not None:
                st_new, e = [1], 1
                for s in reversed(shape[:-1]):
                    st_new.append(ld * e)
                    e *= s
                return B200Array(self._buf, shape, self.dtype, tuple(reversed(st_new)), self._off)
        src = self if self.is_contiguous else self.contiguous()
        return B200Array(src._buf, shape, src.dtype, None, src._off)

    def flatten(self) -> "B200Array":
        return self.copy().reshape(-1)

    def ravel(self) -> "B200Array":
        return self.reshape(-1)

    def transpose(self, *axes) -> "B200Array":
        if len(axes) == 1 and (axes[0] is None or isinstance(axes[0], (tuple, list))):
            axes = axes[0]
        if axes is None or len(axes) == 0:
            axes = tuple(reversed(range(self.ndim)))
        axes = tuple(a + self.ndim if a < 0 else a for a in axes)
        if sorted(axes) != list(range(self.ndim)):
            raise ValueError("axes don't match array")
        return B200Array(self._buf, [self.shape[a] for a in axes], self.dtype, [self.strides[a] for a in axes], self._off)

    def swapaxes(self, a, b) -> "B200Array":
        ax = list(range(self.ndim))
        ax[a], ax[b] = ax[b], ax[a]
        return self.transpose(ax)

    def squeeze(self, axis=None) -> "B200Array":
        if axis is None:
            keep = [i for i, s in enumerate(self.shape) if s != 1]
        else:
            axs = (axis,) if isinstance(axis, numbers.Integral) else tuple(axis)
            axs = tuple(a + self.ndim if a < 0 else a for a in axs)
            for a in axs:
                if self.shape[a] != 1:
                    raise ValueError("cannot select an axis to squeeze out which has size not equal to one")
            keep = [i for i in range(self.ndim) if i not in axs]
        return B200Array(self._buf, [self.shape[i] for i in keep], self.dtype, [self.strides[i] for i in keep], self._off)

    def expand_dims(self, axis) -> "B200Array":
        nd = self.ndim + 1
        if axis < 0:
            axis += nd
        if not 0 <= axis < nd:
            raise ValueError(f"axis {axis} is out of bounds for array of dimension {nd}")
        shape = list(self.shape)
        strides = list(self.strides)
        shape.insert(axis, 1)
        strides.insert(axis, 0)
        return B200Array(self._buf, shape, self.dtype, strides, self._off)

    def broadcast_to(self, shape) -> "B200Array":
        shape = tuple(int(s) for s in shape)
        nd = len(shape)
        if nd < self.ndim:
            raise ValueError("cannot broadcast to fewer dimensions")
        sh = (1,) * (nd - self.ndim) + self.shape
        st = (0,) * (nd - self.ndim) + self.strides
        out_st = []
        for s_have, s_want, stv in zip(sh, shape, st):
            if s_have == s_want:
                out_st.append(stv)
            elif s_have == 1:
                out_st.append(0)
            else:
                raise ValueError(f"operands could not be broadcast together with shapes {self.shape} {shape}")
        return B200Array(self._buf, shape, self.dtype, out_st, self._off)

    def astype(self, dtype, copy=True) -> "B200Array":
        dt = _norm_dtype(dtype)
        if dt == self.dtype or (dt is bfloat16 and self.dtype is bfloat16):
            return self.copy() if copy else self
        src = self.contiguous()
        if dt is bfloat16 and src.dtype != np.dtype(np.float32):
            src = src.astype(np.float32)
        if src.dtype is bfloat16 and dt != np.dtype(np.float32):
            src = src.astype(np.float32)
            if dt == src.dtype:
                return src
        out = B200Array.empty(src.shape, dt)
        if out.size:
            call("nfb_cast", dtype_code(src.dtype), dtype_code(dt), src.ptr, out.ptr, out.size)
        return out

    # ------------------------------------------------------------------ indexing
    def _basic_index(self, key):
        """ints / slices / None / Ellipsis -> (shape, strides, offset) of a view, or None."""
        if not isinstance(key, tuple):
            key = (key,)
        if any(isinstance(k, B200Array) and k.ndim == 0 and k.dtype is not bfloat16 and np.dtype(k.dtype).kind in "iu" for k in key):
            key = tuple(int(k.item()) if (isinstance(k, B200Array) and k.ndim == 0) else k for k in key)   # a device scalar used as an index
        if any(isinstance(k, np.ndarray) and k.ndim == 0 and k.dtype.kind in "iu" for k in key):
            key = tuple(int(k) if (isinstance(k, np.ndarray) and k.ndim == 0) else k for k in key)
        if any(isinstance(k, (list, np.ndarray, B200Array)) for k in key):
            return None
        n_specified = sum(1 for k in key if k is not None and k is not Ellipsis)
        if n_specified > self.ndim:
            raise IndexError("too many indices for array")
        if sum(1 for k in key if k is Ellipsis) > 1:
            raise IndexError("an index can only have a single ellipsis ('...')")
        expanded = []
        for k in key:
            if k is Ellipsis:
                expanded.extend([slice(None)] * (self.ndim - n_specified))
            else:
                expanded.append(k)
        if not any(k is Ellipsis for k in key):
            expanded.extend([slice(None)] * (self.ndim - n_specified))
        shape, strides, off, dim = [], [], self._off, 0
        for k in expanded:
            if k is None:
                shape.append(1)
                strides.append(0)
            elif isinstance(k, slice):
                start, stop, step = k.indices(self.shape[dim])
                n = max(0, (stop - start + (step - (1 if step > 0 else -1))) // step)
                shape.append(n)
                strides.append(self.strides[dim] * step)
                off += start * self.strides[dim]
                dim += 1
            elif isinstance(k, numbers.Integral):
                i = int(k)
                if i < 0:
                    i += self.shape[dim]
                if not 0 <= i < self.shape[dim]:
                    raise IndexError(f"index {k} is out of bounds for axis {dim} with size {self.shape[dim]}")
                off += i * self.strides[dim]
                dim += 1
            else:
                raise IndexError(f"unsupported index type {type(k).__name__}")
        return shape, strides, off

    def __getitem__(self, key) -> "B200Array":
        basic = self._basic_index(key)
        if basic is not None:
            shape, strides, off = basic
            return B200Array(self._buf, shape, self.dtype, strides, off)
        # advanced indexing: a single integer index array on axis 0 (embedding-style row gather), or one integer array per
        # axis of a 2-D array (x[rows, cols]: the loss picks p[i, target_i], functional/loss.py:55-60)
        if isinstance(key, tuple):
            if len(key) == 1:
                key = key[0]
            else:
                flat = self._pair_index(key)
                return take_rows(self.contiguous().reshape(-1, 1), flat).reshape(flat.shape)
        idx = key if isinstance(key, B200Array) else B200Array.from_numpy(np.asarray(key))
        if idx.dtype is bfloat16 or np.dtype(idx.dtype).kind == "b":
            raise IndexError("b200: boolean-mask indexing is not supported on device")
        if np.dtype(idx.dtype).kind not in "iu":
            raise IndexError("arrays used as indices must be of integer (or boolean) type")
        return take_rows(self, idx)

    def _pair_index(self, key) -> "B200Array":
        """x[rows, cols] on a 2-D array with two 1-D integer index arrays (host or device) -> flat int64 element indices."""
        if self.ndim != 2 or len(key) != 2:
            raise IndexError("b200: advanced indexing supports x[index_array] and, on 2-D arrays, x[row_indices, col_indices]")
        parts = []
        for k, n in zip(key, self.shape):
            if isinstance(k, numbers.Integral):
                k = np.asarray([int(k)])
            k = k if isinstance(k, B200Array) else B200Array.from_numpy(np.asarray(k))
            if k.dtype is bfloat16 or np.dtype(k.dtype).kind not in "iu":
                raise IndexError("arrays used as indices must be of integer type")
            k = k.astype(np.int64) if k.dtype != np.dtype(np.int64) else k
            parts.append(where(binary("lt", k, 0), binary("add", k, n), k))     # NumPy-style negative indices
        rows, cols = parts
        return binary("add", binary("mul", rows, self.shape[1]), cols)

    def __setitem__(self, key, value):
        basic = self._basic_index(key)
        if basic is None and isinstance(key, tuple) and len(key) == 2:
            # x[rows, cols] = v with UNIQUE (row, col) pairs (`grad[arange(B), targets] -= 1`, functional/loss.py:93): written as
            # x += scatter(v - x[rows, cols]), which equals the assignment when no pair repeats (with repeats NumPy keeps the
            # last write; the reference never relies on that here)
            if self.dtype != np.dtype(np.float32) or not self.is_contiguous:
                raise IndexError("b200: x[rows, cols] = v is supported on contiguous float32 arrays")
            flat = self._pair_index(key)
            cur = take_rows(self.reshape(-1, 1), flat).reshape(flat.shape)
            val = value if isinstance(value, B200Array) else B200Array.from_numpy(np.broadcast_to(np.asarray(value, np.float32), flat.shape).copy())
            delta = binary("sub", val.astype(np.float32) if val.dtype != np.dtype(np.float32) else val, cur).contiguous()
            call("nfb_scatter_add_rows", dtype_code(flat.dtype), dtype_code(delta.dtype), delta.ptr, flat.contiguous().ptr, self.ptr, flat.size, 1, self.size, 1)
            return
        if basic is None:
            raise IndexError("b200: only basic (slice/int) indexing and x[rows, cols] are supported for assignment on device")
        shape, strides, off = basic
        view = B200Array(self._buf, shape, self.dtype, strides, off)
        view._assign(value)

    def _assign(self, value):
        if self.size == 0:
            return
        if isinstance(value, (numbers.Number, np.generic)):
            if self.is_contiguous:
                call("nfb_fill", dtype_code(self.dtype), self.ptr, self.size, float(value))
                return
            value = B200Array.full((), value, self.dtype)
        if not isinstance(value, B200Array):
            value = B200Array.from_numpy(np.asarray(value))
        if value.dtype != self.dtype:
            value = value.astype(self.dtype)
        src = value.broadcast_to(self.shape)
        call("nfb_copy_strided", self.itemsize, src.ptr, self.ptr, self.ndim, _lib.i64_array(self.shape),
             _lib.i64_array(src.strides), _lib.i64_array(self.strides))

    def fill(self, value):
        self._assign(value)

    def __iter__(self):
        if not self.shape:
            raise TypeError("iteration over a 0-d array")
        if self.ndim == 1 and self.dtype is not bfloat16 and np.dtype(self.dtype).kind in "iub":
            # an index loop (`for i, idx in enumerate(flat_indices)`, nn/embedding.py:52-55) wants host integers: ONE copy
            yield from self.get().tolist()
            return
        for i in range(self.shape[0]):
            yield self[i]

    # ------------------------------------------------------------------ arithmetic
    def _binary(self, other, op: str, reverse=False):
        return binary(op, other, self) if reverse else binary(op, self, other)

    def __add__(self, o): return self._binary(o, "add")
    def __radd__(self, o): return self._binary(o, "add", True)
    def __sub__(self, o): return self._binary(o, "sub")
    def __rsub__(self, o): return self._binary(o, "sub", True)
    def __mul__(self, o): return self._binary(o, "mul")
    def __rmul__(self, o): return self._binary(o, "mul", True)
    def __truediv__(self, o): return self._binary(o, "div")
    def __rtruediv__(self, o): return self._binary(o, "div", True)
    def __pow__(self, o): return self._binary(o, "pow")
    def __rpow__(self, o): return self._binary(o, "pow", True)
    def __neg__(self): return unary("neg", self)
    def __pos__(self): return self
    def __abs__(self): return unary("abs", self)
    def __eq__(self, o): return self._binary(o, "eq")
    def __ne__(self, o): return self._binary(o, "ne")
    def __lt__(self, o): return self._binary(o, "lt")
    def __le__(self, o): return self._binary(o, "le")
    def __gt__(self, o): return self._binary(o, "gt")
    def __ge__(self, o): return self._binary(o, "ge")
    __hash__ = None

    def _inplace(self, other, op):
        res = binary(op, self, other)
        if res.shape != self.shape:
            raise ValueError(f"non-broadcastable output operand with shape {self.shape} doesn't match the broadcast shape {res.shape}")
        self._assign(res if res.dtype == self.dtype else res.astype(self.dtype))
        return self

    def __iadd__(self, o): return self._inplace(o, "add")
    def __isub__(self, o): return self._inplace(o, "sub")
    def __imul__(self, o): return self._inplace(o, "mul")
    def __itruediv__(self, o): return self._inplace(o, "div")

    def __matmul__(self, o): return matmul(self, o)
    def __rmatmul__(self, o): return matmul(o, self)
    def dot(self, o): return dot(self, o)

    # ------------------------------------------------------------------ reductions & friends
    def sum(self, axis=None, keepdims=False, dtype=None): return reduce("sum", self, axis, keepdims)
    def mean(self, axis=None, keepdims=False, dtype=None): return reduce("mean", self, axis, keepdims)
    def max(self, axis=None, keepdims=False): return reduce("max", self, axis, keepdims)
    def min(self, axis=None, keepdims=False): return reduce("min", self, axis, keepdims)
    def prod(self, axis=None, keepdims=False): return reduce("prod", self, axis, keepdims)
    def argmax(self, axis=None): return argreduce(True, self, axis)
    def argmin(self, axis=None): return argreduce(False, self, axis)

    def var(self, axis=None, keepdims=False, ddof=0):
        m = reduce("mean", self, axis, True)
        d = self - m
        n = self.size // max(m.size, 1)
        s = reduce("sum", d * d, axis, keepdims)
        return s / float(n - ddof)

    def std(self, axis=None, keepdims=False, ddof=0):
        return unary("sqrt", self.var(axis, keepdims, ddof))

    def clip(self, a_min=None, a_max=None):
        lo = -math.inf if a_min is None else a_min
        hi = math.inf if a_max is None else a_max
        return unary("clip", self, lo, hi)

    def all(self):
        return bool(reduce("min", self.astype(np.int64) if self.dtype != np.dtype(np.int64) else self, None, False).item() != 0) if self.size else True

    def any(self):
        return bool(reduce("max", self.astype(np.int64) if self.dtype != np.dtype(np.int64) else self, None, False).item() != 0) if self.size else False

    # ------------------------------------------------------------------ NumPy protocols
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__":
            return NotImplemented
        out = kwargs.pop("out", None)
        kwargs.pop("dtype", None)
        kwargs.pop("casting", None)
        if kwargs.pop("where", True) is not True:
            raise NotImplementedError("b200: ufunc where= is not supported")
        fn = _UFUNCS.get(ufunc.__name__)
        if fn is None:
            raise NotImplementedError(f"b200: NumPy ufunc '{ufunc.__name__}' has no device kernel (no CPU fallback)")
        res = fn(*inputs)
        if out is not None:
            tgt = out[0] if isinstance(out, tuple) else out
            if not isinstance(tgt, B200Array):
                raise TypeError("b200: out= must be a B200Array")
            tgt._assign(res)
            return tgt
        return res

    def __array_function__(self, func, types, args, kwargs):
        fn = _FUNCTIONS.get(func.__name__)
        if fn is None:
            raise NotImplementedError(f"b200: NumPy function '{func.__name__}' has no device implementation (no CPU fallback)")
        return fn(*args, **kwargs)


# ======================================================================================
# free functions (the kernels behind the methods)
# ======================================================================================
def _as_array(x, like: B200Array = None):
    if isinstance(x, B200Array):
        return x
    if isinstance(x, (numbers.Number, np.generic)):
        raise TypeError("scalar")
    return B200Array.from_numpy(np.asarray(x))


def _result_dtype(a, b, op):
    """NumPy-2 promotion with weak Python scalars."""
    da = a.dtype if isinstance(a, B200Array) else a
    db = b.dtype if isinstance(b, B200Array) else b
    if da is bfloat16 or db is bfloat16:
        raise TypeError("b200: bf16 arrays are internal GEMM/attention operands; cast to float32 for elementwise math")
    if isinstance(da, (np.generic,)):
        da = da.dtype
    if isinstance(db, (np.generic,)):
        db = db.dtype
    rt = np.result_type(da, db)
    if op == "div" and rt.kind in "iub":
        rt = np.dtype(np.float64)
    if rt not in _CODE:
        if rt.kind == "f":
            rt = np.dtype(np.float32) if rt.itemsize < 4 else np.dtype(np.float64)
        elif rt.kind in "iu":
            rt = np.dtype(np.int64)
        else:
            raise TypeError(f"b200: unsupported result dtype {rt}")
    return rt


def binary(op: str, a, b) -> B200Array:
    a_arr, b_arr = isinstance(a, B200Array), isinstance(b, B200Array)
    if not a_arr and not isinstance(a, (numbers.Number, np.generic)):
        a, a_arr = B200Array.from_numpy(np.asarray(a)), True
    if not b_arr and not isinstance(b, (numbers.Number, np.generic)):
        b, b_arr = B200Array.from_numpy(np.asarray(b)), True
    code = BIN[op]
    is_cmp = code >= 10
    if a_arr and b_arr and (a.size == 1 and a.ndim == 0) and False:
        pass
    if a_arr != b_arr:  # array (op) python scalar
        arr, sc, rev = (a, b, False) if a_arr else (b, a, True)
        rt = _result_dtype(arr, sc, op)
        if op == "pow" and rt.kind in "iu" and float(sc) < 0 and not rev:
            raise ValueError("Integers to negative integer powers are not allowed.")
        if op == "pow" and not rev and rt.kind == "f":
            # ndarray.__pow__ fast paths for scalar exponents (exactly rounded, unlike a generic powf)
            fast = {2: "square", 0.5: "sqrt", -1: "recip"}.get(sc if isinstance(sc, (int, float)) else float(sc))
            if fast is not None:
                return unary(fast, arr.astype(rt) if arr.dtype != rt else arr)
            if sc == 1:
                return (arr.astype(rt) if arr.dtype != rt else arr).copy()
        src = arr.astype(rt) if arr.dtype != rt else arr.contiguous()
        out = B200Array.empty(src.shape, np.bool_ if is_cmp else rt)
        if out.size:
            call("nfb_binary_scalar", code, dtype_code(rt), src.ptr, float(sc), out.ptr, out.size, 1 if rev else 0)
        return out
    if not a_arr:  # scalar (op) scalar
        raise TypeError("binary() needs at least one array operand")
    rt = _result_dtype(a, b, op)
    if a.dtype != rt:
        a = a.astype(rt)
    if b.dtype != rt:
        b = b.astype(rt)
    shape = np.broadcast_shapes(a.shape, b.shape)
    av, bv = a.broadcast_to(shape), b.broadcast_to(shape)
    out = B200Array.empty(shape, np.bool_ if is_cmp else rt)
    if out.size:
        shp, sa, sb = _collapse(shape, av.strides, bv.strides)
        call("nfb_binary", code, dtype_code(rt), av.ptr, bv.ptr, out.ptr, len(shp), _lib.i64_array(shp), _lib.i64_array(sa), _lib.i64_array(sb))
    return out


def _collapse(shape, *stride_sets):
    """Merge adjacent dimensions that are jointly contiguous in every operand (fewer index divisions)."""
    dims = [(s, tuple(st[i] for st in stride_sets)) for i, s in enumerate(shape) if s != 1]
    if not dims:
        return [1], *[[0] for _ in stride_sets]
    merged = [dims[0]]
    for s, sts in dims[1:]:
        ps, psts = merged[-1]
        if all(p == q * s for p, q in zip(psts, sts)):
            merged[-1] = (ps * s, sts)
        else:
            merged.append((s, sts))
    shp = [m[0] for m in merged]
    outs = [[m[1][k] for m in merged] for k in range(len(stride_sets))]
    if len(shp) > 8:
        raise B200Error("b200: more than 8 non-mergeable dimensions")
    return (shp, *outs)


def unary(op: str, x: B200Array, p0=0.0, p1=0.0) -> B200Array:
    if not isinstance(x, B200Array):
        x = B200Array.from_numpy(np.asarray(x))
    if x.dtype is bfloat16:
        raise TypeError("b200: bf16 arrays do not support elementwise math; cast to float32")
    dt = np.dtype(x.dtype)
    int_ok = ("neg", "abs", "sign", "square", "clip", "relu")
    if dt.kind in "iub" and op not in int_ok:
        x = x.astype(np.float64)  # NumPy promotes integer inputs of float ufuncs to float64
        dt = x.dtype
    elif dt.kind == "b":
        x = x.astype(np.int64)
        dt = x.dtype
    src = x.contiguous()
    out = B200Array.empty(src.shape, dt)
    if out.size:
        call("nfb_unary", UN[op], dtype_code(dt), src.ptr, out.ptr, out.size, float(p0), float(p1))
    return out


def unary_bwd(op: str, src: B200Array, g: B200Array) -> B200Array:
    s, gg = src.contiguous(), g.contiguous()
    if s.shape != gg.shape:
        gg = gg.broadcast_to(s.shape).contiguous()
    if s.dtype != gg.dtype:
        raise TypeError("unary_bwd: dtype mismatch")
    out = B200Array.empty(s.shape, s.dtype)
    if out.size:
        call("nfb_unary_bwd", UN[op], dtype_code(s.dtype), s.ptr, gg.ptr, out.ptr, out.size)
    return out


def where(cond, x, y) -> B200Array:
    def arr(v, dt=None):
        if isinstance(v, B200Array):
            return v
        if isinstance(v, (numbers.Number, np.generic)):
            return B200Array.full((), v, dt or (np.float32 if isinstance(v, float) else np.asarray(v).dtype))
        return B200Array.from_numpy(np.asarray(v))
    cond = arr(cond)
    if cond.dtype != np.dtype(np.bool_):
        cond = binary("ne", cond, 0)
    xa, ya = isinstance(x, B200Array), isinstance(y, B200Array)
    if xa and not ya:
        y = arr(y, x.dtype)
    elif ya and not xa:
        x = arr(x, y.dtype)
    else:
        x, y = arr(x), arr(y)
    rt = _result_dtype(x, y, "add")
    x = x.astype(rt) if x.dtype != rt else x
    y = y.astype(rt) if y.dtype != rt else y
    shape = np.broadcast_shapes(cond.shape, x.shape, y.shape)
    c, xv, yv = cond.broadcast_to(shape), x.broadcast_to(shape), y.broadcast_to(shape)
    out = B200Array.empty(shape, rt)
    if out.size:
        shp, sc, sx, sy = _collapse(shape, c.strides, xv.strides, yv.strides)
        call("nfb_where", dtype_code(rt), c.ptr, _lib.i64_array(sc), xv.ptr, _lib.i64_array(sx), yv.ptr, _lib.i64_array(sy),
             out.ptr, len(shp), _lib.i64_array(shp))
    return out


def _norm_axes(axis, ndim):
    if axis is None:
        return tuple(range(ndim))
    if isinstance(axis, numbers.Integral):
        axis = (int(axis),)
    out = []
    for a in axis:
        a = int(a)
        if a < 0:
            a += ndim
        if not 0 <= a < ndim:
            raise ValueError(f"axis {a} is out of bounds for array of dimension {ndim}")
        if a in out:
            raise ValueError("duplicate value in 'axis'")
        out.append(a)
    return tuple(sorted(out))


def _reduce_layout(x: B200Array, axes):
    """Return (contiguous array, outer, R, inner) with the reduced axes forming the middle block."""
    nd = x.ndim
    if not axes:
        return x.contiguous(), x.size, 1, 1
    lo, hi = axes[0], axes[-1]
    if list(axes) == list(range(lo, hi + 1)):
        xc = x.contiguous()
        return xc, _prod(x.shape[:lo]), _prod(x.shape[lo:hi + 1]), _prod(x.shape[hi + 1:])
    kept = [i for i in range(nd) if i not in axes]
    xt = x.transpose(kept + list(axes)).contiguous()
    return xt, _prod([x.shape[i] for i in kept]), _prod([x.shape[i] for i in axes]), 1


def reduce(op: str, x: B200Array, axis=None, keepdims=False) -> B200Array:
    if not isinstance(x, B200Array):
        x = B200Array.from_numpy(np.asarray(x))
    if x.dtype is bfloat16:
        x = x.astype(np.float32)
    axes = _norm_axes(axis, x.ndim)
    dt = np.dtype(x.dtype)
    if dt.kind in "iub":
        if op == "mean":
            x, dt = x.astype(np.float64), np.dtype(np.float64)
        elif dt != np.dtype(np.int64):
            back = dt if op in ("max", "min") and dt.kind != "b" else None
            r = reduce(op, x.astype(np.int64), axis, keepdims)
            if op in ("max", "min") and dt.kind == "b":
                return r.astype(np.bool_)
            return r.astype(back) if back is not None else r
    out_shape_keep = tuple(1 if i in axes else s for i, s in enumerate(x.shape))
    out_shape = out_shape_keep if keepdims else tuple(s for i, s in enumerate(x.shape) if i not in axes)
    if x.size == 0 or _prod([x.shape[i] for i in axes]) == 0:
        if op in ("max", "min") and _prod([x.shape[i] for i in axes]) == 0:
            raise ValueError("zero-size array to reduction operation which has no identity")
        # (max / min over a NON-empty axis of an empty array: an empty result, the fill value is never used)
        return B200Array.full(out_shape, {"sum": 0, "prod": 1, "mean": float("nan"), "max": 0, "min": 0}[op], dt)
    xc, outer, R, inner = _reduce_layout(x, axes)
    out = B200Array.empty(out_shape, dt)
    call("nfb_reduce", RED[op], dtype_code(dt), xc.ptr, out.ptr, outer, R, inner)
    return out


def argreduce(is_max: bool, x: B200Array, axis=None) -> B200Array:
    if x.dtype is bfloat16:
        x = x.astype(np.float32)
    if axis is None:
        xc = x.contiguous()
        if xc.size == 0:
            raise ValueError("attempt to get argmax of an empty sequence")
        out = B200Array.empty((), np.int64)
        call("nfb_argreduce", 1 if is_max else 0, dtype_code(xc.dtype), xc.ptr, out.ptr, 1, xc.size, 1)
        return out
    a = int(axis)
    if a < 0:
        a += x.ndim
    if not 0 <= a < x.ndim:
        raise ValueError(f"axis {axis} is out of bounds for array of dimension {x.ndim}")
    xc = x.contiguous()
    out = B200Array.empty(x.shape[:a] + x.shape[a + 1:], np.int64)
    if x.shape[a] == 0:
        raise ValueError("attempt to get argmax of an empty sequence")
    if out.size:
        call("nfb_argreduce", 1 if is_max else 0, dtype_code(xc.dtype), xc.ptr, out.ptr, _prod(x.shape[:a]), x.shape[a], _prod(x.shape[a + 1:]))
    return out


def take_rows(table: B200Array, idx: B200Array, out_dtype=None) -> B200Array:
    """table[idx] along axis 0 (nn/embedding.py:34-39)."""
    t = table.contiguous()
    if idx.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
        idx = idx.astype(np.int64)
    ic = idx.contiguous()
    row_shape = t.shape[1:]
    d = _prod(row_shape)
    odt = out_dtype or t.dtype
    out = B200Array.empty(ic.shape + row_shape, odt)
    if out.size:
        call("nfb_gather_rows", dtype_code(ic.dtype), dtype_code(t.dtype), dtype_code(odt), t.ptr, ic.ptr, out.ptr, ic.size, d, t.shape[0])
    return out


def matmul(a, b) -> B200Array:
    """np.matmul semantics (numpy_backend.py:134-135): batched, broadcasting, 1-D promotion. fp32 SIMT path."""
    a = a if isinstance(a, B200Array) else B200Array.from_numpy(np.asarray(a))
    b = b if isinstance(b, B200Array) else B200Array.from_numpy(np.asarray(b))
    if a.ndim == 0 or b.ndim == 0:
        raise ValueError("matmul: Input operand does not have enough dimensions")
    rt = _result_dtype(a, b, "mul")
    if rt.kind in "iub":
        raise TypeError(f"b200: matmul supports float32 / float64 operands (got {a.dtype} @ {b.dtype}); cast explicitly")
    f64 = rt == np.dtype(np.float64)     # fp64: a plain SIMT kernel (Backend.matmul on the reference's default float64 arrays)
    a = a.astype(rt) if a.dtype != rt else a
    b = b.astype(rt) if b.dtype != rt else b
    a1, b1 = a.ndim == 1, b.ndim == 1
    if a1:
        a = a.reshape(1, -1)
    if b1:
        b = b.reshape(-1, 1)
    M, K = a.shape[-2], a.shape[-1]
    K2, N = b.shape[-2], b.shape[-1]
    if K != K2:
        raise ValueError(f"matmul: Input operand 1 has a mismatch in its core dimension 0 (size {K2} is different from {K})")
    batch_shape = np.broadcast_shapes(a.shape[:-2], b.shape[:-2])
    nb = _prod(batch_shape)
    out = B200Array.empty(tuple(batch_shape) + (M, N), rt)
    esz = 8 if f64 else 4

    def prep(x, rows, cols):
        """-> (array, trans flag, ld, batch stride) without copying when a matrix is row- or col-major."""
        xb = x.broadcast_to(tuple(batch_shape) + (rows, cols)) if x.shape[:-2] != tuple(batch_shape) else x
        sr, sc = xb.strides[-2], xb.strides[-1]
        bs_shape, bs_str = xb.shape[:-2], xb.strides[:-2]
        # batch dims must collapse to a single stride
        ok_batch, bstride = True, 0
        nz = [(s, st) for s, st in zip(bs_shape, bs_str) if s != 1]
        if nz:
            bstride = nz[-1][1]
            e = bstride
            for s, st in reversed(nz):
                if st != e:
                    ok_batch = False
                    break
                e *= s
        if ok_batch and (sc == 1 or cols == 1) and (sr >= max(cols, 1) or rows == 1):
            return xb, 0, (sr if rows > 1 else max(cols, 1)), bstride
        if ok_batch and (sr == 1 or rows == 1) and (sc >= max(rows, 1) or cols == 1):
            return xb, 1, (sc if cols > 1 else max(rows, 1)), bstride
        xc = xb.contiguous()
        return xc, 0, max(cols, 1), rows * cols

    if out.size and K > 0:
        aa, ta, lda, sa = prep(a, M, K)
        bb, tb, ldb, sb = prep(b, K, N)
        if not f64 and nb == 1 and float(M) * N * K >= (1 << 24):
            # one large fp32 product: kernels.gemm_f32 (the bf16x3 split on the tensor cores when the shape qualifies, else FFMA)
            from . import kernels as _K
            a2 = aa.reshape(M, K) if not ta else aa.reshape(M, K).transpose(1, 0)
            b2 = bb.reshape(K, N) if not tb else bb.reshape(K, N).transpose(1, 0)
            _K.gemm_f32(a2, b2, bool(ta), bool(tb), out=out.reshape(M, N))
            nb = 0   # done
        for b0 in range(0, nb, 65535):
            cnt = min(65535, nb - b0)
            if f64:
                call("nfb_gemm_f64", ta, tb, M, N, K, aa.ptr + b0 * sa * esz, lda, sa, bb.ptr + b0 * sb * esz, ldb, sb,
                     out.ptr + b0 * M * N * esz, N, M * N, cnt)
            else:
                call("nfb_gemm_f32", ta, tb, M, N, K, 1.0, aa.ptr + b0 * sa * 4, lda, sa, bb.ptr + b0 * sb * 4, ldb, sb, 0.0,
                     out.ptr + b0 * M * N * 4, N, M * N, cnt, None)
    elif out.size:
        out.fill(0.0)
    if a1 and b1:
        return out.reshape(batch_shape)
    if a1:
        return out.reshape(tuple(batch_shape) + (N,))
    if b1:
        return out.reshape(tuple(batch_shape) + (M,))
    return out


def dot(a, b) -> B200Array:
    """np.dot for the cases the reference uses: scalars, 1-D.1-D, N-D . 1-D/2-D."""
    if isinstance(a, (numbers.Number, np.generic)) or isinstance(b, (numbers.Number, np.generic)):
        return binary("mul", a, b)
    a = a if isinstance(a, B200Array) else B200Array.from_numpy(np.asarray(a))
    b = b if isinstance(b, B200Array) else B200Array.from_numpy(np.asarray(b))
    if a.ndim == 0 or b.ndim == 0:
        return binary("mul", a, b)
    if b.ndim <= 2:
        if a.ndim <= 2:
            return matmul(a, b)
        lead = a.shape[:-1]
        r = matmul(a.reshape(-1, a.shape[-1]), b)
        return r.reshape(lead + r.shape[1:])
    raise NotImplementedError("b200: np.dot with an N-D (N>2) right operand is not implemented; use matmul/einsum")


def concatenate(arrays, axis=0) -> B200Array:
    arrays = [x if isinstance(x, B200Array) else B200Array.from_numpy(np.asarray(x)) for x in arrays]
    if not arrays:
        raise ValueError("need at least one array to concatenate")
    nd = arrays[0].ndim
    if nd == 0:
        raise ValueError("zero-dimensional arrays cannot be concatenated")
    if axis is None:
        arrays, axis, nd = [x.reshape(-1) for x in arrays], 0, 1
    ax = axis + nd if axis < 0 else axis
    rt = arrays[0].dtype
    for x in arrays[1:]:
        if x.ndim != nd or any(x.shape[i] != arrays[0].shape[i] for i in range(nd) if i != ax):
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
        if x.dtype != rt:
            rt = _result_dtype(B200Array(None, (), rt), x, "add") if rt is not bfloat16 else rt
    shape = list(arrays[0].shape)
    shape[ax] = sum(x.shape[ax] for x in arrays)
    out = B200Array.empty(shape, rt)
    pos = 0
    for x in arrays:
        sl = [slice(None)] * nd
        sl[ax] = slice(pos, pos + x.shape[ax])
        out[tuple(sl)] = x if x.dtype == rt else x.astype(rt)
        pos += x.shape[ax]
    return out


def stack(arrays, axis=0) -> B200Array:
    arrays = [x if isinstance(x, B200Array) else B200Array.from_numpy(np.asarray(x)) for x in arrays]
    if not arrays:
        raise ValueError("need at least one array to stack")
    if any(x.shape != arrays[0].shape for x in arrays):
        raise ValueError("all input arrays must have the same shape")
    nd = arrays[0].ndim + 1
    ax = axis + nd if axis < 0 else axis
    return concatenate([x.expand_dims(ax) for x in arrays], ax)


def split(x: B200Array, indices_or_sections, axis=0):
    ax = axis + x.ndim if axis < 0 else axis
    n = x.shape[ax]
    if isinstance(indices_or_sections, numbers.Integral):
        k = int(indices_or_sections)
        if n % k:
            raise ValueError("array split does not result in an equal division")
        cuts = [i * (n // k) for i in range(1, k)]
    else:
        cuts = [int(i) for i in indices_or_sections]
    outs, prev = [], 0
    for c in cuts + [n]:
        sl = [slice(None)] * x.ndim
        sl[ax] = slice(prev, c)
        outs.append(x[tuple(sl)])
        prev = c
    return outs


def arange(start, stop=None, step=1, dtype=None) -> B200Array:
    if stop is None:
        start, stop = 0, start
    if dtype is None:
        dtype = np.result_type(type(start), type(stop), type(step))
        dtype = np.dtype(np.int64) if np.dtype(dtype).kind in "iu" else np.dtype(np.float64)
    n = max(0, int(math.ceil((stop - start) / step)))
    out = B200Array.empty((n,), dtype)
    if n:
        call("nfb_arange", dtype_code(out.dtype), out.ptr, n, float(start), float(step))
    return out


def tril_mask(rows, cols, k=0) -> B200Array:
    """Boolean lower-triangular mask built on device (rows x cols)."""
    r = arange(0, rows, 1, np.int64).reshape(rows, 1)
    c = arange(0, cols, 1, np.int64).reshape(1, cols)
    return binary("le", c, binary("add", r, k))


# ---- NumPy protocol tables ----------------------------------------------------------------------
def _u(op):
    return lambda x, *a: unary(op, x)


def _b(op):
    return lambda a, b: binary(op, a, b)


_UFUNCS = {
    "add": _b("add"), "subtract": _b("sub"), "multiply": _b("mul"), "divide": _b("div"), "true_divide": _b("div"),
    "power": _b("pow"), "float_power": _b("pow"), "maximum": _b("max"), "minimum": _b("min"),
    "equal": _b("eq"), "not_equal": _b("ne"), "less": _b("lt"), "less_equal": _b("le"), "greater": _b("gt"),
    "greater_equal": _b("ge"),
    "negative": _u("neg"), "exp": _u("exp"), "log": _u("log"), "sqrt": _u("sqrt"), "absolute": _u("abs"), "fabs": _u("abs"),
    "sign": _u("sign"), "tanh": _u("tanh"), "square": _u("square"), "reciprocal": _u("recip"), "sin": _u("sin"),
    "cos": _u("cos"), "floor": _u("floor"), "positive": lambda x: x,
    "matmul": lambda a, b: matmul(a, b),
    "isfinite": lambda x: binary("lt", unary("abs", x), math.inf),
    "isnan": lambda x: binary("ne", x, x),
    "isinf": lambda x: binary("eq", unary("abs", x), math.inf),
    "logical_not": lambda x: binary("eq", x, 0),
}


def _np_clip(a, a_min=None, a_max=None, out=None, **kw):
    r = a.clip(a_min, a_max)
    if out is not None:
        out._assign(r)
        return out
    return r


def _np_nan_to_num(x, copy=True, nan=0.0, posinf=None, neginf=None):
    fi = np.finfo(np.dtype(x.dtype))
    hi = fi.max if posinf is None else posinf
    lo = fi.min if neginf is None else neginf
    r = where(binary("ne", x, x), nan, x)
    r = where(binary("eq", r, math.inf), hi, r)
    r = where(binary("eq", r, -math.inf), lo, r)
    return r


def _np_transpose(a, axes=None):
    return a.transpose(axes)


def _np_sum(a, axis=None, dtype=None, out=None, keepdims=False, **kw):
    return reduce("sum", a, axis, keepdims)


def _np_mean(a, axis=None, dtype=None, out=None, keepdims=False, **kw):
    return reduce("mean", a, axis, keepdims)


def _np_all(a, axis=None, **kw):
    if axis is not None:
        raise NotImplementedError("b200: np.all with axis")
    return a.all()


def _np_any(a, axis=None, **kw):
    if axis is not None:
        raise NotImplementedError("b200: np.any with axis")
    return a.any()


def _np_broadcast_arrays(*arrays, **kw):
    arrs = [x if isinstance(x, B200Array) else B200Array.from_numpy(np.asarray(x)) for x in arrays]
    shape = np.broadcast_shapes(*[x.shape for x in arrs])
    return [x if x.shape == tuple(shape) else x.broadcast_to(tuple(shape)) for x in arrs]


_FUNCTIONS = {
    "broadcast_arrays": _np_broadcast_arrays,
    "sum": _np_sum, "mean": _np_mean,
    "max": lambda a, axis=None, out=None, keepdims=False, **kw: reduce("max", a, axis, keepdims),
    "amax": lambda a, axis=None, out=None, keepdims=False, **kw: reduce("max", a, axis, keepdims),
    "min": lambda a, axis=None, out=None, keepdims=False, **kw: reduce("min", a, axis, keepdims),
    "amin": lambda a, axis=None, out=None, keepdims=False, **kw: reduce("min", a, axis, keepdims),
    "prod": lambda a, axis=None, dtype=None, out=None, keepdims=False, **kw: reduce("prod", a, axis, keepdims),
    "var": lambda a, axis=None, dtype=None, out=None, ddof=0, keepdims=False, **kw: a.var(axis, keepdims, ddof),
    "std": lambda a, axis=None, dtype=None, out=None, ddof=0, keepdims=False, **kw: a.std(axis, keepdims, ddof),
    "argmax": lambda a, axis=None, out=None, **kw: argreduce(True, a, axis),
    "argmin": lambda a, axis=None, out=None, **kw: argreduce(False, a, axis),
    "clip": _np_clip, "nan_to_num": _np_nan_to_num, "where": where,
    "transpose": _np_transpose, "swapaxes": lambda a, x, y: a.swapaxes(x, y),
    "reshape": lambda a, *s, **kw: a.reshape(*s) if s else a.reshape(kw.get("shape", kw.get("newshape"))),
    "squeeze": lambda a, axis=None: a.squeeze(axis), "expand_dims": lambda a, axis: a.expand_dims(axis),
    "broadcast_to": lambda a, shape, **kw: a.broadcast_to(shape),
    "concatenate": lambda arrs, axis=0, **kw: concatenate(arrs, axis), "stack": lambda arrs, axis=0, **kw: stack(arrs, axis),
    "split": split, "matmul": matmul, "dot": dot, "copy": lambda a, **kw: a.copy(),
    "zeros_like": lambda a, dtype=None, **kw: B200Array.full(a.shape, 0, dtype or a.dtype),
    "ones_like": lambda a, dtype=None, **kw: B200Array.full(a.shape, 1, dtype or a.dtype),
    "full_like": lambda a, v, dtype=None, **kw: B200Array.full(a.shape, v, dtype or a.dtype),
    "empty_like": lambda a, dtype=None, **kw: B200Array.empty(a.shape, dtype or a.dtype),
    "shape": lambda a: a.shape, "ndim": lambda a: a.ndim, "size": lambda a, axis=None: a.size if axis is None else a.shape[axis],
    "ravel": lambda a, **kw: a.ravel(), "all": _np_all, "any": _np_any,
    "isscalar": lambda a: False, "iscomplexobj": lambda a: False, "isrealobj": lambda a: True,
    "result_type": lambda *a: np.result_type(*[x.dtype if isinstance(x, B200Array) else x for x in a]),
    "may_share_memory": lambda a, b, **kw: False, "shares_memory": lambda a, b, **kw: False,
    "outer": lambda a, b, **kw: binary("mul", a.reshape(-1, 1), b.reshape(1, -1)),
    "tril": lambda m, k=0: where(tril_mask(m.shape[-2], m.shape[-1], k), m, 0.0 if np.dtype(m.dtype).kind == "f" else 0),
    "allclose": lambda a, b, rtol=1e-5, atol=1e-8, **kw: bool(np.allclose(a.get() if isinstance(a, B200Array) else a,
                                                                         b.get() if isinstance(b, B200Array) else b, rtol, atol)),
    "array_equal": lambda a, b, **kw: bool(np.array_equal(a.get() if isinstance(a, B200Array) else a,
                                                         b.get() if isinstance(b, B200Array) else b)),
}
