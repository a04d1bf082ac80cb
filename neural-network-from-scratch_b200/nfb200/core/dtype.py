"""DType: the five element types of the reference (pinned by the reference's
tests/test_core_dtype_comprehensive.py:20-135).  bf16 is deliberately NOT a Tensor dtype: it is an
internal storage/compute mode of the B200 backend."""
from __future__ import annotations

from enum import Enum

import numpy as np


class DType(Enum):
    FLOAT32 = np.float32
    FLOAT64 = np.float64
    INT32 = np.int32
    INT64 = np.int64
    BOOL = np.bool_

    @property
    def numpy_dtype(self):
        return np.dtype(self.value)

    @property
    def is_floating(self) -> bool:
        return self in (DType.FLOAT32, DType.FLOAT64)

    @property
    def is_integer(self) -> bool:
        return self in (DType.INT32, DType.INT64)

    @property
    def bytes_per_element(self) -> int:
        return self.numpy_dtype.itemsize

    @classmethod
    def from_numpy(cls, dtype) -> "DType":
        if isinstance(dtype, DType):
            return dtype
        try:
            nd = np.dtype(dtype)
        except TypeError:
            raise ValueError(f"Unsupported NumPy dtype: {dtype}") from None
        for m in cls:
            if m.numpy_dtype == nd:
                return m
        raise ValueError(f"Unsupported NumPy dtype: {nd}")

    def __eq__(self, other):
        if isinstance(other, DType):
            return self is other
        try:
            return self.numpy_dtype == np.dtype(other)
        except TypeError:
            return NotImplemented

    def __hash__(self):
        return hash(self.name)

    def __str__(self):
        return self.numpy_dtype.name

    def __repr__(self):
        return f"DType.{self.name}"


_DEFAULT = DType.FLOAT32


def get_default_dtype() -> DType:
    return _DEFAULT


def set_default_dtype(dtype: DType) -> None:
    global _DEFAULT
    if not isinstance(dtype, DType):
        raise TypeError("Expected DType")
    _DEFAULT = dtype
