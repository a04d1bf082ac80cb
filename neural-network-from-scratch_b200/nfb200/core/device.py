"""Device / DeviceType (pinned by the reference's tests/test_device_complete.py:22-215).
Exactly three device types exist; the B200 backend claims ``cuda`` (one process per GPU: ``cuda:LOCAL_RANK``)."""
from __future__ import annotations

from enum import Enum
from typing import Optional, Union


class DeviceType(Enum):
    CPU = "cpu"
    CUDA = "cuda"
    MPS = "mps"

    def __str__(self):
        return self.value


class Device:
    __slots__ = ("type", "index")

    def __init__(self, type: DeviceType, index: Optional[int] = None):
        if not isinstance(type, DeviceType):
            type = DeviceType(str(type).lower())
        if index is not None and index < 0:
            raise ValueError("Device index must be non-negative")
        if type is DeviceType.CPU:
            index = None
        elif index is None:
            index = 0
        self.type, self.index = type, index

    @classmethod
    def cpu(cls) -> "Device":
        return cls(DeviceType.CPU)

    @classmethod
    def cuda(cls, index: int = 0) -> "Device":
        return cls(DeviceType.CUDA, index)

    @classmethod
    def mps(cls, index: int = 0) -> "Device":
        return cls(DeviceType.MPS, index)

    @classmethod
    def from_string(cls, s: str) -> "Device":
        name, _, idx = str(s).strip().lower().partition(":")
        try:
            t = DeviceType(name)
        except ValueError:
            raise ValueError(f"Unsupported device: {s}") from None
        return cls(t, int(idx) if idx else None)

    @property
    def is_cpu(self):
        return self.type is DeviceType.CPU

    @property
    def is_cuda(self):
        return self.type is DeviceType.CUDA

    @property
    def is_mps(self):
        return self.type is DeviceType.MPS

    @property
    def is_gpu(self):
        return self.type in (DeviceType.CUDA, DeviceType.MPS)

    def __eq__(self, other):
        if isinstance(other, str):
            try:
                other = Device.from_string(other)
            except ValueError:
                return False
        return isinstance(other, Device) and self.type is other.type and self.index == other.index

    def __hash__(self):
        return hash((self.type, self.index))

    def __str__(self):
        return self.type.value if self.index is None else f"{self.type.value}:{self.index}"

    def __repr__(self):
        return f"Device('{self.type.value}')" if self.index is None else f"Device('{self.type.value}', {self.index})"


_DEFAULT = Device.cpu()


def get_default_device() -> Device:
    return _DEFAULT


def set_default_device(device: Union[Device, str]) -> None:
    global _DEFAULT
    if isinstance(device, str):
        device = Device.from_string(device)
    if not isinstance(device, Device):
        raise TypeError("Expected Device or str")
    _DEFAULT = device


def as_device(device) -> Device:
    if device is None:
        return get_default_device()
    if isinstance(device, Device):
        return device
    if isinstance(device, str):
        return Device.from_string(device)
    raise TypeError("Expected Device or str")
