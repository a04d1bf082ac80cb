"""Parameter / Module / Optimizer base classes (the reference's core/base.py is missing from the
snapshot, contract pinned by its tests/test_core_base_comprehensive.py and call sites in nn/, optim/, models/)."""
from __future__ import annotations

import abc
from collections import OrderedDict
from typing import Dict, Iterator, Optional, Tuple

import numpy as np

from .device import as_device
from .tensor import Tensor


class Parameter(Tensor):
    """A Tensor that requires a gradient and is registered by Module.__setattr__."""

    def __init__(self, data, name: Optional[str] = None, device=None, dtype=None):
        super().__init__(data, requires_grad=True, dtype=dtype, device=device, name=name)

    def __repr__(self):
        return "Parameter" + super().__repr__()[len("Tensor"):]


class ParameterView:
    """What ``Module.parameters()`` returns: iterating yields the Parameters (list-like, as in
    tests/test_core_base_comprehensive.py:69-81) while ``items()/keys()/values()/[name]/in`` behave like a dict keyed
    by dotted name (nn/transformer.py:101-114, examples/training/gpt2_training.py:322)."""

    def __init__(self, items: "OrderedDict[str, Parameter]"):
        self._d = items

    def __iter__(self) -> Iterator[Parameter]:
        return iter(self._d.values())

    def __len__(self):
        return len(self._d)

    def __contains__(self, key):
        if isinstance(key, str):
            return key in self._d
        return any(p is key for p in self._d.values())

    def __getitem__(self, key):
        if isinstance(key, int):
            return list(self._d.values())[key]
        return self._d[key]

    def items(self):
        return self._d.items()

    def keys(self):
        return self._d.keys()

    def values(self):
        return self._d.values()

    def get(self, k, default=None):
        return self._d.get(k, default)

    def __repr__(self):
        return f"ParameterView({list(self._d.keys())})"


class Module:
    def __init__(self):
        object.__setattr__(self, "_parameters", OrderedDict())
        object.__setattr__(self, "_modules", OrderedDict())
        object.__setattr__(self, "training", True)

    def __setattr__(self, name, value):
        params = self.__dict__.get("_parameters")
        if params is None:
            raise AttributeError("cannot assign attributes before Module.__init__() call")
        mods = self.__dict__["_modules"]
        if isinstance(value, Parameter):
            mods.pop(name, None)
            params[name] = value
        elif isinstance(value, Module):
            params.pop(name, None)
            mods[name] = value
        else:
            if name in params and value is None:
                params.pop(name)
            if name in mods and value is None:
                mods.pop(name)
        object.__setattr__(self, name, value)

    def register_parameter(self, name: str, param: Optional[Parameter]):
        setattr(self, name, param)

    def _register_parameters(self):  # hook point used by some reference layers (nn/conv.py:85-87)
        pass

    def add_module(self, name: str, module: "Module"):
        setattr(self, name, module)

    # ------------------------------------------------------------------ traversal
    def named_parameters(self, prefix: str = "", recurse: bool = True, _memo=None) -> Iterator[Tuple[str, Parameter]]:
        """Tied parameters are de-duplicated by identity (PyTorch semantics; SURVEY 7 'Tied weights')."""
        memo = _memo if _memo is not None else set()
        for n, p in self._parameters.items():
            if p is None or id(p) in memo:
                continue
            memo.add(id(p))
            yield (f"{prefix}.{n}" if prefix else n), p
        if recurse:
            for mn, m in self._modules.items():
                if m is None:
                    continue
                yield from m.named_parameters(f"{prefix}.{mn}" if prefix else mn, True, memo)

    def parameters(self, recurse: bool = True) -> ParameterView:
        return ParameterView(OrderedDict(self.named_parameters("", recurse)))

    def _parameters_iterator(self, recurse: bool = True):
        return iter(self.parameters(recurse))

    def named_modules(self, prefix: str = ""):
        yield prefix, self
        for n, m in self._modules.items():
            if m is not None:
                yield from m.named_modules(f"{prefix}.{n}" if prefix else n)

    def modules(self):
        for _, m in self.named_modules():
            yield m

    def children(self):
        return iter(self._modules.values())

    # ------------------------------------------------------------------ state
    def train(self, mode: bool = True):
        for m in self.modules():
            object.__setattr__(m, "training", bool(mode))
        return self

    def eval(self):
        return self.train(False)

    def zero_grad(self):
        for p in self.parameters():
            p.zero_grad()

    def state_dict(self) -> Dict[str, np.ndarray]:
        """{dotted_name: host ndarray} (nn/module.py:112-123): a device->host snapshot on demand."""
        out = OrderedDict()
        for prefix, m in self.named_modules():
            for n, p in m._parameters.items():
                if p is not None:
                    out[f"{prefix}.{n}" if prefix else n] = p.numpy().copy()
        return out

    def load_state_dict(self, state: Dict[str, np.ndarray], strict: bool = True):
        own = {}
        for prefix, m in self.named_modules():
            for n, p in m._parameters.items():
                if p is not None:
                    own[f"{prefix}.{n}" if prefix else n] = p
        missing = [k for k in own if k not in state]
        unexpected = [k for k in state if k not in own]
        if strict and (missing or unexpected):
            raise KeyError(f"load_state_dict: missing {missing}, unexpected {unexpected}")
        for k, p in own.items():
            if k in state:
                arr = np.asarray(state[k])
                if tuple(arr.shape) != p.shape:
                    raise ValueError(f"shape mismatch for {k}: {arr.shape} vs {p.shape}")
                p.data = arr.astype(np.float32) if arr.dtype.kind == "f" else arr

    def to(self, device):
        dev = as_device(device)
        seen = set()
        for m in self.modules():
            for n, p in list(m._parameters.items()):
                if p is None or id(p) in seen:
                    continue
                seen.add(id(p))
                if p.device != dev:
                    host = p.numpy()
                    p._device = dev
                    from .tensor import _backend_for
                    p._backend = _backend_for(dev)
                    p._data = host if dev.is_cpu else p._backend.from_numpy(host)
                    p._grad = None
                    p._grad_slot = None
        return self

    def cuda(self, index: int = 0):
        from .device import Device
        return self.to(Device.cuda(index))

    def cpu(self):
        from .device import Device
        return self.to(Device.cpu())

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def extra_repr(self) -> str:
        return ""

    def __repr__(self):
        lines = [f"{self.__class__.__name__}({self.extra_repr()}"]
        for n, m in self._modules.items():
            lines.append(f"  ({n}): {repr(m)}")
        return "\n".join(lines) + ")"


class Optimizer(abc.ABC):
    """Optimizer(parameters: dict name -> Parameter, **defaults) (tests/test_core_base_comprehensive.py:381-413)."""

    def __init__(self, parameters, **defaults):
        if not isinstance(parameters, dict):
            if hasattr(parameters, "items"):
                parameters = dict(parameters.items())
            else:
                parameters = {f"param_{i}": p for i, p in enumerate(parameters)}
        self.parameters = parameters
        self.defaults = defaults
        for k, v in defaults.items():
            if not hasattr(self, k):
                setattr(self, k, v)

    @abc.abstractmethod
    def step(self) -> None:
        ...

    def zero_grad(self) -> None:
        for p in self.parameters.values():
            p.zero_grad()
