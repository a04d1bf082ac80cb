"""Backend plugin interface + registry: the drop-in boundary of this repo.

Mirrors the contract of the reference's ``neural_arch.backends`` package:
  * ``Backend`` abstract interface -- 54 abstract members (backends/backend.py:9-322)
  * ``register_backend / get_backend / set_backend / available_backends / current_backend``
    (backends/backend.py:325-375): ``get_backend(name)`` builds a NEW instance per call, raises
    ``ValueError`` for an unknown name and ``RuntimeError`` for an unavailable backend
  * device <-> backend-name helpers (backends/utils.py:10-90)

When the real ``neural_arch.backends`` package is importable (the reference checked out next to this
repo), :func:`install_into_reference` registers the B200 backend into ITS registry under the names
``"b200"`` and ``"cuda"`` so the reference's own Tensor/functional code routes ``Device.cuda(i)``
tensors to it.
"""
from __future__ import annotations

import abc
from typing import Dict, List, Optional

# The abstract surface, grouped as in SURVEY 8(b).  name -> one-line contract.
_ABSTRACT_PROPERTIES = ("name", "is_available", "supports_gradients")
_ABSTRACT_METHODS: Dict[str, str] = {
    # creation
    "array": "array(data, dtype=None) -> backend array holding a copy of `data`",
    "zeros": "zeros(shape, dtype=None) -> float32 by default",
    "ones": "ones(shape, dtype=None)",
    "full": "full(shape, fill_value, dtype=None)",
    "arange": "arange(start, stop, step=1.0, dtype=None)",
    "random_normal": "random_normal(shape, mean=0.0, std=1.0, dtype=None) -> always float32 unless dtype given",
    "random_uniform": "random_uniform(shape, low=0.0, high=1.0, dtype=None)",
    # shape
    "reshape": "reshape(x, shape)", "transpose": "transpose(x, axes=None)", "squeeze": "squeeze(x, axis=None)",
    "expand_dims": "expand_dims(x, axis)",
    # math
    "add": "add(x, y)", "subtract": "subtract(x, y)", "multiply": "multiply(x, y)", "divide": "divide(x, y)",
    "power": "power(x, y)", "matmul": "matmul(x, y) with np.matmul batching", "dot": "dot(x, y)",
    # reductions
    "sum": "sum(x, axis=None, keepdims=False)", "mean": "mean(x, axis=None, keepdims=False)",
    "max": "max(x, axis=None, keepdims=False)", "min": "min(x, axis=None, keepdims=False)",
    "argmax": "argmax(x, axis=None)", "argmin": "argmin(x, axis=None)",
    # unary
    "exp": "exp(x)", "log": "log(x)", "sqrt": "sqrt(x)", "abs": "abs(x)", "sign": "sign(x)",
    "clip": "clip(x, min_val, max_val)",
    # comparisons
    "equal": "equal(x, y)", "not_equal": "not_equal(x, y)", "less": "less(x, y)", "less_equal": "less_equal(x, y)",
    "greater": "greater(x, y)", "greater_equal": "greater_equal(x, y)",
    # manipulation
    "concatenate": "concatenate(arrays, axis=0)", "stack": "stack(arrays, axis=0)",
    "split": "split(x, indices_or_sections, axis=0) -> list",
    # conversion
    "astype": "astype(x, dtype)", "to_numpy": "to_numpy(x) -> host np.ndarray", "from_numpy": "from_numpy(x, dtype=None)",
    # device
    "to_device": "to_device(x, device) ; ValueError for a device this backend does not own",
    "device_of": "device_of(x) -> 'cpu' | 'cuda:N'",
    # utility
    "is_array": "is_array(x) -> bool", "shape": "shape(x) -> tuple", "size": "size(x) -> int", "dtype": "dtype(x)",
    # advanced
    "einsum": "einsum(equation, *operands)", "where": "where(condition, x, y)",
    "unique": "unique(x, return_counts=False)",
}


def _make_abstract(name: str, doc: str):
    def _missing(self, *args, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError(name)

    _missing.__name__ = name
    _missing.__doc__ = doc
    return abc.abstractmethod(_missing)


class Backend(abc.ABC):
    """Abstract compute backend.  A concrete subclass must provide every member listed in
    ``_ABSTRACT_PROPERTIES`` and ``_ABSTRACT_METHODS`` or it cannot be instantiated (same rule as
    the reference, whose ``get_backend`` calls ``backend_class()``)."""


for _p in _ABSTRACT_PROPERTIES:
    setattr(Backend, _p, property(_make_abstract(_p, f"{_p} (property)")))
for _m, _d in _ABSTRACT_METHODS.items():
    setattr(Backend, _m, _make_abstract(_m, _d))
abc.update_abstractmethods(Backend)

ABSTRACT_MEMBERS = tuple(_ABSTRACT_PROPERTIES) + tuple(_ABSTRACT_METHODS)

# ---------------------------------------------------------------------------------------- registry
_BACKENDS: Dict[str, type] = {}
_CURRENT: Optional[Backend] = None


def register_backend(name: str, backend_class: type) -> None:
    _BACKENDS[name] = backend_class


def get_backend(name: Optional[str] = None) -> Backend:
    global _CURRENT
    if name is None:
        if _CURRENT is None:
            set_backend("numpy")
        return _CURRENT
    cls = _BACKENDS.get(name)
    if cls is None:
        raise ValueError(f"Unknown backend: {name}. Available: {list(_BACKENDS.keys())}")
    inst = cls()
    if not inst.is_available:
        raise RuntimeError(f"Backend '{name}' is not available on this system")
    return inst


def set_backend(name: str) -> None:
    global _CURRENT
    _CURRENT = get_backend(name)


def available_backends() -> List[str]:
    names = []
    for name, cls in _BACKENDS.items():
        try:
            if cls().is_available:
                names.append(name)
        except Exception:
            continue
    return names


def current_backend() -> Optional[Backend]:
    return _CURRENT


# ---------------------------------------------------------------------------------- device helpers
def get_backend_for_device(device: str) -> str:
    """'cpu' -> 'numpy', 'cuda[:N]' -> 'cuda' (which this repo registers as the B200 backend)."""
    d = str(device).lower()
    if d == "cpu":
        return "numpy"
    if d.startswith("cuda") or d.startswith("b200"):
        return "cuda"
    if d.startswith("mps"):
        return "mps"
    return "numpy"


def get_device_for_backend(backend_name: str) -> str:
    return {"numpy": "cpu", "cuda": "cuda", "b200": "cuda", "mps": "mps"}.get(backend_name, "cpu")


def auto_select_backend(prefer_gpu: bool = True) -> Backend:
    if prefer_gpu and "cuda" in available_backends():
        return get_backend("cuda")
    return get_backend("numpy")


def install_into_reference() -> bool:
    """Register the B200 backend into the reference's own registry, if that package is importable.

    This is the drop-in step a Neural Forge user performs:
        import nfb200; nfb200.install_into_reference()
    Returns False when ``neural_arch.backends`` cannot be imported (e.g. on the GPU box, where the
    reference checkout does not exist).
    """
    try:
        from neural_arch.backends import backend as ref_backend  # type: ignore
    except Exception:
        return False
    from .b200_backend import make_reference_subclass

    cls = make_reference_subclass(ref_backend.Backend)
    ref_backend.register_backend("b200", cls)
    ref_backend.register_backend("cuda", cls)
    return True
