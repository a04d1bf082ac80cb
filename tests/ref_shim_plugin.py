"""pytest plugin (`-p ref_shim_plugin`) used by tests/test_reference_suite.py: makes `neural_arch.backends` (and its
`backend`, `numpy_backend` and `utils` submodules) resolve to THIS package's mirror of the reference's backend registry
(nfb200.plugin + nfb200.numpy_backend), so that the reference's own backend test files can be run, unmodified and from
where they lie, against it.  `neural_arch.exceptions` is this package's mirror of the exception tree (nfb200/exceptions.py).  Nothing else of `neural_arch` is provided: a test that needs more fails at import and is reported as such."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-network-from-scratch_b200"))

import nfb200  # noqa: E402
from nfb200 import numpy_backend, plugin  # noqa: E402

REGISTRY = ("Backend", "available_backends", "get_backend", "set_backend", "register_backend", "current_backend")
UTILS = ("auto_select_backend", "get_backend_for_device", "get_device_for_backend")


def _module(name, source, names):
    m = types.ModuleType(name)
    for n in names:
        setattr(m, n, getattr(source, n))
    sys.modules[name] = m
    return m


pkg = types.ModuleType("neural_arch")
pkg.__path__ = []                      # a package with no files: submodules come from sys.modules only
sys.modules["neural_arch"] = pkg
backends = _module("neural_arch.backends", plugin, REGISTRY + UTILS)
backends.__path__ = []
backends.NumpyBackend = nfb200.NumpyBackend
backends.B200Backend = nfb200.B200Backend
backends.backend = _module("neural_arch.backends.backend", plugin, REGISTRY + ("_BACKENDS",))
backends.numpy_backend = _module("neural_arch.backends.numpy_backend", numpy_backend, ("NumpyBackend",))
backends.utils = _module("neural_arch.backends.utils", plugin, UTILS)
pkg.backends = backends

import nfb200.exceptions  # noqa: E402

sys.modules["neural_arch.exceptions"] = nfb200.exceptions
pkg.exceptions = nfb200.exceptions
