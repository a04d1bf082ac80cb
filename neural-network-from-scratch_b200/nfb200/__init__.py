"""nfb200 -- the B200-native compute backend for Neural Forge (``neural_arch``).

Host-side mirror of the reference's backend / Tensor / functional / nn / optim / distributed surface
for the data-parallel Transformer training hot path, on top of libnfb200.so
(hand-written sm_100a CUDA behind a C-ABI, include/nfb200.h).
"""
from . import _lib
from ._lib import B200Error, device_available, lib_available
from .array import B200Array, bfloat16
from .plugin import (Backend, available_backends, auto_select_backend, current_backend, get_backend,
                     get_backend_for_device, get_device_for_backend, install_into_reference, register_backend,
                     set_backend)
from . import numpy_backend as _numpy_backend  # registers "numpy"
from . import b200_backend as _b200_backend  # registers "b200" and "cuda"
from .b200_backend import B200Backend
from .numpy_backend import NumpyBackend

__version__ = "0.1.0"
