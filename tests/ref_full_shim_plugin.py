"""pytest plugin (`-p ref_full_shim_plugin`): `neural_arch` and its `core` / `functional` / `nn` / `optim` / `backends`
sub-packages resolve to THIS package's mirrors (nfb200.*), so that the reference's own op-level test files (Tensor,
layers, optimizers) run unmodified on the host path of this package.  Used by tests/test_reference_suite.py."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim_plugin  # noqa: E402,F401  (installs neural_arch, neural_arch.backends.*, neural_arch.exceptions)

import nfb200.core  # noqa: E402
import nfb200.core.base  # noqa: E402
import nfb200.core.device  # noqa: E402
import nfb200.core.dtype  # noqa: E402
import nfb200.core.tensor  # noqa: E402
import nfb200.functional  # noqa: E402
import nfb200.functional.activation  # noqa: E402
import nfb200.functional.arithmetic  # noqa: E402
import nfb200.functional.loss  # noqa: E402
import nfb200.functional.utils  # noqa: E402
import nfb200.nn  # noqa: E402
import nfb200.nn.layers  # noqa: E402
import nfb200.nn.linear  # noqa: E402
import nfb200.optim  # noqa: E402
import nfb200.optim.adam  # noqa: E402
import nfb200.optim.sgd  # noqa: E402

pkg = sys.modules["neural_arch"]
ALIASES = {"core": nfb200.core, "core.base": nfb200.core.base, "core.tensor": nfb200.core.tensor, "core.device": nfb200.core.device,
           "core.dtype": nfb200.core.dtype, "functional": nfb200.functional, "functional.arithmetic": nfb200.functional.arithmetic,
           "functional.activation": nfb200.functional.activation, "functional.loss": nfb200.functional.loss,
           "functional.utils": nfb200.functional.utils, "nn": nfb200.nn, "nn.linear": nfb200.nn.linear, "optim": nfb200.optim,
           "optim.adam": nfb200.optim.adam, "optim.sgd": nfb200.optim.sgd}
for name, mod in ALIASES.items():
    sys.modules["neural_arch." + name] = mod
for sub in ("embedding", "normalization", "activation", "container", "module"):    # the reference splits nn/ into these files
    m = types.ModuleType("neural_arch.nn." + sub)
    m.__dict__.update({k: v for k, v in vars(nfb200.nn.layers).items() if not k.startswith("_")})
    m.__dict__.update({k: v for k, v in vars(nfb200.core.base).items() if not k.startswith("_")})
    sys.modules["neural_arch.nn." + sub] = m
sys.modules["neural_arch.optim.adamw"] = nfb200.optim.adam
for src in (nfb200.core, nfb200.functional, nfb200.nn, nfb200.optim):     # `from neural_arch import Tensor, add, Linear, Adam, ...`
    for k, v in vars(src).items():
        if not k.startswith("_") and not isinstance(v, types.ModuleType):
            setattr(pkg, k, v)
for k in ("core", "functional", "nn", "optim"):
    setattr(pkg, k, ALIASES[k])
