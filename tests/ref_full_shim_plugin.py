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

# every module of this package under the same dotted path below `neural_arch` (one module object, two names): an import such as
# `neural_arch.nn.positional` must not load a second copy of the file under the other name
import importlib  # noqa: E402
import pkgutil  # noqa: E402

import nfb200  # noqa: E402

for info in pkgutil.walk_packages(nfb200.__path__, "nfb200."):
    if info.name.split(".")[-1].startswith("_") and not info.name.endswith("__init__"):
        continue
    try:
        mod = importlib.import_module(info.name)
    except Exception:
        continue
    alias = "neural_arch." + info.name[len("nfb200."):]
    if alias not in sys.modules or alias.startswith("neural_arch.backends") is False:
        sys.modules.setdefault(alias, mod)

# file names of the reference that differ from this package's layout
import nfb200.distributed  # noqa: E402
import nfb200.models  # noqa: E402
import nfb200.models.bert  # noqa: E402
import nfb200.models.gpt2  # noqa: E402
import nfb200.models.vit  # noqa: E402
import nfb200.nn.transformer  # noqa: E402

sys.modules.setdefault("neural_arch.nn.attention", nfb200.nn.transformer)
for sub, mod in (("language", None), ("vision", None)):
    m = types.ModuleType("neural_arch.models." + sub)
    m.__path__ = []
    sys.modules.setdefault("neural_arch.models." + sub, m)
sys.modules.setdefault("neural_arch.models.language.gpt2", nfb200.models.gpt2)
sys.modules.setdefault("neural_arch.models.language.bert", nfb200.models.bert)
sys.modules.setdefault("neural_arch.models.vision.vision_transformer", nfb200.models.vit)
import nfb200.models.clip  # noqa: E402
import nfb200.models.modern_transformer  # noqa: E402
import nfb200.models.roberta  # noqa: E402

sys.modules.setdefault("neural_arch.models.language.modern_transformer", nfb200.models.modern_transformer)
sys.modules.setdefault("neural_arch.models.language.roberta", nfb200.models.roberta)
_mm = types.ModuleType("neural_arch.models.multimodal")
_mm.__path__ = []
sys.modules.setdefault("neural_arch.models.multimodal", _mm)
sys.modules.setdefault("neural_arch.models.multimodal.clip", nfb200.models.clip)
pkg.models, pkg.distributed = nfb200.models, nfb200.distributed
ALIASES = {"core": nfb200.core, "core.base": nfb200.core.base, "core.tensor": nfb200.core.tensor, "core.device": nfb200.core.device,
           "core.dtype": nfb200.core.dtype, "functional": nfb200.functional, "functional.arithmetic": nfb200.functional.arithmetic,
           "functional.activation": nfb200.functional.activation, "functional.loss": nfb200.functional.loss,
           "functional.utils": nfb200.functional.utils, "nn": nfb200.nn, "optim": nfb200.optim,
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

_lin = types.ModuleType("neural_arch.nn.linear")        # nn/linear.py of the reference holds the (in, out) layer under the name Linear
_lin.__dict__.update({k: v for k, v in vars(nfb200.nn.linear).items() if not k.startswith("__")})
_lin.Linear = nfb200.nn.linear.StandardLinear
sys.modules["neural_arch.nn.linear"] = _lin
