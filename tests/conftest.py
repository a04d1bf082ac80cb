import os
import sys

import numpy as np
import pytest

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PKG = os.path.join(_ROOT, "neural-network-from-scratch_b200")
for p in (_ROOT, _PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # a fresh checkout has no libnfb200.so (built artefacts are git-ignored): build it once (nvcc cross-compiles without a
    # GPU).  Only when it is MISSING -- never rebuild on the GPU box, where the in-tree library travels with the snapshot.
    if not os.path.exists(os.path.join(_PKG, "lib", "libnfb200.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(autouse=True)
def _seed():
    np.random.seed(42)


@pytest.fixture(scope="session")
def be():
    """The B200 backend; GPU tests fail loudly (not skip) when the library or device is missing."""
    import nfb200
    b = nfb200.get_backend("b200")
    return b
