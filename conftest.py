import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_ROOT, "neural-network-from-scratch_b200")
for p in (_ROOT, _PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
