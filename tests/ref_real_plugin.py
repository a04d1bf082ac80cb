"""pytest plugin (`-p ref_real_plugin`): imports the UNMODIFIED reference package through the namespace shim of oracle/compat
before the reference's test files put `/root/reference/src` in front of sys.path (whose package `__init__` cannot import in
this snapshot).  Gives tests/test_reference_suite.py the ground truth: which of the reference's own tests pass on
the reference's own backends here."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "compat"), ROOT, os.path.join(ROOT, "neural-network-from-scratch_b200")]

import neural_arch  # noqa: E402,F401
import neural_arch.backends  # noqa: E402,F401
import neural_arch.exceptions  # noqa: E402,F401
