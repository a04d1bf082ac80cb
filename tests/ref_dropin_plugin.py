"""pytest plugin (`-p ref_dropin_plugin`): the UNMODIFIED reference package (NEURAL_ARCH_REFERENCE, normally oracle/_ref) through
the namespace shim of oracle/compat, with the B200 backend registered into the reference's OWN registry (the drop-in step of
INTEGRATION.md: `import nfb200; nfb200.install_into_reference()`) before the reference's test files are collected."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "compat"), ROOT, os.path.join(ROOT, "neural-network-from-scratch_b200")]

import neural_arch  # noqa: E402,F401
import neural_arch.backends  # noqa: E402,F401
import neural_arch.exceptions  # noqa: E402,F401
import nfb200  # noqa: E402

if nfb200.install_into_reference() and "b200" in neural_arch.backends.available_backends():
    sys.stderr.write("B200_REGISTERED\n")
