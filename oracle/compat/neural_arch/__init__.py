"""ORACLE INFRASTRUCTURE (not product code): a namespace shim that lets the UNMODIFIED reference package at
/root/reference/src/neural_arch import in this container.  The reference snapshot is missing its `core` package
 -- its own .gitignore swallowed it -- so this shim supplies `neural_arch.core` (re-exported from the
reconstruction in nfb200.core, whose contract comes from the reference's call sites and tests) and
takes every other sub-package (`backends`, `functional`, `nn`, `optim`, `models`, ...) straight from the reference
tree.  Used only by oracle/validate_against_reference.py and oracle/gen_golden.py; never on the GPU box."""
import os

REFERENCE_SRC = os.environ.get("NEURAL_ARCH_REFERENCE", "/root/reference/src/neural_arch")
__path__ = [os.path.dirname(os.path.abspath(__file__)), REFERENCE_SRC]
