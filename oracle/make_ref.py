#!/usr/bin/env python
"""ORACLE INFRASTRUCTURE (not product code): stage the UNMODIFIED reference where the GPU box can see it.

    python oracle/make_ref.py            # needs /root/reference; writes oracle/_ref/ (git-ignored, NOT gpurun-ignored)

/root/reference does not exist on the GPU box, and the reference is pure Python (no native code to compile), so
"building the checker" is a byte-for-byte copy of the reference's package `src/neural_arch` and of the few test files of its
own suite that the drop-in tests run (tests/test_gpu_reference_dropin.py, bench.py --impl reference) into oracle/_ref/.
Like the built .so files, oracle/_ref/ travels with the working tree and stays out of the history; nothing under it is ever
imported by the product path.  The snapshot's missing `neural_arch.core` is supplied by oracle/compat at import time
(NEURAL_ARCH_REFERENCE selects the tree)."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
TESTS = ["test_backends.py", "test_backend_accuracy.py", "test_backend_utils.py", "test_mathematical_correctness.py"]


def stage(verbose=True) -> bool:
    pkg = os.path.join(SRC, "src", "neural_arch")
    if not os.path.isdir(pkg):
        if verbose:
            print(f"make_ref: {pkg} not found; oracle/_ref left as it is")
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(pkg, os.path.join(DST, "neural_arch"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(DST, "tests"), exist_ok=True)
    for t in TESTS:
        p = os.path.join(SRC, "tests", t)
        if os.path.exists(p):
            shutil.copy2(p, os.path.join(DST, "tests", t))
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified copy of the reference (src/neural_arch + a few of its own test files), staged by oracle/make_ref.py for the\n"
                "GPU box.  Test infrastructure; git-ignored; never imported by the product path.\n")
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"make_ref: staged {n} files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
