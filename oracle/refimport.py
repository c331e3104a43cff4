"""Import the REAL reference (`/root/reference/src/imitation`) through the names-only shim.

TEST INFRASTRUCTURE; only usable in the build container (the GPU box has no
/root/reference).  Used by `oracle/make_golden.py` and `tests/test_oracle_vs_reference.py`.
"""
import os
import sys

REF_SRC = "/root/reference/src"
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shim")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "imitation"))


def load():
    if not available():
        raise RuntimeError("reference sources not present at /root/reference")
    for p in (SHIM, REF_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    import imitation  # noqa: F401

    return imitation
