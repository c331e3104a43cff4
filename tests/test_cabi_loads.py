"""CPU: libimb.so loads and exports every symbol include/imb.h declares (no compute calls)."""
import os
import re

from imitation_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_header_symbols():
    _build.build()
    lib = _lib.lib()
    header = open(os.path.join(ROOT, "include", "imb.h")).read()
    declared = set(re.findall(r"\b(imb_[a-z_0-9]+)\s*\(", header))
    declared -= {"imb_mlp", "imb_disc_desc"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in imb.h but not exported"
    assert set(_lib.SYMBOLS) == declared
    assert lib.imb_version() >= 1


def test_struct_sizes_match_header_layout():
    import ctypes as C

    assert C.sizeof(_lib.Mlp) == 40
    assert C.sizeof(_lib.DiscDesc) == 24 + 40 + 4 + 40 + 12
    assert C.sizeof(_lib.Adam) == 20
    assert C.sizeof(_lib.EnvDesc) == 32
    assert C.sizeof(_lib.PpoHparams) == 44
    assert C.sizeof(_lib.PolicyDesc) == 4 * 20
