"""The C-ABI library loads on a CPU-only box and exports exactly the symbols include/b200cc.h declares
(no compute calls here)."""
import os
import re

from pycc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    txt = open(os.path.join(ROOT, "include", "b200cc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"\b(b200cc_[a-z0-9_]+)\s*\(", txt))


def test_header_and_binding_agree():
    assert declared() == set(_lib.SIGNATURES)


def test_library_exports_every_symbol():
    lib = _lib.load()                      # raises if the .so is missing or a symbol is absent
    assert lib.b200cc_version() == 200
    for name in declared():
        assert hasattr(lib, name), name
    assert lib.b200cc_last_error() is not None
    assert lib.b200cc_launch_count() >= 0
    assert lib.b200cc_t_energy_scratch(16, 3) == 4 * 3
