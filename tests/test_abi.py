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


def test_integration_doc_struct_is_current():
    """INTEGRATION.md shows the ctypes mirror of b200cc_gemm_desc a maintainer would paste into pycc/device.py: its
    field list must be the binding's (which the struct_size guard then checks against the library at run time)"""
    import ctypes as C
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = txt[txt.index("class _GemmDesc(ctypes.Structure)"):]
    block = block[:block.index("_b200.b200cc_dgemm.argtypes")]
    doc = re.findall(r'\("(\w+)", ctypes\.(\w+)\)', block)
    names = {C.c_int: "c_int", C.c_void_p: "c_void_p", C.c_longlong: "c_longlong", C.c_double: "c_double"}
    assert doc == [(n, names[t]) for n, t in _lib.GemmDesc._fields_]
    # and the C struct has the same members in the same order
    hdr = open(os.path.join(ROOT, "include", "b200cc.h")).read()
    body = hdr[hdr.index("typedef struct {"):hdr.index("} b200cc_gemm_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    members = []
    for decl in body.split(";"):
        decl = decl.replace("typedef struct {", "").strip()
        if not decl:
            continue
        decl = re.sub(r"^(const\s+)?(b200cc_i64|double|int)\s*", "", decl)
        members += [m.strip().lstrip("*").strip() for m in decl.split(",")]
    assert members == [n for n, _ in _lib.GemmDesc._fields_]
