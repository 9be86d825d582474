"""ctypes binding of libb200cc.so (the C ABI declared in include/b200cc.h).

The library is built in-tree (``pycc_b200/csrc/build.sh`` / ``__graft_entry__.build()``) and
loaded from ``pycc_b200/libb200cc.so``.  There is no CPU implementation anywhere in this
package: if the shared object is missing, or a tensor handed to a kernel is not a CUDA
tensor, the call raises :class:`B200ccError` -- it never falls back to torch/numpy math.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from .exceptions import PyCCError

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200cc.so")

i64 = C.c_longlong
dptr = C.c_void_p


class B200ccError(PyCCError, RuntimeError):
    """A libb200cc call failed, or the CUDA extension / device is unavailable."""


class GemmDesc(C.Structure):
    """Mirror of ``b200cc_gemm_desc`` (include/b200cc.h)."""
    _fields_ = [
        ("struct_size", C.c_int),
        ("M", C.c_int), ("N", C.c_int),
        ("transA", C.c_int), ("transB", C.c_int),
        ("K1", C.c_int), ("K2", C.c_int),
        ("A1", dptr), ("B1", dptr), ("A2", dptr), ("B2", dptr),
        ("lda1", i64), ("ldb1", i64), ("lda2", i64), ("ldb2", i64),
        ("strideA1", i64), ("strideB1", i64), ("strideA2", i64), ("strideB2", i64),
        ("C", dptr),
        ("ldc", i64), ("strideC", i64),
        ("alpha", C.c_double), ("beta", C.c_double),
        ("batch", C.c_int),
        ("table", dptr),
        ("table_align16", C.c_int),
        ("ksplit", C.c_int),
        ("workspace", dptr),
        ("out_cube_nv", C.c_int),
        ("bcoords", dptr),
        ("nbA1", C.c_int), ("nbB1", C.c_int), ("nbA2", C.c_int), ("nbB2", C.c_int),
        ("config", C.c_int),
        ("K3", C.c_int), ("K4", C.c_int),
        ("A3", dptr), ("B3", dptr), ("A4", dptr), ("B4", dptr),
        ("lda3", i64), ("ldb3", i64), ("lda4", i64), ("ldb4", i64),
        ("strideA3", i64), ("strideB3", i64), ("strideA4", i64), ("strideB4", i64),
        ("nbA3", C.c_int), ("nbB3", C.c_int), ("nbA4", C.c_int), ("nbB4", C.c_int),
    ]


class Gemm3Desc(C.Structure):
    """Mirror of ``b200cc_gemm3_desc`` (include/b200cc.h)."""
    _fields_ = [
        ("struct_size", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("Ahi", dptr), ("Alo", dptr), ("Bhi", dptr), ("Blo", dptr),
        ("lda", i64), ("ldb", i64), ("strideA", i64), ("strideB", i64),
        ("C", dptr),
        ("ldc", i64), ("strideC", i64),
        ("alpha", C.c_double), ("beta", C.c_double),
        ("batch", C.c_int),
        ("kchunk", C.c_int),
        ("config", C.c_int),
        ("K2", C.c_int),
        ("A2hi", dptr), ("A2lo", dptr), ("B2hi", dptr), ("B2lo", dptr),
        ("lda2", i64), ("ldb2", i64), ("strideA2", i64), ("strideB2", i64),
        ("bcoords", dptr),
        ("nbA1", C.c_int), ("nbB1", C.c_int), ("nbA2", C.c_int), ("nbB2", C.c_int),
        ("lockstep", C.c_int),
    ]


class T3dDesc(C.Structure):
    """Mirror of ``b200cc_t3d_desc`` (include/b200cc.h)."""
    _fields_ = [
        ("no", C.c_int), ("nv", C.c_int), ("i", C.c_int), ("j", C.c_int), ("k0", C.c_int), ("nk", C.c_int),
        ("swap_ab", C.c_int),
        ("M3", dptr), ("t1", dptr), ("t2s", dptr), ("oovvs", dptr), ("fov", dptr),
        ("ldf", i64),
        ("eo", dptr), ("ev", dptr),
        ("W2ab", dptr), ("W2n", dptr), ("Pab", dptr), ("Pn", dptr),
        ("Gij", dptr), ("Xij", dptr),
        ("dvv", dptr), ("Dov", dptr), ("S1", dptr),
        ("scratch", dptr),
    ]


class TAbcDesc(C.Structure):
    """Mirror of ``b200cc_t_abc_desc`` (include/b200cc.h)."""
    _fields_ = [
        ("struct_size", C.c_int),
        ("no", C.c_int), ("nv", C.c_int),
        ("nabc", C.c_int), ("nsorted", C.c_int),
        ("abc", dptr), ("sorted", dptr),
        ("G", dptr), ("t2", dptr), ("t2x", dptr), ("Ox", dptr), ("oovvx", dptr), ("t1", dptr),
        ("fov", dptr),
        ("ldf", i64),
        ("eo", dptr), ("ev", dptr),
        ("wtile", dptr), ("partial", dptr), ("et_out", dptr),
        ("accumulate", C.c_int),
        ("grid", C.c_int),
        ("fov_is_zero", C.c_int),
    ]


# name -> (restype, argtypes); every symbol include/b200cc.h declares
SIGNATURES = {
    "b200cc_version": (C.c_int, []),
    "b200cc_last_error": (C.c_char_p, []),
    "b200cc_launch_count": (i64, []),
    "b200cc_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3 + [C.POINTER(i64)] * 2),
    "b200cc_dgemm": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "b200cc_split_tf32": (C.c_int, [dptr, i64, i64, C.c_int, C.c_int, C.c_int, dptr, dptr, i64, C.c_void_p]),
    "b200cc_gemm_tf32x3": (C.c_int, [C.POINTER(Gemm3Desc), C.c_void_p]),
    "b200cc_merge_tf32": (C.c_int, [dptr, dptr, i64, i64, C.c_int, dptr, i64, C.c_void_p]),
    "b200cc_pair_count": (i64, [C.c_int]),
    "b200cc_pack_pairs": (C.c_int, [dptr, i64, i64, i64, i64, C.c_int, C.c_int, C.c_int, dptr, dptr, i64, C.c_void_p]),
    "b200cc_unpack_pairs": (C.c_int, [dptr, dptr, i64, C.c_int, i64, dptr, C.c_void_p]),
    "b200cc_pack_tau": (C.c_int, [dptr, C.c_int, C.c_int, C.c_int, dptr, dptr, i64, C.c_void_p]),
    "b200cc_ladder_unpack": (C.c_int, [dptr, dptr, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, dptr,
                                       C.c_void_p]),
    "b200cc_pack_rows": (C.c_int, [dptr, i64, C.c_int, dptr, dptr, i64, C.c_void_p]),
    "b200cc_pair_rows_unpack": (C.c_int, [dptr, dptr, i64, C.c_int, i64, dptr, i64, C.c_void_p]),
    "b200cc_ring_layouts": (C.c_int, [dptr, C.c_int, C.c_int, dptr, dptr, C.c_void_p]),
    "b200cc_permute": (C.c_int, [C.c_int, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64),
                                 C.c_double, dptr, C.c_double, dptr, C.c_void_p]),
    "b200cc_axpbyz": (C.c_int, [i64, C.c_double, dptr, C.c_double, dptr, dptr, C.c_void_p]),
    "b200cc_build_tau": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_double, dptr, dptr, dptr, C.c_void_p]),
    "b200cc_div_d2": (C.c_int, [C.c_int, C.c_int, dptr, dptr, dptr, dptr, C.c_void_p]),
    "b200cc_div_d1": (C.c_int, [C.c_int, C.c_int, dptr, dptr, dptr, dptr, C.c_void_p]),
    "b200cc_update_amps": (C.c_int, [C.c_int, C.c_int, dptr, dptr, dptr, dptr, C.c_int, C.c_int,
                                     dptr, dptr, dptr, dptr, C.c_void_p]),
    "b200cc_update_amps_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, dptr, dptr, dptr, dptr, dptr, dptr, dptr,
                                          dptr, C.c_void_p]),
    "b200cc_cc_energy_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dptr, i64, dptr, dptr, dptr, dptr,
                                        dptr, C.c_void_p]),
    "b200cc_symmetrize_r2": (C.c_int, [C.c_int, C.c_int, dptr, C.c_void_p]),
    "b200cc_cc_energy": (C.c_int, [C.c_int, C.c_int, dptr, i64, dptr, dptr, dptr, dptr, dptr, C.c_void_p]),
    "b200cc_multi_dot": (C.c_int, [i64, dptr, C.c_int, C.POINTER(dptr), dptr, dptr, C.c_void_p]),
    "b200cc_multi_axpy": (C.c_int, [i64, C.c_int, C.POINTER(C.c_double), C.POINTER(dptr), dptr, C.c_void_p]),
    "b200cc_t_energy_scratch": (i64, [C.c_int, C.c_int]),
    "b200cc_t_q_size": (i64, [C.c_int, C.c_int]),
    "b200cc_t_energy_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, dptr, dptr, C.c_int, dptr, dptr, dptr, dptr, i64,
                                        dptr, dptr, dptr, C.c_int, dptr, C.c_void_p]),
    "b200cc_t3_assemble": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dptr, C.c_int, dptr, dptr, dptr,
                                     dptr, i64, dptr, dptr, C.c_int, dptr, dptr, C.c_void_p]),
    "b200cc_t3_connected_batch": (C.c_int, [C.c_int, C.c_int, C.c_int, dptr, dptr, dptr, dptr, dptr, C.c_void_p]),
    "b200cc_t3_density_scratch": (i64, [C.c_int]),
    "b200cc_t3_density_forms": (C.c_int, [C.POINTER(T3dDesc), C.c_void_p]),
    "b200cc_t_abc_max_no": (C.c_int, []),
    "b200cc_t_abc": (C.c_int, [C.POINTER(TAbcDesc), C.c_void_p]),
}

_LIB = None
# Test seam ONLY (tests/emu.py): lets the host-side logic be driven on CPU tensors against a numpy
# double of the C ABI.  The product never clears it.
REQUIRE_CUDA = True


def load(path=LIB_PATH):
    """dlopen libb200cc.so and attach prototypes.  Raises B200ccError if it is not built."""
    if not os.path.exists(path):
        raise B200ccError(
            "libb200cc.so not found at %s: build it with pycc_b200/csrc/build.sh "
            "(or __graft_entry__.build()).  pycc_b200 has no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


def get():
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = get().b200cc_last_error()
        raise B200ccError("%s failed: %s" % (what, msg.decode() if isinstance(msg, bytes) else msg))


def ptr(t):
    """Device address of a float64/int tensor (validated)."""
    if t is None:
        return None
    if REQUIRE_CUDA and not t.is_cuda:
        raise B200ccError("pycc_b200 kernels need CUDA tensors (got a %s tensor); there is no CPU path"
                          % t.device)
    return t.data_ptr()


def stream():
    if REQUIRE_CUDA:
        return torch.cuda.current_stream().cuda_stream
    return None
