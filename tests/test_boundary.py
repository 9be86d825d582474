"""The drop-in boundary: the reference's own, unmodified ``CCwfn.residuals`` / ``solve_cc`` running with
THIS package's contraction backend plugged into its ``contract`` seam (pycc/device.py:64); swapped-in (perturbed)
integrals through the generic path (ccderiv.py:250-259); the struct_size guard of the C ABI.  `emu` / `cuda` as in
test_ccsd.py; the tests that execute reference code need the reference tree (build container: /root/reference, GPU box:
baseline/_ref) and are skipped without it."""
import ctypes as C
import warnings

import numpy as np
import pytest
import torch

import pycc_b200
from pycc_b200 import _lib
from pycc_b200.synthetic import full_eri
from baseline import refload
from tests import emu
from tests.conftest import GOLDEN, load_golden

DEV = [torch.device("cpu")]
needs_reference = pytest.mark.skipif(refload.reference_root() is None, reason="reference tree not available")


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    if request.param == "emu":
        DEV[0] = torch.device("cpu")
        with emu.install():
            yield DEV[0]
    else:
        assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
        DEV[0] = torch.device("cuda:0")
        yield DEV[0]
        DEV[0] = torch.device("cpu")


def T(x):
    return torch.from_numpy(np.array(x, dtype=np.float64, order="C", copy=True)).to(DEV[0])


@needs_reference
@pytest.mark.parametrize("path", GOLDEN[:2], ids=lambda p: p.split("ref_")[-1][:-4])
def test_reference_residuals_with_our_contraction_backend(dev, path):
    """wfn.contract = pycc_b200.ContractionBackend('GPU'): the reference's residuals / solve_cc, every contraction on the
    b200cc kernels, against the golden vectors the same code produced with opt_einsum on numpy"""
    g, syn = load_golden(path)
    ref = refload.load_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                      # 'GPU' without CUDA falls back to cpu tensors (emu lane)
        w = refload.reference_wfn(ref, syn.F, syn.no, full_eri(syn), device="GPU",
                                  contract=pycc_b200.ContractionBackend("GPU", device1=DEV[0]))
    assert isinstance(w.t2, torch.Tensor) and w.t2.device.type == DEV[0].type
    n0 = w.contract.engine.stats["gemm"]
    r1, r2 = w.residuals(w.H.F, T(g["rand_t1"]), T(g["rand_t2"]))
    assert w.contract.engine.stats["gemm"] > n0 + 30          # every einsum of the reference went through the seam
    assert np.abs(r1.cpu().numpy() - g["rand_r1"]).max() < 1e-12
    assert np.abs(r2.cpu().numpy() - g["rand_r2"]).max() < 1e-12
    secs, en = refload.timed_solve_cc(w, 4)
    assert np.abs(np.array(en[1:]) - g["trace_ecc_rms"][:4, 0]).max() < 1e-11


def test_swapped_integrals_take_the_generic_path(dev):
    """ccderiv.py:250-259 swaps cc.H.ERI / cc.H.L and calls cc.residuals: with the wavefunction's own integrals handed
    back as plain n^4 host arrays the result must be the golden residual; afterwards the fused path is back"""
    g, syn = load_golden(GOLDEN[0])
    cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
    t1, t2 = T(g["rand_t1"]), T(g["rand_t2"])
    ERI = full_eri(syn)
    keep = cc.H.ERI, cc.H.L
    assert not cc._foreign()
    cc.H.ERI, cc.H.L = ERI, 2.0 * ERI - ERI.swapaxes(2, 3)
    try:
        assert cc._foreign()
        r1, r2 = cc.residuals(cc.H.F, t1, t2)
    finally:
        cc.H.ERI, cc.H.L = keep
    assert np.abs(r1.cpu().numpy() - g["rand_r1"]).max() < 1e-12
    assert np.abs(r2.cpu().numpy() - g["rand_r2"]).max() < 1e-12
    assert not cc._foreign()
    r1, r2 = cc.residuals(cc.H.F, t1, t2)
    assert np.abs(r2.cpu().numpy() - g["rand_r2"]).max() < 1e-12
    # the public builders take the integrals as arguments (ccwfn.py:458-715): foreign ones go the same way
    L = 2.0 * ERI - ERI.swapaxes(2, 3)
    o, v = cc.o, cc.v
    assert np.abs(cc.build_Wmbej(o, v, ERI, L, t1, t2).cpu().numpy() - g["rand_Wmbej"]).max() < 1e-12
    assert np.abs(cc.build_Zmbij(o, v, ERI, t1, t2).cpu().numpy() - g["rand_Zmbij"]).max() < 1e-12
    assert np.abs(cc.build_Fae(o, v, cc.H.F, L, t1, t2).cpu().numpy() - g["rand_Fae"]).max() < 1e-12


@needs_reference
@pytest.mark.parametrize("model", ["CCSD", "CCD"])
def test_perturbed_integrals_without_permutational_symmetry(dev, model):
    """integrals with NO 8-fold symmetry (a derivative-integral stand-in): generic path vs the reference's own code"""
    g, syn = load_golden(GOLDEN[0])
    rng = np.random.default_rng(11)
    n = syn.n
    ERI = full_eri(syn) + 0.02 * rng.standard_normal((n, n, n, n))
    L = 2.0 * ERI - ERI.swapaxes(2, 3)
    F = syn.F + 0.01 * rng.standard_normal((n, n))
    ref = refload.load_reference()
    w = refload.reference_wfn(ref, F, syn.no, ERI, L=L, model=model)
    t1 = np.zeros_like(g["rand_t1"]) if model == "CCD" else g["rand_t1"]
    r1_ref, r2_ref = w.residuals(F, t1, g["rand_t2"])
    cc = pycc_b200.ccwfn(syn, model=model, device="GPU", quiet=True)
    cc.H.ERI, cc.H.L = ERI, L
    r1, r2 = cc.residuals(F, T(t1), T(g["rand_t2"]))
    assert np.abs(r1.cpu().numpy() - r1_ref).max() < 1e-12
    assert np.abs(r2.cpu().numpy() - r2_ref).max() < 1e-12


def test_stale_descriptor_layout_is_refused():
    """a binding built against another layout of b200cc_gemm_desc / b200cc_gemm3_desc must get an error, not a mis-read
    struct (no device work happens before the check, so this runs without a GPU)"""
    lib = _lib.load()
    d = _lib.GemmDesc()
    d.struct_size = C.sizeof(_lib.GemmDesc) - 4
    d.M = d.N = d.K1 = 8
    assert lib.b200cc_dgemm(C.byref(d), None) != 0
    assert b"stale binding" in lib.b200cc_last_error()
    d3 = _lib.Gemm3Desc()
    d3.M = d3.N = d3.K = 8
    assert lib.b200cc_gemm_tf32x3(C.byref(d3), None) != 0
    assert b"stale binding" in lib.b200cc_last_error()


@pytest.mark.parametrize("sub,shapes,kinds", [
    ("ijef,abef->ijab", [(3, 3, 5, 5), (4, 4, 5, 5)], "cc"),
    ("imae,mbej->ijab", [(3, 3, 4, 4), (3, 4, 4, 3)], "cr"),
    ("me,ma->ae", [(3, 5), (3, 5)], "rc"),
    ("ijab->jiba", [(2, 3, 4, 5)], "c"),
    ("ie,ma,mbej->ijab", [(3, 4), (3, 4), (3, 4, 4, 3)], "crc"),
])
def test_complex_operands_in_the_contraction_seam(dev, sub, shapes, kinds):
    """device.py:79-83: real and complex operands may be mixed in one contraction (RT-CC); results are complex128"""
    rng = np.random.default_rng(len(sub))
    ops = []
    for shp, k in zip(shapes, kinds):
        x = rng.standard_normal(shp)
        ops.append(x + 1j * rng.standard_normal(shp) if k == "c" else x)
    ct = pycc_b200.ContractionBackend("GPU", device1=DEV[0])
    dev_ops = [torch.from_numpy(np.ascontiguousarray(x)).to(DEV[0]) for x in ops]
    got = ct(sub, *dev_ops)
    assert got.dtype == torch.complex128
    assert np.abs(got.cpu().numpy() - np.einsum(sub, *ops)).max() < 1e-12
