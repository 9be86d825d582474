"""The real-time CC right-hand side (reference ccwfn.py:321-372 as called by rt/rtcc.py:136-141):
``residuals(F, t1, t2, real_time=True)`` with COMPLEX amplitudes and a field-dressed (real symmetric or complex
Hermitian) Fock matrix.  The numpy oracle and the product are checked against outputs of the unmodified reference
(tests/golden/cplx_*.npz).  `emu` / `cuda` as in test_ccsd.py.  FP64 tolerance: 1e-9 max-abs (north_star); the golden
checks are tighter."""
import glob
import os

import numpy as np
import pytest
import torch

import pycc_b200
from pycc_b200.synthetic import Synthetic, blocks_from_factor, make_synthetic
from oracle import ccsd_oracle as co, lambda_oracle as lo
from tests import emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPLX = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "cplx_*.npz")))
DEV = [torch.device("cpu")]


def load(path):
    g = dict(np.load(path))
    tag = os.path.basename(path)[5:-4]
    r = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % tag)))
    syn = Synthetic(int(r["no"]), int(r["nv"]), r["B"], r["F"], float(r["scale"]), int(r["seed"]))
    return g, r, syn


@pytest.fixture(params=CPLX, ids=[os.path.basename(p)[5:-4] for p in CPLX])
def cplx(request):
    return load(request.param)


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
    return torch.from_numpy(np.array(x, order="C", copy=True)).to(DEV[0])


def test_oracle_complex_residuals(cplx):
    g, r, syn = cplx
    P = co.Problem(blocks_from_factor(syn), syn.F, syn.no)
    for name in ("el", "mag"):
        r1, r2 = P.residuals(g["F_" + name], g["t1"], g["t2"])
        assert np.abs(r1 - g["r1_" + name]).max() < 1e-12, name
        assert np.abs(r2 - g["r2_" + name]).max() < 1e-12, name


@pytest.mark.parametrize("field", ["el", "mag"])
def test_complex_residuals_match_reference(cplx, dev, field):
    g, r, syn = cplx
    cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
    r1, r2 = cc.residuals(T(g["F_" + field]), T(g["t1"]), T(g["t2"]), real_time=True)
    assert r1.is_complex() and r2.is_complex()
    assert np.abs(r1.cpu().numpy() - g["r1_" + field]).max() < 1e-11
    assert np.abs(r2.cpu().numpy() - g["r2_" + field]).max() < 1e-11
    # the plane-wise evaluation of the heavy products switched off (samples only) gives the same
    cc.complex_native_heavy = False
    r1c, r2c = cc.residuals(T(g["F_" + field]), T(g["t1"]), T(g["t2"]), real_time=True)
    assert np.abs(r2c.cpu().numpy() - g["r2_" + field]).max() < 1e-11
    cc.complex_native_heavy = True
    # numpy inputs (as rtcc hands them over on the reference's CPU path) are accepted too
    r1b, _ = cc.residuals(g["F_" + field], T(g["t1"]), T(g["t2"]))
    assert np.abs(r1b.cpu().numpy() - g["r1_" + field]).max() < 1e-11


def test_pair_symmetric_complex_amplitudes(cplx, dev):
    """Amplitudes with t2[i,j,a,b] = t2[j,i,b,a] in both planes (what an RT-CC propagation carries) take the (i >= j)
    formulation of ``iterate`` in every sample, the ladder and the o^3v^3 products on the two planes (3M for the ring
    terms); every setting of the switches gives the reference's complex einsum result (oracle pinned by the goldens)."""
    g, r, syn = cplx
    P = co.Problem(blocks_from_factor(syn), syn.F, syn.no)
    t2s = 0.5 * (g["t2"] + g["t2"].transpose(1, 0, 3, 2))
    cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
    assert cc._pair_symmetric(T(t2s.real)) and cc._pair_symmetric(T(t2s.imag))
    assert not cc._pair_symmetric(T(g["t2"].real))
    for field in ("el", "mag"):
        want1, want2 = P.residuals(g["F_" + field], g["t1"], t2s)
        for pair in (True, False):
            for native, heavy in ((True, True), (True, False), (False, False)):
                cc.complex_pair_mode, cc.complex_native_ladder, cc.complex_native_heavy = pair, native, heavy
                r1, r2 = cc.residuals(T(g["F_" + field]), T(g["t1"]), T(t2s), real_time=True)
                assert np.abs(r1.cpu().numpy() - want1).max() < 1e-11, (field, pair, native, heavy)
                assert np.abs(r2.cpu().numpy() - want2).max() < 1e-11, (field, pair, native, heavy)
    # complex F with REAL amplitudes: the imaginary plane of tau is absent
    cc.complex_pair_mode = cc.complex_native_ladder = cc.complex_native_heavy = True
    want1, want2 = P.residuals(g["F_mag"], g["t1"].real, t2s.real)
    r1, r2 = cc.residuals(T(g["F_mag"]), T(g["t1"].real.copy()), T(t2s.real.copy()), real_time=True)
    assert np.abs(r1.cpu().numpy() - want1).max() < 1e-11 and np.abs(r2.cpu().numpy() - want2).max() < 1e-11


def test_real_amplitudes_in_complex_container(cplx, dev):
    g, r, syn = cplx
    cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
    r1, r2 = cc.residuals(cc.H.F, T(r["conv_t1"].astype(complex)), T(r["conv_t2"].astype(complex)), real_time=True)
    assert np.abs(r1.cpu().numpy() - g["r1_realamps"]).max() < 1e-11
    assert np.abs(r2.cpu().numpy() - g["r2_realamps"]).max() < 1e-11
    assert np.abs(r2.cpu().numpy().imag).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("no,nv,seed", [(8, 40, 0), (7, 33, 1)])
def test_medium_size_complex_vs_oracle(no, nv, seed):
    syn = make_synthetic(no, nv, seed=seed, fock_noise=0.01)
    P = co.Problem(blocks_from_factor(syn), syn.F, no)
    rng = np.random.default_rng(seed)
    t1 = 0.05 * (rng.standard_normal((no, nv)) + 1j * rng.standard_normal((no, nv)))
    t2 = 0.05 * (rng.standard_normal((no, no, nv, nv)) + 1j * rng.standard_normal((no, no, nv, nv)))
    m = rng.standard_normal(syn.F.shape)
    F = syn.F + 0.02 * (m + m.T) + 0.02j * (m - m.T)
    r1, r2 = P.residuals(F, t1, t2)
    DEV[0] = torch.device("cuda:0")
    try:
        cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
        q1, q2 = cc.residuals(T(F), T(t1), T(t2), real_time=True)
        assert np.abs(q1.cpu().numpy() - r1).max() < 1e-9
        assert np.abs(q2.cpu().numpy() - r2).max() < 1e-9
    finally:
        DEV[0] = torch.device("cpu")


# ---- the Lambda half of rtcc.f: cclambda.residuals(F, t1, t2, l1, l2) with complex t and complex lambda ------------
def test_oracle_complex_lambda_residuals(cplx):
    g, r, syn = cplx
    P = co.Problem(blocks_from_factor(syn), syn.F, syn.no)
    for name in ("el", "mag"):
        q1, q2 = lo.residuals(P, g["t1"], g["t2"], g["l1"], g["l2"], F=g["F_" + name])
        assert np.abs(q1 - g["rl1_" + name]).max() < 1e-12, name
        assert np.abs(q2 - g["rl2_" + name]).max() < 1e-12, name


@pytest.mark.parametrize("field", ["el", "mag"])
def test_complex_lambda_residuals_match_reference(cplx, dev, field):
    g, r, syn = cplx
    cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
    cc.t1, cc.t2 = T(r["conv_t1"]), T(r["conv_t2"])
    lm = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc))
    q1, q2 = lm.residuals(T(g["F_" + field]), T(g["t1"]), T(g["t2"]), T(g["l1"]), T(g["l2"]))
    assert q1.is_complex() and q2.is_complex()
    assert np.abs(q1.cpu().numpy() - g["rl1_" + field]).max() < 1e-10
    assert np.abs(q2.cpu().numpy() - g["rl2_" + field]).max() < 1e-10


def test_complex_lambda_residuals_every_evaluation(cplx, dev):
    """The Lambda right-hand side on pairs of real planes (default), with the three-product form forced for every complex x
    complex term, and from five real samples of the whole residual (round 1): all three equal the reference's output."""
    from pycc_b200 import planes
    g, r, syn = cplx
    cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
    cc.t1, cc.t2 = T(r["conv_t1"]), T(r["conv_t2"])
    lm = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc))
    args = (T(g["F_mag"]), T(g["t1"]), T(g["t2"]), T(g["l1"]), T(g["l2"]))
    keep = planes.THREE_M_MIN_OUT
    try:
        for on_planes, min_out in ((True, keep), (True, 0), (False, keep)):
            lm.complex_on_planes, planes.THREE_M_MIN_OUT = on_planes, min_out
            q1, q2 = lm.residuals(*args)
            assert np.abs(q1.cpu().numpy() - g["rl1_mag"]).max() < 1e-10, (on_planes, min_out)
            assert np.abs(q2.cpu().numpy() - g["rl2_mag"]).max() < 1e-10, (on_planes, min_out)
    finally:
        lm.complex_on_planes, planes.THREE_M_MIN_OUT = True, keep
    # real t and lambda with a complex Hermitian F: missing imaginary planes
    q1, q2 = lm.residuals(T(g["F_mag"]), T(g["t1"].real.copy()), T(g["t2"].real.copy()), T(g["l1"].real.copy()),
                          T(g["l2"].real.copy()))
    P = co.Problem(blocks_from_factor(syn), syn.F, syn.no)
    w1, w2 = lo.residuals(P, g["t1"].real, g["t2"].real, g["l1"].real, g["l2"].real, F=g["F_mag"])
    assert np.abs(q1.cpu().numpy() - w1).max() < 1e-10 and np.abs(q2.cpu().numpy() - w2).max() < 1e-10
