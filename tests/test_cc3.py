"""CC3 ground-state T equations (reference ccwfn.py:374-430, 947-1120): the numpy oracle against
the reference's golden vectors, and the product (CCwfn(model='CC3'): build_cc3_W*, _cc3_t_residual, residuals,
solve_cc) against both.  `emu` / `cuda` as in test_ccsd.py.  FP64 tolerances: 1e-10 Eh / 1e-9 (north_star)."""
import glob
import os

import numpy as np
import pytest
import torch

import pycc_b200
from pycc_b200.synthetic import Synthetic, blocks_from_factor, make_synthetic
from oracle import ccsd_oracle as co, cc3_oracle as c3
from tests import emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CC3 = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "cc3_*.npz")))
DEV = [torch.device("cpu")]
WS = ("Wmnij", "Wmbij", "Wmnie", "Wamef", "Wabei")


def load(path):
    g = dict(np.load(path))
    tag = os.path.basename(path)[4:-4]
    r = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % tag)))
    syn = Synthetic(int(r["no"]), int(r["nv"]), r["B"], r["F"], float(r["scale"]), int(r["seed"]))
    return g, r, syn


@pytest.fixture(params=CC3, ids=[os.path.basename(p)[4:-4] for p in CC3])
def cc3(request):
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
    return torch.from_numpy(np.array(x, dtype=np.float64, order="C", copy=True)).to(DEV[0])


def test_oracle_cc3(cc3):
    g, r, syn = cc3
    P = co.Problem(blocks_from_factor(syn), syn.F, syn.no)
    W = c3.intermediates(P, g["t1"])
    for k in WS:
        assert np.abs(W[k] - g[k]).max() < 1e-13, k
    X1, X2 = c3.t_residual(P, syn.F, g["t1"], g["t2"], W)
    assert np.abs(X1 - g["X1"]).max() < 1e-13 and np.abs(X2 - g["X2"]).max() < 1e-13
    r1, r2 = c3.residuals(P, syn.F, g["t1"], g["t2"])
    assert np.abs(r1 - g["r1"]).max() < 1e-12 and np.abs(r2 - g["r2"]).max() < 1e-12


def test_oracle_cc3_solve(cc3):
    g, r, syn = cc3
    P = co.Problem(blocks_from_factor(syn), syn.F, syn.no)
    e, t1, t2, trace = c3.solve_cc(P, 1e-12, 1e-12)
    assert len(trace) == len(g["trace_ecc_rms"]) and abs(e - float(g["ecc"])) < 1e-12
    assert np.abs(t2 - g["conv_t2"]).max() < 1e-10


def test_cc3_intermediates_and_triples(cc3, dev):
    g, r, syn = cc3
    cc = pycc_b200.ccwfn(syn, model="CC3", device="GPU", quiet=True)
    o, v, H = cc.o, cc.v, cc.H
    t1, t2 = T(g["t1"]), T(g["t2"])
    Wmnij = cc.build_cc3_Wmnij(o, v, H.ERI, t1)
    got = {"Wmnij": Wmnij, "Wmbij": cc.build_cc3_Wmbij(o, v, H.ERI, t1, Wmnij),
           "Wmnie": cc.build_cc3_Wmnie(o, v, H.ERI, t1), "Wamef": cc.build_cc3_Wamef(o, v, H.ERI, t1),
           "Wabei": cc.build_cc3_Wabei(o, v, H.ERI, t1)}
    for k in WS:
        assert np.abs(got[k].cpu().numpy() - g[k]).max() < 1e-12, k
    Fme = cc.build_Fme(o, v, H.F, H.L, t1)
    X1, X2 = cc._cc3_t_residual(o, v, H.F, H.ERI, H.L, t1, t2, Fme)
    assert np.abs(X1.cpu().numpy() - g["X1"]).max() < 1e-12
    assert np.abs(X2.cpu().numpy() - g["X2"]).max() < 1e-12
    r1, r2 = cc.residuals(H.F, t1, t2)
    assert np.abs(r1.cpu().numpy() - g["r1"]).max() < 1e-12
    assert np.abs(r2.cpu().numpy() - g["r2"]).max() < 1e-12


def test_cc3_solve_trace(cc3, dev):
    g, r, syn = cc3
    cc = pycc_b200.ccwfn(syn, model="CC3", device="GPU", quiet=True)
    e = cc.solve_cc(1e-12, 1e-12)
    ref = g["trace_ecc_rms"]
    tr = np.array(cc.trace)
    assert len(tr) == len(ref)
    assert np.abs(tr[:, 0] - ref[:, 0]).max() < 1e-11
    assert abs(float(e) - float(g["ecc"])) < 1e-11
    assert np.abs(cc.t1.cpu().numpy() - g["conv_t1"]).max() < 1e-10
    assert np.abs(cc.t2.cpu().numpy() - g["conv_t2"]).max() < 1e-10


def test_cc3_odd_sizes(dev):
    """odd o / v: the non-TMA operand path of the t3 GEMMs with dressed blocks; ragged cubes; k-run chunking"""
    from pycc_b200 import cctriples
    for (no, nv, kb) in ((3, 5, None), (5, 6, 2)):
        syn = make_synthetic(no, nv, seed=11, fock_noise=0.01)
        P = co.Problem(blocks_from_factor(syn), syn.F, no)
        rng = np.random.default_rng(2)
        t1 = 0.05 * rng.standard_normal((no, nv))
        t2 = 0.05 * rng.standard_normal((no, no, nv, nv))
        X1, X2 = c3.t_residual(P, syn.F, t1, t2)
        cc = pycc_b200.ccwfn(syn, model="CC3", device="GPU", quiet=True)
        o, v, H = cc.o, cc.v, cc.H
        a1, a2 = T(t1), T(t2)
        Wmnij = cc.build_cc3_Wmnij(o, v, H.ERI, a1)
        W = {"Wmbij": cc.build_cc3_Wmbij(o, v, H.ERI, a1, Wmnij), "Wmnie": cc.build_cc3_Wmnie(o, v, H.ERI, a1),
             "Wamef": cc.build_cc3_Wamef(o, v, H.ERI, a1), "Wabei": cc.build_cc3_Wabei(o, v, H.ERI, a1)}
        Y1, Y2 = cctriples.cc3_t_residual(cc, H.F, a1, a2, cc.build_Fme(o, v, H.F, H.L, a1), W, k_batch=kb)
        assert np.abs(Y1.cpu().numpy() - X1).max() < 1e-12, (no, nv)
        assert np.abs(Y2.cpu().numpy() - X2).max() < 1e-12, (no, nv)


# ---- real-time CC3: explicit-field triples (ccwfn.py:421-423, cctriples.py:679-705), real and complex amplitudes ----
RT = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "rtcc3_*.npz")))


def load_rt(path):
    g = dict(np.load(path))
    tag = os.path.basename(path)[6:-4]
    r = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % tag)))
    syn = Synthetic(int(r["no"]), int(r["nv"]), r["B"], r["F"], float(r["scale"]), int(r["seed"]))
    return g, r, syn


@pytest.fixture(params=RT, ids=[os.path.basename(p)[6:-4] for p in RT])
def rt(request):
    return load_rt(request.param)


def TC(x):
    return torch.from_numpy(np.array(x, order="C", copy=True)).to(DEV[0])


def test_oracle_rtcc3(rt):
    g, r, syn = rt
    P = co.Problem(blocks_from_factor(syn), syn.F, syn.no)
    X1, X2 = c3.t_residual(P, g["F_el"], g["t1"], g["t2"], real_time=True)
    assert np.abs(X1 - g["X1_el"]).max() < 1e-13 and np.abs(X2 - g["X2_el"]).max() < 1e-13
    Y1, Y2 = c3.t_residual(P, g["F_el"], g["t1"], g["t2"], real_time=False)
    assert np.abs(Y2 - g["X2_el"]).max() > 1e-5                      # the field term is not negligible in the fixture
    for name in ("el", "mag"):
        r1, r2 = c3.residuals(P, g["F_" + name], g["c1"], g["c2"], real_time=True)
        assert np.abs(r1 - g["c_r1_" + name]).max() < 1e-12, name
        assert np.abs(r2 - g["c_r2_" + name]).max() < 1e-12, name


def test_rtcc3_real_amplitudes(rt, dev):
    g, r, syn = rt
    cc = pycc_b200.ccwfn(syn, model="CC3", device="GPU", quiet=True)
    o, v, H = cc.o, cc.v, cc.H
    F, t1, t2 = T(g["F_el"]), T(g["t1"]), T(g["t2"])
    Fme = cc.build_Fme(o, v, F, H.L, t1)
    X1, X2 = cc._cc3_t_residual(o, v, F, H.ERI, H.L, t1, t2, Fme, real_time=True)
    assert np.abs(X1.cpu().numpy() - g["X1_el"]).max() < 1e-12
    assert np.abs(X2.cpu().numpy() - g["X2_el"]).max() < 1e-12
    r1, r2 = cc.residuals(F, t1, t2, real_time=True)
    assert np.abs(r1.cpu().numpy() - g["r1_el"]).max() < 1e-12
    assert np.abs(r2.cpu().numpy() - g["r2_el"]).max() < 1e-12
    # without a field (F = H.F) real_time changes nothing
    a1, a2 = cc.residuals(H.F, t1, t2, real_time=True)
    b1, b2 = cc.residuals(H.F, t1, t2)
    assert np.abs((a1 - b1).cpu().numpy()).max() < 1e-13 and np.abs((a2 - b2).cpu().numpy()).max() < 1e-13


@pytest.mark.parametrize("field", ["el", "mag"])
def test_rtcc3_complex_amplitudes(rt, dev, field):
    """rtcc.f (rt/rtcc.py:136-141) with model='CC3': complex amplitudes, Hermitian field; six real residuals."""
    g, r, syn = rt
    cc = pycc_b200.ccwfn(syn, model="CC3", device="GPU", quiet=True)
    r1, r2 = cc.residuals(TC(g["F_" + field]), TC(g["c1"]), TC(g["c2"]), real_time=True)
    assert r1.is_complex() and r2.is_complex()
    assert np.abs(r1.cpu().numpy() - g["c_r1_" + field]).max() < 1e-10
    assert np.abs(r2.cpu().numpy() - g["c_r2_" + field]).max() < 1e-10


def test_t3_pert_public_functions(rt, dev):
    """cctriples.t3_pert_ijk / t3_pert_abc in the reference's signatures (cctriples.py:679-720) against the oracle's
    restatement (itself pinned through (X1, X2) of the rtcc3 goldens); the two forms are the same tensor."""
    from pycc_b200 import cctriples
    g, r, syn = rt
    P = co.Problem(blocks_from_factor(syn), syn.F, syn.no)
    cc = pycc_b200.ccwfn(syn, model="CC3", device="GPU", quiet=True)
    o, v = cc.o, cc.v
    F, t2 = T(g["F_el"]), T(g["t2"])
    V = T(g["F_el"] - syn.F)
    for (i, j, k) in ((2, 1, 0), (0, 2, 2), (1, 1, 1)):
        want = c3.t3_pert_ijk(P, i, j, k, g["t2"], g["F_el"] - syn.F, g["F_el"])
        got = cctriples.t3_pert_ijk(o, v, i, j, k, t2, V, F, cc.contract)
        assert np.abs(got.cpu().numpy() - want).max() < 1e-13, (i, j, k)
        raw = cctriples.t3_pert_ijk(o, v, i, j, k, t2, V, F, cc.contract, WithDenom=False)
        for (a, b, c) in ((0, 1, 2), (3, 3, 1)):
            x = cctriples.t3_pert_abc(o, v, a, b, c, t2, V, F, cc.contract, WithDenom=False)
            assert abs(float(x[i, j, k]) - float(raw[a, b, c])) < 1e-14
            y = cctriples.t3_pert_abc(o, v, a, b, c, t2, V, F, cc.contract)
            assert abs(float(y[i, j, k]) - want[a, b, c]) < 1e-13


def test_rtcc3_odd_sizes_and_refusals(dev):
    """odd o / v (the address-table operand path reads the second hole operand), k-run chunking; a Fock matrix with a
    complex DIAGONAL makes the t3 denominators non-polynomial in the sample parameter: refused, not silently wrong."""
    from pycc_b200 import cctriples
    for (no, nv, kb) in ((3, 5, None), (5, 6, 2)):
        syn = make_synthetic(no, nv, seed=11, fock_noise=0.01)
        P = co.Problem(blocks_from_factor(syn), syn.F, no)
        rng = np.random.default_rng(2)
        t1 = 0.05 * rng.standard_normal((no, nv))
        t2 = 0.05 * rng.standard_normal((no, no, nv, nv))
        m = rng.standard_normal(syn.F.shape)
        F = syn.F + 0.03 * (m + m.T)
        X1, X2 = c3.t_residual(P, F, t1, t2, real_time=True)
        cc = pycc_b200.ccwfn(syn, model="CC3", device="GPU", quiet=True)
        o, v, H = cc.o, cc.v, cc.H
        a1, a2, Fd = T(t1), T(t2), T(F)
        Wmnij = cc.build_cc3_Wmnij(o, v, H.ERI, a1)
        W = {"Wmbij": cc.build_cc3_Wmbij(o, v, H.ERI, a1, Wmnij), "Wmnie": cc.build_cc3_Wmnie(o, v, H.ERI, a1),
             "Wamef": cc.build_cc3_Wamef(o, v, H.ERI, a1), "Wabei": cc.build_cc3_Wabei(o, v, H.ERI, a1)}
        V = T(np.ascontiguousarray((F - syn.F)[:no, no:]))
        Y1, Y2 = cctriples.cc3_t_residual(cc, Fd, a1, a2, cc.build_Fme(o, v, Fd, H.L, a1), W, k_batch=kb, V=V)
        assert np.abs(Y1.cpu().numpy() - X1).max() < 1e-12, (no, nv)
        assert np.abs(Y2.cpu().numpy() - X2).max() < 1e-12, (no, nv)
    Fbad = torch.as_tensor(syn.F).to(torch.complex128)
    Fbad[0, 0] += 0.01j
    with pytest.raises(NotImplementedError):
        cc.residuals(Fbad.to(DEV[0]), a1.to(torch.complex128), a2.to(torch.complex128), real_time=True)


@pytest.mark.gpu
@pytest.mark.parametrize("no,nv,seed", [(6, 26, 0), (7, 33, 1)])
def test_medium_size_cc3_vs_oracle(no, nv, seed):
    syn = make_synthetic(no, nv, seed=seed, fock_noise=0.01)
    P = co.Problem(blocks_from_factor(syn), syn.F, no)
    e_ref, t1, t2, trace = c3.solve_cc(P, 1e-11, 1e-11)
    DEV[0] = torch.device("cuda:0")
    try:
        cc = pycc_b200.ccwfn(syn, model="CC3", device="GPU", quiet=True)
        e = cc.solve_cc(1e-11, 1e-11)
        assert len(cc.trace) == len(trace) and abs(float(e) - e_ref) < 1e-10
        assert np.abs(cc.t2.cpu().numpy() - t2).max() < 1e-9
    finally:
        DEV[0] = torch.device("cpu")
