"""HBAR + Lambda: the numpy oracle against the reference's golden vectors, and the product path
(pycc_b200.cchbar / pycc_b200.cclambda) against both.  `emu`: host logic through the numpy double of the C ABI;
`cuda` (-m gpu): the same assertions through libb200cc.so, plus a medium size against the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

import pycc_b200
from pycc_b200.synthetic import Synthetic
from pycc_b200.wavefunction import IntegralReference
from oracle import ccsd_oracle as co, lambda_oracle as lo
from tests import emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAM = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "lam_*.npz")))
HBAR = ("Hov", "Hvv", "Hoo", "Hoooo", "Hvvvv", "Hvovv", "Hooov", "Hovvo", "Hovov", "Hvvvo", "Hovoo")
DEV = [torch.device("cpu")]


def load(path):
    g = dict(np.load(path))
    tag = os.path.basename(path)[4:].rsplit("_", 1)[0]
    r = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % tag)))
    syn = Synthetic(int(r["no"]), int(r["nv"]), r["B"], r["F"], float(r["scale"]), int(r["seed"]))
    return g, syn, str(g["model"])


@pytest.fixture(params=LAM, ids=[os.path.basename(p)[4:-4] for p in LAM])
def lam(request):
    return load(request.param)


def problem(syn):
    from pycc_b200.synthetic import blocks_from_factor
    return co.Problem(blocks_from_factor(syn), syn.F, syn.no)


# ---------------------------------------------------------------------------------------------------------
# oracle vs the reference's own outputs
# ---------------------------------------------------------------------------------------------------------
def test_oracle_hbar_blocks(lam):
    g, syn, model = lam
    H = lo.Hbar(problem(syn), g["t1"], g["t2"], model)
    for k in HBAR:
        assert np.abs(getattr(H, k) - g[k]).max() < 1e-12, k


def test_oracle_lambda_residuals(lam):
    g, syn, model = lam
    P = problem(syn)
    l1, l2 = lo.guess(g["t1"], g["t2"])
    assert np.abs(l1 - g["guess_l1"]).max() < 1e-14 and np.abs(l2 - g["guess_l2"]).max() < 1e-14
    assert np.abs(lo.Goo(g["t2"], g["rand_l2"]) - g["rand_Goo"]).max() < 1e-13
    assert np.abs(lo.Gvv(g["t2"], g["rand_l2"]) - g["rand_Gvv"]).max() < 1e-13
    # cclambda.residuals (202-256) never adds the (T) sources -- only solve_lambda does (145-148)
    r1, r2 = lo.residuals(P, g["t1"], g["t2"], g["rand_l1"], g["rand_l2"], model)
    assert np.abs(r1 - g["rand_r1"]).max() < 1e-12
    assert np.abs(r2 - g["rand_r2"]).max() < 1e-12
    assert abs(lo.pseudoenergy(P, g["rand_l2"]) - float(g["rand_pseudo"])) < 1e-13


def test_oracle_solve_lambda_trace(lam):
    g, syn, model = lam
    lecc, l1, l2, trace = lo.solve_lambda(problem(syn), g["t1"], g["t2"], 1e-12, 1e-12, 100, model=model,
                                          s1=g.get("S1"), s2=g.get("S2"))
    ref = g["trace_lecc_rms"]
    assert len(trace) == len(ref)
    tr = np.array(trace)
    assert np.abs(tr[:, 0] - ref[:, 0]).max() < 1e-11
    assert abs(lecc - float(g["lecc"])) < 1e-12
    assert np.abs(l1 - g["conv_l1"]).max() < 1e-10 and np.abs(l2 - g["conv_l2"]).max() < 1e-10


# ---------------------------------------------------------------------------------------------------------
# product path vs the reference's goldens (emu on CPU, the CUDA kernels under -m gpu)
# ---------------------------------------------------------------------------------------------------------
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


def wfn(syn, model, g):
    cc = pycc_b200.ccwfn(IntegralReference.from_synthetic(syn), model=model, device='GPU', quiet=True)
    cc.t1, cc.t2 = T(g["t1"]), T(g["t2"])
    if "S1" in g:                                  # CCSD(T): the (T) Lambda sources that t3_density leaves on the wfn
        cc.S1, cc.S2 = T(g["S1"]), T(g["S2"])
    return cc


def test_hbar_blocks(lam, dev):
    g, syn, model = lam
    cc = wfn(syn, model, g)
    hb = pycc_b200.cchbar(cc)
    for k in HBAR:
        assert np.abs(getattr(hb, k).cpu().numpy() - g[k]).max() < 1e-12, k
    # the reference's per-block builders, same signatures (cchbar.py:128-823)
    o, v, H = cc.o, cc.v, cc.H
    assert np.abs(hb.build_Hvv(o, v, H.F, H.L, cc.t1, cc.t2).cpu().numpy() - g["Hvv"]).max() < 1e-12
    x = hb.build_Hvvvo(o, v, H.ERI, H.L, hb.Hov, hb.Hvvvv, cc.t1, cc.t2)
    assert np.abs(x.cpu().numpy() - g["Hvvvo"]).max() < 1e-12
    x = hb.build_Hovoo(o, v, H.ERI, H.L, hb.Hov, hb.Hoooo, cc.t1, cc.t2)
    assert np.abs(x.cpu().numpy() - g["Hovoo"]).max() < 1e-12


def test_lambda_residuals(lam, dev):
    g, syn, model = lam
    cc = wfn(syn, model, g)
    lm = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc))
    assert np.abs(lm.l1.cpu().numpy() - g["guess_l1"]).max() < 1e-14
    assert np.abs(lm.l2.cpu().numpy() - g["guess_l2"]).max() < 1e-14
    l1, l2 = T(g["rand_l1"]), T(g["rand_l2"])
    assert np.abs(lm.build_Goo(cc.t2, l2).cpu().numpy() - g["rand_Goo"]).max() < 1e-13
    assert np.abs(lm.build_Gvv(cc.t2, l2).cpu().numpy() - g["rand_Gvv"]).max() < 1e-13
    r1, r2 = lm.residuals(cc.H.F, cc.t1, cc.t2, l1, l2)
    assert np.abs(r1.cpu().numpy() - g["rand_r1"]).max() < 1e-12
    assert np.abs(r2.cpu().numpy() - g["rand_r2"]).max() < 1e-12
    assert abs(float(lm.pseudoenergy(cc.o, cc.v, cc.H.ERI, l2)) - float(g["rand_pseudo"])) < 1e-13


def test_solve_lambda_trace(lam, dev):
    g, syn, model = lam
    cc = wfn(syn, model, g)
    lm = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc))
    lecc = lm.solve_lambda(1e-12, 1e-12, 100)
    ref = g["trace_lecc_rms"]
    tr = np.array(lm.trace)
    assert len(tr) == len(ref)
    assert np.abs(tr[:, 0] - ref[:, 0]).max() < 1e-11
    assert np.all(np.abs(tr[:, 1] - ref[:, 1]) <= 1e-5 * np.abs(ref[:, 1]) + 1e-13)
    assert abs(float(lecc) - float(g["lecc"])) < 1e-11
    assert np.abs(lm.l1.cpu().numpy() - g["conv_l1"]).max() < 1e-10
    assert np.abs(lm.l2.cpu().numpy() - g["conv_l2"]).max() < 1e-10


def test_lambda_not_converged_and_unsupported(dev):
    from pycc_b200.synthetic import make_synthetic
    syn = make_synthetic(3, 6, seed=4)
    cc = pycc_b200.ccwfn(syn, model="CCSD", quiet=True)
    cc.solve_cc(1e-10, 1e-10)
    lm = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc))
    assert lm.solve_lambda(1e-12, 1e-12, maxiter=2) is None
    from pycc_b200.exceptions import PyCCError
    cct = pycc_b200.ccwfn(syn, model="CCSD(T)", quiet=True)
    with pytest.raises(PyCCError):                 # CCSD(T) Lambda needs S1 / S2 (make_t3_density=True)
        pycc_b200.cclambda(cct, pycc_b200.cchbar(cct))


def test_ccsd_t_chain_solve_cc_t3_density_lambda(dev):
    """The whole CCSD(T) chain of the reference (ccwfn.py:300-304 -> cchbar -> cclambda.py:145-148): amplitudes with
    make_t3_density=True, HBAR, Lambda with the (T) sources -- against the reference's own pseudo-energy / amplitudes."""
    path = [p for p in LAM if p.endswith("ccsdpt.npz")][0]
    g, syn, model = load(path)
    assert model == "CCSD(T)"
    cc = pycc_b200.ccwfn(IntegralReference.from_synthetic(syn), model="CCSD(T)", device='GPU', quiet=True,
                         make_t3_density=True)
    e = cc.solve_cc(1e-12, 1e-12)
    assert abs(float(e) - float(g["ecc"])) < 1e-10
    assert np.abs(cc.S1.cpu().numpy() - g["S1"]).max() < 1e-9 and np.abs(cc.S2.cpu().numpy() - g["S2"]).max() < 1e-9
    lm = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc))
    lecc = lm.solve_lambda(1e-12, 1e-12, 100)
    assert abs(float(lecc) - float(g["lecc"])) < 1e-10
    assert np.abs(lm.l1.cpu().numpy() - g["conv_l1"]).max() < 1e-9
    assert np.abs(lm.l2.cpu().numpy() - g["conv_l2"]).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("no,nv,seed", [(8, 40, 0), (10, 64, 2)])
def test_medium_size_lambda_vs_oracle(no, nv, seed):
    """CUDA path vs the numpy oracle at sizes the oracle does in seconds: HBAR blocks to 1e-11, pseudo-energy to
    1e-10, converged lambda amplitudes to 1e-9."""
    from pycc_b200.synthetic import make_synthetic, blocks_from_factor
    DEV[0] = torch.device("cuda:0")
    try:
        syn = make_synthetic(no, nv, seed=seed)
        P = co.Problem(blocks_from_factor(syn), syn.F, no)
        cc = pycc_b200.ccwfn(syn, model="CCSD", device='GPU', quiet=True)
        cc.solve_cc(1e-11, 1e-11, 100)
        t1, t2 = cc.t1.cpu().numpy(), cc.t2.cpu().numpy()
        Href = lo.Hbar(P, t1, t2, "CCSD")
        hb = pycc_b200.cchbar(cc)
        for k in HBAR:
            ref = getattr(Href, k)
            assert np.abs(getattr(hb, k).cpu().numpy() - ref).max() < 1e-11 * max(1.0, np.abs(ref).max()), k
        lecc_ref, l1_ref, l2_ref, trace = lo.solve_lambda(P, t1, t2, 1e-11, 1e-11, 100)
        lm = pycc_b200.cclambda(cc, hb)
        lecc = lm.solve_lambda(1e-11, 1e-11, 100)
        assert len(lm.trace) == len(trace)
        assert abs(float(lecc) - lecc_ref) < 1e-10
        assert np.abs(lm.l1.cpu().numpy() - l1_ref).max() < 1e-9
        assert np.abs(lm.l2.cpu().numpy() - l2_ref).max() < 1e-9
    finally:
        DEV[0] = torch.device("cpu")
