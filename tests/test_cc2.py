"""CC2 (reference ccwfn.py:596-602, 711-713, 832-884): the numpy oracle against the reference's golden vectors, and the
product (CCwfn(model='CC2')) against both.  `emu` / `cuda` as in test_ccsd.py."""
import glob
import os

import numpy as np
import pytest
import torch

import pycc_b200
from pycc_b200.synthetic import Synthetic, blocks_from_factor, make_synthetic
from oracle import ccsd_oracle as co, cc2_oracle as c2
from tests import emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CC2 = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "cc2_*.npz")))
DEV = [torch.device("cpu")]


def load(path):
    g = dict(np.load(path))
    tag = os.path.basename(path)[4:-4]
    r = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % tag)))
    syn = Synthetic(int(r["no"]), int(r["nv"]), r["B"], r["F"], float(r["scale"]), int(r["seed"]))
    return g, r, syn


@pytest.fixture(params=CC2, ids=[os.path.basename(p)[4:-4] for p in CC2])
def cc2(request):
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


def test_oracle_cc2(cc2):
    g, r, syn = cc2
    P = co.Problem(blocks_from_factor(syn), syn.F, syn.no)
    assert np.abs(c2.Wmnij(P, g["t1"]) - g["Wmnij"]).max() < 1e-13
    assert np.abs(c2.Zmbij(P, g["t1"]) - g["Zmbij"]).max() < 1e-13
    r1, r2 = c2.residuals(P, syn.F, g["t1"], g["t2"])
    assert np.abs(r1 - g["r1"]).max() < 1e-12 and np.abs(r2 - g["r2"]).max() < 1e-12
    e, t1, t2, trace = c2.solve_cc(P, 1e-12, 1e-12)
    assert len(trace) == len(g["trace_ecc_rms"]) and abs(e - float(g["ecc"])) < 1e-12


def test_cc2_residuals_and_solve(cc2, dev):
    g, r, syn = cc2
    cc = pycc_b200.ccwfn(syn, model="CC2", device="GPU", quiet=True)
    o, v, H = cc.o, cc.v, cc.H
    t1, t2 = T(g["t1"]), T(g["t2"])
    assert np.abs(cc.build_Wmnij(o, v, H.ERI, t1, t2).cpu().numpy() - g["Wmnij"]).max() < 1e-12
    assert np.abs(cc.build_Zmbij(o, v, H.ERI, t1, t2).cpu().numpy() - g["Zmbij"]).max() < 1e-12
    assert cc.build_Wmbej(o, v, H.ERI, H.L, t1, t2) is None and cc.build_Wmbje(o, v, H.ERI, t1, t2) is None
    r1, r2 = cc.residuals(H.F, t1, t2)
    assert np.abs(r1.cpu().numpy() - g["r1"]).max() < 1e-12
    assert np.abs(r2.cpu().numpy() - g["r2"]).max() < 1e-12
    # the public r_T1 / r_T2 dispatch on the model like the reference's (ccwfn.py:786-787 -> _r_T2_cc2)
    assert np.abs(cc.r_T2(o, v, H.F, H.ERI, t1, t2).cpu().numpy() - g["r2"]).max() < 1e-12
    assert np.abs(cc.r_T1(o, v, H.F, H.ERI, H.L, t1, t2).cpu().numpy() - g["r1"]).max() < 1e-12
    e = cc.solve_cc(1e-12, 1e-12)
    ref = g["trace_ecc_rms"]
    tr = np.array(cc.trace)
    assert len(tr) == len(ref) and np.abs(tr[:, 0] - ref[:, 0]).max() < 1e-11
    assert abs(float(e) - float(g["ecc"])) < 1e-11
    assert np.abs(cc.t2.cpu().numpy() - g["conv_t2"]).max() < 1e-10


@pytest.mark.gpu
def test_medium_size_cc2_vs_oracle():
    no, nv = 8, 40
    syn = make_synthetic(no, nv, seed=4, fock_noise=0.01)
    P = co.Problem(blocks_from_factor(syn), syn.F, no)
    e_ref, t1, t2, trace = c2.solve_cc(P, 1e-11, 1e-11)
    DEV[0] = torch.device("cuda:0")
    try:
        cc = pycc_b200.ccwfn(syn, model="CC2", device="GPU", quiet=True)
        e = cc.solve_cc(1e-11, 1e-11)
        assert len(cc.trace) == len(trace) and abs(float(e) - e_ref) < 1e-10
        assert np.abs(cc.t2.cpu().numpy() - t2).max() < 1e-9
    finally:
        DEV[0] = torch.device("cpu")
