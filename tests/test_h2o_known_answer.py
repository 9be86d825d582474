"""Known-answer tests on a real molecule: H2O / STO-3G, frozen core -- the numbers the reference's own hot-path tests
hard-code (pycc/tests/test_002_ccsd_energy.py:22-31, test_005_ccsd_t_energy.py:21-36, test_044_ccsd_t_gpu.py:21-38).

The AO integrals come from tests/golden/h2o_sto3g.npz (tests/golden/make_h2o_sto3g.py: s/p McMurchie-Davidson integrals +
RHF in numpy, since psi4 is not installable offline; the unmodified reference, fed with them, reproduces its hard-coded
energies to 1e-14).  Tolerance 1e-11 Eh as in the reference tests.  `emu` / `cuda` as in test_ccsd.py."""
import os

import numpy as np
import pytest
import torch

import pycc_b200
from pycc_b200 import cctriples
from pycc_b200.wavefunction import IntegralReference
from oracle import aomo_oracle as ao
from oracle import ccsd_oracle as co
from oracle import triples_oracle as to
from tests import emu

ECCSD = -0.070616830152761        # test_002_ccsd_energy.py:31
ET = -0.000099957499645           # test_005_ccsd_t_energy.py:33
ECCSD_T = -0.0707167876524093     # test_044_ccsd_t_gpu.py:37
TOL = 1e-11

G = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "h2o_sto3g.npz")))
NO, NV, NFZC = int(G["no"]), int(G["nv"]), int(G["nfzc"])


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    if request.param == "emu":
        with emu.install():
            yield torch.device("cpu")
    else:
        assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
        yield torch.device("cuda:0")


def test_fixture_is_a_converged_rhf():
    S, C, F_ao, eps = G["S"], G["C"], G["F_ao"], G["eps"]
    assert S.shape == (7, 7) and (NO, NV, NFZC) == (4, 2, 1)
    assert np.abs(C.T @ S @ C - np.eye(7)).max() < 1e-12
    assert np.abs(F_ao @ C - S @ C * eps).max() < 1e-11
    # the Fock matrix is the one of its own density
    D = C[:, :5] @ C[:, :5].T
    eri = G["eri_ao"]
    F = G["Hcore"] + 2 * np.einsum("pqrs,rs->pq", eri, D) - np.einsum("prqs,rs->pq", eri, D)
    assert np.abs(F - F_ao).max() < 1e-11
    assert abs(np.sum(D * (G["Hcore"] + F)) + float(G["enuc"]) - float(G["escf"])) < 1e-11
    # 8-fold symmetry of (pq|rs)
    for perm in ((1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)):
        assert np.abs(eri - eri.transpose(perm)).max() < 1e-14


def test_reference_outputs_match_its_hardcoded_numbers():
    """What make_h2o_sto3g.py recorded from the unmodified reference on these integrals."""
    assert abs(float(G["ref_eccsd"]) - ECCSD) < TOL
    assert abs(float(G["ref_et"]) - ET) < TOL
    assert abs(ECCSD + ET - ECCSD_T) < TOL


def test_oracle_known_answer():
    F, ERI, _ = ao.mo_hamiltonian(G["F_ao"], G["eri_ao"], G["C"])
    blocks = co.blocks_from_full(ERI, NO, NFZC)
    P = co.Problem(blocks, F, NO, NFZC)
    ecc, t1, t2, _ = co.solve_cc(P, 1e-12, 1e-12, 75)
    assert abs(ecc - ECCSD) < TOL
    assert np.abs(t1 - G["ref_t1"]).max() < 1e-10 and np.abs(t2 - G["ref_t2"]).max() < 1e-10
    et = to.t_tjl(t1, t2, F, blocks["ovvv"], blocks["ooov"], blocks["oovv"], NFZC)
    assert abs(et - ET) < TOL
    assert abs(to.t_vikings(t1, t2, F, blocks["ovvv"], blocks["ooov"], blocks["oovv"], NFZC) - ET) < TOL


def h2o_reference(kind):
    if kind == "ao":
        return IntegralReference.from_ao(G["F_ao"], G["eri_ao"], G["C"], NO, NFZC)
    F, ERI, _ = ao.mo_hamiltonian(G["F_ao"], G["eri_ao"], G["C"])
    return IntegralReference.from_arrays(F, ERI, NO, NFZC)


@pytest.mark.parametrize("kind", ["ao", "mo"])
def test_ccsd_energy(dev, kind):
    """test_002_ccsd_energy.py:22-31"""
    cc = pycc_b200.ccwfn(h2o_reference(kind), quiet=True)
    eccsd = cc.solve_cc(1e-12, 1e-12, 75)
    assert abs(float(eccsd) - ECCSD) < TOL
    assert np.abs(cc.t2.cpu().numpy() - G["ref_t2"]).max() < 1e-10
    assert np.abs(cc.t1.cpu().numpy() - G["ref_t1"]).max() < 1e-10


def test_ccsd_t_three_formulations(dev):
    """test_005_ccsd_t_energy.py:21-36"""
    cc = pycc_b200.ccwfn(h2o_reference("ao"), model="ccsd(t)", quiet=True)
    total = cc.solve_cc(1e-12, 1e-12, 75)
    assert abs(float(total) - ECCSD_T) < TOL
    for fn in (cctriples.t_vikings, cctriples.t_vikings_inverted, cctriples.t_tjl):
        assert abs(float(fn(cc)) - ET) < TOL, fn.__name__


def test_ccsd_t_gpu_total(dev):
    """test_044_ccsd_t_gpu.py:21-38: device='GPU', the return value must be float()-able."""
    cc = pycc_b200.ccwfn(h2o_reference("ao"), model="CCSD(T)", device="GPU", quiet=True)
    ecc = cc.solve_cc(1e-12, 1e-12, 75)
    assert abs(float(ecc) - ECCSD_T) < TOL


def test_mixed_precision_within_1e6(dev):
    """BASELINE north_star: mixed precision within 1e-6 Eh (the reference's SP test, test_030_sp.py:26-31, uses 1e-7 on
    another basis)."""
    cc = pycc_b200.ccwfn(h2o_reference("ao"), model="CCSD(T)", device="GPU", precision="MP", quiet=True)
    ecc = cc.solve_cc(1e-9, 1e-9, 75)
    assert abs(float(ecc) - ECCSD_T) < 1e-6
