"""Known-answer tests on a real molecule: H2O in STO-3G and cc-pVDZ (BASELINE.json configs[0]) -- the numbers the
reference's own tests hard-code for this path:

    test_002_ccsd_energy.py:31,38   test_003_ccsd_lambda.py:49,61   test_005_ccsd_t_energy.py:33,44
    test_017_ccd.py:19,25           test_020_cc2.py:19              test_030_sp.py:30
    test_031_cc3.py:31              test_044_ccsd_t_gpu.py:37       test_034_ccsd_t_density.py:44,68

The AO integrals come from tests/golden/h2o_*.npz (tests/golden/make_h2o.py: McMurchie-Davidson integrals + RHF in
numpy, since psi4 is not installable offline; the unmodified reference, fed with them, reproduces every one of its
hard-coded energies to <= 1e-13, recorded as ``dev_*`` in the fixtures).  Tolerance 1e-11 Eh as in the reference's
tests.  Oracle tests run on the CPU; product tests through `emu` / `cuda` as in test_ccsd.py."""
import os

import numpy as np
import pytest
import torch

import pycc_b200
from pycc_b200 import cctriples
from pycc_b200.wavefunction import IntegralReference
from oracle import aomo_oracle as ao
from oracle import ccsd_oracle as co
from oracle import cc2_oracle, cc3_oracle
from oracle import lambda_oracle as lo
from oracle import triples_oracle as to
from tests import emu
from tests.golden import gto

TOL = 1e-11
ECCSD_T_STO3G = -0.0707167876524093     # test_044_ccsd_t_gpu.py:37


def load(tag):
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "h2o_%s.npz" % tag)))
    g["eri_ao"] = gto.unpack_eri(g["eri_packed"], g["S"].shape[0])
    return g


FIX = {tag: load(tag) for tag in ("sto3g", "ccpvdz", "teach_ccpvdz", "t034_sto3g", "t034_ccpvdz", "h2_ccpvdz")}


def hard(tag, core, model, what):
    return float(FIX[tag]["hardcoded_%s_%s_%s" % (core, model.lower().replace("(t)", "pt"), what)])


def ref(tag, core, model, what):
    return FIX[tag]["ref_%s_%s_%s" % (core, model.lower().replace("(t)", "pt"), what)]


def sizes(tag, core):
    nfzc = 1 if core == "fc" else 0
    return int(FIX[tag]["ndocc"]) - nfzc, nfzc


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    if request.param == "emu":
        with emu.install():
            yield torch.device("cpu")
    else:
        assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
        yield torch.device("cuda:0")


# ---------------------------------------------------------------------------------------------- fixtures themselves
@pytest.mark.parametrize("tag", sorted(FIX))
def test_fixture_is_a_converged_rhf(tag):
    g = FIX[tag]
    S, C, F_ao, eps, eri = g["S"], g["C"], g["F_ao"], g["eps"], g["eri_ao"]
    n = S.shape[0]
    nd = int(g["ndocc"])
    assert n == (7 if tag.endswith("sto3g") else 10 if tag.startswith("h2_") else 24)
    assert np.abs(C.T @ S @ C - np.eye(n)).max() < 1e-11
    assert np.abs(F_ao @ C - S @ C * eps).max() < 1e-10
    D = C[:, :nd] @ C[:, :nd].T
    F = g["Hcore"] + 2 * np.einsum("pqrs,rs->pq", eri, D) - np.einsum("prqs,rs->pq", eri, D)
    assert np.abs(F - F_ao).max() < 1e-10                                   # the Fock matrix of its own density
    assert abs(np.sum(D * (g["Hcore"] + F)) + float(g["enuc"]) - float(g["escf"])) < 1e-10
    for perm in ((1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)):
        assert np.abs(eri - eri.transpose(perm)).max() == 0.0


def test_reference_outputs_match_its_hardcoded_numbers():
    """What make_h2o.py recorded from the unmodified reference on these integrals."""
    n = 0
    for tag, g in FIX.items():
        for k in g:
            if k.startswith("dev_"):
                assert abs(float(g[k])) < TOL, (tag, k)
                n += 1
    assert n == 15                      # + the CCSD(T) total of test_044 below = sixteen known answers
    assert abs(hard("sto3g", "fc", "CCSD", "ecc") + hard("sto3g", "fc", "CCSD", "et") - ECCSD_T_STO3G) < TOL


# ---------------------------------------------------------------------------------------------------------- oracle
def oracle_problem(tag, core):
    g = FIX[tag]
    no, nfzc = sizes(tag, core)
    F, ERI, _ = ao.mo_hamiltonian(g["F_ao"], g["eri_ao"], g["C"])
    blocks = co.blocks_from_full(ERI, no, nfzc)
    return co.Problem(blocks, F, no, nfzc), blocks, F, nfzc


def oracle_ccd(P, F):
    """CCSD equations with t1 pinned to zero (the CCD branches of ccwfn.py drop every t1 term)."""
    t1, t2 = P.guess()
    ecc = P.cc_energy(F, t1, t2)
    diis = co.Diis(t1, t2, 8)
    for it in range(75):
        last = ecc
        _, r2 = P.residuals(F, t1, t2)
        t2 = t2 + r2 / P.Dijab
        rms = np.sqrt(np.sum((r2 / P.Dijab) ** 2))
        ecc = P.cc_energy(F, t1, t2)
        if abs(ecc - last) < 1e-12 and rms < 1e-12:
            return ecc, t1, t2
        diis.add_error_vector(t1, t2)
        _, t2 = diis.extrapolate(t1, t2)
    raise AssertionError("CCD oracle did not converge")


@pytest.mark.parametrize("tag", ["sto3g", "ccpvdz"])
def test_oracle_ccsd_t_lambda(tag):
    P, b, F, nfzc = oracle_problem(tag, "fc")
    ecc, t1, t2, _ = co.solve_cc(P, 1e-12, 1e-12, 75)
    assert abs(ecc - hard(tag, "fc", "CCSD", "ecc")) < TOL
    assert np.abs(t1 - ref(tag, "fc", "CCSD", "t1")).max() < 1e-10
    assert np.abs(t2 - ref(tag, "fc", "CCSD", "t2")).max() < 1e-10
    et = to.t_tjl(t1, t2, F, b["ovvv"], b["ooov"], b["oovv"], nfzc)
    assert abs(et - hard(tag, "fc", "CCSD", "et")) < TOL
    assert abs(to.t_vikings(t1, t2, F, b["ovvv"], b["ooov"], b["oovv"], nfzc) - et) < 1e-12
    lecc, l1, l2, _ = lo.solve_lambda(P, t1, t2, 1e-12, 1e-12, 75)
    assert abs(lecc - hard(tag, "fc", "CCSD", "lecc")) < TOL
    assert np.abs(l2 - ref(tag, "fc", "CCSD", "l2")).max() < 1e-10


def test_oracle_all_electron_ccsd_ccd_cc2():
    P, b, F, nfzc = oracle_problem("ccpvdz", "ae")
    ecc, t1, t2, _ = co.solve_cc(P, 1e-12, 1e-12, 75)
    assert abs(ecc - hard("ccpvdz", "ae", "CCSD", "ecc")) < TOL           # the reference's SP test asks for 1e-7
    assert abs(lo.solve_lambda(P, t1, t2, 1e-12, 1e-12, 75)[0] - hard("ccpvdz", "ae", "CCSD", "lecc")) < TOL
    eccd, t1, t2 = oracle_ccd(P, F)
    assert abs(eccd - hard("ccpvdz", "ae", "CCD", "ecc")) < TOL
    lecc = lo.solve_lambda(P, t1, t2, 1e-12, 1e-12, 75, model="CCD")[0]
    assert abs(lecc - hard("ccpvdz", "ae", "CCD", "lecc")) < TOL
    assert abs(cc2_oracle.solve_cc(P, 1e-12, 1e-12, 75)[0] - hard("ccpvdz", "ae", "CC2", "ecc")) < TOL


def test_oracle_cc2_h2():
    """no = 1"""
    P, b, F, nfzc = oracle_problem("h2_ccpvdz", "ae")
    assert abs(cc2_oracle.solve_cc(P, 1e-12, 1e-12, 75)[0] - hard("h2_ccpvdz", "ae", "CC2", "ecc")) < TOL


def test_oracle_cc3():
    P, b, F, nfzc = oracle_problem("teach_ccpvdz", "ae")
    assert abs(cc3_oracle.solve_cc(P, 1e-12, 1e-12, 75)[0] - hard("teach_ccpvdz", "ae", "CC3", "ecc")) < TOL


@pytest.mark.parametrize("tag", ["t034_sto3g", "t034_ccpvdz"])
def test_oracle_ccsd_t_density_lambda(tag):
    """CCSD -> t3_density ((T) energy + Lambda sources S1/S2) -> Lambda, all-electron; max_diis=0 in STO-3G as in
    test_034_ccsd_t_density.py:32,36."""
    from oracle import t3density_oracle as td
    P, b, F, nfzc = oracle_problem(tag, "ae")
    md = 0 if tag == "t034_sto3g" else 8
    ecc, t1, t2, _ = co.solve_cc(P, 1e-12, 1e-12, 75, max_diis=md)
    et, d = td.t3_density(t1, t2, F, b["ovvv"], b["ooov"], b["oovv"], nfzc)
    assert abs(ecc + et - float(ref(tag, "ae", "CCSD(T)", "ecc"))) < TOL
    lecc = lo.solve_lambda(P, t1, t2, 1e-12, 1e-12, 75, max_diis=md, model="CCSD(T)", s1=d["S1"], s2=d["S2"])[0]
    assert abs(lecc - hard(tag, "ae", "CCSD(T)", "lecc")) < TOL


# --------------------------------------------------------------------------------------------------------- product
def h2o_reference(tag, core, kind="ao"):
    g = FIX[tag]
    no, nfzc = sizes(tag, core)
    if kind == "ao":
        return IntegralReference.from_ao(g["F_ao"], g["eri_ao"], g["C"], no, nfzc)
    F, ERI, _ = ao.mo_hamiltonian(g["F_ao"], g["eri_ao"], g["C"])
    return IntegralReference.from_arrays(F, ERI, no, nfzc)


@pytest.mark.parametrize("tag,kind", [("sto3g", "ao"), ("sto3g", "mo"), ("ccpvdz", "ao")])
def test_ccsd_energy(dev, tag, kind):
    """test_002_ccsd_energy.py:22-39"""
    cc = pycc_b200.ccwfn(h2o_reference(tag, "fc", kind), quiet=True)
    eccsd = cc.solve_cc(1e-12, 1e-12, 75)
    assert abs(float(eccsd) - hard(tag, "fc", "CCSD", "ecc")) < TOL
    assert np.abs(cc.t2.cpu().numpy() - ref(tag, "fc", "CCSD", "t2")).max() < 1e-10
    assert np.abs(cc.t1.cpu().numpy() - ref(tag, "fc", "CCSD", "t1")).max() < 1e-10


@pytest.mark.parametrize("tag", ["sto3g", "ccpvdz"])
def test_ccsd_t_three_formulations(dev, tag):
    """test_005_ccsd_t_energy.py:21-47"""
    cc = pycc_b200.ccwfn(h2o_reference(tag, "fc"), model="ccsd(t)", quiet=True)
    total = cc.solve_cc(1e-12, 1e-12, 75)
    et = hard(tag, "fc", "CCSD", "et")
    assert abs(float(total) - hard(tag, "fc", "CCSD", "ecc") - et) < TOL
    for fn in (cctriples.t_vikings, cctriples.t_vikings_inverted, cctriples.t_tjl):
        assert abs(float(fn(cc)) - et) < TOL, fn.__name__


def test_ccsd_t_gpu_total(dev):
    """test_044_ccsd_t_gpu.py:21-38: device='GPU', the return value must be float()-able."""
    cc = pycc_b200.ccwfn(h2o_reference("sto3g", "fc"), model="CCSD(T)", device="GPU", quiet=True)
    ecc = cc.solve_cc(1e-12, 1e-12, 75)
    assert abs(float(ecc) - ECCSD_T_STO3G) < TOL


@pytest.mark.parametrize("tag", ["sto3g", "ccpvdz"])
def test_ccsd_lambda(dev, tag):
    """test_003_ccsd_lambda.py:36-63"""
    cc = pycc_b200.ccwfn(h2o_reference(tag, "fc"), quiet=True)
    cc.solve_cc(1e-12, 1e-12, 75)
    lecc = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc)).solve_lambda(1e-12, 1e-12)
    assert abs(float(lecc) - hard(tag, "fc", "CCSD", "lecc")) < TOL


@pytest.mark.parametrize("tag", ["t034_sto3g", "t034_ccpvdz"])
def test_ccsd_t_density_lambda(dev, tag):
    """test_034_ccsd_t_density.py:13-68: make_t3_density=True, then Lambda with the (T) sources (all-electron)."""
    md = 0 if tag == "t034_sto3g" else 8
    cc = pycc_b200.ccwfn(h2o_reference(tag, "ae"), model="ccsd(t)", make_t3_density=True, quiet=True)
    ecc = cc.solve_cc(1e-12, 1e-12, 75, max_diis=md)
    assert abs(float(ecc) - float(ref(tag, "ae", "CCSD(T)", "ecc"))) < TOL
    lecc = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc)).solve_lambda(1e-12, 1e-12, 75, max_diis=md)
    assert abs(float(lecc) - hard(tag, "ae", "CCSD(T)", "lecc")) < TOL


def test_all_electron_ccsd_lambda(dev):
    """test_030_sp.py:26-40 in double precision (the hard-coded numbers are good to 1e-14)."""
    cc = pycc_b200.ccwfn(h2o_reference("ccpvdz", "ae"), quiet=True)
    assert abs(float(cc.solve_cc(1e-12, 1e-12, 75)) - hard("ccpvdz", "ae", "CCSD", "ecc")) < TOL
    lecc = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc)).solve_lambda(1e-12, 1e-12)
    assert abs(float(lecc) - hard("ccpvdz", "ae", "CCSD", "lecc")) < TOL


def test_ccd_and_its_lambda(dev):
    """test_017_ccd.py:11-26 (all-electron)"""
    cc = pycc_b200.ccwfn(h2o_reference("ccpvdz", "ae"), model="CCD", quiet=True)
    assert abs(float(cc.solve_cc(1e-12, 1e-12, 75)) - hard("ccpvdz", "ae", "CCD", "ecc")) < TOL
    lecc = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc)).solve_lambda(1e-12, 1e-12)
    assert abs(float(lecc) - hard("ccpvdz", "ae", "CCD", "lecc")) < TOL


def test_cc2(dev):
    """test_020_cc2.py:11-20 (all-electron)"""
    cc = pycc_b200.ccwfn(h2o_reference("ccpvdz", "ae"), model="CC2", quiet=True)
    assert abs(float(cc.solve_cc(1e-12, 1e-12, 75)) - hard("ccpvdz", "ae", "CC2", "ecc")) < TOL


def test_cc2_h2(dev):
    """test_020_cc2.py:32-41: one occupied orbital."""
    cc = pycc_b200.ccwfn(h2o_reference("h2_ccpvdz", "ae"), model="CC2", quiet=True)
    assert abs(float(cc.solve_cc(1e-12, 1e-12, 75)) - hard("h2_ccpvdz", "ae", "CC2", "ecc")) < TOL


def test_ccsd_h2_one_occupied(dev):
    """Edge case no = 1 through CCSD(T) + Lambda (no reference number: oracle on the same integrals)."""
    P, b, F, nfzc = oracle_problem("h2_ccpvdz", "ae")
    ecc, t1, t2, _ = co.solve_cc(P, 1e-12, 1e-12, 75)
    cc = pycc_b200.ccwfn(h2o_reference("h2_ccpvdz", "ae"), model="CCSD(T)", quiet=True)
    assert abs(float(cc.solve_cc(1e-12, 1e-12, 75)) - ecc) < TOL          # E(T) = 0 with one occupied orbital
    cc2 = pycc_b200.ccwfn(h2o_reference("h2_ccpvdz", "ae"), quiet=True)
    cc2.solve_cc(1e-12, 1e-12, 75)
    lecc = pycc_b200.cclambda(cc2, pycc_b200.cchbar(cc2)).solve_lambda(1e-12, 1e-12)
    assert abs(float(lecc) - lo.solve_lambda(P, t1, t2, 1e-12, 1e-12, 75)[0]) < TOL


def test_cc3(dev):
    """test_031_cc3.py:24-32 (H2O_Teach geometry, all-electron)"""
    cc = pycc_b200.ccwfn(h2o_reference("teach_ccpvdz", "ae"), model="CC3", quiet=True)
    assert abs(float(cc.solve_cc(1e-12, 1e-12, 75)) - hard("teach_ccpvdz", "ae", "CC3", "ecc")) < TOL


def test_sp_all_electron(dev):
    """test_030_sp.py:17-31: precision='SP', all-electron cc-pVDZ, 1e-7."""
    cc = pycc_b200.ccwfn(h2o_reference("ccpvdz", "ae"), precision="SP", quiet=True)
    ecc = cc.solve_cc(1e-7, 1e-7)
    assert abs(float(ecc) - hard("ccpvdz", "ae", "CCSD", "ecc")) < 1e-7
    lecc = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc)).solve_lambda(1e-7, 1e-7)
    assert abs(float(lecc) - hard("ccpvdz", "ae", "CCSD", "lecc")) < 1e-7


@pytest.mark.parametrize("tag", ["sto3g", "ccpvdz"])
def test_mixed_precision_within_1e6(dev, tag):
    """BASELINE north_star: mixed precision within 1e-6 Eh of the FP64 reference."""
    cc = pycc_b200.ccwfn(h2o_reference(tag, "fc"), model="CCSD(T)", device="GPU", precision="MP", quiet=True)
    ecc = cc.solve_cc(1e-7, 1e-7, 75)      # the reference's defaults; the split-TF32 products leave rms ~1e-8
    assert abs(float(ecc) - hard(tag, "fc", "CCSD", "ecc") - hard(tag, "fc", "CCSD", "et")) < 1e-6


def test_mixed_precision_lambda_cc2_cc3(dev):
    """precision='MP' releases the FP64 <ab|ef> block; HBAR / CC2 / CC3 rebuild the rows they need from its TF32 planes
    (b200cc_merge_tf32).  Within 1e-6 Eh of the reference's FP64 numbers."""
    cc = pycc_b200.ccwfn(h2o_reference("ccpvdz", "fc"), precision="MP", quiet=True)
    assert abs(float(cc.solve_cc(1e-7, 1e-7, 75)) - hard("ccpvdz", "fc", "CCSD", "ecc")) < 1e-6
    assert cc._vvvv_released()
    hb = pycc_b200.cchbar(cc)
    cc.H.merge_chunk_bytes = 8 * 19 ** 3 * 3                    # three rows a of <ab|ef> per rebuilt chunk: same block
    assert float((pycc_b200.cchbar(cc).Hvvvo - hb.Hvvvo).abs().max()) < 1e-13
    lecc = pycc_b200.cclambda(cc, hb).solve_lambda(1e-7, 1e-7)
    assert abs(float(lecc) - hard("ccpvdz", "fc", "CCSD", "lecc")) < 1e-6
    cc = pycc_b200.ccwfn(h2o_reference("ccpvdz", "ae"), model="CC2", precision="MP", quiet=True)
    assert abs(float(cc.solve_cc(1e-7, 1e-7, 75)) - hard("ccpvdz", "ae", "CC2", "ecc")) < 1e-6
    cc = pycc_b200.ccwfn(h2o_reference("teach_ccpvdz", "ae"), model="CC3", precision="MP", quiet=True)
    assert abs(float(cc.solve_cc(1e-7, 1e-7, 75)) - hard("teach_ccpvdz", "ae", "CC3", "ecc")) < 1e-6


def test_h2_minimal_basis_one_occupied_one_virtual(dev):
    """Smallest closed-shell problem there is: H2 / STO-3G (no = nv = 1), integrals computed on the fly.  CCSD is exact
    here (two electrons); Szabo & Ostlund quote E_SCF = -1.1167 and E_corr = -0.0206 Eh at R = 1.4 bohr.  The product
    must equal the oracle through CCSD(T) (E(T) = 0) and Lambda.  max_diis=0: with a single non-zero amplitude the DIIS
    error vectors are collinear and the B matrix of utils.py:330-348 is singular (in the reference as well)."""
    sto3g_h = {"H": [(0, [3.42525091, 0.62391373, 0.16885540], [0.15432897, 0.53532814, 0.44463454])]}
    atoms = [("H", np.zeros(3)), ("H", np.array([0.0, 0.0, 1.4]))]
    S, Hc, eri, enuc = gto.integrals(atoms, sto3g_h, {"H": 1.0})
    eel, eps, C, F_ao = gto.rhf(S, Hc, eri, ndocc=1)
    assert abs(eel + enuc + 1.1167) < 1e-4
    F, ERI, _ = ao.mo_hamiltonian(F_ao, eri, C)
    P = co.Problem(co.blocks_from_full(ERI, 1, 0), F, 1, 0)
    ecc, t1, t2, _ = co.solve_cc(P, 1e-12, 1e-12, 200, max_diis=0)
    assert abs(ecc + 0.0206) < 1e-4
    cc = pycc_b200.ccwfn(IntegralReference.from_ao(F_ao, eri, C, 1, 0), model="CCSD(T)", quiet=True)
    assert abs(float(cc.solve_cc(1e-12, 1e-12, 200, max_diis=0)) - ecc) < TOL
    cc = pycc_b200.ccwfn(IntegralReference.from_ao(F_ao, eri, C, 1, 0), quiet=True)
    cc.solve_cc(1e-12, 1e-12, 200, max_diis=0)
    lecc = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc)).solve_lambda(1e-12, 1e-12, 200, max_diis=0)
    assert abs(float(lecc) - lo.solve_lambda(P, t1, t2, 1e-12, 1e-12, 200, max_diis=0)[0]) < TOL


def test_split_plane_cache_budget(dev):
    """kernels.MIXED.cache_bytes bounds the cached TF32 planes of constant operands; past it a constant is split per
    use.  Same numbers either way."""
    from pycc_b200 import kernels as K
    keep = (K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles, K.MIXED.cache_bytes)
    K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles = 0.0, 1, 1       # tiny problem: force the mixed kernel
    try:
        res = []
        for budget in (keep[3], 0):
            K.MIXED.cache_bytes = budget
            n0 = K.MIXED.stats["split_uncached_over_budget"]
            cc = pycc_b200.ccwfn(h2o_reference("ccpvdz", "fc"), precision="MP", quiet=True)
            e = float(cc.solve_cc(1e-7, 1e-7, 75))
            n_ccsd = K.MIXED.stats["split_uncached_over_budget"] - n0
            held = len(cc.H._split_cache)
            hb = pycc_b200.cchbar(cc)                 # HBAR / Lambda never add to the cache and drop what CCSD left
            lecc = float(pycc_b200.cclambda(cc, hb).solve_lambda(1e-7, 1e-7))
            assert (held > 0) == (budget > 0) and len(cc.H._split_cache) == 0
            res.append((e, lecc, n_ccsd))
    finally:
        K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles, K.MIXED.cache_bytes = keep
    assert res[0][2] == 0 and res[1][2] > 0
    assert abs(res[0][0] - res[1][0]) < 1e-10 and abs(res[0][1] - res[1][1]) < 1e-10
    assert abs(res[0][0] - hard("ccpvdz", "fc", "CCSD", "ecc")) < 1e-6
