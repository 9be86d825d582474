#!/usr/bin/env python
"""Known-answer fixtures on the molecule of the reference's hot-path tests: H2O in STO-3G and cc-pVDZ.

    python tests/golden/make_h2o.py        # rewrites tests/golden/h2o_*.npz (all, or the tags given)

The reference's tests pin the path on this molecule with hard-coded numbers (geometry ``moldict["H2O"]`` =
pycc/data/molecules.py:42-46 unless noted, SCF converged to 1e-12 -- pycc/tests/conftest.py:28-36); "fc" = frozen core,
"ae" = all-electron:

    tag            what                          STO-3G                cc-pVDZ               reference test
    sto3g/ccpvdz   fc E_corr(CCSD)               -0.070616830152761    -0.222029814166783    test_002:31,38
                   fc E(T)                       -0.000099957499645    -0.003861236558801    test_005:33,44
                   fc Lambda pseudo-energy       -0.068826452648939    -0.217838951550509    test_003:49,61
                   fc E_corr(CCSD(T))            -0.0707167876524093                         test_044:37 (device='GPU')
    ccpvdz         ae E_corr(CCSD), its Lambda                         -0.223910018703551    test_030:30,39 ('SP', 1e-7)
                                                                       -0.219688229733875
                   ae E_corr(CCD), its Lambda                          -0.222559319034       test_017:19,25
                                                                       -0.218758826700
                   ae E_corr(CC2)                                      -0.215857544656       test_020:19
    teach_ccpvdz   ae E_corr(CC3), moldict["H2O_Teach"]                -0.227888246840310    test_031:31
    t034_*         ae Lambda of CCSD(T) with the     -0.069084521221746    -0.227199866607450    test_034:44,68
                   t3_density sources (geometry of test_034:19-25; STO-3G with max_diis=0)
    h2_ccpvdz      ae E_corr(CC2), moldict["H2"] (no = 1)              -0.026445902512140185 test_020:40

cc-pVDZ is BASELINE.json configs[0].  psi4 (which supplies the integrals to the reference, hamiltonian.py:58-68) is
not installable offline, so the integrals are computed by tests/golden/gto.py (McMurchie-Davidson in numpy), RHF is
run here, and

  * the AO quantities (S, Hcore, packed (pq|rs), C, eps, F_ao, E_nuc, E_SCF) are stored as the fixture,
  * the UNMODIFIED reference (loader of make_golden.py) is run on the MO integrals; its energies and amplitudes are
    stored next to the hard-coded numbers above.

The integrals are therefore NOT psi4's; what anchors them is that the reference's own code, fed with them,
reproduces the reference's published numbers (asserted below; the achieved differences are stored as ``dev_*``).

Basis sets as distributed with psi4 (share/basis/sto-3g.gbs, cc-pvdz.gbs = the EMSL tables; cc-pVDZ with pure d
functions).  Geometry: Z-matrix O / H 1 1.1 / H 1 1.1 2 104 in Angstrom, 1 bohr = 0.52917721067 Angstrom.
"""
import contextlib
import io
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import gto

STO3G = {
    "H": [(0, [3.42525091, 0.62391373, 0.16885540], [0.15432897, 0.53532814, 0.44463454])],
    "O": [(0, [130.7093200, 23.8088610, 6.4436083], [0.15432897, 0.53532814, 0.44463454]),
          (0, [5.0331513, 1.1695961, 0.3803890], [-0.09996723, 0.39951283, 0.70011547]),
          (1, [5.0331513, 1.1695961, 0.3803890], [0.15591627, 0.60768372, 0.39195739])],
}
_OS = [11720.0, 1759.0, 400.8, 113.7, 37.03, 13.27, 5.025, 1.013]
CCPVDZ = {
    "H": [(0, [13.01, 1.962, 0.4446], [0.019685, 0.137977, 0.478148]),
          (0, [0.122], [1.0]),
          (1, [0.727], [1.0])],
    "O": [(0, _OS, [0.000710, 0.005470, 0.027837, 0.104800, 0.283062, 0.448719, 0.270952, 0.015458]),
          (0, _OS, [-0.000160, -0.001263, -0.006267, -0.025716, -0.070924, -0.165411, -0.116955, 0.557368]),
          (0, [0.3023], [1.0]),
          (1, [17.70, 3.854, 1.046], [0.043018, 0.228913, 0.508728]),
          (1, [0.2753], [1.0]),
          (2, [1.185], [1.0])],
}
CHARGE = {"H": 1.0, "O": 8.0}

# (geometry, frozen core, model) -> hard-coded numbers of the reference's tests
HARDCODED = {
    "sto3g": {("fc", "CCSD"): dict(ecc=-0.070616830152761, et=-0.000099957499645, lecc=-0.068826452648939)},
    "ccpvdz": {("fc", "CCSD"): dict(ecc=-0.222029814166783, et=-0.003861236558801, lecc=-0.217838951550509),
               ("ae", "CCSD"): dict(ecc=-0.223910018703551, lecc=-0.219688229733875),   # test_030_sp.py:30,39 ('SP', 1e-7)
               ("ae", "CCD"): dict(ecc=-0.222559319034, lecc=-0.218758826700),      # test_017_ccd.py:19,25
               ("ae", "CC2"): dict(ecc=-0.215857544656)},                            # test_020_cc2.py:19
    "teach_ccpvdz": {("ae", "CC3"): dict(ecc=-0.227888246840310)},                   # test_031_cc3.py:31
    # test_034_ccsd_t_density.py:19-68: CCSD(T) with make_t3_density=True, then Lambda with the (T) sources
    "t034_sto3g": {("ae", "CCSD(T)"): dict(lecc=-0.069084521221746)},                # max_diis=0 (test_034:32,36)
    "t034_ccpvdz": {("ae", "CCSD(T)"): dict(lecc=-0.227199866607450)},
    "h2_ccpvdz": {("ae", "CC2"): dict(ecc=-0.026445902512140185)},                   # test_020_cc2.py:40 (no = 1)
}
MAX_DIIS = {"t034_sto3g": 0}
ECCSD_T_STO3G = -0.0707167876524093      # test_044_ccsd_t_gpu.py:37
TOL = {("ae", "CCSD"): 1e-7}             # everything else 1e-10 here (1e-11 in the reference's tests)


def geometry(tag):
    if tag.startswith("teach"):          # moldict["H2O_Teach"], pycc/data/molecules.py:35-40, bohr
        return [("O", np.array([0.0, -0.143225816552, 0.0])),
                ("H", np.array([1.638036840407, 1.136548822547, 0.0])),
                ("H", np.array([-1.638036840407, 1.136548822547, 0.0]))]
    if tag.startswith("h2_"):            # moldict["H2"], pycc/data/molecules.py:2-6, bohr
        return [("H", np.zeros(3)), ("H", np.array([0.0, 0.0, 1.4]))]
    if tag.startswith("t034"):           # pycc/tests/test_034_ccsd_t_density.py:19-25, bohr
        return [("O", np.array([0.0, 0.0, 0.143225857166674])),
                ("H", np.array([0.0, -1.638037301628121, -1.136549142277225])),
                ("H", np.array([0.0, 1.638037301628121, -1.136549142277225]))]
    r = 1.1 / gto.BOHR                   # moldict["H2O"], pycc/data/molecules.py:42-46, Angstrom
    th = math.radians(104.0)
    return [("O", np.zeros(3)), ("H", np.array([0.0, 0.0, r])),
            ("H", np.array([r * math.sin(th), 0.0, r * math.cos(th)]))]


def run_reference(mods, F, ERI, no, nfzc, model, want, max_diis=8):
    import importlib
    from make_golden import reference_wfn
    ccwfn_mod, cctriples, utils, device_mod = mods
    n = F.shape[0]
    syn = types.SimpleNamespace(no=no, nv=n - nfzc - no, n=n, F=F, eps=np.diag(F).copy(),
                                o=slice(nfzc, nfzc + no), v=slice(nfzc + no, n))
    w = reference_wfn(ccwfn_mod, device_mod, syn, ERI, model=model)
    w.nfzc = nfzc
    w.make_t3_density = model == "CCSD(T)"        # solve_cc then runs t3_density, which leaves S1/S2 for Lambda
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        out["ecc"] = float(w.solve_cc(1e-12, 1e-12, 75, max_diis=max_diis))
        out["t1"], out["t2"] = np.array(w.t1), np.array(w.t2)
        if "et" in want:
            out["et"] = float(cctriples.t_tjl(w))
            assert abs(float(cctriples.t_vikings(w)) - out["et"]) < 1e-12
            assert abs(float(cctriples.t_vikings_inverted(w)) - out["et"]) < 1e-12
        if "lecc" in want:
            hbar = importlib.import_module("pycc.cchbar").cchbar(w)
            lam = importlib.import_module("pycc.cclambda").cclambda(w, hbar)
            out["lecc"] = float(lam.solve_lambda(1e-12, 1e-12, 75, max_diis=max_diis))
            out["l1"], out["l2"] = np.array(lam.l1), np.array(lam.l2)
    return out


def make(tag, shells, mods):
    atoms = geometry(tag)
    ndocc = int(sum(CHARGE[sym] for sym, _ in atoms)) // 2
    S, H, eri, enuc = gto.integrals(atoms, shells, CHARGE)
    escf_el, eps, C, F_ao = gto.rhf(S, H, eri, ndocc=ndocc)
    n = S.shape[0]
    print("%s: n = %d  E_nuc = %.12f  E_SCF = %.12f" % (tag, n, enuc, escf_el + enuc))
    mo = np.einsum("pqrs,pi,qj,rk,sl->ijkl", eri, C, C, C, C, optimize=True)
    ERI = np.ascontiguousarray(mo.swapaxes(1, 2))          # Dirac <pq|rs>, hamiltonian.py:67
    F = C.T @ F_ao @ C
    rec = dict(S=S, Hcore=H, eri_packed=gto.pack_eri(eri), C=C, eps=eps, F_ao=F_ao, enuc=enuc, escf=escf_el + enuc,
               ndocc=ndocc)
    for (core, model), hard in HARDCODED[tag].items():
        nfzc = 1 if core == "fc" else 0
        out = run_reference(mods, F, ERI, ndocc - nfzc, nfzc, model, hard, MAX_DIIS.get(tag, 8))
        key = "%s_%s" % (core, model.lower().replace("(t)", "pt"))
        for k, val in out.items():
            rec["ref_%s_%s" % (key, k)] = val
        for k, val in hard.items():
            rec["hardcoded_%s_%s" % (key, k)] = val
            rec["dev_%s_%s" % (key, k)] = out[k] - val
            print("  reference code, %s %-5s %-4s = %.15f   hard-coded %.15f   diff %.1e" % (core, model, k, out[k], val, out[k] - val))
            assert abs(out[k] - val) < TOL.get((core, model), 1e-10)
    np.savez_compressed(os.path.join(HERE, "h2o_%s.npz" % tag), **rec)
    print("  wrote h2o_%s.npz" % tag)


def main():
    from make_golden import load_reference
    mods = load_reference()
    for tag in (sys.argv[1:] or sorted(HARDCODED)):
        make(tag, STO3G if tag.endswith("sto3g") else CCPVDZ, mods)


if __name__ == "__main__":
    main()
