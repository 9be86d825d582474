#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (CrawfordGroup/pycc).

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference cannot be imported as-is here (psi4 / opt_einsum / qcelemental are
not installed), so it is loaded with the three shims of SURVEY.md Appendix C:
a stub ``psi4`` module (never called on this path), an ``opt_einsum`` whose
``contract`` is an exact ``numpy.einsum``, and a pre-registered empty ``pycc``
package so ``pycc/__init__.py`` (which pulls qcelemental) is skipped.  The
reference source files themselves are executed unmodified from /root/reference.

Every array written here is an OUTPUT OF THE REFERENCE'S OWN CODE on the seeded
synthetic inputs of ``pycc_b200.synthetic`` (inputs are stored too, as the
factor B + F + scale, so the fixtures do not depend on numpy's RNG stream).
"""
import importlib
import io
import re
import os
import sys
import types
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("PYCC_REFERENCE", "/root/reference")


def load_reference():
    for name in ("psi4", "psi4.core"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["psi4"].core = sys.modules["psi4.core"]
    oe = types.ModuleType("opt_einsum")
    oe.contract = lambda sub, *ops, **kw: np.einsum(sub, *ops, optimize=True)
    sys.modules["opt_einsum"] = oe
    pkg = types.ModuleType("pycc")
    pkg.__path__ = [os.path.join(REF, "pycc")]
    sys.modules["pycc"] = pkg
    ccwfn = importlib.import_module("pycc.ccwfn")
    cctriples = importlib.import_module("pycc.cctriples")
    utils = importlib.import_module("pycc.utils")
    device = importlib.import_module("pycc.device")
    return ccwfn, cctriples, utils, device


def reference_wfn(ccwfn_mod, device_mod, syn, ERI, model="CCSD(T)"):
    """A reference CCwfn filled exactly as ccwfn.py:195-211 / wavefunction.py:153-169 do."""
    CCwfn = ccwfn_mod.CCwfn
    w = CCwfn.__new__(CCwfn)
    w.model = w.method = model
    w.make_t3_density = False
    w.store_triples = False
    w.local = None
    w.orbital_basis = "spatial"
    w.eref = 0.0
    w.no, w.nv, w.nmo, w.nfzc = syn.no, syn.nv, syn.n, 0
    w.o, w.v = syn.o, syn.v
    mgr = device_mod.DeviceManager(device="CPU", precision="DP")
    w.device_manager = mgr
    w.device, w.device0, w.device1 = mgr.device, mgr.device0, mgr.device1
    w.precision, w.contract = mgr.precision, mgr.contract
    H = types.SimpleNamespace()
    H.F = syn.F.copy()
    H.eps = syn.eps.copy()
    H.ERI = ERI
    H.L = 2.0 * ERI - ERI.swapaxes(2, 3)
    w.H = H
    eo, ev = H.eps[w.o], H.eps[w.v]
    w.Dijab = eo.reshape(-1, 1, 1, 1) + eo.reshape(-1, 1, 1) - ev.reshape(-1, 1) - ev
    w.Dia = eo.reshape(-1, 1) - ev
    w.t1 = np.zeros((w.no, w.nv))
    w.t2 = ERI[w.o, w.o, w.v, w.v].copy() / w.Dijab
    return w


def case(ccwfn_mod, cctriples, utils, device_mod, no, nv, seed, fock_noise, tag,
         triples=((2, 1, 0), (1, 1, 0), (1, 0, 0), (0, 0, 0)), full_t=True):
    from pycc_b200.synthetic import make_synthetic, full_eri
    syn = make_synthetic(no, nv, seed=seed, fock_noise=fock_noise)
    ERI = full_eri(syn)
    out = dict(no=no, nv=nv, seed=seed, fock_noise=fock_noise,
               B=syn.B, F=syn.F, scale=syn.scale)
    w = reference_wfn(ccwfn_mod, device_mod, syn, ERI)
    o, v = w.o, w.v
    F, L = w.H.F, w.H.L

    # --- (1) every intermediate and both residuals at a generic (random, unsymmetric) point
    rng = np.random.default_rng(1000 + seed)
    t1 = 0.05 * rng.standard_normal((no, nv))
    t2 = 0.05 * rng.standard_normal((no, no, nv, nv))
    out["rand_t1"], out["rand_t2"] = t1, t2
    out["rand_Fae"] = w.build_Fae(o, v, F, L, t1, t2)
    out["rand_Fmi"] = w.build_Fmi(o, v, F, L, t1, t2)
    out["rand_Fme"] = w.build_Fme(o, v, F, L, t1)
    out["rand_Wmnij"] = w.build_Wmnij(o, v, ERI, t1, t2)
    out["rand_Wmbej"] = w.build_Wmbej(o, v, ERI, L, t1, t2)
    out["rand_Wmbje"] = w.build_Wmbje(o, v, ERI, t1, t2)
    out["rand_Zmbij"] = w.build_Zmbij(o, v, ERI, t1, t2)
    for (f1, f2) in ((1.0, 1.0), (1.0, 0.5), (0.5, 1.0)):
        out["rand_tau_%g_%g" % (f1, f2)] = w.build_tau(t1, t2, f1, f2)
    r1, r2 = w.residuals(F, t1, t2)
    out["rand_r1"], out["rand_r2"] = r1, r2
    out["rand_ecc"] = w.cc_energy(o, v, F, L, t1, t2)

    # --- (2) the iteration trace of solve_cc(1e-12,1e-12), captured from its own prints
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ecc = w.solve_cc(1e-12, 1e-12, 100)
    trace = []
    for line in buf.getvalue().splitlines():
        if line.startswith("Iter") and "rms" in line:
            m = re.search(r"Ecorr =\s*(\S+)\s+dE =\s*(\S+)\s+rms =\s*(\S+)", line)
            trace.append((float(m.group(1)), float(m.group(3))))
    out["trace_ecc_rms"] = np.array(trace)
    out["solve_text"] = np.array(buf.getvalue())
    out["e_total_ccsd_t"] = float(ecc)
    out["conv_t1"], out["conv_t2"] = w.t1.copy(), w.t2.copy()
    out["e_ccsd"] = float(w.cc_energy(o, v, F, L, w.t1, w.t2))
    out["e_t_from_solve"] = float(ecc) - out["e_ccsd"]
    r1, r2 = w.residuals(F, w.t1, w.t2)
    out["conv_r1"], out["conv_r2"] = r1, r2

    # --- (3) (T): the three drivers and per-triple W3 / V3 (no denominators) at converged t
    if full_t:
        out["e_t_tjl"] = float(cctriples.t_tjl(w))
        out["e_t_vikings"] = float(cctriples.t_vikings(w))
        out["e_t_vikings_inverted"] = float(cctriples.t_vikings_inverted(w))
    trip = [t for t in triples if max(t) < no]
    out["triples"] = np.array(trip)
    W3s, V3s, T3c, T3d = [], [], [], []
    for (i, j, k) in trip:
        W3 = cctriples.t3c_ijk(o, v, i, j, k, w.t2, ERI[v, v, v, o], ERI[o, v, o, o], F, w.contract, False)
        V3 = cctriples.t3d_ijk(o, v, i, j, k, w.t1, w.t2, ERI[o, o, v, v], F, w.contract, False) + W3
        W3s.append(W3.copy()); V3s.append(V3.copy())
        T3c.append(cctriples.t3c_ijk(o, v, i, j, k, w.t2, ERI[v, v, v, o], ERI[o, v, o, o], F, w.contract, True))
        T3d.append(cctriples.t3d_ijk(o, v, i, j, k, w.t1, w.t2, ERI[o, o, v, v], F, w.contract, True))
    out["W3"], out["V3"] = np.array(W3s), np.array(V3s)
    out["t3c_denom"], out["t3d_denom"] = np.array(T3c), np.array(T3d)
    a, b, c = 2, 1, 0
    out["t3c_abc_210"] = cctriples.t3c_abc(o, v, a, b, c, w.t2, ERI[v, v, v, o], ERI[o, v, o, o], F, w.contract, True)
    out["t3d_abc_210"] = cctriples.t3d_abc(o, v, a, b, c, w.t1, w.t2, ERI[o, o, v, v], F, w.contract, True)

    # --- (4) DIIS: feed the reference helper a deterministic sequence, record every extrapolant
    rng = np.random.default_rng(2000 + seed)
    d1 = [0.1 * rng.standard_normal((no, nv)) * 0.5 ** n for n in range(7)]
    d2 = [0.1 * rng.standard_normal((no, no, nv, nv)) * 0.5 ** n for n in range(7)]
    out["diis_in_t1"], out["diis_in_t2"] = np.array(d1), np.array(d2)
    diis = utils.helper_diis(d1[0], d2[0], 4)
    e1, e2 = [], []
    x1, x2 = d1[0], d2[0]
    for n in range(1, 7):
        # the driver pattern of solve_cc: new iterate = last extrapolant + increment
        x1 = x1 + d1[n]
        x2 = x2 + d2[n]
        diis.add_error_vector(x1, x2)
        x1, x2 = diis.extrapolate(x1, x2)
        e1.append(x1.copy()); e2.append(x2.copy())
    out["diis_out_t1"], out["diis_out_t2"] = np.array(e1), np.array(e2)

    path = os.path.join(HERE, "ref_%s.npz" % tag)
    np.savez_compressed(path, **out)
    print("wrote %s  E(CCSD)=%.15f  E(T)=%.15f  iters=%d" % (
        path, out["e_ccsd"], out.get("e_t_tjl", out["e_t_from_solve"]), len(trace)))


def main():
    mods = load_reference()
    # canonical F, even dims
    case(*mods, no=4, nv=10, seed=0, fock_noise=0.0, tag="o4v10_s0")
    # non-canonical F (off-diagonal noise): exercises Fme / f_kc terms
    case(*mods, no=4, nv=10, seed=1, fock_noise=0.01, tag="o4v10_s1_noise")
    # odd dims (ragged tiles, unaligned leading dimensions)
    case(*mods, no=3, nv=7, seed=2, fock_noise=0.0, tag="o3v7_s2")
    # H2O/cc-pVDZ frozen-core shape (BASELINE configs[0]: o=4, v=19)
    case(*mods, no=4, nv=19, seed=0, fock_noise=0.0, tag="o4v19_s0", full_t=True)


if __name__ == "__main__":
    main()
