#!/usr/bin/env python
"""Golden vectors for the real-time CC3 right-hand side: the UNMODIFIED reference's
``CCwfn(model='CC3').residuals(F, t1, t2, real_time=True)`` (ccwfn.py:321-430) -- the connected triples corrected by the
explicit-field term ``t3 -= t3_pert_ijk(V = F - H.F)`` (ccwfn.py:421-423, cctriples.py:679-705) -- with the shims of
make_golden.py.

    python tests/golden/make_golden_rtcc3.py        # writes tests/golden/rtcc3_<tag>.npz

Inputs are those of the CCSD goldens (ref_<tag>.npz): the generic real point (rand_t1, rand_t2) and seeded complex
amplitudes (rng 7000 + seed); two Fock matrices per case as in make_golden_complex.py (real symmetric field; complex
Hermitian field).  Every stored array is an output of the reference's own code.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def case(mods, tag):
    ccwfn_mod, cctriples, utils, device_mod = mods
    from pycc_b200.synthetic import Synthetic, full_eri
    g = dict(np.load(os.path.join(HERE, "ref_%s.npz" % tag)))
    syn = Synthetic(int(g["no"]), int(g["nv"]), g["B"], g["F"], float(g["scale"]), int(g["seed"]))
    ERI = full_eri(syn)
    w = mg.reference_wfn(ccwfn_mod, device_mod, syn, ERI, model="CC3")
    w.real_time = True
    o, v, L = w.o, w.v, w.H.L
    rng = np.random.default_rng(7000 + int(g["seed"]))
    no, nv, n = syn.no, syn.nv, syn.n
    mu = rng.standard_normal((n, n))
    mu = 0.5 * (mu + mu.T)
    m = rng.standard_normal((n, n))
    m = 0.5 * (m - m.T)
    F_el = syn.F + 0.05 * mu                              # real symmetric
    F_mag = syn.F + 0.05 * mu + 0.03j * m                 # complex Hermitian (real diagonal)
    t1, t2 = g["rand_t1"], g["rand_t2"]
    c1 = g["conv_t1"] + 0.03 * (rng.standard_normal((no, nv)) + 1j * rng.standard_normal((no, nv)))
    c2 = g["conv_t2"] + 0.03 * (rng.standard_normal((no, no, nv, nv)) + 1j * rng.standard_normal((no, no, nv, nv)))
    out = dict(t1=t1, t2=t2, c1=c1, c2=c2, F_el=F_el, F_mag=F_mag)
    # real amplitudes, real field: the explicit-field triples alone and the whole residual
    Fme = w.build_Fme(o, v, F_el, L, t1)
    X1, X2 = w._cc3_t_residual(o, v, F_el, ERI, L, t1, t2, Fme, real_time=True)
    out["X1_el"], out["X2_el"] = np.array(X1), np.array(X2)
    r1, r2 = w.residuals(F_el, t1, t2, real_time=True)
    out["r1_el"], out["r2_el"] = np.array(r1), np.array(r2)
    # complex amplitudes (rtcc.f, rt/rtcc.py:136-141), both fields
    for name, F in (("el", F_el), ("mag", F_mag)):
        r1, r2 = w.residuals(F, c1, c2, real_time=True)
        out["c_r1_" + name], out["c_r2_" + name] = np.array(r1), np.array(r2)
    path = os.path.join(HERE, "rtcc3_%s.npz" % tag)
    np.savez_compressed(path, **out)
    print("wrote %s  |X2_el| = %.6f  |c_r2_el| = %.6f  |c_r2_mag| = %.6f"
          % (path, np.abs(out["X2_el"]).max(), np.abs(out["c_r2_el"]).max(), np.abs(out["c_r2_mag"]).max()))


def main():
    mods = mg.load_reference()
    for tag in ("o4v10_s0", "o4v10_s1_noise", "o3v7_s2"):
        case(mods, tag)


if __name__ == "__main__":
    main()
