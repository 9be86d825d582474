#!/usr/bin/env python
"""Golden vectors for the real-time CC right-hand side: the UNMODIFIED reference's
``CCwfn.residuals(F, t1, t2, real_time=True)`` (ccwfn.py:321-372) evaluated with COMPLEX amplitudes and a
field-dressed Fock matrix, as ``rtcc.f`` calls it (rt/rtcc.py:136-141), with the shims of make_golden.py.

    python tests/golden/make_golden_complex.py        # writes tests/golden/cplx_<tag>.npz

Inputs are those of the CCSD goldens (ref_<tag>.npz) plus seeded complex amplitudes (rng 5000 + seed); two Fock
matrices per case: F + mu V(t) with a real symmetric dipole-like perturbation (electric field), and with an additional
imaginary antisymmetric part (magnetic field: Hermitian complex F).  Every stored array is an output of the
reference's own code.
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
    w = mg.reference_wfn(ccwfn_mod, device_mod, syn, ERI, model="CCSD")
    rng = np.random.default_rng(5000 + int(g["seed"]))
    no, nv, n = syn.no, syn.nv, syn.n
    t1 = g["conv_t1"] + 0.03 * (rng.standard_normal((no, nv)) + 1j * rng.standard_normal((no, nv)))
    t2 = g["conv_t2"] + 0.03 * (rng.standard_normal((no, no, nv, nv)) + 1j * rng.standard_normal((no, no, nv, nv)))
    mu = rng.standard_normal((n, n))
    mu = 0.5 * (mu + mu.T)
    m = rng.standard_normal((n, n))
    m = 0.5 * (m - m.T)
    F_el = syn.F + 0.05 * mu                              # real symmetric
    F_mag = syn.F + 0.05 * mu + 0.03j * m                 # complex Hermitian
    out = dict(t1=t1, t2=t2, F_el=F_el, F_mag=F_mag)
    for name, F in (("el", F_el), ("mag", F_mag)):
        r1, r2 = w.residuals(F, t1, t2, real_time=True)
        out["r1_" + name], out["r2_" + name] = np.array(r1), np.array(r2)
    # the Lambda half of rtcc.f (rt/rtcc.py:143-147): cclambda.residuals with complex t AND complex lambda amplitudes
    import contextlib
    import importlib
    import io
    cchbar = importlib.import_module("pycc.cchbar").cchbar
    cclambda = importlib.import_module("pycc.cclambda").cclambda
    w.t1, w.t2 = g["conv_t1"].copy(), g["conv_t2"].copy()
    with contextlib.redirect_stdout(io.StringIO()):
        lam = cclambda(w, cchbar(w))
    l1 = 2.0 * g["conv_t1"] + 0.03 * (rng.standard_normal((no, nv)) + 1j * rng.standard_normal((no, nv)))
    l2 = 0.03 * (rng.standard_normal((no, no, nv, nv)) + 1j * rng.standard_normal((no, no, nv, nv)))
    l2 = l2 + 2.0 * (2.0 * g["conv_t2"] - g["conv_t2"].swapaxes(2, 3))
    out["l1"], out["l2"] = l1, l2
    for name, F in (("el", F_el), ("mag", F_mag)):
        with contextlib.redirect_stdout(io.StringIO()):
            q1, q2 = lam.residuals(F, t1, t2, l1, l2)
        out["rl1_" + name], out["rl2_" + name] = np.array(q1), np.array(q2)
    # real amplitudes in a complex container must reproduce the real path
    r1, r2 = w.residuals(syn.F, g["conv_t1"].astype(complex), g["conv_t2"].astype(complex), real_time=True)
    out["r1_realamps"], out["r2_realamps"] = np.array(r1), np.array(r2)
    path = os.path.join(HERE, "cplx_%s.npz" % tag)
    np.savez_compressed(path, **out)
    print("wrote %s  |r2_el| = %.6f  |r2_mag| = %.6f" % (path, np.abs(out["r2_el"]).max(), np.abs(out["r2_mag"]).max()))


def main():
    mods = mg.load_reference()
    for tag in ("o4v10_s0", "o4v10_s1_noise", "o3v7_s2"):
        case(mods, tag)


if __name__ == "__main__":
    main()
