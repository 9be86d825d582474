#!/usr/bin/env python
"""Golden vectors for the HBAR build and the Lambda solver, from the UNMODIFIED reference
(pycc/cchbar.py, pycc/cclambda.py) run in the build container with the shims of make_golden.py.

    python tests/golden/make_golden_lambda.py        # writes tests/golden/lam_<tag>.npz

Inputs are those of the CCSD goldens (ref_<tag>.npz: factor B, F, scale, converged t1/t2), so nothing new is
seeded except the random lambda point (rng 3000 + seed).  Every stored array is an output of the reference's own code.
"""
import contextlib
import importlib
import io
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

HBAR = ("Hov", "Hvv", "Hoo", "Hoooo", "Hvvvv", "Hvovv", "Hooov", "Hovvo", "Hovov", "Hvvvo", "Hovoo")


def case(mods, tag, model):
    ccwfn_mod, cctriples, utils, device_mod = mods
    cchbar = importlib.import_module("pycc.cchbar").cchbar
    cclambda = importlib.import_module("pycc.cclambda").cclambda
    from pycc_b200.synthetic import Synthetic, full_eri
    g = dict(np.load(os.path.join(HERE, "ref_%s.npz" % tag)))
    syn = Synthetic(int(g["no"]), int(g["nv"]), g["B"], g["F"], float(g["scale"]), int(g["seed"]))
    ERI = full_eri(syn)
    w = mg.reference_wfn(ccwfn_mod, device_mod, syn, ERI, model=model)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        if model == "CCSD":
            w.t1, w.t2 = g["conv_t1"].copy(), g["conv_t2"].copy()
            ecc = float(g["e_ccsd"])
        else:
            # CCSD(T): the (T) step of solve_cc is t3_density, which leaves the Lambda sources S1 / S2 on the wavefunction
            w.make_t3_density = model == "CCSD(T)"
            ecc = float(w.solve_cc(1e-12, 1e-12, 100))
        hbar = cchbar(w)
    out = dict(model=np.array(model), t1=w.t1.copy(), t2=w.t2.copy(), ecc=ecc)
    if model == "CCSD(T)":
        out["S1"], out["S2"] = np.array(w.S1), np.array(w.S2)
    for k in HBAR:
        out[k] = np.array(getattr(hbar, k))
    lam = cclambda(w, hbar)
    out["guess_l1"], out["guess_l2"] = np.array(lam.l1), np.array(lam.l2)
    # residual pieces at a generic (random, unsymmetric) lambda point
    rng = np.random.default_rng(3000 + int(g["seed"]))
    l1 = 0.05 * rng.standard_normal((syn.no, syn.nv))
    l2 = 0.05 * rng.standard_normal((syn.no, syn.no, syn.nv, syn.nv))
    out["rand_l1"], out["rand_l2"] = l1, l2
    out["rand_Goo"], out["rand_Gvv"] = lam.build_Goo(w.t2, l2), lam.build_Gvv(w.t2, l2)
    with contextlib.redirect_stdout(io.StringIO()):
        r1, r2 = lam.residuals(w.H.F, w.t1, w.t2, l1, l2)
    out["rand_r1"], out["rand_r2"] = np.array(r1), np.array(r2)
    out["rand_pseudo"] = float(lam.pseudoenergy(w.o, w.v, ERI, l2))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        lecc = lam.solve_lambda(1e-12, 1e-12, 100)
    trace = []
    for line in buf.getvalue().splitlines():
        if line.startswith("Iter") and "rms" in line:
            m = re.search(r"PseudoE =\s*(\S+)\s+dE =\s*(\S+)\s+rms =\s*(\S+)", line)
            trace.append((float(m.group(1)), float(m.group(3))))
    out["trace_lecc_rms"] = np.array(trace)
    out["lecc"] = float(lecc)
    out["conv_l1"], out["conv_l2"] = np.array(lam.l1), np.array(lam.l2)
    path = os.path.join(HERE, "lam_%s_%s.npz" % (tag, {"CCSD(T)": "ccsdpt"}.get(model, model.lower())))
    np.savez_compressed(path, **out)
    print("wrote %s  pseudo-E = %.15f  iters = %d" % (path, out["lecc"], len(trace)))


def main():
    mods = mg.load_reference()
    case(mods, "o4v10_s0", "CCSD")
    case(mods, "o4v10_s1_noise", "CCSD")
    case(mods, "o3v7_s2", "CCSD")
    case(mods, "o4v10_s0", "CCD")
    case(mods, "o4v10_s1_noise", "CCSD(T)")


if __name__ == "__main__":
    main()
