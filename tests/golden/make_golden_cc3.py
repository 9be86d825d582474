#!/usr/bin/env python
"""Golden vectors for the CC3 ground-state T equations: the UNMODIFIED reference's
``CCwfn(model='CC3')`` -- the T1-dressed intermediates ``build_cc3_W*`` (ccwfn.py:947-1120), the connected-triples
contribution ``_cc3_t_residual`` (374-430) and ``solve_cc`` -- with the shims of make_golden.py.

    python tests/golden/make_golden_cc3.py        # writes tests/golden/cc3_<tag>.npz

Inputs are those of the CCSD goldens (ref_<tag>.npz).  Every stored array is an output of the reference's own code.
"""
import contextlib
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


def case(mods, tag):
    ccwfn_mod, cctriples, utils, device_mod = mods
    from pycc_b200.synthetic import Synthetic, full_eri
    g = dict(np.load(os.path.join(HERE, "ref_%s.npz" % tag)))
    syn = Synthetic(int(g["no"]), int(g["nv"]), g["B"], g["F"], float(g["scale"]), int(g["seed"]))
    ERI = full_eri(syn)
    w = mg.reference_wfn(ccwfn_mod, device_mod, syn, ERI, model="CC3")
    w.real_time = False
    o, v, F, L = w.o, w.v, w.H.F, w.H.L
    t1, t2 = g["rand_t1"], g["rand_t2"]                 # the generic (unsymmetric) point of the CCSD goldens
    out = dict(t1=t1, t2=t2)
    Wmnij = w.build_cc3_Wmnij(o, v, ERI, t1)
    out["Wmnij"] = Wmnij
    out["Wmbij"] = w.build_cc3_Wmbij(o, v, ERI, t1, Wmnij)
    out["Wmnie"] = w.build_cc3_Wmnie(o, v, ERI, t1)
    out["Wamef"] = w.build_cc3_Wamef(o, v, ERI, t1)
    out["Wabei"] = w.build_cc3_Wabei(o, v, ERI, t1)
    Fme = w.build_Fme(o, v, F, L, t1)
    X1, X2 = w._cc3_t_residual(o, v, F, ERI, L, t1, t2, Fme)
    out["X1"], out["X2"] = np.array(X1), np.array(X2)
    r1, r2 = w.residuals(F, t1, t2)
    out["r1"], out["r2"] = np.array(r1), np.array(r2)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ecc = w.solve_cc(1e-12, 1e-12, 100)
    trace = []
    for line in buf.getvalue().splitlines():
        if line.startswith("Iter") and "rms" in line:
            m = re.search(r"Ecorr =\s*(\S+)\s+dE =\s*(\S+)\s+rms =\s*(\S+)", line)
            trace.append((float(m.group(1)), float(m.group(3))))
    out["trace_ecc_rms"] = np.array(trace)
    out["ecc"] = float(ecc)
    out["conv_t1"], out["conv_t2"] = w.t1.copy(), w.t2.copy()
    path = os.path.join(HERE, "cc3_%s.npz" % tag)
    np.savez_compressed(path, **out)
    print("wrote %s  E(CC3) = %.15f  iters = %d  (E(CCSD) = %.15f)" % (path, out["ecc"], len(trace), float(g["e_ccsd"])))


def main():
    mods = mg.load_reference()
    for tag in ("o4v10_s0", "o4v10_s1_noise", "o3v7_s2"):
        case(mods, tag)


if __name__ == "__main__":
    main()
