#!/usr/bin/env python
"""Golden vectors for the CC2 model (reference ccwfn.py:596-602 Wmnij, 711-713 Zmbij, 832-884 _r_T2_cc2, solve_cc) from
the UNMODIFIED reference with the shims of make_golden.py.

    python tests/golden/make_golden_cc2.py        # writes tests/golden/cc2_<tag>.npz
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
    w = mg.reference_wfn(ccwfn_mod, device_mod, syn, ERI, model="CC2")
    o, v, F, L = w.o, w.v, w.H.F, w.H.L
    t1, t2 = g["rand_t1"], g["rand_t2"]
    out = dict(t1=t1, t2=t2)
    out["Wmnij"] = w.build_Wmnij(o, v, ERI, t1, t2)
    out["Zmbij"] = w.build_Zmbij(o, v, ERI, t1, t2)
    assert w.build_Wmbej(o, v, ERI, L, t1, t2) is None and w.build_Wmbje(o, v, ERI, t1, t2) is None
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
    path = os.path.join(HERE, "cc2_%s.npz" % tag)
    np.savez_compressed(path, **out)
    print("wrote %s  E(CC2) = %.15f  iters = %d" % (path, out["ecc"], len(trace)))


def main():
    mods = mg.load_reference()
    for tag in ("o4v10_s0", "o4v10_s1_noise", "o3v7_s2"):
        case(mods, tag)


if __name__ == "__main__":
    main()
