#!/usr/bin/env python
"""Golden vectors for the (T) densities / Lambda sources, from the UNMODIFIED reference
(pycc/cctriples.py:1063-1157 ``t3_density``; ``CCwfn.t3_density`` ccwfn.py:1819-1829) run in the build container with
the shims of make_golden.py.

    python tests/golden/make_golden_t3density.py        # writes tests/golden/t3d_<tag>.npz

Inputs are those of the CCSD goldens (ref_<tag>.npz: factor B, F, scale, converged t1/t2).  Every stored array is an
output of the reference's own code.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

NAMES = ("Doo", "Dvv", "Dov", "Goovv", "Gooov", "Gvvvo", "S1", "S2")


def case(mods, tag):
    ccwfn_mod, cctriples, utils, device_mod = mods
    from pycc_b200.synthetic import Synthetic, full_eri
    g = dict(np.load(os.path.join(HERE, "ref_%s.npz" % tag)))
    syn = Synthetic(int(g["no"]), int(g["nv"]), g["B"], g["F"], float(g["scale"]), int(g["seed"]))
    ERI = full_eri(syn)
    w = mg.reference_wfn(ccwfn_mod, device_mod, syn, ERI, model="CCSD(T)")
    w.t1, w.t2 = g["conv_t1"].copy(), g["conv_t2"].copy()
    et = w.t3_density()                      # ccwfn.py:1819-1829 -> cctriples.t3_density, caches the pieces on w
    out = dict(t1=w.t1.copy(), t2=w.t2.copy(), et=float(et), e_t_tjl=float(g["e_t_tjl"]))
    for k in NAMES:
        out[k] = np.array(getattr(w, k))
    path = os.path.join(HERE, "t3d_%s.npz" % tag)
    np.savez_compressed(path, **out)
    print("wrote %s  E(T) = %.15f  (t_tjl %.15f)" % (path, out["et"], out["e_t_tjl"]))


def main():
    mods = mg.load_reference()
    for tag in ("o4v10_s0", "o4v10_s1_noise", "o3v7_s2"):
        case(mods, tag)


if __name__ == "__main__":
    main()
