"""CPU ORACLE (test infrastructure, NOT a product path) for the (T) one-/two-particle density increments and Lambda
sources -- SURVEY.md 8(f) "next" #2.

Plain-numpy restatement of CrawfordGroup/pycc ``cctriples.t3_density`` (cctriples.py:1063-1157; called by
``CCwfn.t3_density`` ccwfn.py:1819-1829, which ``solve_cc`` uses instead of ``t_tjl`` when ``make_t3_density`` is
set, ccwfn.py:300-304), written against the stored Dirac blocks:

  ERI[v,o,v,v][:,k][d,b,c] = ovvv[k,d,c,b]       ERI[o,v,v,v][k][d,c,b] = ovvv[k,d,c,b]
  ERI[o,o,o,v][j,k][l,c]   = ooov[j,k,l,c]       L[o,o,v,v][j,k][b,c]   = 2 oovv[j,k,b,c] - oovv[j,k,c,b]

Per (i,j,k) (full o^3 loop, cctriples.py:1119-1148): M3 = connected t3 / D, N3 = disconnected t3 / D,
X3 = sym(M3), Y3 = sym(N3) with sym(A) = 8A - 4A_bac - 4A_acb - 4A_cba + 2A_cab + 2A_bca (1124-1125), then the
accumulations of lines 1128-1148 and the closing lines 1150-1156.

PARITY PINNED: ``tests/test_t3density.py::test_oracle_*`` checks ET and all eight returned arrays against outputs of
the reference's own, unmodified code (``tests/golden/t3d_*.npz`` from ``tests/golden/make_golden_t3density.py``).

Only ``tests/`` may import this module, as the checker.
"""
from __future__ import annotations

import numpy as np

from .triples_oracle import es, t3c_ijk, t3d_ijk

NAMES = ("Doo", "Dvv", "Dov", "Goovv", "Gooov", "Gvvvo", "S1", "S2")


def sym(A):
    """8 A_abc - 4 A_bac - 4 A_acb - 4 A_cba + 2 A_cab + 2 A_bca   (cctriples.py:1124)"""
    return (8.0 * A - 4.0 * A.transpose(1, 0, 2) - 4.0 * A.transpose(0, 2, 1) - 4.0 * A.transpose(2, 1, 0)
            + 2.0 * A.transpose(2, 0, 1) + 2.0 * A.transpose(1, 2, 0))


def t3_density(t1, t2, F, ovvv, ooov, oovv, nfzc=0, js=None):
    """(ET, dict) exactly as cctriples.t3_density returns them.  ``js`` restricts the middle loop index j (the
    product shards j over ranks; the returned pieces are then partial sums, S2 NOT yet symmetrised, ET None)."""
    no, nv = t1.shape
    Fov = F[nfzc:nfzc + no, nfzc + no:]
    Loovv = 2.0 * oovv - oovv.transpose(0, 1, 3, 2)
    dvv, doo = np.zeros(nv), np.zeros(no)
    Dov = np.zeros((no, nv))
    Goovv = np.zeros_like(t2)
    Gooov = np.zeros((no, no, no, nv))
    Gvvvo = np.zeros((nv, nv, nv, no))
    S1 = np.zeros_like(t1)
    S2 = np.zeros_like(t2)
    X2 = np.zeros_like(t2)
    for i in range(no):
        for j in (range(no) if js is None else js):
            for k in range(no):
                M3 = t3c_ijk(i, j, k, t2, ovvv, ooov, F, True, nfzc)
                N3 = t3d_ijk(i, j, k, t1, t2, oovv, F, True, nfzc)
                X3, Y3 = sym(M3), sym(N3)
                U = M3 - M3.transpose(2, 1, 0)
                P = 2.0 * M3 - M3.transpose(0, 2, 1) - M3.transpose(2, 1, 0)
                W2 = 2.0 * X3 + Y3
                # doubles part of E(T)                                       (1128-1130)
                X2[i, j] += es("abc,c->ab", U, Fov[k])
                X2[i, j] += es("abc,dcb->ad", P, ovvv[k])
                X2[i] -= es("abc,lc->lab", P, ooov[j, k])
                # one-particle increments                                    (1133-1137)
                dvv += 0.5 * es("acd,acd->a", M3, X3 + Y3)
                doo[i] -= 0.5 * np.sum(M3 * (X3 + Y3))
                Dov[i] += es("abc,bc->a", U, 4.0 * t2[j, k] - 2.0 * t2[j, k].T)
                # two-particle increments                                    (1140-1143)
                Z3 = 2.0 * (M3 - M3.transpose(0, 2, 1)) - (M3.transpose(1, 0, 2) - M3.transpose(2, 0, 1))
                Goovv[i, j] += 4.0 * es("c,abc->ab", t1[k], Z3)
                Gooov[j, i] -= es("abc,lbc->la", W2, t2[:, k])
                Gvvvo[:, :, :, j] += es("abc,cd->abd", W2, t2[k, i])
                # Lambda sources                                             (1146-1148)
                S1[i] += es("abc,bc->a", 2.0 * (M3 - M3.transpose(1, 0, 2)), Loovv[j, k])
                S2[i] -= es("abc,lc->lab", W2, ooov[j, k])
                S2[i, j] += es("abc,dcb->ad", W2, ovvv[k])
    pieces = {"Doo": np.diag(doo), "Dvv": np.diag(dvv), "Dov": Dov, "Goovv": Goovv, "Gooov": Gooov, "Gvvvo": Gvvvo,
              "S1": S1, "S2": S2, "X2": X2}
    if js is not None:
        return None, pieces
    pieces["S2"] = S2 + S2.transpose(1, 0, 3, 2)                               # (1150)
    et = np.sum(t1 * S1) + np.sum((4.0 * t2 - 2.0 * t2.transpose(0, 1, 3, 2)) * X2)   # (1153-1154)
    del pieces["X2"]
    return et, pieces
