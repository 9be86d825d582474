"""CPU ORACLE (test infrastructure, NOT a product path) for the CC3 ground-state T equations -- SURVEY.md 8(f) "next" #4.

Plain-numpy restatement of CrawfordGroup/pycc, closed-shell spatial-orbital path:

  build_cc3_Wmnij   ccwfn.py:947-977      build_cc3_Wamef   ccwfn.py:1033-1052
  build_cc3_Wmbij   ccwfn.py:979-1008     build_cc3_Wabei   ccwfn.py:1054-1120
  build_cc3_Wmnie   ccwfn.py:1010-1031    _cc3_t_residual   ccwfn.py:374-430
  t3c_ijk with arbitrary (non-symmetric) W blocks          cctriples.py:27-72
  residuals (CC3 branch: r1 += X1, r2 += X2 + X2^T)        ccwfn.py:364-367

written against the six stored Dirac blocks through ``lambda_oracle.eri`` (any <pq|rs> pattern is a permuted view).

PARITY PINNED: ``tests/test_cc3.py::test_oracle_*`` checks the five intermediates, (X1, X2), the CC3 residuals at a
generic point and the full ``solve_cc`` trace against outputs of the reference's own, unmodified code
(``tests/golden/cc3_*.npz`` from ``tests/golden/make_golden_cc3.py``); the real-time variant (explicit-field triples,
real and complex amplitudes) against ``tests/golden/rtcc3_*.npz`` (``make_golden_rtcc3.py``).

Only ``tests/`` may import this module, as the checker.
"""
from __future__ import annotations

import numpy as np

from .ccsd_oracle import Diis, es
from .lambda_oracle import eri, lint


def intermediates(P, t1):
    """The five T1-dressed CC3 intermediates in the reference's index orders."""
    W = {}
    tmp = es("ijma,na->ijmn", eri(P, "ooov"), t1)
    x = es("ia,mnaf->mnif", t1, eri(P, "oovv"))
    W["Wmnij"] = eri(P, "oooo") + tmp + tmp.transpose(1, 0, 3, 2) + es("mnif,jf->mnij", x, t1)
    x = eri(P, "ovvo") + es("mbef,jf->mbej", eri(P, "ovvv"), t1)
    W["Wmbij"] = (eri(P, "ovoo") - es("mnij,nb->mbij", W["Wmnij"], t1) + es("mbie,je->mbij", eri(P, "ovov"), t1)
                  + es("ie,mbej->mbij", t1, x))
    W["Wmnie"] = eri(P, "ooov") + es("if,mnfe->mnie", t1, eri(P, "oovv"))
    W["Wamef"] = eri(P, "vovv") - es("na,nmef->amef", t1, eri(P, "oovv"))
    # W_abei = Z_abei + Z_eiab^T; the reference's symmetric + antisymmetric split of <ab|ef> sums back to the block
    Zeiam = eri(P, "vovo") + es("amef,if->amei", eri(P, "vovv"), t1).transpose(2, 3, 0, 1)
    Zmnei = eri(P, "oovo") + es("mnef,if->mnei", eri(P, "oovv"), t1)
    Z_eiab = (eri(P, "vovv") + es("if,abef->eiab", t1, eri(P, "vvvv")) - es("eiam,mb->eiab", Zeiam, t1)
              + es("anei,nb->eiab", es("ma,mnei->anei", t1, Zmnei), t1))
    Zmbei = eri(P, "ovvo") + es("mbef,if->mbei", eri(P, "ovvv"), t1)
    W["Wabei"] = -es("ma,mbei->abei", t1, Zmbei) + Z_eiab.transpose(2, 3, 0, 1)
    return W


def t3c_ijk(P, i, j, k, t2, Wvvvo, Wovoo, F):
    """Connected t3 with denominators for arbitrary W blocks (cctriples.py:50-70)."""
    Wv = lambda x: Wvvvo[:, :, :, x]
    t3 = es("bae,ce->abc", Wv(i), t2[k, j]) + es("cae,be->abc", Wv(i), t2[j, k]) + es("ace,be->abc", Wv(k), t2[j, i])
    t3 += es("bce,ae->abc", Wv(k), t2[i, j]) + es("cbe,ae->abc", Wv(j), t2[i, k]) + es("abe,ce->abc", Wv(j), t2[k, i])
    t3 -= es("mc,mab->abc", Wovoo[:, :, j, k], t2[i]) + es("mb,mac->abc", Wovoo[:, :, k, j], t2[i])
    t3 -= es("mb,mca->abc", Wovoo[:, :, i, j], t2[k]) + es("ma,mcb->abc", Wovoo[:, :, j, i], t2[k])
    t3 -= es("ma,mbc->abc", Wovoo[:, :, k, i], t2[j]) + es("mc,mba->abc", Wovoo[:, :, i, k], t2[j])
    eps = np.diagonal(F)
    eo, ev = eps[P.o], eps[P.v]
    return t3 / ((eo[i] + eo[j] + eo[k]) - (ev[:, None, None] + ev[None, :, None] + ev[None, None, :]))


def t3_pert_ijk(P, i, j, k, t2, V, F):
    """Explicit-field coupling of the connected triples, with denominators (cctriples.py:679-705):
    D t3[a,b,c] = V_ld t2[i,j,a,d] t2[k,l,c,b] -- ONE term, no permutations, as the reference writes it."""
    t3 = es("al,lcb->abc", es("ld,ad->al", V[P.o, P.v], t2[i, j]), t2[k])
    eps = np.diagonal(F)
    eo, ev = eps[P.o], eps[P.v]
    return t3 / ((eo[i] + eo[j] + eo[k]) - (ev[:, None, None] + ev[None, :, None] + ev[None, None, :]))


def t_residual(P, F, t1, t2, W=None, real_time=False):
    """(X1, X2) of _cc3_t_residual (ccwfn.py:374-430); ``real_time``: t3 -= t3_pert_ijk(V = F - H.F) (421-423)."""
    W = intermediates(P, t1) if W is None else W
    no = P.no
    Fme = F[P.o, P.v] + es("nf,mnef->me", t1, lint(P, "oovv"))            # build_Fme, ccwfn.py:563-564
    Loovv = lint(P, "oovv")
    Wkdbc = W["Wamef"].transpose(1, 0, 2, 3)                               # Wamef.swapaxes(0,1)[k] = [d,b,c]
    X1, X2 = np.zeros_like(t1), np.zeros_like(t2)
    for i in range(no):
        for j in range(no):
            for k in range(no):
                t3 = t3c_ijk(P, i, j, k, t2, W["Wabei"], W["Wmbij"], F)
                if real_time:
                    t3 = t3 - t3_pert_ijk(P, i, j, k, t2, F - P.F, F)
                u = t3 - t3.transpose(2, 1, 0)
                p = 2.0 * t3 - t3.transpose(0, 2, 1) - t3.transpose(2, 1, 0)
                X1[i] += es("abc,bc->a", u, Loovv[j, k])
                X2[i, j] += es("abc,dbc->ad", p, Wkdbc[k])
                X2[i] -= es("lc,abc->lab", W["Wmnie"][j, k], p)
                X2[i, j] += es("abc,c->ab", u, Fme[k])
    return X1, X2


def residuals(P, F, t1, t2, real_time=False):
    """CC3 residuals: the CCSD ones plus the connected-triples terms (ccwfn.py:358-367)."""
    r1, r2 = P.residuals(F, t1, t2)
    X1, X2 = t_residual(P, F, t1, t2, real_time=real_time)
    return r1 + X1, r2 + X2 + X2.transpose(1, 0, 3, 2)


def solve_cc(P, e_conv=1e-7, r_conv=1e-7, maxiter=100, max_diis=8, start_diis=1):
    """solve_cc (ccwfn.py:216-319) for model='CC3'.  Returns (ecc, t1, t2, trace[(ecc, rms)])."""
    F = P.F
    t1, t2 = P.guess()
    ecc = P.cc_energy(F, t1, t2)
    diis = Diis(t1, t2, max_diis)
    trace = []
    for niter in range(1, maxiter + 1):
        last = ecc
        r1, r2 = residuals(P, F, t1, t2)
        t1 = t1 + r1 / P.Dia
        t2 = t2 + r2 / P.Dijab
        rms = np.sqrt(np.sum((r1 / P.Dia) ** 2) + np.sum((r2 / P.Dijab) ** 2))
        ecc = P.cc_energy(F, t1, t2)
        trace.append((float(ecc), float(rms)))
        if abs(ecc - last) < e_conv and rms < r_conv:
            return float(ecc), t1, t2, trace
        diis.add_error_vector(t1, t2)
        if niter >= start_diis:
            t1, t2 = diis.extrapolate(t1, t2)
    return None, t1, t2, trace
