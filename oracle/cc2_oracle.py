"""CPU ORACLE (test infrastructure, NOT a product path) for the closed-shell CC2 model.

Plain-numpy restatement of CrawfordGroup/pycc: the CC2 branches of ``build_Wmnij`` (ccwfn.py:596-602) and
``build_Zmbij`` (711-713), ``_r_T2_cc2`` (832-884) and ``r_T2``'s symmetrisation (788-790); r_T1 and Fae/Fmi/Fme are the
CCSD ones (``ccsd_oracle.Problem``).

PARITY PINNED: ``tests/test_cc2.py::test_oracle_*`` against outputs of the reference's own code
(``tests/golden/cc2_*.npz`` from ``tests/golden/make_golden_cc2.py``).  Only ``tests/`` may import this module.
"""
from __future__ import annotations

import numpy as np

from .ccsd_oracle import Diis, es
from .lambda_oracle import eri


def Wmnij(P, t1):
    W = eri(P, "oooo") + es("je,mnie->mnij", t1, eri(P, "ooov")) + es("ie,mnej->mnij", t1, eri(P, "oovo"))
    return W + es("mnif,fj->mnij", es("mnef,ei->mnif", eri(P, "oovv"), t1.T), t1.T)


def Zmbij(P, t1):
    return es("mbif,fj->mbij", es("mbef,ie->mbif", eri(P, "ovvv"), t1), t1.T)


def r2_half(P, F, t1, t2):
    """_r_T2_cc2 (ccwfn.py:868-881), before the symmetrisation"""
    o, v = P.o, P.v
    r = 0.5 * eri(P, "vvoo").transpose(2, 3, 0, 1)
    tmp = F[v, v] - 0.5 * es("me,ma->ae", F[o, v], t1)
    r = r + es("ijae,eb->ijab", t2, tmp.T)
    r = r - 0.5 * es("ijae,eb->ijab", t2, es("mb,me->be", t1, F[o, v]).T)
    r = r - es("imab,mj->ijab", t2, F[o, o] + 0.5 * es("ie,me->mi", t1, F[o, v]))
    r = r - 0.5 * es("imab,jm->ijab", t2, es("je,me->jm", t1, F[o, v]))
    r = r + 0.5 * es("ma,mbij->ijab", t1, es("nb,mnij->mbij", t1, Wmnij(P, t1)))
    r = r + 0.5 * es("jf,abif->ijab", t1, es("ie,abef->abif", t1, eri(P, "vvvv")))
    r = r - es("ma,mbij->ijab", t1, Zmbij(P, t1))
    r = r - es("ma,mbij->ijab", t1, es("ie,mbej->mbij", t1, eri(P, "ovvo")))
    r = r - es("mb,maji->ijab", t1, es("ie,maje->maji", t1, eri(P, "ovov")))
    r = r + es("ie,abej->ijab", t1, eri(P, "vvvo"))
    r = r - es("ma,mbij->ijab", t1, eri(P, "ovoo"))
    return r


def residuals(P, F, t1, t2):
    r1 = P.r1(F, t1, t2, P.Fae(F, t1, t2), P.Fme(F, t1), P.Fmi(F, t1, t2))
    h = r2_half(P, F, t1, t2)
    return r1, h + h.transpose(1, 0, 3, 2)


def solve_cc(P, e_conv=1e-7, r_conv=1e-7, maxiter=100, max_diis=8, start_diis=1):
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
