"""CPU ORACLE (test infrastructure, NOT a product path) for the similarity-transformed Hamiltonian (HBAR) and the
closed-shell CCSD / CCD Lambda equations -- SURVEY.md 8(f) "next" #1.

Plain-numpy restatement of pycc/cchbar.py (spatial-orbital builders) and pycc/cclambda.py (solve_lambda, residuals,
build_Goo/Gvv, r_L1, r_L2, pseudoenergy), written against the six unique Dirac blocks of ``ccsd_oracle.Problem``:
any <pq|rs> occupied/virtual pattern is a permuted view of a stored block (8-fold symmetry), and
L_pqrs = 2<pq|rs> - <pq|sr> (hamiltonian.py:70).  Only ``tests/`` may import this module, as the checker.

PARITY PINNED: ``tests/test_lambda.py::test_oracle_*`` checks every HBAR block, Goo/Gvv, r_L1, r_L2, the
pseudo-energy and the full ``solve_lambda`` iteration trace against outputs of the reference's own, unmodified code
(``tests/golden/lam_*.npz``, produced by ``tests/golden/make_golden_lambda.py``), to <= 1e-12.

Reference lines restated (all in /root/reference/pycc/):
  build_Hov    cchbar.py:128-153     build_Hvovv  cchbar.py:431-458     build_Goo    cclambda.py:258-281
  build_Hvv    cchbar.py:177-210     build_Hooov  cchbar.py:481-508     build_Gvv    cclambda.py:283-306
  build_Hoo    cchbar.py:240-273     build_Hovvo  cchbar.py:531-569     r_L1         cclambda.py:308-370
  build_Hoooo  cchbar.py:303-339     build_Hovov  cchbar.py:598-630     r_L2         cclambda.py:408-497
  build_Hvvvv  cchbar.py:367-403     build_Hvvvo  cchbar.py:632-698     pseudoenergy cclambda.py:547-570
  build_Hovoo  cchbar.py:755-823     guess        cclambda.py:62-66      solve_lambda cclambda.py:69-200
(CC2 / CC3 / spin-orbital branches are outside the accelerated path and are not restated.)
"""
from __future__ import annotations

import itertools

import numpy as np

from .ccsd_oracle import Diis, es

_STORED = ("oooo", "ooov", "oovv", "ovov", "ovvv", "vvvv")
# the eight index permutations that leave a real <pq|rs> invariant: <pq|rs> = <qp|sr> = <rs|pq> = <rq|ps> = ...
_SYM = []
for swap_el, swap_12, swap_34 in itertools.product((0, 1), repeat=3):
    g = [0, 1, 2, 3]
    if swap_12:
        g[0], g[2] = g[2], g[0]          # <pq|rs> = <rq|ps>
    if swap_34:
        g[1], g[3] = g[3], g[1]          # <pq|rs> = <ps|rq>
    if swap_el:
        g = [g[1], g[0], g[3], g[2]]     # <pq|rs> = <qp|sr>
    _SYM.append(tuple(g))


def eri(P, pat):
    """<pq|rs> for an occupied/virtual pattern such as 'vovv', as a view of one of the six stored blocks."""
    for g in _SYM:
        for name in _STORED:
            if all(pat[t] == name[g[t]] for t in range(4)):
                return getattr(P, name).transpose(g)
    raise KeyError(pat)


def lint(P, pat):
    """L_pqrs = 2<pq|rs> - <pq|sr>   (hamiltonian.py:70)"""
    swapped = pat[0] + pat[1] + pat[3] + pat[2]
    return 2.0 * eri(P, pat) - eri(P, swapped).transpose(0, 1, 3, 2)


def tau(t1, t2):
    return t2 + es("ia,jb->ijab", t1, t1)


class Hbar:
    """All eleven HBAR blocks for amplitudes (t1, t2); model 'CCSD' or 'CCD' (cchbar.py:54-99)."""

    def __init__(self, P, t1, t2, model="CCSD", F=None):
        ccd = model == "CCD"                      # 'CCSD(T)' builds the CCSD blocks (cchbar.py has no (T) branch)
        o, v = P.o, P.v
        F = P.F if F is None else F               # cclambda.residuals rebuilds HBAR with the passed (field-dressed) F
        tt = tau(t1, t2)
        oovv, Loovv = eri(P, "oovv"), lint(P, "oovv")
        # ---- one-body (cchbar.py:147-153, 201-210, 264-273)
        self.Hov = F[o, v].copy() if ccd else F[o, v] + es("nf,mnef->me", t1, Loovv)
        if ccd:
            self.Hvv = F[v, v] - es("mnfa,mnfe->ae", t2, Loovv)
            self.Hoo = F[o, o] + es("inef,mnef->mi", t2, Loovv)
        else:
            self.Hvv = (F[v, v] - es("me,ma->ae", F[o, v], t1) + es("mf,amef->ae", t1, lint(P, "vovv"))
                        - es("mnfa,mnfe->ae", tt, Loovv))
            self.Hoo = (F[o, o] + es("ie,me->mi", t1, F[o, v]) + es("ne,mnie->mi", t1, lint(P, "ooov"))
                        + es("inef,mnef->mi", tt, Loovv))
        # ---- Hoooo, Hvvvv (cchbar.py:330-339, 394-403)
        if ccd:
            self.Hoooo = eri(P, "oooo") + es("ijef,mnef->mnij", t2, oovv)
            self.Hvvvv = eri(P, "vvvv") + es("mnab,mnef->abef", t2, oovv)
        else:
            x = es("je,mnie->mnij", t1, eri(P, "ooov"))
            self.Hoooo = eri(P, "oooo") + x + x.transpose(1, 0, 3, 2) + es("ijef,mnef->mnij", tt, oovv)
            x = es("mb,amef->abef", t1, eri(P, "vovv"))
            self.Hvvvv = eri(P, "vvvv") - x - x.transpose(1, 0, 3, 2) + es("mnab,mnef->abef", tt, oovv)
        # ---- Hvovv, Hooov (cchbar.py:450-458, 500-508)
        self.Hvovv = eri(P, "vovv").copy()
        self.Hooov = eri(P, "ooov").copy()
        if not ccd:
            self.Hvovv = self.Hvovv - es("na,nmef->amef", t1, oovv)
            self.Hooov = self.Hooov + es("if,nmef->mnie", t1, oovv)
        # ---- Hovvo, Hovov (cchbar.py:554-569, 617-630)
        if ccd:
            self.Hovvo = eri(P, "ovvo") - es("jnfb,mnef->mbej", t2, oovv) + es("njfb,mnef->mbej", t2, Loovv)
            self.Hovov = eri(P, "ovov") - es("jnfb,nmef->mbje", t2, oovv)
        else:
            self.Hovvo = (eri(P, "ovvo") + es("jf,mbef->mbej", t1, eri(P, "ovvv"))
                          - es("nb,mnej->mbej", t1, eri(P, "oovo")) - es("jnfb,mnef->mbej", tt, oovv)
                          + es("njfb,mnef->mbej", t2, Loovv))
            self.Hovov = (eri(P, "ovov") + es("jf,bmef->mbje", t1, eri(P, "vovv"))
                          - es("nb,mnje->mbje", t1, eri(P, "ooov")) - es("jnfb,nmef->mbje", tt, oovv))
        # ---- Hvvvo (cchbar.py:662-698)
        vovv, Lvovv = eri(P, "vovv"), lint(P, "vovv")
        H = (eri(P, "vvvo") - es("me,miab->abei", self.Hov, t2) + es("mnab,mnei->abei", tt, eri(P, "oovo"))
             - es("imfa,bmfe->abei", t2, vovv) - es("imfb,amef->abei", t2, vovv) + es("mifb,amef->abei", t2, Lvovv))
        if not ccd:
            H = H + es("if,abef->abei", t1, self.Hvvvv)
            x = eri(P, "vovo") - es("infa,mnfe->amei", t2, oovv)
            H = H - es("mb,amei->abei", t1, x)
            x = eri(P, "voov") - es("infb,mnef->bmie", t2, oovv) + es("nifb,mnef->bmie", t2, Loovv)
            H = H - es("ma,bmie->abei", t1, x)
        self.Hvvvo = H
        # ---- Hovoo (cchbar.py:785-823)
        ooov, Looov = eri(P, "ooov"), lint(P, "ooov")
        H = (eri(P, "ovoo") + es("me,ijeb->mbij", self.Hov, t2) + es("ijef,mbef->mbij", tt, eri(P, "ovvv"))
             - es("ineb,nmje->mbij", t2, ooov) - es("jneb,mnie->mbij", t2, ooov) + es("njeb,mnie->mbij", t2, Looov))
        if not ccd:
            H = H - es("nb,mnij->mbij", t1, self.Hoooo)
            x = eri(P, "ovov") - es("infb,mnfe->mbie", t2, oovv)
            H = H + es("je,mbie->mbij", t1, x)
            x = eri(P, "voov") - es("jnfb,mnef->bmje", t2, oovv) + es("njfb,mnef->bmje", t2, Loovv)
            H = H + es("ie,bmje->mbij", t1, x)
        self.Hovoo = H


def guess(t1, t2):
    """l1 = 2 t1, l2 = 2 (2 t2 - t2^T)     (cclambda.py:65-66)"""
    return 2.0 * t1, 2.0 * (2.0 * t2 - t2.transpose(0, 1, 3, 2))


def Goo(t2, l2):
    return es("mjab,ijab->mi", t2, l2)                      # cclambda.py:281


def Gvv(t2, l2):
    return -es("ijeb,ijab->ae", t2, l2)                     # cclambda.py:306


def r_L1(H, l1, l2, gvv, goo, model="CCSD", s1=None):
    """cclambda.py:344-370; ``s1``: the (T) source cc.S1 of a CCSD(T) wavefunction (350-352)"""
    if model == "CCD":
        return np.zeros_like(l1)
    r = 2.0 * H.Hov
    if s1 is not None:
        r = r + s1
    r = r + es("ie,ea->ia", l1, H.Hvv) - es("ma,im->ia", l1, H.Hoo)
    r = r + es("imef,efam->ia", l2, H.Hvvvo) - es("mnae,iemn->ia", l2, H.Hovoo)
    r = r + es("me,ieam->ia", l1, 2.0 * H.Hovvo - H.Hovov.transpose(0, 1, 3, 2))
    r = r - 2.0 * es("ef,eifa->ia", gvv, H.Hvovv) + es("ef,eiaf->ia", gvv, H.Hvovv)
    r = r - 2.0 * es("mn,mina->ia", goo, H.Hooov) + es("mn,imna->ia", goo, H.Hooov)
    return r


def r_L2(P, H, l1, l2, gvv, goo, model="CCSD", s2=None):
    """cclambda.py:447-497; ``s2``: the (T) source cc.S2, added as 1/2 s2 before the symmetrisation (470-474)"""
    Loovv = lint(P, "oovv")
    r = Loovv.copy()
    if s2 is not None:
        r = r + 0.5 * s2
    if model != "CCD":
        r = r + 2.0 * es("ia,jb->ijab", l1, H.Hov) - es("ja,ib->ijab", l1, H.Hov)
        r = r + 2.0 * es("ie,ejab->ijab", l1, H.Hvovv) - es("ie,ejba->ijab", l1, H.Hvovv)
        r = r - 2.0 * es("mb,jima->ijab", l1, H.Hooov) + es("mb,ijma->ijab", l1, H.Hooov)
    r = r + es("ijeb,ea->ijab", l2, H.Hvv) - es("mjab,im->ijab", l2, H.Hoo)
    r = r + 0.5 * es("mnab,ijmn->ijab", l2, H.Hoooo) + 0.5 * es("ijef,efab->ijab", l2, H.Hvvvv)
    r = r + es("mjeb,ieam->ijab", l2, 2.0 * H.Hovvo - H.Hovov.transpose(0, 1, 3, 2))
    r = r - es("mibe,jema->ijab", l2, H.Hovov) - es("mieb,jeam->ijab", l2, H.Hovvo)
    r = r + es("ae,ijeb->ijab", gvv, Loovv) - es("mi,mjab->ijab", goo, Loovv)
    return r + r.transpose(1, 0, 3, 2)


def pseudoenergy(P, l2):
    return 0.5 * es("ijab,ijab->", eri(P, "oovv"), l2)      # cclambda.py:570


def residuals(P, t1, t2, l1, l2, model="CCSD", s1=None, s2=None, F=None):
    """cclambda.py:202-256: HBAR rebuilt from (F, t1, t2), then r_L1 / r_L2 (amplitudes may be complex: RT-CC)"""
    H = Hbar(P, t1, t2, model, F)
    gvv, goo = Gvv(t2, l2), Goo(t2, l2)
    return r_L1(H, l1, l2, gvv, goo, model, s1), r_L2(P, H, l1, l2, gvv, goo, model, s2)


def solve_lambda(P, t1, t2, e_conv=1e-7, r_conv=1e-7, maxiter=100, max_diis=8, start_diis=1, model="CCSD", s1=None,
                 s2=None):
    """cclambda.py:69-200.  Returns (pseudo-energy, l1, l2, trace[(lecc, rms)]); None as energy if not converged."""
    H = Hbar(P, t1, t2, model)
    l1, l2 = guess(t1, t2)
    lecc = pseudoenergy(P, l2)
    diis = Diis(l1, l2, max_diis)
    trace = []
    for niter in range(1, maxiter + 1):
        last = lecc
        gvv, goo = Gvv(t2, l2), Goo(t2, l2)
        r1 = r_L1(H, l1, l2, gvv, goo, model, s1)
        r2 = r_L2(P, H, l1, l2, gvv, goo, model, s2)
        l1 = l1 + r1 / P.Dia
        l2 = l2 + r2 / P.Dijab
        rms = np.sqrt(es("ia,ia->", r1 / P.Dia, r1 / P.Dia) + es("ijab,ijab->", r2 / P.Dijab, r2 / P.Dijab))
        lecc = pseudoenergy(P, l2)
        trace.append((float(lecc), float(rms)))
        if abs(lecc - last) < e_conv and abs(rms) < r_conv:
            return float(lecc), l1, l2, trace
        diis.add_error_vector(l1, l2)
        if niter >= start_diis:
            l1, l2 = diis.extrapolate(l1, l2)
    return None, l1, l2, trace
