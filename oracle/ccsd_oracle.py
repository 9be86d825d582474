"""CPU ORACLE (test infrastructure, NOT a product path) for the RHF-CCSD hot path.

A plain-numpy restatement of the algorithm of CrawfordGroup/pycc's closed-shell
CCSD amplitude iteration, written against the SIX UNIQUE Dirac integral blocks
(oooo, ooov, oovv, ovov, ovvv, vvvv) instead of the reference's full n^4
``ERI``/``L`` arrays, so it also runs at sizes where 2 x n^4 doubles do not fit.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module, and only as the
checker / CPU baseline.  ``pycc_b200`` never imports it.

PARITY PINNED: ``tests/test_oracle_golden.py`` checks every function below
against outputs of the reference's own, unmodified code
(``tests/golden/ref_*.npz``, produced by ``tests/golden/make_golden.py`` which
imports /root/reference): all seven intermediates, r1, r2, the energy, the full
``solve_cc`` iteration trace and the DIIS extrapolants, to <= 1e-12.

Reference lines restated (all in /root/reference/pycc/):
  build_tau      ccwfn.py:432-455      build_Wmbej  ccwfn.py:607-646
  build_Fae      ccwfn.py:458-498      build_Wmbje  ccwfn.py:649-684
  build_Fmi      ccwfn.py:501-534      build_Zmbij  ccwfn.py:687-715
  build_Fme      ccwfn.py:537-565      r_T1         ccwfn.py:718-761
  build_Wmnij    ccwfn.py:568-604      r_T2         ccwfn.py:764-791, 886-944
  cc_energy      ccwfn.py:1122-1162    solve_cc     ccwfn.py:216-319
  helper_diis    utils.py:257-361      L = 2<pq|rs>-<pq|sr>  hamiltonian.py:70

Block identities used (8-fold symmetry of real integrals):
  <mn|ej> = ooov[n,m,j,e]    <mb|ej> = oovv[m,j,e,b]    <ab|ej> = ovvv[j,e,b,a]
  <mb|ij> = ooov[i,j,m,b]    <ab|ij> = oovv[i,j,a,b]
"""
from __future__ import annotations

import numpy as np

BLOCK_NAMES = ("oooo", "ooov", "oovv", "ovov", "ovvv", "vvvv")


def es(sub, *ops):
    return np.einsum(sub, *ops, optimize=True)


def blocks_from_full(ERI, no, nfzc=0):
    """Slice the six unique blocks out of a full Dirac <pq|rs> array."""
    n = ERI.shape[0]
    sl = {"o": slice(nfzc, nfzc + no), "v": slice(nfzc + no, n)}
    return {k: np.ascontiguousarray(ERI[sl[k[0]], sl[k[1]], sl[k[2]], sl[k[3]]])
            for k in BLOCK_NAMES}


class Problem:
    """Blocks + Fock of one closed-shell problem; derives the L blocks once."""

    def __init__(self, blocks, F, no, nfzc=0, need_vvvv=True):
        self.no = no
        n = F.shape[0]
        self.nv = n - no - nfzc
        self.o = slice(nfzc, nfzc + no)
        self.v = slice(nfzc + no, n)
        self.F = F
        self.eps = np.diagonal(F).copy()
        b = blocks
        self.oooo, self.ooov, self.oovv = b["oooo"], b["ooov"], b["oovv"]
        self.ovov, self.ovvv = b["ovov"], b["ovvv"]
        self.vvvv = b.get("vvvv")
        # L_pqrs = 2<pq|rs> - <pq|sr>            (hamiltonian.py:70)
        self.Loovv = 2.0 * self.oovv - self.oovv.transpose(0, 1, 3, 2)
        self.Lovvv = 2.0 * self.ovvv - self.ovvv.transpose(0, 1, 3, 2)
        # L[m,n,i,e] = 2<mn|ie> - <mn|ei>,  <mn|ei> = ooov[n,m,i,e]
        self.Looov = 2.0 * self.ooov - self.ooov.transpose(1, 0, 2, 3)
        eo, ev = self.eps[self.o], self.eps[self.v]
        self.Dia = eo[:, None] - ev[None, :]
        self.Dijab = (eo[:, None, None, None] + eo[None, :, None, None]
                      - ev[None, None, :, None] - ev[None, None, None, :])

    # ---- amplitudes -------------------------------------------------------
    def guess(self):
        """t1 = 0, t2 = <ij|ab>/D_ijab     (ccwfn.py:210-211)"""
        return np.zeros((self.no, self.nv)), self.oovv / self.Dijab

    @staticmethod
    def tau(t1, t2, f1=1.0, f2=1.0):
        """tau = f1 t2 + f2 t1 (x) t1      (ccwfn.py:455)"""
        return f1 * t2 + f2 * es("ia,jb->ijab", t1, t1)

    # ---- one-body intermediates ---------------------------------------------
    def Fae(self, F, t1, t2):
        o, v = self.o, self.v
        X = F[v, v] - 0.5 * es("me,ma->ae", F[o, v], t1)
        X += es("mf,mafe->ae", t1, self.Lovvv)
        X -= es("mnaf,mnef->ae", self.tau(t1, t2, 1.0, 0.5), self.Loovv)
        return X

    def Fmi(self, F, t1, t2):
        o, v = self.o, self.v
        X = F[o, o] + 0.5 * es("ie,me->mi", t1, F[o, v])
        X += es("ne,mnie->mi", t1, self.Looov)
        X += es("inef,mnef->mi", self.tau(t1, t2, 1.0, 0.5), self.Loovv)
        return X

    def Fme(self, F, t1):
        return F[self.o, self.v] + es("nf,mnef->me", t1, self.Loovv)

    # ---- two-body intermediates ---------------------------------------------
    def Wmnij(self, t1, t2):
        X = self.oooo + es("je,mnie->mnij", t1, self.ooov)
        X += es("ie,nmje->mnij", t1, self.ooov)          # <mn|ej> = ooov[n,m,j,e]
        X += es("ijef,mnef->mnij", self.tau(t1, t2), self.oovv)
        return X

    def Wmbej(self, t1, t2):
        # <mb|ej> = oovv[m,j,e,b]   (out-of-place first term: the result takes the amplitudes' dtype, real or complex)
        X = self.oovv.transpose(0, 3, 2, 1) + es("jf,mbef->mbej", t1, self.ovvv)
        X -= es("nb,nmje->mbej", t1, self.ooov)          # <mn|ej> = ooov[n,m,j,e]
        X -= es("jnfb,mnef->mbej", self.tau(t1, t2, 0.5, 1.0), self.oovv)
        X += 0.5 * es("njfb,mnef->mbej", t2, self.Loovv)
        return X

    def Wmbje(self, t1, t2):
        X = -self.ovov - es("jf,mbfe->mbje", t1, self.ovvv)
        X += es("nb,mnje->mbje", t1, self.ooov)
        X += es("jnfb,mnfe->mbje", self.tau(t1, t2, 0.5, 1.0), self.oovv)
        return X

    def Zmbij(self, t1, t2):
        return es("mbef,ijef->mbij", self.ovvv, self.tau(t1, t2))

    # ---- residuals -----------------------------------------------------------
    def r1(self, F, t1, t2, Fae, Fme, Fmi):
        o, v = self.o, self.v
        s = 2.0 * t2 - t2.transpose(0, 1, 3, 2)
        X = F[v, o].T + es("ie,ae->ia", t1, Fae)
        X -= es("mi,ma->ia", Fmi, t1)
        X += es("imae,me->ia", s, Fme)
        # L[n,a,f,i] = 2<na|fi> - <na|if> = 2 oovv[n,i,f,a] - ovov[n,a,i,f]
        X += 2.0 * es("nf,nifa->ia", t1, self.oovv) - es("nf,naif->ia", t1, self.ovov)
        X += es("mief,maef->ia", s, self.ovvv)
        # L[n,m,e,i] = 2<nm|ei> - <nm|ie> = 2 ooov[m,n,i,e] - ooov[n,m,i,e]
        X -= es("mnae,mnie->ia", t2, 2.0 * self.ooov - self.ooov.transpose(1, 0, 2, 3))
        return X

    def r2_terms(self, F, t1, t2, Fae, Fme, Fmi, Wmnij, Wmbej, Wmbje, Zmbij):
        """The unsymmetrised half-residual of ccwfn.py:922-940, term by term."""
        tau = self.tau(t1, t2)
        T = {}
        T["drive"] = 0.5 * self.oovv                                     # 922
        T["Fae"] = es("ijae,be->ijab", t2, Fae)                          # 923
        T["FmeT1v"] = -0.5 * es("ijae,mb,me->ijab", t2, t1, Fme)         # 924-925
        T["Fmi"] = -es("imab,mj->ijab", t2, Fmi)                         # 926
        T["FmeT1o"] = -0.5 * es("imab,je,me->ijab", t2, t1, Fme)         # 927-928
        T["Wmnij"] = 0.5 * es("mnab,mnij->ijab", tau, Wmnij)             # 930
        T["ladder"] = 0.5 * es("ijef,abef->ijab", tau, self.vvvv)        # 931
        T["Zmbij"] = -es("ma,mbij->ijab", t1, Zmbij)                     # 932
        T["ringA"] = es("imae,mbej->ijab", t2 - t2.transpose(0, 1, 3, 2), Wmbej)          # 933
        T["ringB"] = es("imae,mbej->ijab", t2, Wmbej + Wmbje.transpose(0, 1, 3, 2))       # 934
        T["ringC"] = es("mjae,mbie->ijab", t2, Wmbje)                    # 935
        T["t1t1a"] = -es("ie,ma,mjeb->ijab", t1, t1, self.oovv)          # 936-937 <mb|ej>=oovv[m,j,e,b]
        T["t1t1b"] = -es("ie,mb,maje->ijab", t1, t1, self.ovov)          # 938
        T["vvvo"] = es("ie,jeba->ijab", t1, self.ovvv)                   # 939 <ab|ej>=ovvv[j,e,b,a]
        T["ovoo"] = -es("ma,ijmb->ijab", t1, self.ooov)                  # 940 <mb|ij>=ooov[i,j,m,b]
        return T

    def residuals(self, F, t1, t2, parts=False):
        """(r1, r2) of ccwfn.py:321-372 (CCSD branch)."""
        Fae = self.Fae(F, t1, t2)
        Fmi = self.Fmi(F, t1, t2)
        Fme = self.Fme(F, t1)
        Wmnij = self.Wmnij(t1, t2)
        Wmbej = self.Wmbej(t1, t2)
        Wmbje = self.Wmbje(t1, t2)
        Zmbij = self.Zmbij(t1, t2)
        r1 = self.r1(F, t1, t2, Fae, Fme, Fmi)
        T = self.r2_terms(F, t1, t2, Fae, Fme, Fmi, Wmnij, Wmbej, Wmbje, Zmbij)
        half = sum(T.values())
        r2 = half + half.transpose(1, 0, 3, 2)                           # 790
        if parts:
            inter = dict(Fae=Fae, Fmi=Fmi, Fme=Fme, Wmnij=Wmnij, Wmbej=Wmbej,
                         Wmbje=Wmbje, Zmbij=Zmbij)
            return r1, r2, inter, T
        return r1, r2

    def cc_energy(self, F, t1, t2):
        """E = 2 f_ia t_ia + tau_ijab L_ijab     (ccwfn.py:1160-1161)"""
        return 2.0 * np.sum(F[self.o, self.v] * t1) + np.sum(self.tau(t1, t2) * self.Loovv)


class Diis:
    """Restatement of helper_diis (utils.py:272-361)."""

    def __init__(self, t1, t2, max_diis):
        self.old = (t1.copy(), t2.copy())
        self.vals = [(t1.copy(), t2.copy())]
        self.errors = []
        self.max_diis = max_diis
        self.last_c = None

    def add_error_vector(self, t1, t2):
        self.vals.append((t1.copy(), t2.copy()))
        self.errors.append(np.concatenate(((t1 - self.old[0]).ravel(), (t2 - self.old[1]).ravel())))
        self.old = (t1.copy(), t2.copy())

    def extrapolate(self, t1, t2):
        if self.max_diis == 0:
            return t1, t2
        if len(self.errors) > self.max_diis:
            del self.vals[0]
            del self.errors[0]
        m = len(self.errors)
        B = -np.ones((m + 1, m + 1))
        B[-1, -1] = 0.0
        for p in range(m):
            for q in range(p, m):
                B[p, q] = B[q, p] = np.dot(self.errors[p], self.errors[q])
        B[:-1, :-1] /= np.abs(B[:-1, :-1]).max()
        rhs = np.zeros(m + 1)
        rhs[-1] = -1.0
        c = np.linalg.solve(B, rhs)
        self.last_c = c
        n1 = np.zeros_like(self.old[0])
        n2 = np.zeros_like(self.old[1])
        for p in range(m):
            n1 += c[p] * self.vals[p + 1][0]
            n2 += c[p] * self.vals[p + 1][1]
        self.old = (n1.copy(), n2.copy())
        return n1, n2


def solve_cc(P, e_conv=1e-7, r_conv=1e-7, maxiter=100, max_diis=8, start_diis=1, F=None):
    """The iteration of ccwfn.py:216-319.  Returns (ecc, t1, t2, trace[(ecc, rms)])."""
    F = P.F if F is None else F
    t1, t2 = P.guess()
    ecc = P.cc_energy(F, t1, t2)
    diis = Diis(t1, t2, max_diis)
    trace = []
    for niter in range(1, maxiter + 1):
        last = ecc
        r1, r2 = P.residuals(F, t1, t2)
        d1, d2 = r1 / P.Dia, r2 / P.Dijab
        t1 = t1 + d1
        t2 = t2 + d2
        rms = np.sqrt(np.sum(d1 * d1) + np.sum(d2 * d2))
        ecc = P.cc_energy(F, t1, t2)
        trace.append((ecc, rms))
        if abs(ecc - last) < e_conv and rms < r_conv:
            return ecc, t1, t2, trace
        diis.add_error_vector(t1, t2)
        if niter >= start_diis:
            t1, t2 = diis.extrapolate(t1, t2)
    return None, t1, t2, trace
