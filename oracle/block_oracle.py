"""CPU ORACLE (test infrastructure, NOT a product path): the CCSD residual at sizes where the oracle cannot hold the
integral blocks -- r1 in full and r2 for INDIVIDUAL occupied pairs (i,j), evaluated straight from the factor of the
synthetic integrals (pycc_b200/synthetic.py: <pq|rs> = scale * sum_P B[P,p,r] B[P,q,s]).

It restates, for a fixed pair, exactly the terms of ``ccsd_oracle.Problem`` (reference lines in /root/reference/pycc/):
  build_Fae/Fmi/Fme  ccwfn.py:458-565      build_Wmnij  ccwfn.py:568-604      build_Wmbej  ccwfn.py:607-646
  build_Wmbje        ccwfn.py:649-684      build_Zmbij  ccwfn.py:687-715      r_T1         ccwfn.py:718-761
  _r_T2_ccsd         ccwfn.py:886-944 and the symmetrisation r2 += r2.swapaxes(0,1).swapaxes(2,3) (ccwfn.py:790)
with <mb|ef> and <ab|ef> never formed: every contraction with them goes through B (two thin products instead of one).
At o=40, v=300 one pair costs ~5e11 flop on the host (seconds), where the full residual would need 64.8 GB of <ab|ef>.

PARITY PINNED through ``ccsd_oracle`` (itself pinned to the reference's goldens): ``tests/test_block_oracle.py`` checks
every pair and r1 against ``Problem.residuals`` at the golden sizes (<= 1e-13), for symmetric and generic amplitudes.

Only ``tests/`` may import this module, as the checker.
"""
from __future__ import annotations

import numpy as np


def es(sub, *ops):
    return np.einsum(sub, *ops, optimize=True)


class FactorProblem:
    """``syn``: pycc_b200.synthetic.Synthetic (B, F, scale, no, nv)."""

    def __init__(self, syn):
        self.no, self.nv, self.s = syn.no, syn.nv, float(syn.scale)
        no = self.no
        B = syn.B
        self.Boo = np.ascontiguousarray(B[:, :no, :no])
        self.Bov = np.ascontiguousarray(B[:, :no, no:])
        self.Bvv = np.ascontiguousarray(B[:, no:, no:])
        self.F = syn.F
        self.o, self.v = slice(0, no), slice(no, no + self.nv)
        blk = lambda X, Y: np.tensordot(X, Y, axes=(0, 0)).transpose(0, 2, 1, 3) * self.s     # [p,r],[q,s] -> <pq|rs>
        self.oooo = blk(self.Boo, self.Boo)
        self.ooov = blk(self.Boo, self.Bov)                    # <mn|ie> = (mi|ne)
        self.oovv = blk(self.Bov, self.Bov)                    # <mn|ef> = (me|nf)
        self.ovov = np.tensordot(self.Boo, self.Bvv, axes=(0, 0)).transpose(0, 2, 1, 3) * self.s   # <mb|je> = (mj|be)
        self.Loovv = 2.0 * self.oovv - self.oovv.transpose(0, 1, 3, 2)
        self.Looov = 2.0 * self.ooov - self.ooov.transpose(1, 0, 2, 3)
        self._cache = {}

    # ---- amplitude-dependent state ------------------------------------------------------------------------------
    def bind(self, F, t1, t2):
        """Fix (F, t1, t2); builds the one-body intermediates (ccwfn.py:458-565)."""
        o, v, no, nv, s = self.o, self.v, self.no, self.nv, self.s
        self.Fk, self.t1, self.t2 = F, t1, t2
        self.tau = t2 + es("ia,jb->ijab", t1, t1)
        self.tauh = 0.5 * t2 + es("ia,jb->ijab", t1, t1)            # build_tau(t1, t2, 0.5, 1.0)
        tau12 = t2 + 0.5 * es("ia,jb->ijab", t1, t1)                # build_tau(t1, t2, 1.0, 0.5)
        self._cache = {}
        # Fme = f_me + t_nf L_mnef
        self.Fme = F[o, v] + es("nf,mnef->me", t1, self.Loovv)
        # Fae = f_ae - 1/2 f_me t_ma + t_mf L_mafe - tau(1,1/2)_mnaf L_mnef,  L_mafe = 2<ma|fe> - <ma|ef>
        c = es("mf,Pmf->P", t1, self.Bov)
        Y = es("mf,Paf->Pma", t1, self.Bvv)
        tL = 2.0 * s * es("P,Pae->ae", c, self.Bvv) - s * es("Pme,Pma->ae", self.Bov, Y)
        self.Fae = F[v, v] - 0.5 * t1.T @ F[o, v] + tL
        A = tau12.transpose(2, 0, 1, 3).reshape(nv, no * no * nv)          # [a,(m,n,f)]
        Lm = self.Loovv.transpose(0, 1, 3, 2).reshape(no * no * nv, nv)    # [(m,n,f),e]
        self.Fae = self.Fae - A @ Lm
        # Fmi = f_mi + 1/2 t_ie f_me + t_ne L_mnie + tau(1,1/2)_inef L_mnef
        self.Fmi = (F[o, o] + 0.5 * F[o, v] @ t1.T + es("ne,mnie->mi", t1, self.Looov)
                    + es("inef,mnef->mi", tau12, self.Loovv))
        return self

    def _t1_ovvv_f(self, q):
        """[m,b,e] = t_qf <mb|ef>"""
        x = self.Bvv @ self.t1[q]                                          # [P,b]
        return self.s * es("Pme,Pb->mbe", self.Bov, x)

    def _t1_ovvv_e(self, q):
        """[m,b,e] = t_qf <mb|fe>"""
        x = self.Bov @ self.t1[q]                                          # [P,m]
        return self.s * es("Pm,Pbe->mbe", x, self.Bvv)

    def Wmbej_col(self, q):
        """Wmbej[:, :, :, q] as [m,b,e] (ccwfn.py:641-645)."""
        key = ("ej", q)
        if key not in self._cache:
            no, nv = self.no, self.nv
            t1, t2 = self.t1, self.t2
            W = self.oovv[:, q].transpose(0, 2, 1) + self._t1_ovvv_f(q)                     # <mb|ej> = oovv[m,j,e,b]
            W = W - es("nb,nme->mbe", t1, self.ooov[:, :, q, :])                     # <mn|ej> = ooov[n,m,j,e]
            A = self.oovv.transpose(0, 2, 1, 3).reshape(no * nv, no * nv)                   # [(m,e),(n,f)]
            W = W - (A @ self.tauh[q].transpose(0, 1, 2).reshape(no * nv, nv)).reshape(no, nv, nv).transpose(0, 2, 1)
            AL = self.Loovv.transpose(0, 2, 1, 3).reshape(no * nv, no * nv)
            W = W + 0.5 * (AL @ t2[:, q].reshape(no * nv, nv)).reshape(no, nv, nv).transpose(0, 2, 1)
            self._cache[key] = W
        return self._cache[key]

    def Wmbje_col(self, q):
        """Wmbje[:, :, q, :] as [m,b,e] (ccwfn.py:680-683)."""
        key = ("je", q)
        if key not in self._cache:
            no, nv = self.no, self.nv
            t1 = self.t1
            W = -self.ovov[:, :, q, :] - self._t1_ovvv_e(q)
            W = W + es("nb,mne->mbe", t1, self.ooov[:, :, q, :])
            A = self.oovv.transpose(0, 3, 1, 2).reshape(no * nv, no * nv)                   # [(m,e),(n,f)] = <mn|fe>
            W = W + (A @ self.tauh[q].reshape(no * nv, nv)).reshape(no, nv, nv).transpose(0, 2, 1)
            self._cache[key] = W
        return self._cache[key]

    # ---- r1 in full (ccwfn.py:754-760) --------------------------------------------------------------------------
    def r1(self):
        o, v, no, nv = self.o, self.v, self.no, self.nv
        F, t1, t2 = self.Fk, self.t1, self.t2
        s2 = 2.0 * t2 - t2.transpose(0, 1, 3, 2)
        X = F[v, o].T + t1 @ self.Fae.T - self.Fmi.T @ t1
        X = X + es("imae,me->ia", s2, self.Fme)
        X = X + 2.0 * es("nf,nifa->ia", t1, self.oovv) - es("nf,naif->ia", t1, self.ovov)
        # (2 t2 - t2^T)_mief <ma|ef>,  <ma|ef> = s B[P,m,e] B[P,a,f]
        U = (self.Bov.reshape(-1, no * nv) @ s2.transpose(0, 2, 1, 3).reshape(no * nv, no * nv)).reshape(-1, no, nv)
        X = X + self.s * es("Pif,Paf->ia", U, self.Bvv)
        X = X - es("mnae,mnie->ia", t2, self.Looov)
        return X

    # ---- one pair of r2 -----------------------------------------------------------------------------------------
    def half(self, p, q):
        """The unsymmetrised half-residual block [a,b] of the ordered pair (p,q) (ccwfn.py:922-940)."""
        no, nv, s = self.no, self.nv, self.s
        t1, t2, tau = self.t1, self.t2, self.tau
        tpq = tau[p, q]
        H = 0.5 * self.oovv[p, q]                                                           # 922
        H = H + t2[p, q] @ self.Fae.T                                                       # 923
        H = H - 0.5 * t2[p, q] @ (t1.T @ self.Fme).T                                        # 924-925
        H = H - es("mab,m->ab", t2[p], self.Fmi[:, q])                               # 926
        H = H - 0.5 * es("mab,m->ab", t2[p], self.Fme @ t1[q])                       # 927-928
        Wmn = (self.oooo[:, :, p, q] + self.ooov[:, :, p, :] @ t1[q] + (self.ooov[:, :, q, :] @ t1[p]).T
               + es("ef,mnef->mn", tpq, self.oovv))                                  # 596-603
        H = H + 0.5 * es("mnab,mn->ab", tau, Wmn)                                    # 930
        M = self.Bvv @ tpq                                                                  # [P,a,f]
        H = H + 0.5 * s * np.tensordot(M, self.Bvv, axes=([0, 2], [0, 2]))                  # 931: tau_pqef <ab|ef>
        Z = s * np.tensordot(self.Bov @ tpq, self.Bvv, axes=([0, 2], [0, 2]))               # Z[m,b] = <mb|ef> tau_pqef
        H = H - t1.T @ Z                                                                    # 932
        Wej, Wje_q, Wje_p = self.Wmbej_col(q), self.Wmbje_col(q), self.Wmbje_col(p)
        tp = t2[p]                                                                          # [m,a,e]
        A = (tp - tp.transpose(0, 2, 1)).transpose(1, 0, 2).reshape(nv, no * nv)            # [a,(m,e)]
        H = H + A @ Wej.transpose(0, 2, 1).reshape(no * nv, nv)                             # 933
        A = tp.transpose(1, 0, 2).reshape(nv, no * nv)
        H = H + A @ (Wej + Wje_q).transpose(0, 2, 1).reshape(no * nv, nv)                   # 934
        A = t2[:, q].transpose(1, 0, 2).reshape(nv, no * nv)                                # t2[m,q,a,e] -> [a,(m,e)]
        H = H + A @ Wje_p.transpose(0, 2, 1).reshape(no * nv, nv)                           # 935
        H = H - t1.T @ es("e,meb->mb", t1[p], self.oovv[:, q])                       # 936-937
        H = H - (t1.T @ es("e,mae->ma", t1[p], self.ovov[:, :, q, :])).T             # 938
        x = es("e,Pea->Pa", t1[p], self.Bvv)                                         # <ab|ej> = s B[P,q,b] B[P,e,a]
        H = H + s * es("Pa,Pb->ab", x, self.Bov[:, q, :])                            # 939
        H = H - t1.T @ self.ooov[p, q]                                                      # 940
        return H

    def r2_block(self, i, j):
        """r2[i,j,:,:] after the symmetrisation of ccwfn.py:790."""
        return self.half(i, j) + self.half(j, i).T
