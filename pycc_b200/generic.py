"""The CCSD / CCD residual for ARBITRARY integral objects, through the contraction seam.

The fused residual of ``ccwfn.CCwfn`` works on the six unique blocks of the wavefunction's own, 8-fold-symmetric
integrals.  Some callers of the reference swap other integrals in: ``CCderiv`` temporarily replaces ``cc.H.ERI`` /
``cc.H.L`` by perturbed (derivative) integrals and calls ``cc.residuals(df, cc.t1, cc.t2)`` (pycc/ccderiv.py:250-259),
and the public ``build_*`` / ``r_T1`` / ``r_T2`` methods take ``ERI`` / ``L`` as arguments.  Perturbed integrals have no
block-storage contract (no 8-fold symmetry is assumed, they may be host numpy arrays), so for them the equations are
evaluated term by term exactly as the reference writes them (pycc/ccwfn.py:458-944), every contraction going through
``ContractionBackend.__call__`` -- i.e. the DMMA GEMM / permutation kernels, operands uploaded per call as the reference
does (pycc/device.py:70-74).  Slower than the fused path (nine o^3v^3 products, permuted copies), but general.

``ERI`` / ``L`` must support ``X[o, v, v, v]``-style slicing with the wavefunction's slices (numpy arrays, torch tensors,
or block views).  Term tables: (factor, subscripts, left operand, right operand); operand names starting with ``E:`` /
``L:`` are integral slices.
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K

F64 = torch.float64

# F_ae (ccwfn.py:494-497), F_mi (530-533), F_me (563-564); CCD rows: 491-492, 527-528
_FAE = [(-0.5, "me,ma->ae", "Fov", "t1"), (1.0, "mf,mafe->ae", "t1", "L:ovvv"), (-1.0, "mnaf,mnef->ae", "tau_h", "L:oovv")]
_FMI = [(0.5, "ie,me->mi", "t1", "Fov"), (1.0, "ne,mnie->mi", "t1", "L:ooov"), (1.0, "inef,mnef->mi", "tau_h", "L:oovv")]
_FME = [(1.0, "nf,mnef->me", "t1", "L:oovv")]
# W_mnij (596-603), W_mbej (641-645), W_mbje (680-683), Z_mbij (715)
_WMNIJ = [(1.0, "je,mnie->mnij", "t1", "E:ooov"), (1.0, "ie,mnej->mnij", "t1", "E:oovo"),
          (1.0, "ijef,mnef->mnij", "tau", "E:oovv")]
_WMBEJ = [(1.0, "jf,mbef->mbej", "t1", "E:ovvv"), (-1.0, "nb,mnej->mbej", "t1", "E:oovo"),
          (-1.0, "jnfb,mnef->mbej", "tau_t", "E:oovv"), (0.5, "njfb,mnef->mbej", "t2", "L:oovv")]
_WMBJE = [(-1.0, "jf,mbfe->mbje", "t1", "E:ovvv"), (1.0, "nb,mnje->mbje", "t1", "E:ooov"),
          (1.0, "jnfb,mnfe->mbje", "tau_t", "E:oovv")]
# r1 (754-760)
_R1 = [(1.0, "ie,ae->ia", "t1", "Fae"), (-1.0, "mi,ma->ia", "Fmi", "t1"), (1.0, "imae,me->ia", "s2", "Fme"),
       (1.0, "nf,nafi->ia", "t1", "L:ovvo"), (1.0, "mief,maef->ia", "s2", "E:ovvv"), (-1.0, "mnae,nmei->ia", "t2", "L:oovo")]
# r2, unsymmetrised (922-940); the two (t1 x t1) terms (936-938) as written, through the t1 t1 product
_R2 = [(1.0, "ijae,be->ijab", "t2", "Fae"), (-0.5, "ijae,be->ijab", "t2", "X_be"), (-1.0, "imab,mj->ijab", "t2", "Fmi"),
       (-0.5, "imab,mj->ijab", "t2", "X_mj"), (0.5, "mnab,mnij->ijab", "tau", "Wmnij"),
       (0.5, "ijef,abef->ijab", "tau", "E:vvvv"), (-1.0, "ma,mbij->ijab", "t1", "Zmbij"),
       (1.0, "imae,mbej->ijab", "d2", "Wmbej"), (1.0, "imae,mbej->ijab", "t2", "Wsum"),
       (1.0, "mjae,mbie->ijab", "t2", "Wmbje"), (-1.0, "imea,mbej->ijab", "tt", "E:ovvo"),
       (-1.0, "imeb,maje->ijab", "tt", "E:ovov"), (1.0, "ie,abej->ijab", "t1", "E:vvvo"),
       (-1.0, "ma,mbij->ijab", "t1", "E:ovoo")]
_SINGLES_ONLY = {"t1", "Fov", "Fme", "X_be", "X_mj", "Zmbij", "tt"}


class GenericResidual:
    """Evaluates the tables above for one wavefunction (sizes, slices, contraction backend) and one (ERI, L) pair."""

    def __init__(self, w, ERI, L):
        self.w, self.ERI, self.L = w, ERI, L
        self.ct = w.contract                      # ContractionBackend: uploads host operands, runs the GEMM kernels
        self.sl = {"o": w.o, "v": w.v}
        self.ccd = w.model == "CCD"

    def _dev(self, x):
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        return x.to(self.w.device1, dtype=F64)

    def _ints(self, name):
        src = self.ERI if name[0] == "E" else self.L
        return src[tuple(self.sl[c] for c in name[2:])]

    def _env(self, F, t1, t2):
        w = self.w
        F = w._check_F(F)
        t1, t2 = self._dev(t1).contiguous(), self._dev(t2).contiguous()
        env = {"F": F, "Fov": F[w.o, w.v], "t1": t1, "t2": t2}
        z = 0.0 if self.ccd else 1.0
        env["tau"] = K.build_tau(t1, t2, 1.0, z)
        env["tau_h"] = K.build_tau(t1, t2, 1.0, 0.5 * z)
        env["tau_t"] = K.build_tau(t1, t2, 0.5, z)
        s2 = K.permuted(t2, (0, 1, 2, 3), 2.0)                    # 2 t2 - t2^T(ab)
        env["s2"] = K.strided_axpby(s2, t2.permute(0, 1, 3, 2), -1.0, 1.0)
        d2 = K.permuted(t2, (0, 1, 2, 3))                         # t2 - t2^T(ab)
        env["d2"] = K.strided_axpby(d2, t2.permute(0, 1, 3, 2), -1.0, 1.0)
        return env

    def _sum(self, out, terms, env):
        for alpha, sub, a, b in terms:
            if self.ccd and (a in _SINGLES_ONLY or b in _SINGLES_ONLY):
                continue
            A = env[a] if a in env else self._ints(a)
            B = env[b] if b in env else self._ints(b)
            self.ct(sub, A, B, out=out, alpha=alpha, beta=1.0)
        return out

    def _start(self, x):
        x = self._dev(x)
        return K.permuted(x, tuple(range(x.dim())))

    # ---- intermediates with the reference's layouts --------------------------------------------------------------
    def intermediates(self, F, t1, t2):
        w, env = self.w, self._env(F, t1, t2)
        F = env["F"]
        env["Fae"] = self._sum(self._start(F[w.v, w.v]), _FAE, env)
        env["Fmi"] = self._sum(self._start(F[w.o, w.o]), _FMI, env)
        if not self.ccd:
            env["Fme"] = self._sum(self._start(F[w.o, w.v]), _FME, env)
        env["Wmnij"] = self._sum(self._start(self._ints("E:oooo")), _WMNIJ, env)
        env["Wmbej"] = self._sum(self._start(self._ints("E:ovvo")), _WMBEJ, env)
        Wmbje = self._start(self._ints("E:ovov"))
        Wmbje = K.strided_axpby(Wmbje, Wmbje, -1.0, 0.0)
        env["Wmbje"] = self._sum(Wmbje, _WMBJE, env)
        if not self.ccd:
            env["Zmbij"] = self.ct("mbef,ijef->mbij", self._ints("E:ovvv"), env["tau"])
        return env

    def residuals(self, F, t1, t2):
        """(r1, unsymmetrised half of r2)"""
        w = self.w
        env = self.intermediates(F, t1, t2)
        F, t1 = env["F"], env["t1"]
        r1 = K.permuted(F[w.v, w.o], (1, 0))
        if self.ccd:
            r1.zero_()
        else:
            self._sum(r1, _R1, env)
            env["X_be"] = self.ct("mb,me->be", t1, env["Fme"])
            env["X_mj"] = self.ct("je,me->mj", t1, env["Fme"])
            env["tt"] = self.ct("ie,ma->imea", t1, t1)                              # ccwfn.py:936
        Wsum = K.permuted(env["Wmbej"], (0, 1, 2, 3))
        env["Wsum"] = K.strided_axpby(Wsum, env["Wmbje"].permute(0, 1, 3, 2), 1.0, 1.0)     # Wmbej + Wmbje^T (934)
        vvoo = self._dev(self._ints("E:vvoo")).permute(2, 3, 0, 1)                   # 1/2 <ab|ij> as [i,j,a,b]   (922)
        half = K.strided_axpby(torch.empty(tuple(vvoo.shape), dtype=F64, device=w.device1), vvoo, 0.5, 0.0)
        self._sum(half, _R2, env)
        return r1, half
