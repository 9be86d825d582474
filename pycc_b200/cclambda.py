"""Closed-shell CCSD / CCD / CCSD(T) Lambda-amplitude solver on B200 -- drop-in for ``pycc.cclambda`` on the spatial-orbital
path (reference: pycc/cclambda.py:27-66 ctor/guess, 69-200 solve_lambda, 202-256 residuals, 258-306 Goo/Gvv,
308-370 r_L1, 408-497 r_L2, 547-570 pseudoenergy).  SURVEY 8(f) "next" #1.

Same surface: ``cclambda(ccwfn, hbar).solve_lambda(e_conv, r_conv, maxiter, max_diis, start_diis)``, ``residuals``,
``build_Goo``, ``build_Gvv``, ``r_L1``, ``r_L2``, ``pseudoenergy``, attributes ``l1``, ``l2``.  The loop skeleton is
the one of ``solve_cc`` and reuses its kernels: residual terms accumulate in place through the contraction backend
(term tables, see cchbar.py), the r2 symmetrisation + Jacobi update + sum((r/D)^2) is the ONE fused pass
``b200cc_update_amps`` (cclambda.py:496 + 163-168), the pseudo-energy is a ``b200cc_multi_dot`` and DIIS is the shared
``helper_diis``.  One device->host sync per iteration.
"""
from __future__ import annotations

import time

import torch

from . import kernels as K
from .exceptions import PyCCError
from .utils import helper_diis, title, iteration, converged

F64 = torch.float64

# r_l1_ia = 2 H_ia + l1_ie H_ea - l1_ma H_im + l2_imef H_efam - l2_mnae H_iemn + l1_me (2 H_ieam - H_iema)
#           - 2 G_ef H_eifa + G_ef H_eiaf - 2 G_mn H_mina + G_mn H_imna                      cclambda.py:344-370
_R1 = [(1.0, "ie,ea->ia", "l1", "Hvv"), (-1.0, "ma,im->ia", "l1", "Hoo"),
       (1.0, "imef,efam->ia", "l2", "Hvvvo"), (-1.0, "mnae,iemn->ia", "l2", "Hovoo"),
       (1.0, "me,ieam->ia", "l1", "W"),
       (-2.0, "ef,eifa->ia", "Gvv", "Hvovv"), (1.0, "ef,eiaf->ia", "Gvv", "Hvovv"),
       (-2.0, "mn,mina->ia", "Goo", "Hooov"), (1.0, "mn,imna->ia", "Goo", "Hooov")]
# unsymmetrised half of r_l2 (cclambda.py:447-495); the l1 terms only with singles
_R2_SINGLES = [(2.0, "ia,jb->ijab", "l1", "Hov"), (-1.0, "ja,ib->ijab", "l1", "Hov"),
               (2.0, "ie,ejab->ijab", "l1", "Hvovv"), (-1.0, "ie,ejba->ijab", "l1", "Hvovv"),
               (-2.0, "mb,jima->ijab", "l1", "Hooov"), (1.0, "mb,ijma->ijab", "l1", "Hooov")]
_R2 = [(1.0, "ijeb,ea->ijab", "l2", "Hvv"), (-1.0, "mjab,im->ijab", "l2", "Hoo"),
       (0.5, "mnab,ijmn->ijab", "l2", "Hoooo"),
       (1.0, "mjeb,ieam->ijab", "l2", "W"), (-1.0, "mibe,jema->ijab", "l2", "Hovov"),
       (-1.0, "mieb,jeam->ijab", "l2", "Hovvo"),
       (1.0, "ae,ijeb->ijab", "Gvv", "Loovv"), (-1.0, "mi,mjab->ijab", "Goo", "Loovv")]


class cclambda(object):
    def __init__(self, ccwfn, hbar):
        if ccwfn.model not in ("CCSD", "CCD", "CCSD(T)"):
            raise NotImplementedError("the Lambda equations are accelerated for closed-shell CCD / CCSD / CCSD(T); "
                                      "CC2/CC3 stay with the reference implementation")
        if ccwfn.model == "CCSD(T)" and not (hasattr(ccwfn, "S1") and hasattr(ccwfn, "S2")):
            raise PyCCError("CCSD(T) Lambda needs the (T) sources S1/S2: solve the amplitudes with "
                            "make_t3_density=True (or call ccwfn.t3_density()) first")
        self.ccwfn, self.hbar = ccwfn, hbar
        self.contract = ccwfn.contract
        # the amplitudes HBAR was built from (its snapshot), not whatever the wavefunction holds by now
        t1, t2 = getattr(hbar, "t1", ccwfn.t1), getattr(hbar, "t2", ccwfn.t2)
        self.t1, self.t2 = t1, t2
        # l1 = 2 t1, l2 = 2 (2 t2 - t2^T)                                              cclambda.py:65-66
        self.l1 = torch.empty_like(t1)
        K.strided_axpby(self.l1, t1, 2.0, 0.0)
        self.l2 = torch.empty_like(t2)
        K.strided_axpby(self.l2, t2, 4.0, 0.0)
        K.strided_axpby(self.l2, t2.permute(0, 1, 3, 2), -2.0, 1.0)

    # ---- building blocks with the reference's signatures ----------------------------------------------------
    def build_Goo(self, t2, l2):
        return self.ccwfn._ct("mjab,ijab->mi", t2.contiguous(), l2.contiguous())            # cclambda.py:281

    def build_Gvv(self, t2, l2):
        return self.ccwfn._ct("ijeb,ijab->ae", t2.contiguous(), l2.contiguous(), alpha=-1.0)  # cclambda.py:306

    @staticmethod
    def _w(Hovvo, Hovov):
        """2 H_ieam - H_iema (used by r_L1 and r_L2; constant during a solve)"""
        W = K.permuted(Hovvo, (0, 1, 2, 3), 2.0)
        return K.strided_axpby(W, Hovov.permute(0, 1, 3, 2), -1.0, 1.0)

    def _accumulate(self, out, terms, env):
        ct = self.ccwfn._ct
        with K.mixed_mode(getattr(self.ccwfn, "mixed", False), cache=False):
            for alpha, sub, a, b in terms:
                ct(sub, env[a], env[b], out=out, alpha=alpha, beta=1.0)
        return out

    def r_L1(self, o, v, l1, l2, Hov, Hvv, Hoo, Hovvo, Hovov, Hvvvo, Hovoo, Hvovv, Hooov, Gvv, Goo, s1=None, W=None):
        if self.ccwfn.model == "CCD":
            return torch.zeros_like(l1)
        env = dict(l1=l1.contiguous(), l2=l2.contiguous(), Hvv=Hvv, Hoo=Hoo, Hvvvo=Hvvvo, Hovoo=Hovoo, Hvovv=Hvovv,
                   Hooov=Hooov, Gvv=Gvv, Goo=Goo, W=W if W is not None else self._w(Hovvo, Hovov))
        out = K.permuted(Hov, (0, 1), 2.0)
        if s1 is not None:                                   # (T) source cc.S1          cclambda.py:350-352
            K.strided_axpby(out, s1, 1.0, 1.0)
        return self._accumulate(out, _R1, env)

    def _ladder(self, half, l2, Hvvvv=None, t1=None, t2=None, symmetric=False):
        """half += 1/2 l2_ijef H_efab (cclambda.py:468) WITHOUT the v^4 tensor H_efab:
             1/2 l2_ijef <ef|ab>                      the ladder GEMM of the CCSD residual on <ab|ef> in place
           - 1/2 (l2_ijef t_mf) <em|ab> - 1/2 (l2_ijef t_me) <fm|ba>                     two o^3v^3 products
           + 1/2 (l2_ijef tau_mnef) <mn|ab>                                              two o^4v^2 products
        (cchbar.py:394-403 substituted).  A caller that hands in a materialised ``Hvvvv`` gets the literal term.
        ``t1, t2``: the amplitudes HBAR is built from (default: HBAR's own snapshot)."""
        w, ct = self.ccwfn, self.ccwfn._ct
        t1 = self.t1 if t1 is None else t1
        t2 = self.t2 if t2 is None else t2
        if Hvvvv is not None:
            return ct("ijef,efab->ijab", l2, Hvvvv, out=half, alpha=0.5, beta=1.0)
        with K.mixed_mode(getattr(w, "mixed", False), cache=False):
            if w.part.size > 1:
                # <ab|ef> is a-sharded: every rank adds the rows it holds, one all-reduce (o^2v^2) sums the pieces;
                # all other terms are replicated, so l1 / l2 stay identical on all ranks
                piece = torch.zeros_like(half)
                w._ladder(l2, piece, symmetric=symmetric)
                w.part.all_reduce_sum(piece)
                K.strided_axpby(half, piece, 1.0, 1.0)
            else:
                w._ladder(l2, half, symmetric=symmetric)
            o, v = w.o, w.v
            oovv = w.H.ERI[o, o, v, v]
            if w.model == "CCD":
                tau = t2.contiguous()
            else:
                t1 = t1.contiguous()
                vovv = w.H.ERI[v, o, v, v]
                if symmetric:
                    # l2[i,j,e,f] = l2[j,i,f,e]: the second product is the first with (i,a) <-> (j,b) exchanged, and the
                    # caller adds half^T to half -- ONE o^3v^3 GEMM with twice the weight gives the same symmetrised sum
                    ct("ijem,emab->ijab", ct("ijef,mf->ijem", l2, t1), vovv, out=half, alpha=-1.0, beta=1.0)
                else:
                    ct("ijem,emab->ijab", ct("ijef,mf->ijem", l2, t1), vovv, out=half, alpha=-0.5, beta=1.0)
                    ct("ijfm,fmba->ijab", ct("ijef,me->ijfm", l2, t1), vovv, out=half, alpha=-0.5, beta=1.0)
                tau = K.build_tau(t1, t2.contiguous(), 1.0, 1.0)
            ct("ijmn,mnab->ijab", ct("ijef,mnef->ijmn", l2, tau), oovv, out=half, alpha=0.5, beta=1.0)
        return half

    def _r_L2_half(self, l1, l2, Hov, Hvv, Hoo, Hoooo, Hvvvv, Hovvo, Hovov, Hvovv, Hooov, Gvv, Goo, W, s2=None,
                   t1=None, t2=None, symmetric=False):
        Loovv = self.ccwfn.H.derived("Loovv")
        l2 = l2.contiguous()
        env = dict(l1=l1.contiguous(), l2=l2, Hov=Hov, Hvv=Hvv, Hoo=Hoo, Hoooo=Hoooo,
                   Hovvo=Hovvo, Hovov=Hovov, Hvovv=Hvovv, Hooov=Hooov, Gvv=Gvv, Goo=Goo, W=W, Loovv=Loovv)
        terms = _R2 if self.ccwfn.model == "CCD" else _R2_SINGLES + _R2
        half = K.permuted(Loovv, (0, 1, 2, 3))
        if s2 is not None:                                   # (T) source: + 1/2 cc.S2 before P_ij^ab   cclambda.py:470-474
            K.strided_axpby(half, s2, 0.5, 1.0)
        self._accumulate(half, terms, env)
        return self._ladder(half, l2, Hvvvv, t1, t2, symmetric=symmetric)

    def r_L2(self, o, v, l1, l2, L, Hov, Hvv, Hoo, Hoooo, Hvvvv, Hovvo, Hovov, Hvvvo, Hovoo, Hvovv, Hooov, Gvv, Goo,
             s2=None):
        self.ccwfn._own(L=L)
        half = self._r_L2_half(l1, l2, Hov, Hvv, Hoo, Hoooo, Hvvvv, Hovvo, Hovov, Hvovv, Hooov, Gvv, Goo,
                               self._w(Hovvo, Hovov), s2=s2)
        return K.symmetrize_r2(half)                                                       # cclambda.py:496

    def _sources(self):
        """(S1, S2) of a CCSD(T) wavefunction (cclambda.py:145-146), else (None, None)"""
        w = self.ccwfn
        return (w.S1, w.S2) if w.model == "CCSD(T)" else (None, None)

    def pseudoenergy(self, o, v, ERI, l2):
        """1/2 <ij|ab> l2_ijab as a 0-d device tensor                                       cclambda.py:570"""
        self.ccwfn._own(ERI)
        oovv = self.ccwfn.H.block("oovv")
        return 0.5 * K.multi_dot(oovv.reshape(-1), [l2.contiguous().reshape(-1)])[0]

    def residuals(self, F, t1, t2, l1, l2):
        """(r1, r2) with HBAR rebuilt from (F, t1, t2)                                      cclambda.py:202-256
        Complex amplitudes / a complex Hermitian F (the Lambda half of rtcc.f, rt/rtcc.py:143-147): HBAR and the residual
        are carried as pairs of real planes through the same term tables (planes.py, ``_residuals_planes``);
        ``complex_on_planes = False`` selects the round-1 evaluation from five real samples of the whole residual
        (utils.complex_from_real_samples; the residual is exactly quartic in the joint scaling)."""
        from .utils import complex_from_real_samples, is_complex
        if any(is_complex(x) for x in (F, t1, t2, l1, l2)):
            if self.complex_on_planes:
                return self._residuals_planes(F, t1, t2, l1, l2)
            return tuple(complex_from_real_samples(self.residuals, (F, t1, t2, l1, l2), self.l2.device))
        hb = self.hbar.build_all(F, t1, t2)
        Goo, Gvv = self.build_Goo(t2, l2), self.build_Gvv(t2, l2)
        W = self._w(hb["Hovvo"], hb["Hovov"])
        # as in the reference, this entry point never adds the (T) sources (only solve_lambda does, cclambda.py:145-148)
        r1 = self.r_L1(self.ccwfn.o, self.ccwfn.v, l1, l2, hb["Hov"], hb["Hvv"], hb["Hoo"], hb["Hovvo"], hb["Hovov"],
                       hb["Hvvvo"], hb["Hovoo"], hb["Hvovv"], hb["Hooov"], Gvv, Goo, W=W)
        half = self._r_L2_half(l1, l2, hb["Hov"], hb["Hvv"], hb["Hoo"], hb["Hoooo"], None, hb["Hovvo"],
                               hb["Hovov"], hb["Hvovv"], hb["Hooov"], Gvv, Goo, W, t1=t1, t2=t2)
        return r1, K.symmetrize_r2(half)

    complex_on_planes = True

    def _residuals_planes(self, F, t1, t2, l1, l2):
        """cclambda.py:202-256 for complex arguments on pairs of real planes.  Cost in real GEMMs: every amplitude x
        integral product of HBAR (incl. the three o^2v^4 terms of H_abei) and the lambda2 ladder twice (once per plane),
        the three o^3v^3 products lambda2 x {W, H_ovov, H_ovvo} three times (3M), everything smaller four times --
        against five evaluations of everything when the residual is sampled."""
        from .planes import Planes, cterm, cprod, copy_real, copy_planes
        from .utils import real_planes, planes_to_complex
        w, ct = self.ccwfn, self.ccwfn._ct
        dev = self.l2.device
        Fp, t1p, t2p, l1p, l2p = (Planes(*real_planes(x, dev)) for x in (F, t1, t2, l1, l2))
        hb = self.hbar.build_all_planes(Fp, t1p, t2p)
        Goo = cprod(ct, 1.0, "mjab,ijab->mi", t2p, l2p)                                     # cclambda.py:281
        Gvv = cprod(ct, -1.0, "ijeb,ijab->ae", t2p, l2p)                                    # cclambda.py:306
        W = Planes(self._w(hb["Hovvo"].re, hb["Hovov"].re), self._w(hb["Hovvo"].im, hb["Hovov"].im))
        env = dict(hb)
        env.update(l1=l1p, l2=l2p, Gvv=Gvv, Goo=Goo, W=W, Loovv=w.H.derived("Loovv"))
        ccd = w.model == "CCD"
        with K.mixed_mode(getattr(w, "mixed", False), cache=False):
            r1 = copy_planes(hb["Hov"], 2.0).full()
            if ccd:
                r1 = Planes(torch.zeros_like(r1.re), torch.zeros_like(r1.re))
            else:
                for alpha, sub, a, b in _R1:
                    cterm(ct, alpha, sub, env[a], env[b], r1)
            # H_abei / H_mbij (both planes: 2 x o v^3 doubles) are read by r_L1 only
            for k in ("Hvvvo", "Hovoo"):
                hb.pop(k, None)
                env.pop(k, None)
            half = copy_real(env["Loovv"])
            for alpha, sub, a, b in (_R2 if ccd else _R2_SINGLES + _R2):
                cterm(ct, alpha, sub, env[a], env[b], half)
        self._ladder_planes(half, l2p, t1p, t2p)
        K.symmetrize_r2(half.re)
        K.symmetrize_r2(half.im)
        return planes_to_complex(r1.re, r1.im), planes_to_complex(half.re, half.im)

    def _ladder_planes(self, half, l2, t1, t2):
        """``_ladder`` for Planes: the bare ladder once per plane of lambda2 (<ab|ef> is real), the H_abef corrections
        with their small complex x complex first factors."""
        from .planes import Planes, cterm, cprod, complex_tau
        w, ct = self.ccwfn, self.ccwfn._ct
        with K.mixed_mode(getattr(w, "mixed", False), cache=False):
            for src, dst in ((l2.re, half.re), (l2.im, half.im)):
                if src is None:
                    continue
                if w.part.size > 1:
                    piece = torch.zeros_like(dst)
                    w._ladder(src, piece)
                    w.part.all_reduce_sum(piece)
                    K.strided_axpby(dst, piece, 1.0, 1.0)
                else:
                    w._ladder(src, dst)
            o, v = w.o, w.v
            oovv = w.H.ERI[o, o, v, v]
            if w.model == "CCD":
                tau = t2
            else:
                vovv = w.H.ERI[v, o, v, v]
                cterm(ct, -0.5, "ijem,emab->ijab", cprod(ct, 1.0, "ijef,mf->ijem", l2, t1), vovv, half)
                cterm(ct, -0.5, "ijfm,fmba->ijab", cprod(ct, 1.0, "ijef,me->ijfm", l2, t1), vovv, half)
                tau = complex_tau(t1, t2)
            cterm(ct, 0.5, "ijmn,mnab->ijab", cprod(ct, 1.0, "ijef,mnef->ijmn", l2, tau), oovv, half)
        return half

    # ---- solve_lambda (cclambda.py:69-200) -------------------------------------------------------------------
    def solve_lambda(self, e_conv=1e-7, r_conv=1e-7, maxiter=100, max_diis=8, start_diis=1):
        t0 = time.time()
        w, hb = self.ccwfn, self.hbar
        o, v = w.o, w.v
        say = (lambda *a: None) if getattr(w, "quiet", False) else print
        lecc = float(self.pseudoenergy(o, v, w.H.ERI, self.l2))
        name = "Lambda-amplitudes (%s)" % w.model
        say(title(name))
        say(iteration(0, energy=lecc, de=-lecc, e_label="LCC PseudoE"))
        diis = helper_diis(self.l1, self.l2, max_diis, w.precision)
        W = self._w(hb.Hovvo, hb.Hovov)
        s1, s2 = self._sources()
        self.trace = []
        for niter in range(1, maxiter + 1):
            last = lecc
            Goo, Gvv = self.build_Goo(self.t2, self.l2), self.build_Gvv(self.t2, self.l2)
            r1 = self.r_L1(o, v, self.l1, self.l2, hb.Hov, hb.Hvv, hb.Hoo, hb.Hovvo, hb.Hovov, hb.Hvvvo, hb.Hovoo,
                           hb.Hvovv, hb.Hooov, Gvv, Goo, s1=s1, W=W)
            half = self._r_L2_half(self.l1, self.l2, hb.Hov, hb.Hvv, hb.Hoo, hb.Hoooo, None, hb.Hovvo, hb.Hovov,
                                   hb.Hvovv, hb.Hooov, Gvv, Goo, W, s2=s2, symmetric=True)   # l2[i,j,a,b] = l2[j,i,b,a]
            # r2 = half + half^T, l += r/D, sum (r/D)^2 in one pass; then the pseudo-energy
            ssq = K.update_amps(r1, half, w.eps_o, w.eps_v, self.l1, self.l2, symmetrize=True, write_r2=False)
            e_dev = self.pseudoenergy(o, v, w.H.ERI, self.l2)
            ssq_h, lecc = torch.stack((ssq[0], e_dev)).tolist()
            rms = ssq_h ** 0.5
            self.trace.append((lecc, rms))
            say(iteration(niter, energy=lecc, de=lecc - last, rms=rms, e_label="LCC PseudoE"))
            if abs(lecc - last) < e_conv and abs(rms) < r_conv:
                say(converged(name, time.time() - t0))
                return torch.tensor(lecc, dtype=F64, device=self.l2.device)
            diis.add_error_vector(self.l1, self.l2)
            if niter >= start_diis:
                self.l1, self.l2 = diis.extrapolate(self.l1, self.l2)
        return None          # the reference falls off the loop (cclambda.py:69-200)
