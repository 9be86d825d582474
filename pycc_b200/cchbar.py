"""Similarity-transformed Hamiltonian (HBAR) for closed-shell CCSD / CCD on B200 -- drop-in for ``pycc.cchbar``
(reference: pycc/cchbar.py:24-110 and the spatial-orbital ``build_*`` methods 128-823).  SURVEY 8(f) "next" #1.

Same eleven blocks with the reference's names and index orders (``Hov, Hvv, Hoo, Hoooo, Hvvvv, Hvovv, Hooov,
Hovvo, Hovov, Hvvvo, Hovoo``), same ``build_*`` signatures.  Underneath, each block is a TERM TABLE
``(alpha, subscripts, operand, operand)`` accumulated IN PLACE into one output tensor by the contraction backend
(``Contractor``: TTGT planning on strides -> ``b200cc_dgemm`` / ``b200cc_permute``), so no ``X = X + contract(...)``
temporaries exist, and every <pq|rs> / L_pqrs operand is a permuted VIEW of one of the six device-resident integral
blocks (``BlockHamiltonian.ERI`` / ``.L``), never an n^4 array.  The reference's ``tmp + tmp.swapaxes(0,1).swapaxes(2,3)``
pairs are written as the two contractions they are.

This is the correctness-first slice of the row (parity against the reference's golden vectors and the numpy oracle);
the products go through the generic planner, not yet through fused layouts as the CCSD residual does.  ``Hvvvv`` (v^4
doubles, 64.8 GB at v=300) is never needed as a tensor on the Lambda path and is materialised lazily, see EAGER.
CC2 / CC3 / spin-orbital branches of the reference are outside the accelerated path.
"""
from __future__ import annotations

import time

import torch

from . import kernels as K
from .utils import timing

F64 = torch.float64


def _view(H, name):
    """'E:vovv' -> <am|ef> view, 'L:vovv' -> 2<am|ef> - <am|fe> (materialised once per Hamiltonian)."""
    kind, pat = name.split(":")
    key = tuple(H.o if c == "o" else H.v for c in pat)
    return (H.ERI if kind == "E" else H.L)[key]


# block -> (initial value, terms for every model, extra terms when singles are present (CCSD))
# operands: t1, t2, tau (= t2 + t1 t1; t2 itself for CCD), Fov, previously built blocks, 'E:pqrs' / 'L:pqrs' integrals
_TABLE = {
    # H_me = f_me + t_nf L_mnef                                                  cchbar.py:147-153
    "Hov": ("F:ov", [], [(1.0, "nf,mnef->me", "t1", "L:oovv")]),
    # H_ae = f_ae - f_me t_ma + t_mf L_amef - tau_mnfa L_mnfe                      cchbar.py:201-210
    "Hvv": ("F:vv", [(-1.0, "mnfa,mnfe->ae", "tau", "L:oovv")],
            [(-1.0, "me,ma->ae", "Fov", "t1"), (1.0, "mf,amef->ae", "t1", "L:vovv")]),
    # H_mi = f_mi + t_ie f_me + t_ne L_mnie + tau_inef L_mnef                      cchbar.py:264-273
    "Hoo": ("F:oo", [(1.0, "inef,mnef->mi", "tau", "L:oovv")],
            [(1.0, "ie,me->mi", "t1", "Fov"), (1.0, "ne,mnie->mi", "t1", "L:ooov")]),
    # H_mnij = <mn|ij> + t_je <mn|ie> + t_ie <nm|je> + tau_ijef <mn|ef>            cchbar.py:330-339
    "Hoooo": ("E:oooo", [(1.0, "ijef,mnef->mnij", "tau", "E:oovv")],
              [(1.0, "je,mnie->mnij", "t1", "E:ooov"), (1.0, "ie,nmje->mnij", "t1", "E:ooov")]),
    # H_abef = <ab|ef> - t_mb <am|ef> - t_ma <bm|fe> + tau_mnab <mn|ef>            cchbar.py:394-403
    "Hvvvv": ("E:vvvv", [(1.0, "mnab,mnef->abef", "tau", "E:oovv")],
              [(-1.0, "mb,amef->abef", "t1", "E:vovv"), (-1.0, "ma,bmfe->abef", "t1", "E:vovv")]),
    # H_amef = <am|ef> - t_na <nm|ef>                                            cchbar.py:450-458
    "Hvovv": ("E:vovv", [], [(-1.0, "na,nmef->amef", "t1", "E:oovv")]),
    # H_mnie = <mn|ie> + t_if <nm|ef>                                            cchbar.py:500-508
    "Hooov": ("E:ooov", [], [(1.0, "if,nmef->mnie", "t1", "E:oovv")]),
    # H_mbej = <mb|ej> + t_jf <mb|ef> - t_nb <mn|ej> - tau_jnfb <mn|ef> + t2_njfb L_mnef     cchbar.py:554-569
    "Hovvo": ("E:ovvo", [(-1.0, "jnfb,mnef->mbej", "tau", "E:oovv"), (1.0, "njfb,mnef->mbej", "t2", "L:oovv")],
              [(1.0, "jf,mbef->mbej", "t1", "E:ovvv"), (-1.0, "nb,mnej->mbej", "t1", "E:oovo")]),
    # H_mbje = <mb|je> + t_jf <bm|ef> - t_nb <mn|je> - tau_jnfb <nm|ef>            cchbar.py:617-630
    "Hovov": ("E:ovov", [(-1.0, "jnfb,nmef->mbje", "tau", "E:oovv")],
              [(1.0, "jf,bmef->mbje", "t1", "E:vovv"), (-1.0, "nb,mnje->mbje", "t1", "E:ooov")]),
    # H_abei                                                                     cchbar.py:662-698
    "Hvvvo": ("E:vvvo", [(-1.0, "me,miab->abei", "Hov", "t2"), (1.0, "mnab,mnei->abei", "tau", "E:oovo"),
                         (-1.0, "imfa,bmfe->abei", "t2", "E:vovv"), (-1.0, "imfb,amef->abei", "t2", "E:vovv"),
                         (1.0, "mifb,amef->abei", "t2", "L:vovv")],
              # t_if H_abef without H_abef (64.8 GB at v=300): t_if <ab|ef> - t_mb (t_if <am|ef>) - t_ma (t_if <bm|fe>)
              # + tau_mnab (t_if <mn|ef>); the three small products are the single-term intermediates I:* below
              [(1.0, "if,abef->abei", "t1", "E:vvvv"), (-1.0, "mb,amei->abei", "t1", "I:amei"),
               (-1.0, "ma,bmei->abei", "t1", "I:bmei"), (1.0, "mnab,mnei->abei", "tau", "I:mnei"),
               (-1.0, "mb,amei->abei", "t1", "X:vovo"), (-1.0, "ma,bmie->abei", "t1", "X:voov")]),
    # H_mbij                                                                     cchbar.py:785-823
    "Hovoo": ("E:ovoo", [(1.0, "me,ijeb->mbij", "Hov", "t2"), (1.0, "ijef,mbef->mbij", "tau", "E:ovvv"),
                         (-1.0, "ineb,nmje->mbij", "t2", "E:ooov"), (-1.0, "jneb,mnie->mbij", "t2", "E:ooov"),
                         (1.0, "njeb,mnie->mbij", "t2", "L:ooov")],
              [(-1.0, "nb,mnij->mbij", "t1", "Hoooo"), (1.0, "je,mbie->mbij", "t1", "X:ovov"),
               (1.0, "ie,bmje->mbij", "t1", "X:voov2")]),
}
# t2-dressed integrals used only inside Hvvvo / Hovoo (the reference's `tmp`, cchbar.py:689-696, 812-820)
_DRESSED = {
    "X:vovo": ("E:vovo", [(-1.0, "infa,mnfe->amei", "t2", "E:oovv")]),
    "X:voov": ("E:voov", [(-1.0, "infb,mnef->bmie", "t2", "E:oovv"), (1.0, "nifb,mnef->bmie", "t2", "L:oovv")]),
    "X:ovov": ("E:ovov", [(-1.0, "infb,mnfe->mbie", "t2", "E:oovv")]),
    "X:voov2": ("E:voov", [(-1.0, "jnfb,mnef->bmje", "t2", "E:oovv"), (1.0, "njfb,mnef->bmje", "t2", "L:oovv")]),
}
# single-term intermediates of the factorised t1.Hvvvv (singles only)
_INTER = {
    "I:amei": (1.0, "if,amef->amei", "t1", "E:vovv"),
    "I:bmei": (1.0, "if,bmfe->bmei", "t1", "E:vovv"),
    "I:mnei": (1.0, "if,mnef->mnei", "t1", "E:oovv"),
}
ORDER = ("Hov", "Hvv", "Hoo", "Hoooo", "Hvvvv", "Hvovv", "Hooov", "Hovvo", "Hovov", "Hvvvo", "Hovoo")
# H_abef is v^4 doubles -- as large as <ab|ef> itself.  Nothing on the Lambda path needs it as a tensor (cclambda
# applies it to l2 as one ladder GEMM against <ab|ef> plus three o^3v^3 / o^4v^2 corrections; Hvvvo uses the same
# factorisation), so it is built only when somebody reads ``hbar.Hvvvv`` / calls ``build_Hvvvv``.
EAGER = tuple(k for k in ORDER if k != "Hvvvv")


class cchbar(object):
    """See module docstring.  ``cchbar(ccwfn)`` builds all blocks from the wavefunction's current amplitudes."""

    def __init__(self, ccwfn):
        t0 = time.time()
        if ccwfn.model not in ("CCSD", "CCD", "CCSD(T)"):
            raise NotImplementedError("HBAR is accelerated for the closed-shell CCD / CCSD amplitudes only")
        self.ccwfn = ccwfn
        self.contract = ccwfn.contract
        self.o, self.v = ccwfn.o, ccwfn.v
        self.no, self.nv = ccwfn.no, ccwfn.nv
        if getattr(ccwfn, "mixed", False):
            # the planes the CCSD iterations cached (33 GB at o=40,v=300) are of no use to HBAR / Lambda
            ccwfn.H.release_split_cache()
        # HBAR belongs to the amplitudes it was built from: keep a snapshot, not references -- solve_cc / update_amps
        # mutate ccwfn.t1 / t2 in place, and the lazily built Hvvvv and the Lambda ladder (which substitutes Hvvvv by
        # its definition in t1, t2) must see the same amplitudes as the eagerly built blocks
        self.t1 = K.permuted(ccwfn.t1, (0, 1))
        self.t2 = K.permuted(ccwfn.t2, (0, 1, 2, 3))
        blocks = self.build_all(ccwfn.H.F, self.t1, self.t2)
        for k in EAGER:
            setattr(self, k, blocks[k])
        self._Hvvvv = None
        self._amps = (K.permuted(ccwfn.H.F, (0, 1)), self.t1, self.t2)
        if not getattr(ccwfn, "quiet", False):
            print(timing("HBAR", time.time() - t0))

    @property
    def Hvvvv(self):
        """H_abef (cchbar.py:394-403), materialised on first access (v^4 doubles)."""
        if self._Hvvvv is None:
            if self.ccwfn.part.size > 1:
                raise NotImplementedError("H_abef is not materialised when <ab|ef> is sharded over ranks")
            self._Hvvvv = self._one("Hvvvv", *self._amps)
        return self._Hvvvv

    # ---- evaluation of the tables -----------------------------------------------------------------------
    def _env(self, F, t1, t2):
        w = self.ccwfn
        F = w._check_F(F)
        ccd = w.model == "CCD"
        t1, t2 = t1.contiguous(), t2.contiguous()
        env = {"t1": t1, "t2": t2, "F": F, "Fov": F[self.o, self.v]}
        env["tau"] = t2 if ccd else K.build_tau(t1, t2, 1.0, 1.0)
        return env

    def _operand(self, env, name):
        if name in env:
            return env[name]
        if name.startswith("I:"):
            alpha, sub, a, b = _INTER[name]
            return self.ccwfn._ct(sub, self._operand(env, a), self._operand(env, b), alpha=alpha)
        if name.startswith("X:"):
            init, terms = _DRESSED[name]
            x = K.permuted(_view(self.ccwfn.H, init), (0, 1, 2, 3))
            for alpha, sub, a, b in terms:
                self.ccwfn._ct(sub, self._operand(env, a), self._operand(env, b), out=x, alpha=alpha, beta=1.0)
            return x                       # not cached: used once
        return _view(self.ccwfn.H, name)

    def _build(self, key, env):
        init, terms, singles = _TABLE[key]
        w = self.ccwfn
        if init.startswith("F:"):
            sl = {"o": self.o, "v": self.v}
            out = K.permuted(env["F"][sl[init[2]], sl[init[3]]], (0, 1))
        else:
            src = _view(w.H, init)
            out = K.permuted(src, tuple(range(src.dim())))
        todo = list(terms) + ([] if w.model == "CCD" else list(singles))
        sharded = w.part.size > 1
        with K.mixed_mode(getattr(w, "mixed", False), cache=False):
            for alpha, sub, a, b in todo:
                if sharded and "E:vvvv" in (a, b):
                    if sub != "if,abef->abei" or b != "E:vvvv":
                        raise NotImplementedError("sharded <ab|ef> in HBAR term %r" % sub)
                    self._vvvv_term_sharded(out, alpha, sub, self._operand(env, a))
                    continue
                if "E:vvvv" in (a, b) and w._vvvv_released():
                    # precision='MP' / 'SP': only the TF32 planes of <ab|ef> are resident
                    if sub != "if,abef->abei" or b != "E:vvvv":
                        raise NotImplementedError("HBAR term %r needs the FP64 <ab|ef> block (precision='MP' released "
                                                  "it); read hbar.Hvvvv in double precision" % sub)
                    w._t1_vvvv(self._operand(env, a), out, alpha)
                    continue
                w._ct(sub, self._operand(env, a), self._operand(env, b), out=out, alpha=alpha, beta=1.0)
        env[key] = out
        return out

    def _vvvv_term_sharded(self, out, alpha, sub, t1):
        """A term whose integral operand is <ab|ef> when that block is sharded over the ranks (parallel.py): each rank
        contracts the pairs (a >= b) it holds -- giving the (a,b) and (b,a) elements -- the pieces are summed with one
        all-reduce and added to the (replicated) block.  Only 't_if <ab|ef> -> abei' (Hvvvo) occurs."""
        w = self.ccwfn
        a_lo, a_hi = w.part.a_range(w.nv)
        r_lo, r_hi = w.H.a_range
        if (r_lo, r_hi) != (a_lo, a_hi) and (r_lo, r_hi) != (0, w.nv):
            raise NotImplementedError("<ab|ef> rows resident on this rank are neither its share nor the whole block")
        piece = torch.zeros_like(out)
        if (r_lo, r_hi) == (a_lo, a_hi):
            w._t1_vvvv(t1, piece, alpha)                                     # this rank's pairs, from the packed form
            w.part.all_reduce_sum(piece)
        elif w.H.has("vvvv"):                                                # every rank holds the whole FP64 block
            w._ct(sub, t1, w.H.block("vvvv"), out=piece, alpha=alpha, beta=0.0)
        else:
            w._t1_vvvv(t1, piece, alpha)                                     # whole packed form on every rank
        K.strided_axpby(out, piece, 1.0, 1.0)

    # ---- the same tables for COMPLEX (F, t1, t2) carried as pairs of real planes (planes.py) -------------------------
    def _operand_p(self, env, name):
        from .planes import Planes, cprod, cterm, copy_real
        if name in env:
            return env[name]
        ct = self.ccwfn._ct
        if name.startswith("I:"):
            alpha, sub, a, b = _INTER[name]
            return cprod(ct, alpha, sub, self._operand_p(env, a), self._operand_p(env, b))
        if name.startswith("X:"):
            init, terms = _DRESSED[name]
            x = copy_real(_view(self.ccwfn.H, init))
            for alpha, sub, a, b in terms:
                cterm(ct, alpha, sub, self._operand_p(env, a), self._operand_p(env, b), x)
            return x
        return _view(self.ccwfn.H, name)                          # a real integral block

    def _build_p(self, key, env):
        from .planes import Planes, cterm, copy_real, copy_planes
        init, terms, singles = _TABLE[key]
        w = self.ccwfn
        if init.startswith("F:"):
            sl = {"o": self.o, "v": self.v}
            out = copy_planes(env["F"].view(lambda x: x[sl[init[2]], sl[init[3]]]))
        else:
            out = copy_real(_view(w.H, init))
        todo = list(terms) + ([] if w.model == "CCD" else list(singles))
        sharded = w.part.size > 1
        with K.mixed_mode(getattr(w, "mixed", False), cache=False):
            for alpha, sub, a, b in todo:
                if "E:vvvv" in (a, b):
                    # t_if <ab|ef> is linear in t1 and <ab|ef> is real: once per plane, through whichever form of the
                    # block this rank holds (sharded pairs + all-reduce, TF32 planes / packed form, or the full block)
                    if sub != "if,abef->abei" or b != "E:vvvv":
                        raise NotImplementedError("HBAR term %r with complex amplitudes" % sub)
                    t1p = self._operand_p(env, a)
                    for src, dst in ((t1p.re, out.re), (t1p.im, out.im)):
                        if src is None:
                            continue
                        if sharded:
                            self._vvvv_term_sharded(dst, alpha, sub, src)
                        elif w._vvvv_released():
                            w._t1_vvvv(src, dst, alpha)
                        else:
                            w._ct(sub, src, w.H.block("vvvv"), out=dst, alpha=alpha, beta=1.0)
                    continue
                cterm(w._ct, alpha, sub, self._operand_p(env, a), self._operand_p(env, b), out)
        env[key] = out
        return out

    def build_all_planes(self, F, t1, t2):
        """The eager blocks for complex (F, t1, t2) given as ``planes.Planes``: every "amplitude x integral" term runs
        once per plane on the real kernels (two products where five real samples of the whole build take five), the
        few amplitude x amplitude-dependent terms (all o^2v^3 or smaller) take four.  Returns a dict of Planes."""
        from .planes import Planes, complex_tau
        w = self.ccwfn
        env = {"t1": t1, "t2": t2, "F": F, "Fov": F.view(lambda x: x[self.o, self.v])}
        env["tau"] = t2 if w.model == "CCD" else complex_tau(t1, t2)
        return {k: self._build_p(k, env) for k in EAGER}

    def build_all(self, F, t1, t2, with_vvvv=False):
        """Every block in dependency order (Hvvvo needs Hov, Hovoo needs Hov and Hoooo); H_abef only on request."""
        env = self._env(F, t1, t2)
        return {k: self._build(k, env) for k in (ORDER if with_vvvv else EAGER)}

    # ---- the reference's builders (signatures of cchbar.py:128-823; integrals must be the wavefunction's own) ----
    def _one(self, key, F, t1, t2, **have):
        env = self._env(F, t1, t2)
        env.update({k: x for k, x in have.items() if x is not None})
        for dep in ("Hov", "Hoooo"):
            if dep not in env and any(dep in (a, b) for _, _, a, b in _TABLE[key][1] + _TABLE[key][2]):
                self._build(dep, env)
        return self._build(key, env)

    def build_Hov(self, o, v, F, L, t1):
        self.ccwfn._own(L=L)
        return self._one("Hov", F, t1, self.t2)

    def build_Hvv(self, o, v, F, L, t1, t2):
        self.ccwfn._own(L=L)
        return self._one("Hvv", F, t1, t2)

    def build_Hoo(self, o, v, F, L, t1, t2):
        self.ccwfn._own(L=L)
        return self._one("Hoo", F, t1, t2)

    def build_Hoooo(self, o, v, ERI, t1, t2):
        self.ccwfn._own(ERI)
        return self._one("Hoooo", self.ccwfn.H.F, t1, t2)

    def build_Hvvvv(self, o, v, ERI, t1, t2):
        self.ccwfn._own(ERI)
        return self._one("Hvvvv", self.ccwfn.H.F, t1, t2)

    def build_Hvovv(self, o, v, ERI, t1):
        self.ccwfn._own(ERI)
        return self._one("Hvovv", self.ccwfn.H.F, t1, self.t2)

    def build_Hooov(self, o, v, ERI, t1):
        self.ccwfn._own(ERI)
        return self._one("Hooov", self.ccwfn.H.F, t1, self.t2)

    def build_Hovvo(self, o, v, ERI, L, t1, t2):
        self.ccwfn._own(ERI, L)
        return self._one("Hovvo", self.ccwfn.H.F, t1, t2)

    def build_Hovov(self, o, v, ERI, t1, t2):
        self.ccwfn._own(ERI)
        return self._one("Hovov", self.ccwfn.H.F, t1, t2)

    def build_Hvvvo(self, o, v, ERI, L, Hov, Hvvvv, t1, t2):
        """``Hvvvv`` is accepted for signature compatibility (cchbar.py:632) and not read: t1.Hvvvv is evaluated from
        <ab|ef> directly."""
        self.ccwfn._own(ERI, L)
        return self._one("Hvvvo", self.ccwfn.H.F, t1, t2, Hov=Hov)

    def build_Hovoo(self, o, v, ERI, L, Hov, Hoooo, t1, t2):
        self.ccwfn._own(ERI, L)
        return self._one("Hovoo", self.ccwfn.H.F, t1, t2, Hov=Hov, Hoooo=Hoooo)
