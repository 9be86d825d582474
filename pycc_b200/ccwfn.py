"""Closed-shell RHF-CCSD amplitude solver on B200: drop-in for ``pycc.ccwfn`` on the CCSD / CCSD(T)
energy path (reference: pycc/ccwfn.py:76-372, 432-944, 1122-1162).

Same public surface -- ``ccwfn(scf_wfn, model=..., device='GPU').solve_cc(e_conv, r_conv, maxiter,
max_diis, start_diis)``, ``residuals(F, t1, t2)``, ``build_tau/Fae/Fmi/Fme/Wmnij/Wmbej/Wmbje/Zmbij``,
``r_T1``, ``r_T2``, ``cc_energy`` and the attributes downstream pycc modules read
(``t1, t2, Dia, Dijab, H, o, v, no, nv, contract, ecc ...``) -- but a different machine underneath:

* integrals are six device-resident blocks (hamiltonian.BlockHamiltonian), never the n^4 arrays;
* every contraction is a ``b200cc_dgemm`` (FP64 DMMA) launch; the ring/ladder algebra is laid out so
  the 64.8 GB <ab|ef> and the 8.6 GB <mb|ef> blocks are only ever *streamed in place* (no permuted
  copies of them exist), and only o^2v^2-sized tensors are permuted (HBM-bound, <1 % of a step);
* the two (t1 x t1) x <ov|vo> terms are factorised to o^3v^2 work (7 instead of 9 o^3v^3 GEMMs);
* tau / r2-symmetrisation / Jacobi update / rms / energy are fused elementwise kernels;
* nothing on this path runs on the CPU or through torch math: without libb200cc.so and a CUDA
  device every call raises.

Index conventions follow the reference: t1[i,a], t2[i,j,a,b]; intermediates returned by the public
``build_*`` methods have the reference's layouts (Wmbej[m,b,e,j], Wmbje[m,b,j,e], Zmbij[m,b,i,j]).
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import kernels as K
from ._lib import B200ccError
from .contract import Contractor
from .device import DeviceManager
from .exceptions import InvalidKeywordError, PyCCError
from .hamiltonian import BlockHamiltonian, _BlockView
from .utils import helper_diis, title, iteration, converged, timing, solve_params
from .wavefunction import resolve_reference
from .parallel import Serial

F64 = torch.float64


def _is_complex(x):
    return x.is_complex() if isinstance(x, torch.Tensor) else np.iscomplexobj(x)


class CCwfn(object):
    """See module docstring.  Constructor mirrors pycc/ccwfn.py:76-213 for the closed-shell path."""

    VALID_MODELS = ['CCD', 'CC2', 'CCSD', 'CCSD(T)', 'CC3']
    SUPPORTED_MODELS = ['CCD', 'CC2', 'CCSD', 'CCSD(T)', 'CC3']

    def __init__(self, scf_wfn, **kwargs):
        t0 = time.time()
        model = kwargs.pop('model', 'CCSD').upper()
        if model not in self.VALID_MODELS:
            raise InvalidKeywordError('model', model, self.VALID_MODELS)
        if model not in self.SUPPORTED_MODELS:
            raise NotImplementedError("pycc_b200 accelerates the closed-shell CCD/CC2/CCSD/CCSD(T)/CC3 energy path; "
                                      "model %r stays with the reference implementation" % model)
        self.model = self.method = model
        self.e_conv, self.r_conv, self.maxiter = 1e-7, 1e-7, 100
        self.need_singles = ['CCSD', 'CCSD(T)', 'CC2', 'CC3']
        self.make_t3_density = kwargs.pop('make_t3_density', False)
        if self.make_t3_density not in (True, False):
            raise InvalidKeywordError('make_t3_density', self.make_t3_density, [True, False])
        local = kwargs.pop('local', None)
        if local is not None:
            raise NotImplementedError("local correlation is CPU-only in the reference and not accelerated")
        self.local = None
        orbital_basis = kwargs.pop('orbital_basis', None)
        if orbital_basis not in (None, 'spatial'):
            raise NotImplementedError("only the spatial (closed-shell) path is accelerated")
        self.orbital_basis = 'spatial'
        self.quiet = kwargs.pop('quiet', False)
        device = kwargs.pop('device', 'GPU')
        precision = kwargs.pop('precision', 'DP')
        if 'frozen_core' in kwargs:
            raise TypeError("the 'frozen_core' argument was removed; the frozen core comes from the reference "
                            "wavefunction")
        self.comm = kwargs.pop('comm', None)          # parallel.Comm for multi-GPU runs (None = single GPU)
        self.part = self.comm if self.comm is not None else Serial()
        if kwargs:
            raise PyCCError("Unexpected keyword argument(s): %s" % sorted(kwargs))

        self.device_manager = mgr = DeviceManager(device=device, precision=precision)
        self.precision, self.device = mgr.precision, mgr.device
        self.device0, self.device1 = mgr.device0, mgr.device1
        self.contract = mgr.contract
        self._ct = mgr.contract.engine                 # the Contractor (in-place / alpha-beta form)

        ref = resolve_reference(scf_wfn)
        self.ref = scf_wfn
        self.eref = ref.eref
        self.mixed = mgr.mixed
        self.H = ref.hamiltonian(self.device1, comm=self.comm, mixed=self.mixed)
        self.nfzc, self.no, self.nv, self.nmo = self.H.nfzc, self.H.no, self.H.nv, self.H.nmo
        self.nact = self.no + self.nv
        self.o, self.v = self.H.o, self.H.v
        self.eps_o = self.H.eps[self.o].contiguous()
        self.eps_v = self.H.eps[self.v].contiguous()
        self._Dia = self._Dijab = None

        # t1 = 0, t2 = <ij|ab>/Dijab     (ccwfn.py:210-211)
        self.t1 = torch.zeros((self.no, self.nv), dtype=F64, device=self.device1)
        self.t2 = K.div_d2(self.H.block("oovv"), self.eps_o, self.eps_v)
        self.ecc = None
        if not self.quiet:
            print(timing("CCwfn", time.time() - t0))

    # the reference materialises these (ccwfn.py:195-198); here they are built only if someone asks,
    # from eps by strided broadcasts (the kernels divide by eps sums on the fly instead)
    @property
    def Dia(self):
        if self._Dia is None:
            no, nv = self.no, self.nv
            out = torch.empty((no, nv), dtype=F64, device=self.device1)
            K.strided_axpby(out, self.eps_o.view(-1, 1).expand(no, nv), 1.0, 0.0)
            K.strided_axpby(out, self.eps_v.view(1, -1).expand(no, nv), -1.0, 1.0)
            self._Dia = out
        return self._Dia

    @property
    def Dijab(self):
        if self._Dijab is None:
            no, nv = self.no, self.nv
            out = torch.empty((no, no, nv, nv), dtype=F64, device=self.device1)
            K.strided_axpby(out, self.eps_o.view(-1, 1, 1, 1).expand(no, no, nv, nv), 1.0, 0.0)
            K.strided_axpby(out, self.eps_o.view(1, -1, 1, 1).expand(no, no, nv, nv), 1.0, 1.0)
            K.strided_axpby(out, self.eps_v.view(1, 1, -1, 1).expand(no, no, nv, nv), -1.0, 1.0)
            K.strided_axpby(out, self.eps_v.view(1, 1, 1, -1).expand(no, no, nv, nv), -1.0, 1.0)
            self._Dijab = out
        return self._Dijab

    # =============================================================================================
    # solve_cc (ccwfn.py:216-319)
    # =============================================================================================
    def solve_cc(self, e_conv=1e-7, r_conv=1e-7, maxiter=100, max_diis=8, start_diis=1):
        tstart = time.time()
        self.e_conv, self.r_conv, self.maxiter = e_conv, r_conv, maxiter
        o, v, F = self.o, self.v, self.H.F
        say = (lambda *a: None) if self.quiet else print

        ecc = float(self.cc_energy(o, v, F, self.H.L, self.t1, self.t2))
        name = "T-amplitudes (%s)" % self.model
        say(solve_params(self.model, e_conv, r_conv, maxiter, max_diis, start_diis))
        say(title(name))
        say(iteration(0, energy=ecc, de=-ecc, e_label="CC Ecorr", note="MP2"))

        diis = self.make_diis(max_diis)
        self.trace = []
        for niter in range(1, maxiter + 1):
            ecc_last = ecc
            ecc, rms = self.iterate(F)
            ediff = ecc - ecc_last
            self.trace.append((ecc, rms))
            say(iteration(niter, energy=ecc, de=ediff, rms=rms, e_label="CC Ecorr"))
            if abs(ediff) < e_conv and abs(rms) < r_conv:
                self._gather_rows()
                say(converged(name, time.time() - tstart))
                say("E(REF)  = %20.15f" % self.eref)
                ecc_t = torch.tensor(ecc, dtype=F64, device=self.device1)
                if self.model == 'CCSD(T)':
                    say("E(CCSD) = %20.15f" % ecc)
                    from .cctriples import t_tjl
                    # ccwfn.py:300-308: the density-producing (T) when make_t3_density is set, else Lee-Rendell
                    et = self.t3_density() if self.make_t3_density is True else t_tjl(self)
                    say("E(T)    = %20.15f" % float(et))
                    ecc_t = ecc_t + et
                else:
                    say("E(%s) = %20.15f" % (self.model, ecc))
                self.ecc = ecc_t
                say("E(TOT)  = %20.15f" % (float(ecc_t) + self.eref))
                return ecc_t
            self.diis_step(diis, niter >= start_diis)
        # not converged: the reference falls off the loop and returns None (ccwfn.py:268-319)
        self._gather_rows()
        return None

    # with several ranks the HBM-bound tail of an iteration (update, energy, DIIS) is sharded over the rows of t2 and
    # the rows are all-gathered once per iteration (parallel.py); False = every rank repeats the whole tail
    shard_update = True

    def make_diis(self, max_diis=8):
        """The DIIS object solve_cc uses: sharded over this wavefunction's ranks when the update is (see shard_update)."""
        sharded = self.part.size > 1 and self.shard_update
        return helper_diis(self.t1, self.t2, max_diis, self.precision, comm=self.comm if sharded else None)

    def _gather_rows(self):
        """Make t2 complete on every rank again after a sharded update (no-op otherwise)."""
        if getattr(self, "_rows_stale", False):
            with K.PHASES("all-gather t2"):
                self.part.all_gather_rows(self.t2)
            self._rows_stale = False

    def t3_density(self):
        """(T) contributions to the Lambda residuals and the one-/two-particle densities (ccwfn.py:1819-1829):
        delegates to cctriples.t3_density, caches the returned pieces on the wavefunction (Doo, Dvv, Dov, Goovv,
        Gooov, Gvvvo, S1, S2) and returns the (T) energy."""
        from . import cctriples
        et, dens = cctriples.t3_density(self.o, self.v, self.no, self.nv, self.t1, self.t2, self.H.F, self.H.ERI,
                                        self.H.L, self.contract, comm=self.comm, mixed=self.mixed)
        for name, value in dens.items():
            setattr(self, name, value)
        return et

    # =============================================================================================
    # CC3 (ccwfn.py:374-430, 947-1120): T1-dressed intermediates and the connected-triples terms
    # =============================================================================================
    def _E(self, pat):
        return self.H.ERI[tuple(self.o if c == 'o' else self.v for c in pat)]

    def _vvvv_released(self):
        """The full FP64 <ab|ef> block is not resident: only its pair-packed form (or the TF32 planes of that,
        precision='MP' / 'SP') is."""
        return not self.H.has("vvvv")

    def _t1_vvvv(self, t1, out_abei, alpha=1.0):
        """out_abei[a,b,e,i] += alpha * sum_f t_if <ab|ef> from the pair-packed <ab|ef> (cchbar.py:654 / ccwfn.py:880,
        1104 when the full block is not resident).  ``out_abei``: any strided (v,v,v,o) view with that index order.
        Every RESIDENT pair (a >= b) contributes its (a,b) element and, through <ba|ef> = <ab|fe>, its (b,a) image --
        with <ab|ef> sharded over ranks the result is this rank's piece (the caller all-reduces)."""
        t1 = t1.contiguous()
        ct = self._ct
        with K.mixed_mode(False):                       # the rebuilt chunks are temporaries: no TF32 split of them
            for a0, a1, X in self.H.vvvv_pair_chunks():
                d1 = ct('pef,if->pei', X, t1)
                d2 = ct('pfe,if->pei', X, t1)
                off = 0
                for a in range(a0, a1):
                    K.strided_axpby(out_abei[a, 0:a + 1], d1[off:off + a + 1], alpha, 1.0)
                    if a > 0:
                        K.strided_axpby(out_abei[0:a, a], d2[off:off + a], alpha, 1.0)
                    off += a + 1
                del X, d1, d2
        return out_abei

    def build_cc3_Wmnij(self, o, v, ERI, t1):
        """W_mnij = <mn|ij> + t_ja <mn|ia> + t_ia <nm|ja> + t_ie t_jf <mn|ef>          (ccwfn.py:947-977)"""
        self._own(ERI)
        ct, t1 = self._ct, t1.contiguous()
        W = K.permuted(self._E('oooo'), (0, 1, 2, 3))
        tmp = ct('ijma,na->ijmn', self._E('ooov'), t1)
        K.strided_axpby(W, tmp, 1.0, 1.0)
        K.strided_axpby(W, tmp.permute(1, 0, 3, 2), 1.0, 1.0)
        ct('mnif,jf->mnij', ct('ia,mnaf->mnif', t1, self._E('oovv')), t1, out=W, alpha=1.0, beta=1.0)
        return W

    def build_cc3_Wmbij(self, o, v, ERI, t1, Wmnij):
        """W_mbij = <mb|ij> - W_mnij t_nb + t_je <mb|ie> + t_ie (<mb|ej> + t_jf <mb|ef>)   (ccwfn.py:979-1008)"""
        self._own(ERI)
        ct, t1 = self._ct, t1.contiguous()
        W = K.permuted(self._E('ovoo'), (0, 1, 2, 3))
        ct('mnij,nb->mbij', Wmnij, t1, out=W, alpha=-1.0, beta=1.0)
        ct('mbie,je->mbij', self._E('ovov'), t1, out=W, alpha=1.0, beta=1.0)
        tmp = K.permuted(self._E('ovvo'), (0, 1, 2, 3))
        ct('mbef,jf->mbej', self._E('ovvv'), t1, out=tmp, alpha=1.0, beta=1.0)
        ct('ie,mbej->mbij', t1, tmp, out=W, alpha=1.0, beta=1.0)
        return W

    def build_cc3_Wmnie(self, o, v, ERI, t1):
        """W_mnie = <mn|ie> + t_if <mn|fe>                                           (ccwfn.py:1010-1031)"""
        self._own(ERI)
        W = K.permuted(self._E('ooov'), (0, 1, 2, 3))
        return self._ct('if,mnfe->mnie', t1.contiguous(), self._E('oovv'), out=W, alpha=1.0, beta=1.0)

    def build_cc3_Wamef(self, o, v, ERI, t1):
        """W_amef = <am|ef> - t_na <nm|ef>                                           (ccwfn.py:1033-1052)"""
        self._own(ERI)
        W = K.permuted(self._E('vovv'), (0, 1, 2, 3))
        return self._ct('na,nmef->amef', t1.contiguous(), self._E('oovv'), out=W, alpha=-1.0, beta=1.0)

    def build_cc3_Wabei(self, o, v, ERI, t1):
        """W_abei = Z_abei + Z_eiab^T with                                           (ccwfn.py:1054-1120)
             Z_eiab = <ei|ab> + t_if <ab|ef> - t_mb (<ei|am> + t_if <am|ef>) + t_ma t_nb (<mn|ei> + t_if <mn|ef>)
             Z_abei = -t_ma (<mb|ei> + t_if <mb|ef>)
        (the reference's symmetric + antisymmetric split of <ab|ef> sums back to the block: ONE pass over <ab|ef>)."""
        self._own(ERI)
        ct, t1 = self._ct, t1.contiguous()
        Z = K.permuted(self._E('vovv'), (0, 1, 2, 3))                                   # [e,i,a,b]
        if self.part.size > 1 and self.H.a_range != (0, self.nv):
            # <ab|ef> a-sharded over the ranks (parallel.py): each contracts the pairs it holds, ONE all-reduce of the
            # o v^3 piece; every other term of W_abei is replicated
            if self.H.a_range != tuple(self.part.a_range(self.nv)):
                raise NotImplementedError("<ab|ef> rows resident on this rank are neither its share nor the whole block")
            piece = torch.zeros((self.nv, self.nv, self.nv, self.no), dtype=F64, device=self.device1)
            self._t1_vvvv(t1, piece)
            self.part.all_reduce_sum(piece)
            K.strided_axpby(Z, piece.permute(2, 3, 0, 1), 1.0, 1.0)
            del piece
        elif self._vvvv_released():
            self._t1_vvvv(t1, Z.permute(2, 3, 0, 1))
        else:
            ct('if,abef->eiab', t1, self._E('vvvv'), out=Z, alpha=1.0, beta=1.0)
        Zeiam = K.permuted(self._E('vovo'), (0, 1, 2, 3))
        K.strided_axpby(Zeiam, ct('amef,if->amei', self._E('vovv'), t1).permute(2, 3, 0, 1), 1.0, 1.0)
        ct('eiam,mb->eiab', Zeiam, t1, out=Z, alpha=-1.0, beta=1.0)
        Zmnei = K.permuted(self._E('oovo'), (0, 1, 2, 3))
        ct('mnef,if->mnei', self._E('oovv'), t1, out=Zmnei, alpha=1.0, beta=1.0)
        ct('anei,nb->eiab', ct('ma,mnei->anei', t1, Zmnei), t1, out=Z, alpha=1.0, beta=1.0)
        Zmbei = K.permuted(self._E('ovvo'), (0, 1, 2, 3))
        ct('mbef,if->mbei', self._E('ovvv'), t1, out=Zmbei, alpha=1.0, beta=1.0)
        W = ct('ma,mbei->abei', t1, Zmbei, alpha=-1.0)
        return K.strided_axpby(W, Z.permute(2, 3, 0, 1), 1.0, 1.0)

    def _cc3_t_residual(self, o, v, F, ERI, L, t1, t2, Fme, real_time=False, reduce=True):
        """(X1, X2): the connected-triples contributions to the T1 / T2 residuals (ccwfn.py:374-430).  ``real_time``:
        every t3 is corrected by the explicit-field term t3_pert_ijk(V = F - H.F) (ccwfn.py:421-423).  With several
        ranks the (i,j) pairs of the triples loop are dealt round-robin; ``reduce=False`` returns this rank's partial
        sums (the caller adds them to a buffer it all-reduces anyway)."""
        self._own(ERI, L)
        from . import cctriples
        t1 = t1.contiguous()
        F = self._check_F(F)
        V = None
        if real_time:
            V = K.strided_axpby(K.permuted(F[o, v], (0, 1)), self._check_F(self.H.F)[o, v], -1.0, 1.0)
        Wmnij = self.build_cc3_Wmnij(o, v, ERI, t1)
        W = {"Wmbij": self.build_cc3_Wmbij(o, v, ERI, t1, Wmnij), "Wmnie": self.build_cc3_Wmnie(o, v, ERI, t1),
             "Wamef": self.build_cc3_Wamef(o, v, ERI, t1), "Wabei": self.build_cc3_Wabei(o, v, ERI, t1)}
        comm = self.part if self.part.size > 1 else None
        return cctriples.cc3_t_residual(self, F, t1, t2, Fme, W, V=V, comm=comm, reduce=reduce)

    def iterate(self, F=None):
        """One Jacobi step of solve_cc (ccwfn.py:272-286): residuals, then ONE fused pass doing
        r2 = half + half^T (790), t += r/D and sum (r/D)^2 (281-284), then the energy (286).
        Returns (ecc, rms) as host floats -- the single device->host sync of the iteration."""
        F = self.H.F if F is None else F
        self._gather_rows()
        # the amplitudes solve_cc iterates are pair-symmetric, t2[i,j,a,b] = t2[j,i,b,a] (the guess <ij|ab>/D is, the
        # update adds a symmetrised residual, DIIS mixes iterates linearly): the ladder runs on rows (i >= j) only
        r1, half = self._residuals_half(F, self.t1, self.t2, symmetric=True)
        if self.part.size > 1 and self.shard_update:
            # every rank holds the summed half residual; it updates ITS rows of t2 (and the small, replicated t1),
            # the partial sums of squares / energies are added over the ranks; the rows are gathered by diis_step
            i0, i1 = self.part.occ_range(self.no)
            Fc = self._check_F(F)
            with K.PHASES("update+energy"):
                ss = K.update_amps_rows(r1, half, self.eps_o, self.eps_v, self.t1, self.t2, i0, i1)
                e_p = K.cc_energy_rows(Fc[self.o, self.v], self.t1, self.t2, self.H.derived("Loovv"), i0, i1,
                                       self.part.rank == 0)
                both = torch.stack((ss[0], e_p[0]))
                self.part.all_reduce_sum(both)
                ssq2, ecc, ssq1 = torch.cat((both, ss[1:2])).tolist()
            self._rows_stale = True
            return ecc, (ssq2 + ssq1) ** 0.5
        with K.PHASES("update+energy"):
            ssq = K.update_amps(r1, half, self.eps_o, self.eps_v, self.t1, self.t2, symmetrize=True,
                                write_r2=False)
            e_dev = self.cc_energy(self.o, self.v, F, self.H.L, self.t1, self.t2)
            ssq_h, ecc = torch.stack((ssq[0], e_dev)).tolist()
        return ecc, ssq_h ** 0.5

    def diis_step(self, diis, extrapolate=True):
        """ccwfn.py:317-319.  After a sharded update (iterate with several ranks) t2 is complete only in this rank's rows:
        a sharded DIIS object (make_diis) reads just those and all-gathers the extrapolant; otherwise the rows are
        gathered first."""
        with K.PHASES("diis"):
            if getattr(diis, "comm", None) is None or not extrapolate or diis.max_diis == 0:
                self._gather_rows()
            diis.add_error_vector(self.t1, self.t2)
            if extrapolate:
                self.t1, self.t2 = diis.extrapolate(self.t1, self.t2)
                if getattr(diis, "comm", None) is not None and diis.max_diis != 0:
                    self._rows_stale = False                # the extrapolant was all-gathered

    # =============================================================================================
    # residuals (ccwfn.py:321-372)
    # =============================================================================================
    def residuals(self, F, t1, t2, real_time=False):
        """(r1, r2) for the given Fock matrix and amplitudes; r2 symmetrised as in r_T2 (ccwfn.py:790)."""
        if _is_complex(t1) or _is_complex(t2) or _is_complex(F):
            return self._residuals_complex(F, t1, t2, real_time)
        if self._foreign():
            # someone swapped H.ERI / H.L (perturbed integrals, ccderiv.py:250-259): term-by-term generic evaluation
            r1, half = self._generic().residuals(F, t1, t2)
            return r1, K.symmetrize_r2(half)
        r1, half = self._residuals_half(F, t1, t2, real_time=real_time)
        K.symmetrize_r2(half)
        return r1, half

    def _residuals_complex(self, F, t1, t2, real_time=False):
        """(r1, r2) for COMPLEX t1, t2 (and optionally a complex Hermitian F) as torch.complex128 tensors: the RT-CC
        right-hand side (rt/rtcc.py:136-141), from REAL evaluations of the fused kernels (no complex kernel exists):

        * everything but the ladder: five real residuals, see utils.complex_from_real_samples (six for CC3);
        * the v^4 ladder (ccwfn.py:931) is LINEAR in tau = t2 + t1 t1, and <ab|ef> is real: its contribution is
          L(Re tau) + i L(Im tau) -- two ladder GEMMs instead of one per sample;
        * pair-symmetric amplitudes (t2[i,j,a,b] = t2[j,i,b,a] in both planes -- what an RT-CC propagation carries)
          take the (i >= j) formulation of ``iterate`` in every sample (four o^3v^3 GEMMs instead of six, half the
          ladder rows, Z in pair form), detected with one o^2v^2 pass per plane.
        With the same rank sharding as the real residual (the two ladder pieces take one all-reduce each)."""
        from .utils import complex_from_real_samples, real_planes, planes_to_complex
        degree = 4
        if self.model == 'CC3':
            # the CC3 triples terms are of degree 5 in (F, t1, t2) scaled together (t3 ~ Wabei(t1^3) t2, contracted with
            # Wamef(t1) / Fme; the explicit-field term V t2 t2 W(t1) stays below): six samples.  The t3 denominators hold
            # diag(F): they are the same in every sample only if that diagonal is real (a Hermitian field, rtcc.py:136)
            degree = 5
            if _is_complex(F):
                Fc = F if isinstance(F, torch.Tensor) else torch.as_tensor(np.asarray(F))
                if float(torch.diagonal(Fc).imag.abs().max()) != 0.0:
                    raise NotImplementedError("complex CC3 residuals need a Fock matrix with a real diagonal "
                                              "(the t3 denominators are not polynomial in Im f_pp)")
        dev = self.device1
        P = [real_planes(x, dev) for x in (F, t1, t2)]
        (x1, y1), (x2, y2) = P[1], P[2]
        symmetric = self.complex_pair_mode and self._pair_symmetric(x2) and (y2 is None or self._pair_symmetric(y2))
        native_ladder = self.complex_native_ladder and self.model != 'CC2'
        # every o^3v^3 product on the planes as well: single rank, both amplitude planes present, t1 in play
        native_heavy = (native_ladder and self.complex_native_heavy and self.part.size == 1 and y1 is not None
                        and y2 is not None and self.model in ('CCSD', 'CCSD(T)', 'CC3'))
        lin = []

        def real_residual(Fs, t1s, t2s):
            r1, half = self._residuals_half(Fs, t1s, t2s, symmetric=symmetric, real_time=real_time,
                                            ladder=not native_ladder, heavy=not native_heavy)
            if native_heavy:
                I = self._last_I
                self._last_I = None
                return r1, half.view(t2s.shape), I["W1"], I["W2"]
            return r1, half.view(t2s.shape)

        out = complex_from_real_samples(real_residual, None, dev, degree=degree, as_planes=True, planes=P)
        (r1re, r1im), (hre, him) = out[0], out[1]
        if native_ladder:
            # Re tau = x2 + x1 x1 - y1 y1,  Im tau = y2 + x1 y1 + y1 x1 = y2 + (x1+y1)(x1+y1) - x1 x1 - y1 y1
            tre = K.build_tau(x1, x2, 1.0, 1.0)
            tim = None
            if y1 is not None or y2 is not None:
                z2 = torch.zeros_like(x2) if y2 is None else y2
                if y1 is None:
                    tim = z2
                else:
                    yy = K.build_tau(y1, x2, 0.0, 1.0)                                   # y1 y1
                    xx = K.build_tau(x1, x2, 0.0, 1.0)                                   # x1 x1
                    s1 = K.axpbyz(1.0, x1, 1.0, y1, torch.empty_like(x1))
                    tim = K.build_tau(s1, z2, 1.0, 1.0)                                  # y2 + (x1+y1)(x1+y1)
                    K.axpbyz(1.0, tim.view(-1), -1.0, xx.view(-1), tim.view(-1))
                    K.axpbyz(1.0, tim.view(-1), -1.0, yy.view(-1), tim.view(-1))
                    K.axpbyz(1.0, tre.view(-1), -1.0, yy.view(-1), tre.view(-1))
                    del xx, yy
            with K.mixed_mode(self.mixed):
                for tau, acc in ((tre, hre), (tim, him)):
                    if tau is None:
                        continue
                    piece = torch.zeros_like(tau)
                    self._ladder(tau, piece, symmetric=symmetric)
                    if self.part.size > 1:
                        self.part.all_reduce_sum(piece)
                    K.axpbyz(1.0, acc.view(-1), 1.0, piece.view(-1), acc.view(-1))
                    del piece
                if native_heavy:
                    self._complex_heavy((x1, y1), (x2, y2), (tre, tim), out[2], out[3], (hre, him), symmetric)
        K.symmetrize_r2(hre)
        K.symmetrize_r2(him)
        return planes_to_complex(r1re, r1im), planes_to_complex(hre, him)

    def _complex_heavy(self, t1p, t2p, taup, W1lin, W2lin, half, symmetric):
        """The o^3v^3 products of the residual for complex amplitudes, on planes (single rank; see _residuals_complex):

        * W_mbej / W_mbje (ccwfn.py:641-645, 680-683): their linear parts (integrals + t1 terms) arrive as planes
          ``W1lin``, ``W2lin`` (interpolated from the samples: they are linear in t1); the products of t2 / tau(1/2,1)
          with the REAL blocks <mn|ef>, L_mnef are plane-wise -- two GEMMs each instead of one per sample;
        * the ring terms (933-935) are complex x complex: the 3M form (three real GEMMs per product);
        * Z_mbij = <mb|ef> tau_ijef (715) plane-wise, then - t_ma Z_mbij (932) as four small batched products.
        Pair-symmetric amplitudes use the {D = W1 + W2/2, W2} form of ``_r2_half`` (two ring products instead of three)
        and Z in pair form."""
        H, ct = self.H, self._ct
        no, nv = self.no, self.nv
        oovv_menf, oovv_mfne, L_menf = H.derived("oovv_menf"), H.derived("oovv_mfne"), H.derived("Loovv_menf")
        fused = bool(symmetric) and self.fuse_rings

        def add(a, x, b, y):                                       # a x + b y on equally shaped contiguous tensors
            return K.axpbyz(a, x.view(-1), b, y.view(-1), torch.empty_like(x).view(-1)).view(x.shape)

        def cprod(sub, Xp, Yp):
            """(X_re + i X_im)(Y_re + i Y_im) contracted as ``sub``: the 3M form, three real GEMMs"""
            p1 = ct(sub, Xp[0], Yp[0])
            p2 = ct(sub, Xp[1], Yp[1])
            p3 = ct(sub, add(1.0, Xp[0], 1.0, Xp[1]), add(1.0, Yp[0], 1.0, Yp[1]))
            K.axpbyz(1.0, p3.view(-1), -1.0, p1.view(-1), p3.view(-1))
            K.axpbyz(1.0, p3.view(-1), -1.0, p2.view(-1), p3.view(-1))            # Im = P3 - P1 - P2
            K.axpbyz(1.0, p1.view(-1), -1.0, p2.view(-1), p1.view(-1))            # Re = P1 - P2
            return p1, p3

        # ---- W1 / W2 (or D / W2) on planes: quadratic parts from tau(1/2,1) = tau - t2/2 and t2 ------------------
        W1, W2, lay = [], [], []
        for p in range(2):
            taut = add(1.0, taup[p], -0.5, t2p[p])                                 # tau(1/2, 1) = t2/2 + t1 t1
            taut_jbnf = K.permuted(taut, (0, 3, 1, 2))                             # [j,b,n,f] = tau(1/2,1)[j,n,f,b]
            del taut
            t2_jbnf = K.permuted(t2p[p], (1, 3, 0, 2))                             # [j,b,n,f] = t2[n,j,f,b]
            A, B = W1lin[p], W2lin[p]
            if fused:
                K.strided_axpby(A, B, 0.5, 1.0)                                    # lin(W1) + 1/2 lin(W2)
                ct("jbnf,menf->jbme", taut_jbnf, oovv_mfne, out=B, alpha=1.0, beta=1.0)
                K.strided_axpby(t2_jbnf, taut_jbnf, -1.0, 1.0)
                ct("jbnf,menf->jbme", t2_jbnf, L_menf, out=A, alpha=0.5, beta=1.0)      # D
            else:
                ct("jbnf,menf->jbme", taut_jbnf, oovv_menf, out=A, alpha=-1.0, beta=1.0)
                ct("jbnf,menf->jbme", t2_jbnf, L_menf, out=A, alpha=0.5, beta=1.0)
                ct("jbnf,menf->jbme", taut_jbnf, oovv_mfne, out=B, alpha=1.0, beta=1.0)
            del taut_jbnf, t2_jbnf
            W1.append(A)
            W2.append(B)
            if fused:
                lay.append(K.ring_layouts(t2p[p]))                                 # (s_iame, t2_jame)
            else:
                s = K.permuted(t2p[p], (0, 2, 1, 3), 2.0)
                K.strided_axpby(s, t2p[p].permute(0, 3, 1, 2), -1.0, 1.0)
                lay.append((s, K.permuted(t2p[p], (1, 2, 0, 3)), K.permuted(t2p[p], (0, 2, 1, 3))))
        # ---- ring terms (see _r2_half for the two forms) -----------------------------------------------------------
        s_iame = (lay[0][0], lay[1][0])
        t2_jame = (lay[0][1], lay[1][1])
        if fused:
            R = cprod("iame,jbme->iajb", s_iame, W1)                               # u . D
            X = cprod("jame,ibme->jaib", t2_jame, W2)
            for p in range(2):
                K.strided_axpby(R[p], X[p], 0.5, 1.0)
                K.strided_axpby(half[p], R[p].permute(0, 2, 1, 3), 1.0, 1.0)
                K.strided_axpby(half[p], X[p].permute(2, 0, 1, 3), 1.0, 1.0)
        else:
            t2_iame = (lay[0][2], lay[1][2])
            R = cprod("iame,jbme->iajb", s_iame, W1)
            R2 = cprod("iame,jbme->iajb", t2_iame, W2)
            X = cprod("jame,ibme->jaib", t2_jame, W2)
            for p in range(2):
                K.strided_axpby(R[p], R2[p], 1.0, 1.0)
                K.strided_axpby(half[p], R[p].permute(0, 2, 1, 3), 1.0, 1.0)
                K.strided_axpby(half[p], X[p].permute(2, 0, 1, 3), 1.0, 1.0)
            del R2
        del R, X, W1, W2, lay, s_iame, t2_jame
        # ---- Z and - t_ma Z_mbij -----------------------------------------------------------------------------------
        Z = []
        for p in range(2):
            A = {"tau": taup[p]}
            if symmetric and self.pair_z:
                A["Tpm"] = K.pack_tau(taup[p], True)
            Z.append(self._zgemm(A, 0, no))                                        # [i, j, m, b]
        for (zp, tp, hp, alpha) in ((0, 0, 0, -1.0), (1, 1, 0, 1.0), (1, 0, 1, -1.0), (0, 1, 1, -1.0)):
            # half[hp][i,j,a,b] += alpha sum_m t1[tp][m,a] Z[zp][i,j,m,b]
            K.dgemm(nv, nv, no, t1p[tp], nv, 1, Z[zp], nv, 1, half[hp], nv, alpha, 1.0,
                    batch=no * no, sA=0, sB=no * nv, sC=nv * nv)

    # switches of the complex path (tests compare the settings): the (i >= j) formulation for pair-symmetric amplitudes,
    # the ladder evaluated on the two planes of tau instead of inside every sample
    complex_pair_mode = True
    complex_native_ladder = True
    complex_native_heavy = True

    def _pair_symmetric(self, t2, tol=1e-14):
        """t2[i,j,a,b] == t2[j,i,b,a] to ``tol`` of its norm (one pass with the permute and dot kernels)."""
        d = K.permuted(t2, (0, 1, 2, 3))
        K.strided_axpby(d, t2.permute(1, 0, 3, 2), -1.0, 1.0)
        dd = float(K.multi_dot(d.view(-1), [d.view(-1)])[0])
        if dd == 0.0:
            return True
        flat = t2.contiguous().view(-1)
        return dd <= tol * tol * float(K.multi_dot(flat, [flat])[0])

    def _check_F(self, F):
        if not isinstance(F, torch.Tensor):
            F = torch.as_tensor(np.asarray(F), dtype=F64)
        F = F.to(self.device1, dtype=F64)
        return F if F.is_contiguous() else F.contiguous()

    def _residuals_half(self, F, t1, t2, symmetric=False, real_time=False, ladder=True, heavy=True):
        """r1 and the UNsymmetrised half of r2 (ccwfn.py:922-940), fused formulation.  With several ranks
        each computes its share of r2 (see parallel.py) and ONE all-reduce sums them.  precision='MP': the large
        K-major contractions inside run on the split-TF32 tcgen05 kernel (kernels.mixed_mode).
        ``symmetric``: the caller vouches that t2[i,j,a,b] = t2[j,i,b,a] (halves the ladder once more)."""
        with K.mixed_mode(self.mixed):
            return self._residuals_half_impl(F, t1, t2, symmetric, real_time, ladder, heavy)

    def _residuals_half_impl(self, F, t1, t2, symmetric=False, real_time=False, ladder=True, heavy=True):
        F = self._check_F(F)
        t1 = t1.contiguous()
        t2 = t2.contiguous()
        cc2 = self.model == 'CC2'
        with K.PHASES("intermediates"):
            I = self._intermediates(F, t1, t2, rings=not cc2, symmetric=symmetric, heavy=heavy)
        # heavy=False (the complex path): the caller picks up the linear parts of W1 / W2 from here, then drops the dict
        self._last_I = I if not heavy else None
        # one flat buffer [ r2 half | rank-partial part of r1 ] so that ONE all-reduce carries both
        n2, n1 = t2.numel(), t1.numel()
        buf = torch.empty(n2 + n1, dtype=F64, device=self.device1)
        half = buf[:n2].view(t2.shape)
        r1p = buf[n2:].view(t1.shape)
        with K.PHASES("r1"):
            r1 = self._r1(F, t1, t2, I, r1p)
        if cc2:
            self._r2_half_cc2(F, t1, t2, half)
        else:
            self._r2_half(F, t1, t2, I, half, symmetric=symmetric, ladder=ladder, rings=heavy)
        if self.model == 'CC3':
            # connected triples (ccwfn.py:364-367): r1 += X1, r2 += X2 + X2^T -- X2 joins the unsymmetrised half, the
            # rank's partial sums (its (i,j) pairs of the triples loop) join the rank-partial buffers before the all-reduce
            Fme = self.build_Fme(self.o, self.v, F, self.H.L, t1)
            X1, X2 = self._cc3_t_residual(self.o, self.v, F, self.H.ERI, self.H.L, t1, t2, Fme, real_time=real_time,
                                          reduce=False)
            K.strided_axpby(r1p, X1, 1.0, 1.0)
            K.strided_axpby(half, X2, 1.0, 1.0)
            del X1, X2
        if self.part.size > 1:
            with K.PHASES("all-reduce r2"):
                self.part.all_reduce_sum(buf)
        K.strided_axpby(r1, r1p, 1.0, 1.0)
        if self.model == 'CCD':
            r1.zero_()
        return r1, half

    # ---- shared per-iteration rearrangements of the amplitudes -----------------------------------
    def _amps(self, t1, t2, symmetric=False):
        """o^2v^2 permutations of t2 / tau reused by several contractions ("ring layout" [i,a,m,e]).  ``symmetric``
        (pair-symmetric t2): also T+- of tau over the pairs (e >= f), rows (i >= j) -- shared by the ladder and Z."""
        A = {}
        A["tau"] = K.build_tau(t1, t2, 1.0, 1.0)                 # t2 + t1 t1
        if symmetric:
            A["Tpm"] = K.pack_tau(A["tau"], True)
        if symmetric and self.fuse_rings:
            # both operands of the two ring products in ONE pass over t2 (pair symmetry: t2[m,j,a,e] = t2[j,m,e,a])
            A["s_iame"], A["t2_jame"] = K.ring_layouts(t2)
            return A
        A["t2_iame"] = K.permuted(t2, (0, 2, 1, 3))               # [i,a,m,e] = t2[i,m,a,e]
        s = K.permuted(t2, (0, 2, 1, 3), 2.0)                     # s~[i,a,m,e] = 2 t2[i,m,a,e] - t2[i,m,e,a]
        K.strided_axpby(s, t2.permute(0, 3, 1, 2), -1.0, 1.0)
        A["s_iame"] = s
        return A

    def _intermediates(self, F, t1, t2, full=False, rings=True, symmetric=False, heavy=True):
        """Fae, Fmi, Fme (replicated) and the LOCAL slices of Wmnij (rows i_g), W1/W2 (columns j_g) and
        Z' (rows i_g).  ``full=True`` ignores the rank partition (public build_* methods).  ``heavy=False`` (the complex
        path, _residuals_complex): the o^3v^3 GEMMs are left out -- W1 / W2 come back with their linear parts only
        (integral + t1 terms, I["W1"], I["W2"], unmixed) and Z' as zeros; the caller evaluates those products on the
        real and imaginary planes itself."""
        H, ct = self.H, self._ct
        o, v, no, nv = self.o, self.v, self.no, self.nv
        i0, i1 = (0, no) if full else self.part.occ_range(no)
        ni = i1 - i0
        with K.PHASES("  amps (tau, ring layouts)"):
            A = self._amps(t1, t2, symmetric)
        I = {"amps": A, "occ": (i0, i1)}
        ccd = self.model == 'CCD'
        Fov = F[o, v]
        Loovv = H.derived("Loovv")

        K.PHASES.mark("  Fme, Fae, Fmi")
        # ---------------- Fme = f_me + t_nf L_mnef                       (ccwfn.py:563-564)
        Fme = K.permuted(Fov, (0, 1))
        if not ccd:
            ct("menf,nf->me", H.derived("Loovv_menf"), t1, out=Fme, alpha=1.0, beta=1.0)
        I["Fme"] = Fme

        # ---------------- Fae, Fmi (ccwfn.py:495-497, 531-533).  Their o^2v^3 / o^3v^2 sums run over an occupied
        # index: each rank sums its slice m (resp. n) in [i0,i1) into a packed [Fae|Fmi] buffer, one small
        # all-reduce (v^2 + o^2 doubles) completes them; the cheap t1 / Fock terms are added on every rank.
        tauh = K.build_tau(t1, t2, 1.0, 0.0 if ccd else 0.5)
        pk = torch.zeros(nv * nv + no * no, dtype=F64, device=self.device1)
        Fae_p, Fmi_p = pk[:nv * nv].view(nv, nv), pk[nv * nv:].view(no, no)
        P1 = P2 = None
        if ni > 0 and rings and not ccd:
            # ONE pair of passes over this rank's slabs <m_g b|ef> serves four terms of the reference:
            #   P1[m,b,e,j] = sum_f <mb|ef> t_jf,   P2[m,b,j,e] = sum_f t_jf <mb|fe>          (for ALL j)
            # -> the t1 terms of W_mbej / W_mbje (642, 681), Fae's t_mf L_mafe = 2 P2[m,a,m,e] - P1[m,a,e,m] (496) and
            #    t_ie <ab|ej> = P1[j,a,b,i] (939: <ab|ej> = <ja|be>) -- instead of four separate sweeps of the 8.6 GB block
            K.PHASES.mark("  P1, P2 = t1.<mb|ef> (two passes over the rank's slabs)")
            ovvv = H.block("ovvv")
            P1 = ct("mbef,jf->mbej", ovvv[i0:i1], t1)
            P2 = torch.empty((ni, nv, no, nv), dtype=F64, device=self.device1)
            K.dgemm(no, nv, nv, t1, nv, 0, (ovvv, i0 * nv ** 3), nv, 1, P2, nv, 1.0, 0.0,
                    batch=ni * nv, sA=0, sB=nv * nv, sC=no * nv)
            I["P1"] = P1
            K.PHASES.mark("  Fme, Fae, Fmi")
        if ni > 0:
            ct("mnaf,emnf->ae", tauh[i0:i1], self._Loovv_emnf(i0, i1), out=Fae_p, alpha=-1.0, beta=1.0)
            ct("inef,mnef->mi", tauh[:, i0:i1], Loovv[:, i0:i1], out=Fmi_p, alpha=1.0, beta=1.0)
            if P1 is not None:
                for ml in range(ni):                                  # the j = m diagonal of the two pieces
                    K.strided_axpby(Fae_p, P2[ml, :, i0 + ml, :], 2.0, 1.0)
                    K.strided_axpby(Fae_p, P1[ml, :, :, i0 + ml], -1.0, 1.0)
            elif not ccd:
                self._fae_ovvv(t1, Fae_p, i0, i1)
        if self.part.size > 1 and not full:
            self.part.all_reduce_sum(pk)
        Fae = K.permuted(F[v, v], (0, 1))
        K.strided_axpby(Fae, Fae_p, 1.0, 1.0)
        Fmi = K.permuted(F[o, o], (0, 1))
        K.strided_axpby(Fmi, Fmi_p, 1.0, 1.0)
        if not ccd:
            ct("me,ma->ae", Fov, t1, out=Fae, alpha=-0.5, beta=1.0)
            ct("ie,me->mi", t1, Fov, out=Fmi, alpha=0.5, beta=1.0)
            ct("ne,mnie->mi", t1, H.derived("Looov"), out=Fmi, alpha=1.0, beta=1.0)
        I["Fae"], I["Fmi"] = Fae, Fmi
        del tauh
        K.PHASES.mark(None)
        if ni == 0 or not rings:           # rings=False: the one-body intermediates only (CC2's r1)
            return I

        K.PHASES.mark("  Wmnij")
        # ---------------- Wmnij[m,n,i_g,j]                               (ccwfn.py:596-603)
        ooov = H.block("ooov")
        Wfull = K.permuted(H.block("oooo"), (0, 1, 2, 3))
        if not ccd:
            ct("je,mnie->mnij", t1, ooov, out=Wfull, alpha=1.0, beta=1.0)
            ct("ie,nmje->mnij", t1, ooov, out=Wfull, alpha=1.0, beta=1.0)     # <mn|ej> = <nm|je>
        Wmnij = K.permuted(Wfull[:, :, i0:i1, :], (0, 1, 2, 3))
        del Wfull
        ct("ijef,mnef->mnij", A["tau"][i0:i1], H.block("oovv"), out=Wmnij, alpha=1.0, beta=1.0)
        I["Wmnij"] = Wmnij

        # ---------------- ring intermediates in [j_g,b,m,e] layout (every o^3v^3 GEMM is then K-major x K-major)
        #   W1[j,b,m,e] = Wmbej[m,b,e,j]    (ccwfn.py:641-645)
        #   W2[j,b,m,e] = Wmbje[m,b,j,e]    (ccwfn.py:680-683)
        K.PHASES.mark("  W1, W2: three o3v3 GEMMs + layouts")
        taut_jbnf = t2_jbnf = None
        if heavy:
            taut = K.build_tau(t1, t2, 0.5, 0.0 if ccd else 1.0)
            taut_jbnf = K.permuted(taut[i0:i1], (0, 3, 1, 2))     # [j,b,n,f] = tau(1/2,1)[j,n,f,b]
            del taut
            t2_jbnf = K.permuted(t2[:, i0:i1], (1, 3, 0, 2))      # [j,b,n,f] = t2[n,j,f,b]
        oovv_menf = H.derived("oovv_menf")
        # Pair-symmetric amplitudes (solve_cc): the residual only needs D = W1 + 1/2 W2 and W2 (see _r2_half), and
        #   D = lin(W1) + 1/2 lin(W2) + 1/2 sum_nf (t2[n,j,f,b] - tau(1/2,1)[j,n,f,b]) L_mnef
        # (the two quadratic terms of W1 and half the one of W2 combine, L = 2<mn|ef> - <mn|fe>): TWO o^3v^3 GEMMs build
        # {D, W2} where {W1, W2} take three.  The linear parts must be complete before they are mixed, so the GEMMs come
        # after the t1 terms on this branch.
        fused = bool(symmetric) and self.fuse_rings and not full and heavy
        W1 = K.permuted(oovv_menf[:, :, i0:i1, :], (2, 3, 0, 1))  # <mb|ej> = <mj|eb> -> [j,b,m,e]
        W2 = K.permuted(H.derived("ovov_mejb")[:, :, i0:i1, :], (2, 3, 0, 1), -1.0)
        if not fused and heavy:
            ct("jbnf,menf->jbme", taut_jbnf, oovv_menf, out=W1, alpha=-1.0, beta=1.0)
            ct("jbnf,menf->jbme", t2_jbnf, H.derived("Loovv_menf"), out=W1, alpha=0.5, beta=1.0)
            ct("jbnf,menf->jbme", taut_jbnf, H.derived("oovv_mfne"), out=W2, alpha=1.0, beta=1.0)
        if not fused:
            del t2_jbnf, taut_jbnf
        K.PHASES.mark("  W1, W2: t1 terms (from P1, P2)")
        if not ccd:
            if self.part.size == 1 or full:
                # [m,b,e,j] -> [j,b,m,e]  and  [m,b,j,e] -> [j,b,m,e]
                K.strided_axpby(W1, P1[:, :, :, i0:i1].permute(3, 1, 0, 2), 1.0, 1.0)
                K.strided_axpby(W2, P2[:, :, i0:i1, :].permute(2, 1, 0, 3), -1.0, 1.0)
            else:
                # Several ranks: W1 / W2 are column-sharded (j_g) but these two terms need every slab <mb|ef> for each j.
                # Each rank has contracted ITS slabs m_g with ALL of t1 (an o x o/N share of the work, 1/N of the block
                # read); the ranks swap pieces: rank d receives the columns j_d of everybody's slabs (2 o^2v^2 / N
                # doubles per rank over NVLink, both pieces in one message) and adds them to its W1 / W2.
                bounds = [self.part.occ_range_of(no, r) for r in range(self.part.size)]
                send, recv = [], []
                for (a, b) in bounds:
                    buf = torch.empty((2, ni, nv, nv, b - a), dtype=F64, device=self.device1)
                    K.strided_axpby(buf[0], P1[:, :, :, a:b], 1.0, 0.0)                      # [m_g,b,e,j_d]
                    K.strided_axpby(buf[1], P2[:, :, a:b, :].permute(0, 1, 3, 2), 1.0, 0.0)  # [m_g,b,e,j_d] = P2[m,b,j,e]
                    send.append(buf)
                    recv.append(torch.empty((2, b - a, nv, nv, ni), dtype=F64, device=self.device1))
                self.part.exchange(send, recv)
                for (a, b), got in zip(bounds, recv):                                        # got[.][m_r,b,e,j_g]
                    if b > a:
                        K.strided_axpby(W1[:, :, a:b, :], got[0].permute(3, 1, 0, 2), 1.0, 1.0)
                        K.strided_axpby(W2[:, :, a:b, :], got[1].permute(3, 1, 0, 2), -1.0, 1.0)
                del send, recv
            del P2
            # - t_nb <mn|ej> = - t_nb ooov[n,m,j,e]  and  + t_nb <mn|je>
            ct("nb,nmje->jbme", t1, ooov[:, :, i0:i1, :], out=W1, alpha=-1.0, beta=1.0)
            ct("nb,mnje->jbme", t1, ooov[:, :, i0:i1, :], out=W2, alpha=1.0, beta=1.0)
        if fused:
            K.PHASES.mark("  D, W2: two o3v3 GEMMs")
            K.strided_axpby(W1, W2, 0.5, 1.0)                                     # lin(W1) + 1/2 lin(W2)
            ct("jbnf,menf->jbme", taut_jbnf, H.derived("oovv_mfne"), out=W2, alpha=1.0, beta=1.0)
            K.strided_axpby(t2_jbnf, taut_jbnf, -1.0, 1.0)                        # t2[n,j,f,b] - tau(1/2,1)[j,n,f,b]
            ct("jbnf,menf->jbme", t2_jbnf, H.derived("Loovv_menf"), out=W1, alpha=0.5, beta=1.0)
            del t2_jbnf, taut_jbnf
            I["D"], I["W2"] = W1, W2
        else:
            I["W1"], I["W2"] = W1, W2

        # ---------------- Z'[i_g,j,m,b] = Zmbij[m,b,i,j] = <mb|ef> tau_ijef   (ccwfn.py:715)
        # Sharded over m, not over i: a rank then reads only ITS slabs <m_g b|ef> of the 8.6 GB block (all (i,j) rows of tau)
        # and the contraction with t_ma below gives a partial sum over m_g for every r2 row -- the all-reduce of r2 adds
        # the partial sums.  (Sharded over i, every rank streamed the whole block.)
        K.PHASES.mark("  Z = tau.<mb|ef> (o3v3)")
        if not ccd and not heavy:
            I["Zijmb"] = torch.zeros((no, no, ni, nv), dtype=F64, device=self.device1)
        elif not ccd:
            I["Zijmb"] = self._zgemm(A, i0, i1)
        K.PHASES.mark(None)
        return I

    def _zgemm(self, A, i0, i1):
        """Z'[i,j,m_g,b] = Zmbij[m,b,i,j] = <mb|ef> tau_ijef (ccwfn.py:715) for the slabs m in [i0,i1); ``A``: the dict of
        _amps (tau, and T+- when tau is pair-symmetric: Z in pair form like the ladder -- S = T+ X+^T, A = T- X-^T over
        the rows (i >= j) with X+-[(m,b),Q] = <mb|ef> +- <mb|fe> packed once per Hamiltonian; Z[i,j] = S + A,
        Z[j,i] = S - A: half the flops)."""
        no, nv = self.no, self.nv
        ni = i1 - i0
        if "Tpm" in A and self.pair_z:
            T = A["Tpm"]
            M, ldq = T.shape[1], T.shape[2]
            X = self._ovvv_packed(i0, i1)
            ncols = ni * nv
            lds = (ncols + 1) // 2 * 2
            SA = torch.empty((2, M, lds), dtype=F64, device=self.device1)
            K.dgemm(M, ncols, K.pair_count(nv), T, ldq, 0, X, ldq, 0, SA, lds, 1.0, 0.0,
                    batch=2, sA=M * ldq, sB=ncols * ldq, sC=M * lds)
            Z = torch.empty((no, no, ni, nv), dtype=F64, device=self.device1)
            return K.pair_rows_unpack(SA[0], SA[1], lds, no, ncols, Z, ncols)
        return self._ct("ijef,mbef->ijmb", A["tau"], self.H.block("ovvv")[i0:i1])

    # Z_mbij in pair form when tau is pair-symmetric (solve_cc's iterations); False = always the dense o^3v^3 product
    pair_z = os.environ.get("B200CC_PAIR_Z", "1") != "0"
    # ring terms through {D = W1 + 1/2 W2, W2} (four o^3v^3 GEMMs instead of six) when t2 is pair-symmetric
    fuse_rings = os.environ.get("B200CC_FUSE_RINGS", "1") != "0"

    def _ovvv_packed(self, m0, m1):
        """Constant X+-[2, (m,b), ldq] of the slabs <mb|ef>, m in [m0,m1): built once per Hamiltonian (8.7 GB for all m at
        o=40, v=300; a rank packs only its own slabs)."""
        key = ("ovvv_packed", m0, m1)
        if key not in self.H._derived:
            nv = self.nv
            self.H._derived[key] = K.pack_rows((self.H.block("ovvv"), m0 * nv ** 3), (m1 - m0) * nv, nv)
        return self.H._derived[key]

    def _Loovv_emnf(self, m0, m1):
        """Constant [e, m, n, f] copy of Loovv[m0:m1] (K-major operand of the Fae contraction), built once."""
        key = ("Loovv_emnf", m0, m1)
        if key not in self.H._derived:
            self.H._derived[key] = K.permuted(self.H.derived("Loovv")[m0:m1], (2, 0, 1, 3))
        return self.H._derived[key]

    def _fae_ovvv(self, t1, Fae, m0, m1):
        """Fae += sum_{m in [m0,m1)} sum_f t_mf (2<ma|fe> - <ma|ef>)  (ccwfn.py:496): <mb|ef> streamed in place."""
        no, nv = self.no, self.nv
        nm = m1 - m0
        ovvv = self.H.block("ovvv")
        tmp = torch.empty((nm, nv, nv), dtype=F64, device=self.device1)
        # tmp[m,a,e] = - sum_f <ma|ef> t_mf : batch m, (a,e) x f  times  f x 1
        K.dgemm(nv * nv, 1, nv, (ovvv, m0 * nv ** 3), nv, 0, (t1, m0 * nv), nv, 0, tmp, 1, -1.0, 0.0,
                batch=nm, sA=nv ** 3, sB=nv, sC=nv * nv)
        # tmp[m,a,e] += 2 sum_f t_mf <ma|fe> : per m, batch a: (1 x f) times (f x e)
        for m in range(m0, m1):
            K.dgemm(1, nv, nv, (t1, m * nv), nv, 0, (ovvv, m * nv ** 3), nv, 1, (tmp, (m - m0) * nv * nv), nv,
                    2.0, 1.0, batch=nv, sA=0, sB=nv * nv, sC=nv)
        ones = torch.ones(nm, dtype=F64, device=self.device1)
        self._ct("m,mae->ae", ones, tmp, out=Fae, alpha=1.0, beta=1.0)

    # ---- r1 (ccwfn.py:754-760), replicated on every rank ------------------------------------------------
    def _r1(self, F, t1, t2, I, r1p):
        """Returns the replicated (cheap) part of r1; the two o^2v^3 / o^3v^2 sums over an occupied index m are
        evaluated for this rank's m in [i0,i1) only and written to ``r1p`` (summed by the r2 all-reduce)."""
        H, ct = self.H, self._ct
        o, v, no, nv = self.o, self.v, self.no, self.nv
        A = I["amps"]
        i0, i1 = I["occ"]
        ni = i1 - i0
        r1 = K.permuted(F[v, o], (1, 0))                                        # f_ai
        r1p.zero_()
        if self.model == 'CCD':
            return r1
        ct("ie,ae->ia", t1, I["Fae"], out=r1, alpha=1.0, beta=1.0)
        ct("mi,ma->ia", I["Fmi"], t1, out=r1, alpha=-1.0, beta=1.0)
        ct("iame,me->ia", A["s_iame"], I["Fme"], out=r1, alpha=1.0, beta=1.0)
        # t_nf L_nafi = 2 t_nf <in|af> - t_nf <if|na>
        ct("ianf,nf->ia", H.derived("oovv_menf"), t1, out=r1, alpha=2.0, beta=1.0)
        t1T = K.permuted(t1, (1, 0))
        ovov = H.block("ovov")
        K.dgemm(nv, 1, nv * no, ovov, nv, 1, t1T, nv * no, 0, r1, 1, -1.0, 1.0,
                batch=no, sA=nv * no * nv, sB=0, sC=nv)
        if ni == 0:
            return r1
        # (2 t2 - t2^T)_mief <ma|ef> : per m a (o x v^2)(v^2 x v) product, summed over this rank's m
        t2m = t2[i0:i1]
        s_mief = torch.empty_like(t2m)
        K.strided_axpby(s_mief, t2m, 2.0, 0.0)
        K.strided_axpby(s_mief, t2m.permute(0, 1, 3, 2), -1.0, 1.0)
        tmp = torch.empty((ni, no, nv), dtype=F64, device=self.device1)
        K.dgemm(no, nv, nv * nv, s_mief, nv * nv, 0, (H.block("ovvv"), i0 * nv ** 3), nv * nv, 0, tmp, nv, 1.0, 0.0,
                batch=ni, sA=no * nv * nv, sB=nv ** 3, sC=no * nv)
        ones = torch.ones(ni, dtype=F64, device=self.device1)
        ct("m,mia->ia", ones, tmp, out=r1p, alpha=1.0, beta=1.0)
        del s_mief, tmp
        # - t2_mnae L_nmei,  L_nmei = 2<mn|ie> - <nm|ie> = Looov[m,n,i,e]
        ct("mnae,mnie->ia", t2m, H.derived("Looov")[i0:i1], out=r1p, alpha=-1.0, beta=1.0)
        return r1

    # ---- CC2 (ccwfn.py:596-602, 711-713, 832-884): doubles residual with bare-Fock Fae/Fmi and t1-only W / Z ----------
    def _cc2_Wmnij(self, t1):
        """<mn|ij> + t_je <mn|ie> + t_ie <mn|ej> + t_ie t_jf <mn|ef>                      (ccwfn.py:596-602)"""
        ct = self._ct
        W = K.permuted(self._E('oooo'), (0, 1, 2, 3))
        ct('je,mnie->mnij', t1, self._E('ooov'), out=W, alpha=1.0, beta=1.0)
        ct('ie,mnej->mnij', t1, self._E('oovo'), out=W, alpha=1.0, beta=1.0)
        return ct('mnif,jf->mnij', ct('mnef,ie->mnif', self._E('oovv'), t1), t1, out=W, alpha=1.0, beta=1.0)

    def _cc2_Zmbij(self, t1):
        """<mb|ef> t_ie t_jf                                                              (ccwfn.py:711-713)"""
        return self._ct('mbif,jf->mbij', self._ct('mbef,ie->mbif', self._E('ovvv'), t1), t1)

    def _r2_half_cc2(self, F, t1, t2, half):
        """The unsymmetrised CC2 doubles residual (ccwfn.py:868-881), term by term through the contraction backend.
        t_ie t_jf <ab|ef> is ONE pass over <ab|ef> (M = v^3, N = o, K = v) followed by an o^2v^3 product -- 2 o v^4
        flop, not the 2 o^2 v^4 of the CCSD ladder."""
        ct = self._ct
        o, v = self.o, self.v
        if self.part.size > 1:
            raise NotImplementedError("CC2 is single-GPU: its doubles terms are not rank-partitioned and t_ie t_jf <ab|ef> "
                                      "needs the whole <ab|ef> block")
        Fov = F[o, v]
        K.strided_axpby(half, self.H.block("oovv"), 0.5, 0.0)                       # 1/2 <ab|ij>
        # t2_ijae (f_be - 1/2 f_me t_mb) - 1/2 t2_ijae f_me t_mb  =  t2_ijae (f_be - f_me t_mb)
        Y = K.permuted(F[v, v], (0, 1))
        ct('mb,me->be', t1, Fov, out=Y, alpha=-1.0, beta=1.0)
        ct('ijae,be->ijab', t2, Y, out=half, alpha=1.0, beta=1.0)
        # - t2_imab (f_mj + 1/2 f_me t_je) - 1/2 t2_imab f_me t_je  =  - t2_imab (f_mj + f_me t_je)
        X = K.permuted(F[o, o], (0, 1))
        ct('me,je->mj', Fov, t1, out=X, alpha=1.0, beta=1.0)
        ct('imab,mj->ijab', t2, X, out=half, alpha=-1.0, beta=1.0)
        ct('ma,mbij->ijab', t1, ct('nb,mnij->mbij', t1, self._cc2_Wmnij(t1)), out=half, alpha=0.5, beta=1.0)
        if self._vvvv_released():
            # sum_e t_ie <ab|ef> = X[b,a,f,i] with X_abei = sum_f t_if <ab|ef>   (<ab|ef> = <ba|fe>)
            Y = torch.zeros((self.nv, self.nv, self.no, self.nv), dtype=F64, device=self.device1)       # [a,b,i,f]
            self._t1_vvvv(t1, Y.permute(1, 0, 3, 2))
        else:
            Y = ct('ie,abef->abif', t1, self._E('vvvv'))
        ct('jf,abif->ijab', t1, Y, out=half, alpha=0.5, beta=1.0)
        ct('ma,mbij->ijab', t1, self._cc2_Zmbij(t1), out=half, alpha=-1.0, beta=1.0)
        ct('ma,mbij->ijab', t1, ct('ie,mbej->mbij', t1, self._E('ovvo')), out=half, alpha=-1.0, beta=1.0)
        ct('mb,maji->ijab', t1, ct('ie,maje->maji', t1, self._E('ovov')), out=half, alpha=-1.0, beta=1.0)
        ct('ie,abej->ijab', t1, self._E('vvvo'), out=half, alpha=1.0, beta=1.0)
        ct('ma,mbij->ijab', t1, self._E('ovoo'), out=half, alpha=-1.0, beta=1.0)
        return half

    # ---- r2, unsymmetrised half (ccwfn.py:922-940): this rank's share -------------------------------------
    def _r2_half(self, F, t1, t2, I, r2=None, symmetric=False, ladder=True, rings=True):
        H, ct = self.H, self._ct
        o, v, no, nv = self.o, self.v, self.no, self.nv
        A = I["amps"]
        ccd = self.model == 'CCD'
        i0, i1 = I["occ"]
        ni = i1 - i0
        oovv = H.block("oovv")
        whole = ni == no
        if r2 is None:
            r2 = torch.empty_like(t2)
        if not whole:
            r2.zero_()
        # 1/2 tau_ijef <ab|ef>  -- the particle-particle ladder, local a rows, all (i,j)      931
        if ni > 0:
            K.strided_axpby(r2[i0:i1], oovv[i0:i1], 0.5, 0.0)                      # 1/2 <ab|ij>       922
        if ladder:                                   # (the complex path evaluates the ladder on the two planes of tau)
            with K.PHASES("ladder"):
                self._ladder(A["tau"], r2, symmetric=symmetric, T=A.get("Tpm") if symmetric else None)
        if ni == 0:
            return r2
        rg = r2[i0:i1]                                                             # rows i_g (contiguous)
        t2g, t1g = t2[i0:i1], t1[i0:i1]
        K.PHASES.mark("r2: F and Wmnij terms")
        # t2_ijae (F_be - 1/2 t_mb F_me)                                           923-925
        Fx = I["Fae"]
        if not ccd:
            Fx = K.permuted(Fx, (0, 1))
            ct("mb,me->be", t1, I["Fme"], out=Fx, alpha=-0.5, beta=1.0)
        ct("ijae,be->ijab", t2g, Fx, out=rg, alpha=1.0, beta=1.0)
        # - t2_imab (F_mj + 1/2 t_je F_me)                                         926-928
        Fy = I["Fmi"]
        if not ccd:
            Fy = K.permuted(Fy, (0, 1))
            ct("je,me->mj", t1, I["Fme"], out=Fy, alpha=0.5, beta=1.0)
        K.dgemm(no, nv * nv, no, Fy, no, 1, t2g, nv * nv, 1, rg, nv * nv, -1.0, 1.0,
                batch=ni, sA=0, sB=no * nv * nv, sC=no * nv * nv)
        # 1/2 tau_mnab W_mnij                                                       930
        ct("mnij,mnab->ijab", I["Wmnij"], A["tau"], out=rg, alpha=0.5, beta=1.0)
        # ring terms, columns j_g, in [i,a,j,b] layout                              933-935
        K.PHASES.mark("r2: ring terms (three o3v3 GEMMs)")
        t2_jame = None
        if rings:
            t2_jame = A["t2_jame"] if "t2_jame" in A else K.permuted(t2, (1, 2, 0, 3))  # [j,a,m,e] = t2[m,j,a,e]
        if not rings:
            R = None                             # (the complex path evaluates the ring products on the planes)
        elif "D" in I:
            # pair-symmetric t2: with u = 2t2 - t2^T and t2 = (u + t2^T)/2, lines 933-935 are
            #   u.(W1 + 1/2 W2) + 1/2 X[i,a,j,b] + X[j,a,i,b],   X[x,a,y,b] = sum_me t2[m,x,a,e] W_mbye
            # (t2[i,m,e,a] = t2[m,i,a,e]): TWO o^3v^3 GEMMs instead of three -- the closed-shell (1/2 + P_ij) form
            R = ct("iame,jbme->iajb", A["s_iame"], I["D"])
            X = ct("jame,ibme->jaib", t2_jame, I["W2"])                            # [x, a, y in j_g, b]
            K.strided_axpby(R, X, 0.5, 1.0)
            K.strided_axpby(r2[:, i0:i1], R.permute(0, 2, 1, 3), 1.0, 1.0)
            K.strided_axpby(rg, X.permute(2, 0, 1, 3), 1.0, 1.0)
            del X
        else:
            R = ct("iame,jbme->iajb", A["s_iame"], I["W1"])          # (2t2 - t2^T) W_mbej
            ct("iame,jbme->iajb", A["t2_iame"], I["W2"], out=R, alpha=1.0, beta=1.0)   # t2 W_mbje^T
            K.strided_axpby(r2[:, i0:i1], R.permute(0, 2, 1, 3), 1.0, 1.0)
            ct("jame,ibme->jaib", t2_jame, I["W2"], out=R, alpha=1.0, beta=0.0)        # t2_mjae W_mbie
            K.strided_axpby(rg, R.permute(2, 0, 1, 3), 1.0, 1.0)
        del R, t2_jame
        K.PHASES.mark("r2: t1 terms (Z, <mb|ej>, <ma|je>, <ab|ej>)")
        if not ccd:
            ooov, ovov = H.block("ooov"), H.block("ovov")
            # - t_ma ( Z_mbij + <mb|ij> + t_ie <mb|ej> )  as one batched product    932, 940, 936-937
            # for this rank's m in [i0,i1) and ALL rows (i,j): a partial sum over m, completed by the all-reduce of r2
            Zs = I["Zijmb"]                                                        # [i, j, m_g, b]
            K.strided_axpby(Zs, ooov[:, :, i0:i1, :], 1.0, 1.0)                    # <mb|ij> = ooov[i,j,m,b]
            # Y1[i,j,m,b] = sum_e t_ie <jm|be>: batch j, N = (m_g, b)
            K.dgemm(no, ni * nv, nv, t1, nv, 0, (oovv, i0 * nv * nv), nv, 0, Zs, no * ni * nv, 1.0, 1.0,
                    batch=no, sA=0, sB=no * nv * nv, sC=ni * nv)
            K.dgemm(nv, nv, ni, (t1, i0 * nv), nv, 1, Zs, nv, 1, r2, nv, -1.0, 1.0,
                    batch=no * no, sA=0, sB=ni * nv, sC=nv * nv)
            # - t_ie t_mb <ma|je>                                                    938
            Y2 = torch.empty((ni, no, no, nv), dtype=F64, device=self.device1)       # [i,j,m,a]
            K.dgemm(ni, no * nv, nv, t1g, nv, 0, ovov, no * nv, 0, Y2, no * no * nv, 1.0, 0.0,
                    batch=no, sA=0, sB=nv, sC=no * nv)
            K.dgemm(nv, nv, no, Y2, nv, 1, t1, nv, 1, rg, nv, -1.0, 1.0,
                    batch=ni * no, sA=no * nv, sB=0, sC=nv * nv)
            # t_ie <ab|ej> = sum_e t_ie <ja|be> = P1[j,a,b,i]: all rows i, this rank's COLUMNS j (its own slabs)      939
            K.strided_axpby(r2[:, i0:i1], I["P1"].permute(3, 0, 1, 2), 1.0, 1.0)
        K.PHASES.mark(None)
        return r2

    def _ladder(self, tau, r2, symmetric=False, T=None):
        """r2[i,j,a,b] += 1/2 sum_ef tau[i,j,e,f] <ab|ef>  (ccwfn.py:931) in symmetric / antisymmetric pair form
        (csrc/pairs.cu): T+- = (tau_ef +- tau_fe)/2 over pairs (e >= f), S = T+ V+^T and A = T- V-^T as ONE batched
        GEMM (batch 2, N = K = v(v+1)/2), then r2[..,a,b] += (S+A)/2, r2[..,b,a] += (S-A)/2.  ``symmetric``: tau is
        pair-symmetric, rows (i >= j) only.  Executed flop: o^2 v^4 (general) or o^2 v^4 / 2 instead of 2 o^2 v^4;
        V+- (v^4/2 doubles, a-sharded over ranks by pair count) is streamed once, in place."""
        no, nv = self.no, self.nv
        H = self.H
        r_lo, r_hi = H.a_range                      # rows resident on this device
        a_lo, a_hi = self.part.a_range(nv)          # rows this rank is responsible for
        if a_lo < r_lo or a_hi > r_hi:
            raise B200ccError("<ab|ef> rows [%d,%d) needed but only [%d,%d) are resident" % (a_lo, a_hi, r_lo, r_hi))
        if a_hi == a_lo:
            return
        nq = K.pair_count(nv)
        npl = K.pair_count(a_hi) - K.pair_count(a_lo)           # pairs (columns of S / A) this rank computes
        row0 = K.pair_count(a_lo) - K.pair_count(r_lo)           # first of them among the resident packed rows
        nres = H.npairs_local
        tri = bool(symmetric)
        if T is None:
            T = K.pack_tau(tau, tri)                             # [2, M, ldq]
        M, ldq = T.shape[1], T.shape[2]
        lds = (npl + 1) // 2 * 2
        SA = torch.empty((2, M, lds), dtype=F64, device=self.device1)
        self.ladder_flops = 2.0 * 2 * M * npl * nq               # executed by the last call (bench roofline)
        if K.MIXED.on:
            # precision='MP': V+- live as TF32 planes [2, nres, ldp]; T+- are split per call (a 0.6-1.2 GB pass)
            hi, lo, ldp = H.to_mixed(drop=False)
            th, tl, lpt = K.split_tf32(T, 2 * M, nq, ldq)
            del T
            # One launch over all rows of a 30+ GB operand runs ~25 % slower than the same work in row slices of a few
            # thousand rows (profiles/mp_ladder_slices_r01.json): a launch then sweeps a few GB of address space.
            nsl = max(1, min(16, npl // 4096))
            rows = -(-npl // nsl)
            rows = -(-rows // 128) * 128
            for r0 in range(0, npl, rows):
                n = min(rows, npl - r0)
                off = (row0 + r0) * ldp
                K.gemm_tf32x3(M, n, nq, th, tl, lpt, (hi, off), (lo, off), ldp, (SA, r0), lds, 1.0, 0.0,
                              batch=2, sA=M * lpt, sB=nres * ldp, sC=M * lds)
        else:
            V, ldv = H.packed()
            K.dgemm(M, npl, nq, T, ldq, 0, (V, row0 * ldv), ldv, 0, SA, lds, 1.0, 0.0,
                    batch=2, sA=M * ldq, sB=nres * ldv, sC=M * lds)
        K.ladder_unpack(SA[0], SA[1], lds, no, nv, tri, a_lo, a_hi, 0.5, r2)

    # =============================================================================================
    # the reference's public building blocks, reference layouts (used by tests and downstream code)
    # =============================================================================================
    def build_tau(self, t1, t2, fact1=1.0, fact2=1.0):
        return K.build_tau(t1.contiguous(), t2.contiguous(), fact1, fact2)

    def _own(self, ERI=None, L=None):
        """The fused (block) formulation only applies to the wavefunction's own integrals; CC2 / CC3 have no generic
        path (the CCSD / CCD methods dispatch on ``_foreign`` before they get here)."""
        if self._foreign(ERI, L):
            raise NotImplementedError("swapped-in (perturbed) integrals with model %r are outside the accelerated path; "
                                      "use the wavefunction's own H.ERI / H.L" % self.model)

    def _foreign(self, ERI=None, L=None):
        """True when the integrals to use are not this wavefunction's own block views: a caller passed other objects,
        or swapped ``self.H.ERI`` / ``self.H.L`` (ccderiv.py:250-259 does, around ``residuals``)."""
        H = self.H
        for given, kind in ((ERI, "ERI"), (L, "L")):
            cur = getattr(H, kind)
            own = isinstance(cur, _BlockView) and cur.H is H and cur.kind == kind
            if not own or (given is not None and given is not cur):
                return True
        return False

    def _generic(self, ERI=None, L=None):
        """The term-by-term evaluator for foreign integrals (generic.py); L defaults to 2 ERI - ERI^T(rs) of the ERI in
        use when only ERI is foreign (hamiltonian.py:70)."""
        from .generic import GenericResidual
        if self.model not in ('CCSD', 'CCSD(T)', 'CCD'):
            raise NotImplementedError("swapped-in (perturbed) integrals with model %r are outside the accelerated path"
                                      % self.model)
        ERI = self.H.ERI if ERI is None else ERI
        L = self.H.L if L is None else L
        return GenericResidual(self, ERI, L)

    def _I(self, F, t1, t2):
        with K.mixed_mode(self.mixed):
            return self._intermediates(self._check_F(F), t1.contiguous(), t2.contiguous(), full=True)

    def build_Fae(self, o, v, F, L, t1, t2):
        if self._foreign(L=L) and self.model != 'CC2':
            return self._generic(L=L).intermediates(F, t1, t2)["Fae"]
        self._own(L=L)
        return self._I(F, t1, t2)["Fae"]

    def build_Fmi(self, o, v, F, L, t1, t2):
        if self._foreign(L=L) and self.model != 'CC2':
            return self._generic(L=L).intermediates(F, t1, t2)["Fmi"]
        self._own(L=L)
        return self._I(F, t1, t2)["Fmi"]

    def build_Fme(self, o, v, F, L, t1):
        if self.model == 'CCD':
            return None
        if self._foreign(L=L) and self.model != 'CC2':
            return self._generic(L=L).intermediates(F, t1, self.t2)["Fme"]
        self._own(L=L)
        F = self._check_F(F)
        Fme = K.permuted(F[self.o, self.v], (0, 1))
        self._ct("menf,nf->me", self.H.derived("Loovv_menf"), t1.contiguous(), out=Fme, alpha=1.0, beta=1.0)
        return Fme

    def build_Wmnij(self, o, v, ERI, t1, t2):
        if self._foreign(ERI) and self.model != 'CC2':
            return self._generic(ERI).intermediates(self.H.F, t1, t2)["Wmnij"]
        self._own(ERI)
        if self.model == 'CC2':
            return self._cc2_Wmnij(t1.contiguous())
        return self._I(self.H.F, t1, t2)["Wmnij"]

    def build_Wmbej(self, o, v, ERI, L, t1, t2):
        if self._foreign(ERI, L) and self.model != 'CC2':
            return self._generic(ERI, L).intermediates(self.H.F, t1, t2)["Wmbej"]
        self._own(ERI, L)
        if self.model == 'CC2':
            return None                                                           # ccwfn.py:638-639
        return K.permuted(self._I(self.H.F, t1, t2)["W1"], (2, 1, 3, 0))          # [j,b,m,e] -> [m,b,e,j]

    def build_Wmbje(self, o, v, ERI, t1, t2):
        if self._foreign(ERI) and self.model != 'CC2':
            return self._generic(ERI).intermediates(self.H.F, t1, t2)["Wmbje"]
        self._own(ERI)
        if self.model == 'CC2':
            return None                                                           # ccwfn.py:677-678
        return K.permuted(self._I(self.H.F, t1, t2)["W2"], (2, 1, 0, 3))          # [j,b,m,e] -> [m,b,j,e]

    def build_Zmbij(self, o, v, ERI, t1, t2):
        if self.model == 'CCD':
            return None
        if self._foreign(ERI) and self.model != 'CC2':
            return self._generic(ERI).intermediates(self.H.F, t1, t2)["Zmbij"]
        self._own(ERI)
        if self.model == 'CC2':
            return self._cc2_Zmbij(t1.contiguous())
        return K.permuted(self._I(self.H.F, t1, t2)["Zijmb"], (2, 3, 0, 1))       # [i,j,m,b] -> [m,b,i,j]

    def r_T1(self, o, v, F, ERI, L, t1, t2, Fae=None, Fme=None, Fmi=None):
        """T1 residual (ccwfn.py:718-761).  The intermediates are rebuilt internally in the fused layouts;
        the Fae/Fme/Fmi arguments are accepted for signature compatibility."""
        if self._foreign(ERI, L) and self.model != 'CC2':
            return self._generic(ERI, L).residuals(F, t1, t2)[0]
        self._own(ERI, L)
        F = self._check_F(F)
        t1, t2 = t1.contiguous(), t2.contiguous()
        r1p = torch.empty_like(t1)
        with K.mixed_mode(self.mixed):
            r1 = self._r1(F, t1, t2, self._intermediates(F, t1, t2, rings=self.model != 'CC2'), r1p)
        if self.part.size > 1:
            self.part.all_reduce_sum(r1p)
        K.strided_axpby(r1, r1p, 1.0, 1.0)
        if self.model == 'CCD':
            r1.zero_()
        return r1

    def r_T2(self, o, v, F, ERI, t1, t2, Fae=None, Fme=None, Fmi=None, Wmnij=None, Wmbej=None, Wmbje=None,
             Zmbij=None):
        """Symmetrised T2 residual (ccwfn.py:764-791)."""
        if self._foreign(ERI) and self.model != 'CC2':
            return K.symmetrize_r2(self._generic(ERI).residuals(F, t1, t2)[1])
        self._own(ERI)
        F = self._check_F(F)
        t1, t2 = t1.contiguous(), t2.contiguous()
        with K.mixed_mode(self.mixed):
            if self.model == 'CC2':                     # ccwfn.py:786-787 -> _r_T2_cc2 (832-884)
                half = self._r2_half_cc2(F, t1, t2, torch.empty_like(t2))
            else:
                half = self._r2_half(F, t1, t2, self._intermediates(F, t1, t2))
        if self.part.size > 1:
            self.part.all_reduce_sum(half)
        return K.symmetrize_r2(half)

    def cc_energy(self, o, v, F, L, t1, t2):
        """E = 2 f_ia t_ia + tau_ijab L_ijab as a 0-d device tensor (ccwfn.py:1156-1162); CCD: t1 = 0."""
        self._own(L=L)
        F = self._check_F(F)
        e = K.cc_energy(F[self.o, self.v], t1.contiguous(), t2.contiguous(), self.H.derived("Loovv"))
        return e[0]


ccwfn = CCwfn      # the reference exports both names (ccwfn.py:1845)
