"""Device-resident MO integrals for the closed-shell CC path.

The reference builds the full n^4 ``ERI`` and ``L = 2*ERI - ERI.swapaxes(2,3)`` on the host
(pycc/hamiltonian.py:67-70) and re-uploads slices per contraction (pycc/device.py:70-74).
Here only the SIX unique Dirac blocks of a real, 8-fold-symmetric integral tensor live in HBM,

    oooo[m,n,i,j]  ooov[m,n,i,e]  oovv[m,n,e,f]  ovov[m,b,j,e]  ovvv[m,b,e,f]  vvvv[a,b,e,f]

(107 GB x 2 on the host at o=40,v=300 becomes 75 GB on the device, `vvvv` optionally a-sharded),
and ``H.ERI[o,v,v,o]`` / ``H.L[o,o,v,v]`` style slicing -- the data contract of the reference's
``Hamiltonian`` (attributes F, eps, ERI, L) -- is served lazily from those blocks through the
symmetry relations <pq|rs> = <qp|sr> = <rs|pq> = <rq|ps> = <ps|rq> (+ products).
"""
from __future__ import annotations

import itertools

import numpy as np
import torch

from . import kernels as K
from ._lib import B200ccError
from .contract import Contractor

STORED = ("oooo", "ooov", "oovv", "ovov", "ovvv", "vvvv")

# the 8 index permutations that leave a real <pq|rs> invariant, as position maps new[t] = old[g[t]]
_SYM = set()
_gens = [(2, 1, 0, 3), (0, 3, 2, 1), (1, 0, 3, 2)]
_frontier = [(0, 1, 2, 3)]
while _frontier:
    g = _frontier.pop()
    if g in _SYM:
        continue
    _SYM.add(g)
    for h in _gens:
        _frontier.append(tuple(g[h[t]] for t in range(4)))
_SYM = sorted(_SYM)
assert len(_SYM) == 8


def block_source(pattern):
    """('stored name', perm) such that ERI[pattern] == stored.permute(perm)."""
    for g in _SYM:
        # new[t] = old[g[t]]  =>  pattern[t] = name[g[t]]
        for name in STORED:
            if all(pattern[t] == name[g[t]] for t in range(4)):
                return name, g
    raise B200ccError("no stored block for ERI pattern %r" % pattern)


class _BlockView:
    """``H.ERI`` / ``H.L``: supports ``X[o,v,v,o]`` with the wavefunction's own o/v slices."""

    def __init__(self, H, kind):
        self.H, self.kind = H, kind
        self._cache = {}

    def _pattern(self, key):
        if not (isinstance(key, tuple) and len(key) == 4):
            raise B200ccError("block integrals are indexed as X[o,v,v,o] with the wavefunction's o/v slices")
        pat = ""
        for s in key:
            if s == self.H.o:
                pat += "o"
            elif s == self.H.v:
                pat += "v"
            else:
                raise B200ccError("only the active occupied/virtual slices are available on the device "
                                  "(got %r)" % (s,))
        return pat

    def __getitem__(self, key):
        pat = self._pattern(key)
        if self.kind == "ERI":
            name, g = block_source(pat)
            return self.H.block(name).permute(*g)
        if pat not in self._cache:
            # L_pqrs = 2<pq|rs> - <pq|sr>
            a = self.H.ERI[key]
            swapped = (key[0], key[1], key[3], key[2])
            b = self.H.ERI[swapped].permute(0, 1, 3, 2)
            out = torch.empty(tuple(a.shape), dtype=a.dtype, device=a.device)
            K.strided_axpby(out, a, 2.0, 0.0)
            K.strided_axpby(out, b, -1.0, 1.0)
            self._cache[pat] = out
        return self._cache[pat]


class _ConstDict(dict):
    """``H._derived``: every tensor stored here is registered as a constant GEMM operand (its TF32 planes are cached)."""

    def __init__(self, H):
        super().__init__()
        self._H = H

    def __setitem__(self, key, t):
        super().__setitem__(key, t)
        K.register_constant(t, self._H._split_cache)


class BlockHamiltonian:
    """F, eps and the six integral blocks on one device (float64)."""

    def __init__(self, F, blocks, no, nfzc=0, device="cuda", a_range=None):
        self.device = torch.device(device)
        F = torch.as_tensor(np.asarray(F) if not isinstance(F, torch.Tensor) else F, dtype=torch.float64)
        self.F = F.to(self.device).contiguous()
        self.nmo = self.F.shape[0]
        self.no, self.nfzc = int(no), int(nfzc)
        self.nv = self.nmo - self.no - self.nfzc
        self.o = slice(self.nfzc, self.nfzc + self.no)
        self.v = slice(self.nfzc + self.no, self.nmo)
        self.eps = torch.diagonal(self.F).clone()
        self._blocks = {}
        # vvvv may hold only rows a in [a_lo, a_hi) (multi-GPU ladder sharding)
        self.a_range = (0, self.nv) if a_range is None else (int(a_range[0]), int(a_range[1]))
        for name in STORED:
            if name in blocks and blocks[name] is not None:
                t = blocks[name]
                if not isinstance(t, torch.Tensor):
                    t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float64))
                self._blocks[name] = t.to(self.device, dtype=torch.float64).contiguous()
        self.ERI = _BlockView(self, "ERI")
        self.L = _BlockView(self, "L")
        self._derived = _ConstDict(self)
        # precision='MP': TF32 (hi, lo) planes of constant GEMM operands, computed once (kernels._split_operand);
        # <ab|ef> as planes [(a,b), ldp] (the FP64 block may then be released, see to_mixed)
        self._split_cache = {}
        self.vvvv_planes = None
        for t in self._blocks.values():
            K.register_constant(t, self._split_cache)

    def __del__(self):
        try:
            K.unregister_constants(self._split_cache)
        except Exception:
            pass

    def block(self, name):
        try:
            return self._blocks[name]
        except KeyError:
            if name == "vvvv" and self.vvvv_planes is not None:
                raise B200ccError("the FP64 <ab|ef> block was released for precision='MP' (only its TF32 planes are "
                                  "resident)")
            raise B200ccError("integral block %r is not resident" % name)

    def to_mixed(self, drop=True):
        """Split <ab|ef> into TF32 planes (precision='MP'); ``drop`` releases the FP64 block (64.8 GB at v=300)."""
        if self.vvvv_planes is None:
            vvvv = self.block("vvvv")
            na, nv = vvvv.shape[0], self.nv
            hi, lo, ldp = K.split_tf32(vvvv, na * nv, nv * nv, nv * nv)
            self.vvvv_planes = (hi[0], lo[0], ldp)
        if drop and "vvvv" in self._blocks:
            # the released storage may be handed to a per-iteration tensor next: it must not look constant any more
            K.unregister_tensor(self._blocks["vvvv"])
            for key in [k for k in self._split_cache if k[0] == self._blocks["vvvv"].data_ptr()]:
                del self._split_cache[key]
            del self._blocks["vvvv"]
        return self.vvvv_planes

    def has(self, name):
        return name in self._blocks

    def release_split_cache(self):
        """Drop the cached TF32 planes of the constant operands (precision='MP'); they are rebuilt on demand."""
        self._split_cache.clear()

    merge_chunk_bytes = 2 << 30        # size of the FP64 row chunks rebuilt from the TF32 planes of <ab|ef>

    def vvvv_fp64_chunks(self, chunk_bytes=None):
        """(a0, a1, FP64 [a1-a0, v, v, v]) over the RESIDENT rows of <ab|ef> (local row numbers): the block itself, or --
        precision='MP' after the FP64 block was released -- chunks rebuilt from its TF32 planes (hi + lo: 2^-22
        relative, the accuracy of the mode).  For the few FP64-only consumers of <ab|ef> outside the ladder
        (t_if <ab|ef> in HBAR / CC2 / CC3: 2ov^4 flop, one pass per call)."""
        if "vvvv" in self._blocks:
            blk = self._blocks["vvvv"]
            yield 0, blk.shape[0], blk
            return
        if self.vvvv_planes is None:
            raise B200ccError("integral block 'vvvv' is not resident")
        hi, lo, ldp = self.vvvv_planes
        nv = self.nv
        na = hi.shape[0] // nv
        chunk_bytes = self.merge_chunk_bytes if chunk_bytes is None else chunk_bytes
        step = int(max(1, min(na, chunk_bytes // max(8 * nv ** 3, 1))))
        for a0 in range(0, na, step):
            a1 = min(na, a0 + step)
            rows = (a1 - a0) * nv
            blk = K.merge_tf32((hi, a0 * nv * ldp), (lo, a0 * nv * ldp), ldp, rows, nv * nv)
            yield a0, a1, blk.view(a1 - a0, nv, nv, nv)

    # ---- constructors ---------------------------------------------------------------------------
    @classmethod
    def from_full(cls, F, ERI, no, nfzc=0, device="cuda"):
        """From a host n^4 Dirac array (small molecules / the reference's own Hamiltonian.ERI)."""
        ERI = np.asarray(ERI)
        n = ERI.shape[0]
        sl = {"o": slice(nfzc, nfzc + no), "v": slice(nfzc + no, n)}
        blocks = {k: np.ascontiguousarray(ERI[sl[k[0]], sl[k[1]], sl[k[2]], sl[k[3]]]) for k in STORED}
        return cls(F, blocks, no, nfzc, device)

    @classmethod
    def from_factor(cls, syn, device="cuda", a_range=None, chunk_bytes=4 << 30, names=STORED, mixed=False):
        """From a factorised synthetic problem (pycc_b200.synthetic): every block is contracted on the
        device with the package's own GEMM, <pq|rs> = scale * sum_P B[P,p,r] B[P,q,s]; `vvvv` in row
        chunks so no temporary larger than ``chunk_bytes`` exists."""
        dev = torch.device(device)
        B = torch.from_numpy(syn.B).to(dev)
        no, nv = syn.no, syn.nv
        ct = Contractor()
        sl = {"o": slice(0, no), "v": slice(no, no + nv)}
        blocks = {}
        for name in names:
            p, q, r, s = (sl[c] for c in name)
            if name != "vvvv":
                blocks[name] = ct("Ppr,Pqs->pqrs", B[:, p, r], B[:, q, s], alpha=syn.scale)
                continue
            a_lo, a_hi = (0, nv) if a_range is None else a_range
            na = a_hi - a_lo
            Bqs = K.permuted(B[:, q, s], (1, 2, 0))                  # [(b,f), P]  K-major
            rows = max(1, min(na, int(chunk_bytes // (8 * nv ** 3))))
            if mixed:
                # precision='MP': the FP64 block is never materialised, each row chunk goes straight to TF32 planes
                ldp = (nv * nv + 3) // 4 * 4
                hi = torch.empty((na * nv, ldp), dtype=torch.float32, device=dev)
                lo = torch.empty((na * nv, ldp), dtype=torch.float32, device=dev)
                out = torch.empty((min(rows, na), nv, nv, nv), dtype=torch.float64, device=dev)
                planes = (hi, lo, ldp)
            else:
                out = torch.empty((na, nv, nv, nv), dtype=torch.float64, device=dev)
            for a0 in range(0, na, rows):
                a1 = min(na, a0 + rows)
                Bpr = K.permuted(B[:, no + a_lo + a0:no + a_lo + a1, r], (1, 2, 0))   # [(a,e), P]
                tmp = ct("aeP,bfP->aebf", Bpr, Bqs, alpha=syn.scale)
                dst = out[:a1 - a0] if mixed else out[a0:a1]
                K.strided_axpby(dst, tmp.permute(0, 2, 1, 3), 1.0, 0.0)
                del tmp, Bpr
                if mixed:
                    K.split_tf32(dst, (a1 - a0) * nv, nv * nv, nv * nv,
                                 out=((hi, a0 * nv * ldp), (lo, a0 * nv * ldp), ldp))
            if not mixed:
                blocks[name] = out
            del out
        H = cls(syn.F, blocks, no, 0, dev, a_range)
        if mixed and "vvvv" in names:
            H.vvvv_planes = planes
        return H

    @classmethod
    def from_ao(cls, F_ao, eri_ao, C, no, nfzc=0, device="cuda", a_range=None, chunk_bytes=4 << 30):
        """AO -> MO integral staging straight into the six device blocks (reference:
        hamiltonian.py:54-70, which asks psi4 for the full n^4 MO array on the host and forms ERI and L there).

        ``F_ao`` (nbf,nbf) AO Fock matrix, ``eri_ao`` (nbf,)*4 AO repulsion integrals in CHEMIST order (mu lam|nu sig)
        (psi4 ``MintsHelper.ao_eri()``), ``C`` (nbf,nmo) MO coefficients in energy order with the frozen core in
        columns [0:nfzc] (wavefunction.py:304-315).  Computes, with the package's own GEMM kernel only,

            F_pq    = C_mu,p F_mu,nu C_nu,q                                            (hamiltonian.py:55-56)
            <pq|rs> = (pr|qs) = C_mu,p C_lam,r C_nu,q C_sig,s (mu lam|nu sig)          (hamiltonian.py:67-68)

        as four quarter transformations per block family.  The first two indices (p,r) are transformed once per
        family -- (o,o) feeds oooo/ooov/ovov, (o,v) feeds oovv/ovvv -- and <ab|ef> is produced in row chunks of a
        (``chunk_bytes``), directly for this rank's ``a_range``: neither the n^4 MO array nor L ever exists, on
        the host or on the device.  The AO array itself is held on the device (nbf^4 doubles)."""
        dev = torch.device(device)
        ct = Contractor()
        def as_dev(x):
            if isinstance(x, torch.Tensor):
                return x.to(dev, dtype=torch.float64).contiguous()
            return torch.as_tensor(np.ascontiguousarray(np.asarray(x), dtype=np.float64)).to(dev)
        C = as_dev(C)
        nbf, nmo = C.shape
        nv = nmo - no - nfzc
        Co = K.permuted(C[:, nfzc:nfzc + no], (0, 1))
        Cv = K.permuted(C[:, nfzc + no:], (0, 1))
        F = ct("mp,mn,nq->pq", C, as_dev(F_ao), C)
        AO = as_dev(eri_ao)
        if tuple(AO.shape) != (nbf,) * 4:
            raise B200ccError("from_ao: eri_ao must have shape (nbf,nbf,nbf,nbf) = %r" % ((nbf,) * 4,))

        def finish(X2, Cq, Cs):
            """X2[p,r,nu,sig] -> <pq|rs>[p,q,r,s]"""
            Y = ct("prns,nq->prqs", X2, Cq)
            Z = ct("prqs,st->prqt", Y, Cs)
            del Y
            return K.permuted(Z, (0, 2, 1, 3))

        blocks = {}
        X1 = ct("mp,mlns->plns", Co, AO)                        # first index -> occupied
        X2 = ct("plns,lr->prns", X1, Co)                        # (p,r) = (o,o)
        blocks["oooo"] = finish(X2, Co, Co)
        blocks["ooov"] = finish(X2, Co, Cv)                     # <mn|ie> = (mi|ne)
        blocks["ovov"] = finish(X2, Cv, Cv)                     # <mb|je> = (mj|be)
        X2 = ct("plns,lr->prns", X1, Cv)                        # (p,r) = (o,v)
        del X1
        blocks["oovv"] = finish(X2, Co, Cv)                     # <mn|ef> = (me|nf)
        blocks["ovvv"] = finish(X2, Cv, Cv)                     # <mb|ef> = (me|bf)
        del X2
        a_lo, a_hi = (0, nv) if a_range is None else (int(a_range[0]), int(a_range[1]))
        na = a_hi - a_lo
        vvvv = torch.empty((na, nv, nv, nv), dtype=torch.float64, device=dev)
        rows = max(1, min(max(na, 1), int(chunk_bytes // (8 * max(nbf, 1) ** 3))))
        for a0 in range(0, na, rows):
            a1 = min(na, a0 + rows)
            Ca = K.permuted(Cv[:, a_lo + a0:a_lo + a1], (0, 1))
            X1 = ct("mp,mlns->plns", Ca, AO)
            X2 = ct("plns,lr->prns", X1, Cv)                    # (a,e)
            del X1
            Y = ct("prns,nq->prqs", X2, Cv)
            del X2
            Z = ct("prqs,st->prqt", Y, Cv)                      # [a,e,b,f]
            del Y
            K.strided_axpby(vvvv[a0:a1], Z.permute(0, 2, 1, 3), 1.0, 0.0)
            del Z
        blocks["vvvv"] = vvvv
        del AO
        return cls(F, blocks, no, nfzc, dev, a_range)

    # ---- derived constant layouts (built once, cached) ---------------------------------------------
    def derived(self, key):
        """Constant rearrangements of the blocks used by the fused residual (see ccwfn.py):
           Loovv       [m,n,e,f] = 2<mn|ef> - <mn|fe>
           Looov       [m,n,i,e] = 2<mn|ie> - <nm|ie>
           oovv_menf   [m,e,n,f] = <mn|ef>         Loovv_menf [m,e,n,f] = Loovv[m,n,e,f]
           oovv_mfne   [m,e,n,f] = <mn|fe>         ovov_mejb  [m,e,j,b] = <mb|je>
        """
        if key in self._derived:
            return self._derived[key]
        oovv = self.block("oovv")
        if key == "Loovv":
            t = torch.empty_like(oovv)
            K.strided_axpby(t, oovv, 2.0, 0.0)
            K.strided_axpby(t, oovv.permute(0, 1, 3, 2), -1.0, 1.0)
        elif key == "Looov":
            ooov = self.block("ooov")
            t = torch.empty_like(ooov)
            K.strided_axpby(t, ooov, 2.0, 0.0)
            K.strided_axpby(t, ooov.permute(1, 0, 2, 3), -1.0, 1.0)
        elif key == "oovv_menf":
            t = K.permuted(oovv, (0, 2, 1, 3))
        elif key == "Loovv_menf":
            t = K.permuted(self.derived("Loovv"), (0, 2, 1, 3))
        elif key == "oovv_mfne":
            t = K.permuted(oovv, (0, 3, 1, 2))
        elif key == "ovov_mejb":
            t = K.permuted(self.block("ovov"), (0, 3, 2, 1))
        else:
            raise B200ccError("unknown derived block %r" % key)
        self._derived[key] = t
        return t
