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
import os

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
        # <ab|ef> in pair-packed form (csrc/pairs.cu): FP64 [2, npl, ldq] (V+, V-) over the pairs (a >= b) of the
        # resident rows a, and / or its TF32 planes (hi, lo: FP32 [2, npl, ldp]) for precision='MP'.  The ladder only
        # ever reads these; the full FP64 block is optional (kept when small, see keep_vvvv_bytes).
        self.vvvv_packed = None
        self.vvvv_planes = None
        self.owned = True          # False: a caller's object (wavefunction.resolve_reference) -- never release its blocks
        for t in self._blocks.values():
            K.register_constant(t, self._split_cache)

    # FP64 <ab|ef> blocks up to this size stay resident next to the packed form (small molecules: ``H.ERI[v,v,v,v]``
    # and the other direct consumers keep working on it); larger ones are released once packed when the Hamiltonian
    # belongs to the wavefunction.  B200CC_VVVV_KEEP_GB overrides.
    keep_vvvv_bytes = int(float(os.environ.get("B200CC_VVVV_KEEP_GB", "4")) * (1 << 30))

    @property
    def npairs_local(self):
        a_lo, a_hi = self.a_range
        return K.pair_count(a_hi) - K.pair_count(a_lo)

    def packed(self, drop=None):
        """(V [2, npl, ldq], ldq): the pair-packed FP64 <ab|ef> of the resident rows, built on first use from the FP64
        block in row chunks.  ``drop``: release the FP64 block afterwards (default: when it exceeds keep_vvvv_bytes)."""
        if self.vvvv_packed is None:
            if self.vvvv_planes is not None:
                raise B200ccError("<ab|ef> is resident as TF32 planes only (precision='MP')")
            vvvv = self.block("vvvv")
            nv = self.nv
            a_lo, a_hi = self.a_range
            ldq = K.pair_ld(nv)
            V = torch.empty((2, self.npairs_local, ldq), dtype=torch.float64, device=self.device)
            step = max(1, int((1 << 30) // max(8 * nv ** 3, 1)))
            for a0 in range(a_lo, a_hi, step):
                a1 = min(a_hi, a0 + step)
                row = K.pair_count(a0) - K.pair_count(a_lo)
                K.pack_pairs(vvvv[a0 - a_lo:a1 - a_lo], nv, a0, a1, (V[0], row * ldq), (V[1], row * ldq), ldq)
            self.vvvv_packed = (V, ldq)
            K.register_constant(V, self._split_cache)
        if drop is None:
            drop = (self.owned and "vvvv" in self._blocks
                    and self._blocks["vvvv"].numel() * 8 > self.keep_vvvv_bytes)
        if drop:
            self._release_fp64_vvvv()
        return self.vvvv_packed

    def _release_fp64_vvvv(self):
        if "vvvv" in self._blocks:
            # the released storage may be handed to a per-iteration tensor next: it must not look constant any more
            K.unregister_tensor(self._blocks["vvvv"])
            for key in [k for k in self._split_cache if k[0] == self._blocks["vvvv"].data_ptr()]:
                del self._split_cache[key]
            del self._blocks["vvvv"]

    def __del__(self):
        try:
            K.unregister_constants(self._split_cache)
        except Exception:
            pass

    def block(self, name):
        try:
            return self._blocks[name]
        except KeyError:
            if name == "vvvv" and (self.vvvv_planes is not None or self.vvvv_packed is not None):
                return self.materialize_vvvv()
            raise B200ccError("integral block %r is not resident" % name)

    def materialize_vvvv(self):
        """The full FP64 <ab|ef> block rebuilt from the pair-packed form (or its TF32 planes) -- for callers that index
        ``H.ERI[v,v,v,v]`` directly.  Only when every row is resident and the block fits keep_vvvv_bytes; the ladder and
        the t1-dressed consumers never need it (they work on pair chunks, vvvv_pair_chunks)."""
        nv = self.nv
        if self.a_range != (0, nv):
            raise B200ccError("<ab|ef> is sharded over ranks (rows %r of %d): the full block is not available"
                              % (self.a_range, nv))
        if 8 * nv ** 4 > self.keep_vvvv_bytes:
            raise B200ccError("the full FP64 <ab|ef> block (%.1f GB) is not kept resident: only its pair-packed form "
                              "is (raise B200CC_VVVV_KEEP_GB, or use H.vvvv_pair_chunks())" % (8 * nv ** 4 / 2 ** 30))
        out = torch.empty((nv, nv, nv, nv), dtype=torch.float64, device=self.device)
        for a0, a1, X in self.vvvv_pair_chunks():
            off = 0
            for a in range(a0, a1):
                K.strided_axpby(out[a, 0:a + 1], X[off:off + a + 1], 1.0, 0.0)
                if a > 0:                                   # <ba|ef> = <ab|fe>
                    K.strided_axpby(out[0:a, a], X[off:off + a].permute(0, 2, 1), 1.0, 0.0)
                off += a + 1
        self._blocks["vvvv"] = out
        K.register_constant(out, self._split_cache)
        return out

    def to_mixed(self, drop=True):
        """TF32 planes (hi, lo, ldp), each FP32 [2, npl, ldp], of the pair-packed <ab|ef> (precision='MP'); ``drop``
        releases the FP64 forms (block and packed)."""
        if self.vvvv_planes is None:
            V, ldq = self.packed(drop=drop)
            npl, nq = V.shape[1], K.pair_count(self.nv)
            hi, lo, ldp = K.split_tf32(V, 2 * npl, nq, ldq)
            self.vvvv_planes = (hi[0].view(2, npl, ldp), lo[0].view(2, npl, ldp), ldp)
        if drop:
            self._release_fp64_vvvv()
            if self.vvvv_packed is not None:
                K.unregister_tensor(self.vvvv_packed[0])
                self.vvvv_packed = None
        return self.vvvv_planes

    def has(self, name):
        return name in self._blocks

    def release_split_cache(self):
        """Drop the cached TF32 planes of the constant operands (precision='MP'); they are rebuilt on demand."""
        self._split_cache.clear()

    merge_chunk_bytes = 2 << 30        # size of the FP64 pair chunks rebuilt from the packed form / its TF32 planes

    def vvvv_pair_chunks(self, chunk_bytes=None):
        """(a0, a1, X) over the RESIDENT rows a of <ab|ef> in chunks of whole rows: X[p, e, f] = <ab|ef> (FP64) for the
        pairs p = pair(a,b) - pair(a0,0), a0 <= a < a1, b <= a.  Built from the pair-packed form (exact up to one
        rounding), or -- precision='MP' -- from its TF32 planes (hi + lo: 2^-22 relative, the accuracy of the mode).
        For the FP64 consumers of <ab|ef> outside the ladder (t_if <ab|ef> in HBAR / CC2 / CC3: 2ov^4 flop, one pass per
        call); the (b, a) image of a pair follows from <ba|ef> = <ab|fe>."""
        nv = self.nv
        a_lo, a_hi = self.a_range
        nq = K.pair_count(nv)
        if self.vvvv_packed is None and self.vvvv_planes is None:
            self.packed()
        chunk_bytes = self.merge_chunk_bytes if chunk_bytes is None else chunk_bytes
        budget = max(1, chunk_bytes // max(8 * nv * nv, 1))           # pairs per chunk
        a0 = a_lo
        while a0 < a_hi:
            a1 = a0 + 1
            while a1 < a_hi and K.pair_count(a1 + 1) - K.pair_count(a0) <= budget:
                a1 += 1
            row = K.pair_count(a0) - K.pair_count(a_lo)
            n = K.pair_count(a1) - K.pair_count(a0)
            if self.vvvv_packed is not None:
                V, ldq = self.vvvv_packed
                X = K.unpack_pairs((V[0], row * ldq), (V[1], row * ldq), ldq, nv, n)
            else:
                hi, lo, ldp = self.vvvv_planes
                tmp = torch.empty((2, n, nq), dtype=torch.float64, device=self.device)
                for s_ in (0, 1):
                    K.merge_tf32((hi[s_], row * ldp), (lo[s_], row * ldp), ldp, n, nq, out=tmp[s_])
                X = K.unpack_pairs(tmp[0], tmp[1], nq, nv, n)
                del tmp
            yield a0, a1, X
            a0 = a1

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
            rows = max(1, min(max(na, 1), int(chunk_bytes // (8 * nv ** 3))))
            # <ab|ef> goes straight into its pair-packed form (FP64 V+/V-, or TF32 planes of them for precision='MP'):
            # each row chunk of the generating GEMM ([a,e,b,f] order) is packed in place through a strided view.  The
            # full FP64 block is kept as well only while it is small (keep_vvvv_bytes).
            nq, ldq = K.pair_count(nv), K.pair_ld(nv)
            npl = K.pair_count(a_hi) - K.pair_count(a_lo)
            keep_full = (not mixed) and 8 * na * nv ** 3 <= cls.keep_vvvv_bytes
            full = torch.empty((na, nv, nv, nv), dtype=torch.float64, device=dev) if keep_full else None
            if mixed:
                ldp = (nq + 3) // 4 * 4
                hi = torch.empty((2, npl, ldp), dtype=torch.float32, device=dev)
                lo = torch.empty((2, npl, ldp), dtype=torch.float32, device=dev)
            else:
                V = torch.empty((2, npl, ldq), dtype=torch.float64, device=dev)
            for a0 in range(a_lo, a_hi, rows):
                a1 = min(a_hi, a0 + rows)
                Bpr = K.permuted(B[:, no + a0:no + a1, r], (1, 2, 0))   # [(a,e), P]
                tmp = ct("aeP,bfP->aebf", Bpr, Bqs, alpha=syn.scale)
                view = tmp.permute(0, 2, 1, 3)                              # indexed [a, b, e, f]
                row = K.pair_count(a0) - K.pair_count(a_lo)
                n = K.pair_count(a1) - K.pair_count(a0)
                if mixed:
                    Vc = torch.empty((2, n, ldq), dtype=torch.float64, device=dev)
                    K.pack_pairs(view, nv, a0, a1, Vc[0], Vc[1], ldq)
                    for s_ in (0, 1):
                        K.split_tf32(Vc[s_], n, nq, ldq, out=((hi[s_], row * ldp), (lo[s_], row * ldp), ldp))
                    del Vc
                else:
                    K.pack_pairs(view, nv, a0, a1, (V[0], row * ldq), (V[1], row * ldq), ldq)
                if full is not None:
                    K.strided_axpby(full[a0 - a_lo:a1 - a_lo], view, 1.0, 0.0)
                del tmp, Bpr, view
            if full is not None:
                blocks[name] = full
        H = cls(syn.F, blocks, no, 0, dev, a_range)
        if "vvvv" in names:
            if mixed:
                H.vvvv_planes = (hi, lo, ldp)
            else:
                H.vvvv_packed = (V, ldq)
                K.register_constant(V, H._split_cache)
        return H

    @classmethod
    def from_ao(cls, F_ao, eri_ao, C, no, nfzc=0, device="cuda", a_range=None, chunk_bytes=4 << 30, stream_ao=None,
                slab_bytes=1 << 30):
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
        the host or on the device.

        The AO array (nbf^4 doubles) is held on the device when it fits next to the blocks.  Otherwise (``stream_ao``:
        None = decide from the free device memory, True / False = force) it STAYS ON THE HOST -- a numpy array or
        ``np.memmap`` -- and every first quarter transformation, the only step that reads it, streams it in slabs of the
        first AO index (``slab_bytes`` each, through two pinned staging buffers so that the copy of slab s+1 overlaps the
        GEMM of slab s) and accumulates  X1[p,lam,nu,sig] += C[mu_slab,p]^T (mu_slab lam|nu sig):  one pass for the
        occupied family, one pass per <ab|ef> row chunk (raise ``chunk_bytes`` for fewer passes).  nbf is then bounded by
        the host memory (or the file), not by HBM."""
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
        if tuple(eri_ao.shape) != (nbf,) * 4:
            raise B200ccError("from_ao: eri_ao must have shape (nbf,nbf,nbf,nbf) = %r" % ((nbf,) * 4,))
        on_device = isinstance(eri_ao, torch.Tensor) and eri_ao.device.type == dev.type == "cuda"
        if stream_ao is None:
            stream_ao = False
            if not on_device and dev.type == "cuda":
                # resident AO array + the <ab|ef> rows + one o nbf^3 intermediate and its successor must fit
                nv_, na_ = nmo - no - nfzc, (nmo - no - nfzc) if a_range is None else int(a_range[1]) - int(a_range[0])
                need = 8 * (nbf ** 4 + na_ * nv_ ** 3 + 2 * max(no, 1) * nbf ** 3) + chunk_bytes
                stream_ao = need > int(torch.cuda.mem_get_info(dev)[0] * 0.9)
        if stream_ao and on_device:
            raise B200ccError("from_ao: stream_ao=True needs the AO array on the host")
        AO = None if stream_ao else as_dev(eri_ao)
        host = None
        if stream_ao:
            host = eri_ao.detach().cpu().numpy() if isinstance(eri_ao, torch.Tensor) else eri_ao
            nslab = max(1, min(nbf, int(slab_bytes // (8 * nbf ** 3))))
            stage = [torch.empty((nslab, nbf, nbf, nbf), dtype=torch.float64,
                                 pin_memory=(dev.type == "cuda")) for _ in range(2)]
            dslab = [torch.empty((nslab, nbf, nbf, nbf), dtype=torch.float64, device=dev) for _ in range(2)]
            copy_stream = torch.cuda.Stream(dev) if dev.type == "cuda" else None
            copied = [None, None]          # event: the H2D copy out of stage[b] / into dslab[b] has completed
            consumed = [None, None]        # event: the GEMM that read dslab[b] has been issued on the compute stream

        def upload(bounds, idx):
            """host rows -> pinned stage[b] (host copy) -> dslab[b] (async, copy stream); returns the slab's event"""
            m0, m1 = bounds[idx]
            b = idx & 1
            if copied[b] is not None:
                copied[b].synchronize()                              # stage[b] is free again
            rows = host[m0:m1]
            if isinstance(rows, np.ndarray) and rows.dtype == np.float64 and rows.flags.writeable \
                    and rows.flags.c_contiguous and not isinstance(rows, np.memmap):
                stage[b][:m1 - m0].copy_(torch.from_numpy(rows))        # multi-threaded host copy (1.6x np.copyto)
            else:
                np.copyto(stage[b][:m1 - m0].numpy(), np.asarray(rows, dtype=np.float64))
            if copy_stream is None:
                dslab[b][:m1 - m0].copy_(stage[b][:m1 - m0])
                return None
            with torch.cuda.stream(copy_stream):
                if consumed[b] is not None:
                    copy_stream.wait_event(consumed[b])              # dslab[b] is free again
                dslab[b][:m1 - m0].copy_(stage[b][:m1 - m0], non_blocking=True)
                copied[b] = torch.cuda.Event()
                copied[b].record(copy_stream)
            return copied[b]

        def first_quarter(Cx):
            """X1[p,lam,nu,sig] = C[mu,p] (mu lam|nu sig): ONE GEMM on the resident AO array, or a sweep over host slabs
            (pinned staging + copy stream: the host copy and the H2D copy of slab s+1 run under the GEMM of slab s)."""
            if not stream_ao:
                return ct("mp,mlns->plns", Cx, AO)
            X1 = torch.zeros((Cx.shape[1], nbf, nbf, nbf), dtype=torch.float64, device=dev)
            bounds = [(m0, min(nbf, m0 + nslab)) for m0 in range(0, nbf, nslab)]
            ready = upload(bounds, 0)
            for idx, (m0, m1) in enumerate(bounds):
                b = idx & 1
                if ready is not None:
                    torch.cuda.current_stream(dev).wait_event(ready)
                ct("mp,mlns->plns", K.permuted(Cx[m0:m1], (0, 1)), dslab[b][:m1 - m0], out=X1, alpha=1.0, beta=1.0)
                if copy_stream is not None:
                    consumed[b] = torch.cuda.Event()
                    consumed[b].record(torch.cuda.current_stream(dev))
                if idx + 1 < len(bounds):
                    ready = upload(bounds, idx + 1)
            return X1

        def finish(X2, Cq, Cs):
            """X2[p,r,nu,sig] -> <pq|rs>[p,q,r,s]"""
            Y = ct("prns,nq->prqs", X2, Cq)
            Z = ct("prqs,st->prqt", Y, Cs)
            del Y
            return K.permuted(Z, (0, 2, 1, 3))

        blocks = {}
        X1 = first_quarter(Co)                                  # first index -> occupied
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
            X1 = first_quarter(Ca)
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
        if stream_ao:
            del stage, dslab
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
