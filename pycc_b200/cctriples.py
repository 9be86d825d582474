"""(T) correction on B200 (reference: pycc/cctriples.py:27-354).

``t_tjl(ccwfn)`` -- the (T) that ``solve_cc`` calls -- is re-built around two kernels:

1. the connected t3 numerator of a *batch* of (i>=j>=k) triples is six two-segment FP64 DMMA GEMMs
   per triple in ONE launch (``b200cc_dgemm`` in address-table mode):
       Q1[(a,b),c] = sum_e <ib|ea>... i.e.  ovvv[i][e,(a,b)]^T t2[k,j][c,e]^T  -  t2[i][m,(a,b)]^T <jk|mc>
   particle term (K = v) and hole term (K = o) chained into the same accumulators; <mb|ef> and t2 are
   read in their natural layouts (no permuted copies), only the small <mn|ie> block is pre-permuted;
2. ``b200cc_t_energy_batch`` reads Q1..Q6 once and does everything else of cctriples.py:208-237 on the
   fly (disconnected part, 1/(1+delta), X3/Y3/Z3, denominators, a>=b>=c reduction).

Triples are independent: with a ``comm`` (parallel.Comm) they are dealt round-robin to the ranks and
the energy is summed with one scalar all-reduce.

The other public names of the reference's triples module that sit on the energy path keep their
signatures: ``t3c_ijk``, ``t3d_ijk`` (generic in the passed W blocks), ``t_vikings`` and
``t_vikings_inverted`` (cross-check formulations, same E(T)).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import kernels as K
from ._lib import B200ccError

F64 = torch.float64
_QCACHE = {}
# t_tjl sums the six t3 products of a triple pairwise inside the GEMM (TriplesEngine(paired=True)); B200CC_T_PAIRED=0
# restores the six-array form
PAIRED = os.environ.get("B200CC_T_PAIRED", "1") != "0"
# 'auto' takes the fused (a,b,c)-driven kernel when it applies: measured on the B200 it is the faster formulation at both
# bench shapes (o=40,v=300: 37 s vs 43 s for the whole job; o=30,v=280: 13.2 s vs 14.4 s; profiles/t_abc_probe_r02.json)
AUTO_ABC = True

# per term q: (occupied index of the <mb|ef> slab and of the t2[x] slab,
#              (p,q) of t2[p,q] in the particle GEMM, (p,q) of Y[p,q] in the hole GEMM)
# as positions into (i,j,k); see DESIGN.md "(T)" for the derivation from cctriples.py:50-62
_TERMS = (
    (0, (2, 1), (1, 2)),   # Q1[(a,b),c]: ovvv[i], t2[k,j] ; t2[i], <jk|mc>
    (0, (1, 2), (2, 1)),   # Q2[(a,c),b]: ovvv[i], t2[j,k] ; t2[i], <kj|mb>
    (2, (1, 0), (0, 1)),   # Q3[(c,a),b]: ovvv[k], t2[j,i] ; t2[k], <ij|mb>
    (2, (0, 1), (1, 0)),   # Q4[(c,b),a]: ovvv[k], t2[i,j] ; t2[k], <ji|ma>
    (1, (0, 2), (2, 0)),   # Q5[(b,c),a]: ovvv[j], t2[i,k] ; t2[j], <ki|ma>
    (1, (2, 0), (0, 2)),   # Q6[(b,a),c]: ovvv[j], t2[k,i] ; t2[j], <ik|mc>
)


def triples_list(no):
    """All (i >= j >= k), in the reference's loop order (cctriples.py:204-206)."""
    return [(i, j, k) for i in range(no) for j in range(i + 1) for k in range(j + 1)]


class TriplesEngine:
    """Owns the constant operands of the (T) GEMMs for one wavefunction state."""

    def __init__(self, ccwfn, t1=None, t2=None, q_bytes=None, use_tma=True, cube_q=False, dressed=None, paired=False,
                 pert=None):
        """``dressed = (Wvvvo, Wovoo)``: build the t3 numerators of cctriples.py:50-62 from these blocks ([a,b,e,i] and
        [m,b,i,j], no permutational symmetry assumed -- the T1-dressed CC3 intermediates) instead of the integrals.
        ``pert = V_ov`` (real-time CC3, with ``dressed``): the numerators also carry the explicit-field term of
        t3_pert_ijk (cctriples.py:694-695), -V_ld t2[i,j,a,d] t2[k,l,c,b].  It has the shape of the fourth hole term
        (cctriples.py:60, -W_ovoo[m,a,j,i] t2[k,m,c,b]) and of no other, so ONLY the Q4 product reads a second copy of the
        hole operand, Y'[j,i,a,m] = Y[j,i,a,m] - V_md t2[i,j,a,d], stacked behind Y: no extra pass over the t3 tile."""
        self.w = ccwfn
        H = ccwfn.H
        self.no, self.nv = ccwfn.no, ccwfn.nv
        self.t1 = (ccwfn.t1 if t1 is None else t1).contiguous()
        self.t2 = (ccwfn.t2 if t2 is None else t2).contiguous()
        self.oovv = H.block("oovv")
        self.fov = H.F[ccwfn.o, ccwfn.v]
        self.dev = self.t2.device
        # TMA path (K-major operands for cp.async.bulk.tensor): needs even o, v (16-byte pitches) and one
        # transposed copy of the particle-term operand,  G[i,a,b,e] = Wvvvo[b,a,e,i]  (= ovvv[i,e,a,b] for the
        # integrals: constant, built once per H)
        self.tma = (self.nv % 2 == 0) and (self.no % 2 == 0) and use_tma
        if dressed is None:
            self.ovvv = H.block("ovvv")
            # Y[j,k,c,m] = -<mc|jk> = -ooov[j,k,m,c]
            self.Y = K.permuted(H.block("ooov"), (0, 1, 3, 2), -1.0)
        else:
            Wvvvo, Wovoo = dressed
            # the natural-layout operand of the non-TMA path is [i][e][(a,b)] = Wvvvo[b,a,e,i]
            self.ovvv = None if self.tma else K.permuted(Wvvvo, (3, 2, 1, 0))
            self.Y = K.permuted(Wovoo, (2, 3, 1, 0), -1.0)        # Y[j,k,c,m] = -Wovoo[m,c,j,k]
        self.ypert = 0
        if pert is not None:
            if dressed is None:
                raise B200ccError("TriplesEngine: the explicit-field term is built for the dressed (CC3) numerators only")
            no, nv = self.no, self.nv
            X = torch.empty((no, no, nv, no), dtype=F64, device=self.dev)           # X[i,j,a,m] = t2[i,j,a,d] V[m,d]
            K.dgemm(no * no * nv, no, nv, self.t2, nv, 0, pert.contiguous(), nv, 0, X, no)
            Ycat = torch.empty((2, no, no, nv, no), dtype=F64, device=self.dev)
            K.strided_axpby(Ycat[0], self.Y, 1.0, 0.0)
            K.strided_axpby(Ycat[1], self.Y, 1.0, 0.0)
            K.strided_axpby(Ycat[1], X.permute(1, 0, 2, 3), -1.0, 1.0)
            self.Y, self.ypert = Ycat, no * no
            del X
        if self.tma:
            if dressed is not None:
                self.G = K.permuted(dressed[0], (3, 1, 0, 2))
            else:
                if "ovvv_iabe" not in H._derived:
                    H._derived["ovvv_iabe"] = K.permuted(self.ovvv, (0, 2, 3, 1))
                self.G = H._derived["ovvv_iabe"]
            self.t2p = K.permuted(self.t2, (0, 2, 3, 1))          # [i,a,b,m] = t2[i,m,a,b]
        # precision='MP': the (T) GEMMs run on the split-TF32 tcgen05 kernel.  Amplitudes and <mc|jk> are constant for
        # the lifetime of this engine, so their TF32 planes are split once (kernels._split_operand cache); close()
        # drops them -- t2 is updated in place by the next CCSD iteration.
        self.mixed = bool(getattr(ccwfn, "mixed", False)) and self.tma
        self._planes = {}
        if self.mixed:
            for t in (self.t2, self.t2p, self.Y):
                K.register_constant(t, self._planes)
        # PAIRED products (the energy path, FP64 TMA kernels): W needs Q1 and Q6 (Q2/Q3, Q4/Q5) at TRANSPOSED row pairs with
        # the same column index, so the GEMM sums each pair in its accumulators -- four K segments per output, the second
        # product reading constant copies of its operands with the pair transposed, GT[i,a,b,e] = G[i,b,a,e] (one more
        # <mb|ef>-sized block per Hamiltonian) and t2pT[i,a,b,m] = t2p[i,b,a,m]: three arrays per triple instead of six go
        # through HBM (48 v^3 bytes written + read per triple instead of 96 v^3), half as many GEMM units with twice the K.
        self.paired = bool(paired) and self.tma and not self.mixed and dressed is None
        if self.paired:
            if "ovvv_ibae" not in H._derived:
                H._derived["ovvv_ibae"] = K.permuted(self.ovvv, (0, 3, 2, 1))
            self.GT = H._derived["ovvv_ibae"]
            self.t2pT = K.permuted(self.t2, (0, 3, 2, 1))         # [i,a,b,m] = t2[i,m,b,a]
        self.nq = 3 if self.paired else 6
        nv = self.nv
        # optional: the TMA GEMM can write Q as contiguous 8x8x8 cubes (4 KB runs for the energy kernel).  Measured
        # on B200 this is SLOWER (1.35 vs 2.66 TB/s in the energy kernel), so the plain (v,v,v) layout is the default.
        self.cube = bool(cube_q) and self.tma and not self.mixed
        self.qsz = K.q_size(nv, self.cube)
        self.qflags = (1 if self.cube else 0) | (2 if self.paired else 0)
        per = self.nq * self.qsz * 8
        if q_bytes is None:
            q_bytes = 8 << 30
            if self.dev.type == "cuda":
                free, _ = torch.cuda.mem_get_info(self.dev)
                q_bytes = max(per, min(q_bytes * 2, int(free * 0.5)))
        self.nb_max = int(max(1, min(q_bytes // max(per, 1), 65535 // 6, 4096)))

    def close(self):
        K.unregister_constants(self._planes)
        self._planes.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def qbuf(self, nb):
        """The Q workspace ([nb][6][v^3] doubles) is kept per device across engines (grow-only), so repeated
        (T) evaluations do not pay a multi-GB cudaMalloc each."""
        need = nb * self.nq * self.qsz
        buf = _QCACHE.get(self.dev)
        if buf is None or buf.numel() < need:
            _QCACHE.pop(self.dev, None)
            buf = None
            buf = torch.empty(need, dtype=F64, device=self.dev)
            _QCACHE[self.dev] = buf
        return buf

    def table(self, trip, Q):
        """int64 [6*nb, 5] device table of {A1, B1, A2, B2, C} addresses for b200cc_dgemm."""
        T = np.asarray(trip, dtype=np.int64).reshape(-1, 3)
        nb = T.shape[0]
        no, nv = self.no, self.nv
        v2, v3 = nv * nv, nv ** 3
        p_ovvv, p_t2, p_Y, p_Q = (x.data_ptr() for x in (self.ovvv, self.t2, self.Y, Q))
        tab = np.empty((nb, 6, 5), dtype=np.int64)
        for q, (x, (p1, q1), (p2, q2)) in enumerate(_TERMS):
            tab[:, q, 0] = p_ovvv + 8 * T[:, x] * v3
            tab[:, q, 1] = p_t2 + 8 * (T[:, p1] * no + T[:, q1]) * v2
            tab[:, q, 2] = p_t2 + 8 * T[:, x] * no * v2
            tab[:, q, 3] = p_Y + 8 * (T[:, p2] * no + T[:, q2] + (self.ypert if q == 3 else 0)) * nv * no
            tab[:, q, 4] = p_Q + 8 * (np.arange(nb) * 6 + q) * self.qsz
        aligned = bool(np.all(tab % 16 == 0))
        return torch.from_numpy(tab.reshape(nb * 6, 5)).to(self.dev), aligned

    def build_q(self, trip, Q=None):
        """Launch the 6*len(trip) two-segment GEMMs; returns the Q buffer ([nb][6][v^3])."""
        nb = len(trip)
        Q = self.qbuf(nb) if Q is None else Q
        no, nv = self.no, self.nv
        if self.paired:
            T = np.asarray(trip, dtype=np.int64).reshape(-1, 3)
            co = np.empty((nb, 3, 8), dtype=np.int32)
            for r, (qa, qb) in enumerate(((0, 5), (1, 2), (3, 4))):       # R1 = Q1 + Q6^T, R2 = Q2 + Q3^T, R3 = Q4 + Q5^T
                for base, q in ((0, qa), (4, qb)):
                    x, (p1, q1), (p2, q2) = _TERMS[q]
                    co[:, r, base + 0] = T[:, x]
                    co[:, r, base + 1] = T[:, p1] * no + T[:, q1]
                    co[:, r, base + 2] = T[:, x]
                    co[:, r, base + 3] = T[:, p2] * no + T[:, q2]
            co = torch.from_numpy(co.reshape(nb * 3, 8)).to(self.dev)
            v2, v3 = nv * nv, nv ** 3
            K.dgemm(v2, nv, nv, self.G, nv, 0, self.t2, nv, 0, Q, nv, 1.0, 0.0, batch=3 * nb,
                    sA=v3, sB=v2, sC=self.qsz, seg2=(self.t2p, no, self.Y, no, no, no * v2, nv * no),
                    seg3=(self.GT, nv, self.t2, nv, nv, v3, v2), seg4=(self.t2pT, no, self.Y, no, no, no * v2, nv * no),
                    bcoords=co, nbatch=(no, no * no) * 4, ksplit=1, out_cube_nv=nv if self.cube else 0)
            return Q
        if self.tma:
            T = np.asarray(trip, dtype=np.int64).reshape(-1, 3)
            co = np.empty((nb, 6, 4), dtype=np.int32)
            for q, (x, (p1, q1), (p2, q2)) in enumerate(_TERMS):
                co[:, q, 0] = T[:, x]
                co[:, q, 1] = T[:, p1] * no + T[:, q1]
                co[:, q, 2] = T[:, x]
                co[:, q, 3] = T[:, p2] * no + T[:, q2] + (self.ypert if q == 3 else 0)
            co = torch.from_numpy(co.reshape(nb * 6, 4)).to(self.dev)
            with K.mixed_mode(self.mixed):
                K.dgemm(nv * nv, nv, nv, self.G, nv, 0, self.t2, nv, 0, Q, nv, 1.0, 0.0, batch=6 * nb,
                        sA=nv ** 3, sB=nv * nv, sC=self.qsz, seg2=(self.t2p, no, self.Y, no, no, no * nv * nv, nv * no),
                        bcoords=co, nbatch=(no, no * no, no, no * no + self.ypert), ksplit=1,
                        out_cube_nv=nv if self.cube else 0,
                        mp_kchunk=512)      # K = v + o fits one FP32 run (<= 64 MMAs: bias < 1e-6 of E(T))
            return Q
        tab, aligned = self.table(trip, Q)
        K.dgemm(nv * nv, nv, nv, self.ovvv, nv * nv, 1, self.t2, nv, 0, Q, nv, 1.0, 0.0,
                batch=6 * nb, seg2=(self.t2, nv * nv, self.Y, no, no, 0, 0),
                table=tab, table_align16=aligned, ksplit=1)
        return Q

    def energy(self, trip):
        """Sum of the Lee-Rendell contributions of the given triples, as a 1-element device tensor."""
        et = torch.zeros(1, dtype=F64, device=self.dev)
        w = self.w
        for s in range(0, len(trip), self.nb_max):
            chunk = trip[s:s + self.nb_max]
            Q = self.build_q(chunk)
            ijk = torch.tensor(np.asarray(chunk, dtype=np.int32).reshape(-1, 3), dtype=torch.int32).to(self.dev)
            K.t_energy_batch(self.no, self.nv, ijk, Q, self.t1, self.t2, self.oovv, self.fov,
                             w.eps_o, w.eps_v, et, accumulate=True, blocked=self.qflags)
        return et

    def t3_parts(self, i, j, k, with_denom):
        """(connected, disconnected) t3 numerators of one triple as (v,v,v) tensors."""
        Q = self.build_q([(i, j, k)])
        return K.t3_assemble(self.no, self.nv, i, j, k, Q, self.t1, self.t2, self.oovv, self.fov,
                             self.w.eps_o, self.w.eps_v, with_denom, blocked=self.qflags)


# ---- fused (a,b,c)-driven (T): the o^3 tile of a virtual triple stays on the chip (csrc/triples_abc.cu) -------------
# 'abc' / 'ijk' force a formulation, 'auto' takes the fused one whenever its kernel applies (o <= 64, v <= 1023, FP64)
ALGO = os.environ.get("B200CC_T_ALGO", "auto")


def _pack3(x0, x1, x2):
    return (x0 | (x1 << 10) | (x2 << 20)).astype(np.int32)


def abc_list(nv):
    """All a >= b >= c except a = b = c (which contributes exactly zero), packed a | b << 10 | c << 20, c fastest --
    consecutive entries share the slabs of a (and mostly b), so the CTAs running side by side meet in L2."""
    out = []
    for a in range(nv):
        b, c = np.tril_indices(a + 1)
        keep = ~((b == a) & (c == a))
        out.append(_pack3(np.full(int(keep.sum()), a, dtype=np.int64), b[keep].astype(np.int64), c[keep].astype(np.int64)))
    return np.concatenate(out) if out else np.zeros(0, dtype=np.int32)


def sorted_occ(no):
    """The occupied triples i >= j >= k (i = j = k left out) the energy phase of a tile visits, packed like abc_list."""
    t = np.asarray([t for t in triples_list(no) if not (t[0] == t[1] == t[2])], dtype=np.int64).reshape(-1, 3)
    return _pack3(t[:, 0], t[:, 1], t[:, 2])


class FusedTriples:
    """(T) in the (a,b,c)-driven form of the reference (t3c_abc / t3d_abc, cctriples.py:75-105, 149-173) with the
    Lee-Rendell bracket of t_tjl (208-237) in exchanged roles: one launch of ``b200cc_t_abc`` does the twelve
    contractions, the disconnected part, the denominators and the energy of every listed (a,b,c) without writing a t3
    array.  Constant operands: ``G[l,x,y,e] = <le|xy>`` (shared with the (i,j,k)-driven engine), ``t2x[x,y,l,m]``,
    ``Ox[z,p,q,m] = -<mz|pq>`` and ``oovvx[x,y,i,j]``."""

    def __init__(self, ccwfn, t1=None, t2=None, grid=None):
        self.w = ccwfn
        H = ccwfn.H
        no, nv = ccwfn.no, ccwfn.nv
        self.no_real, self.nv_real = no, nv
        # The kernel wants even extents (16-byte TMA pitches).  Odd ones are padded by one index whose amplitudes and
        # integrals are zero and whose orbital energy is far away: every t3 numerator it touches is exactly zero, so the
        # padded problem has the same E(T); the padded constants are built once per Hamiltonian.
        self.no, self.nv = no + (no & 1), nv + (nv & 1)
        po, pv = self.no, self.nv
        padded = (po, pv) != (no, nv)
        t1 = (ccwfn.t1 if t1 is None else t1).contiguous()
        t2 = (ccwfn.t2 if t2 is None else t2).contiguous()
        self.dev = t2.device
        fov = H.F[ccwfn.o, ccwfn.v]
        # canonical reference: the f_kc terms of t3d_abc (cctriples.py:161-163) vanish identically and are not evaluated
        self.fov_is_zero = not bool(torch.count_nonzero(fov))

        def const(key, build):
            key = key + ("_pad%dx%d" % (po, pv) if padded else "")
            if key not in H._derived:
                H._derived[key] = build()
            return H._derived[key]

        def embed(shape, src, alpha=1.0):
            """zero tensor of ``shape`` with alpha * (permuted view ``src``) copied into its leading corner"""
            out = torch.zeros(shape, dtype=F64, device=self.dev)
            K.strided_axpby(out[tuple(slice(0, n) for n in src.shape)], src, alpha, 0.0)
            return out

        if not padded:
            self.t1, self.t2, self.fov = t1, t2, fov
            self.eo, self.ev = ccwfn.eps_o, ccwfn.eps_v
            self.G = const("ovvv_iabe", lambda: K.permuted(H.block("ovvv"), (0, 2, 3, 1)))
            self.Ox = const("ooov_zpqm_neg", lambda: K.permuted(H.block("ooov"), (3, 0, 1, 2), -1.0))   # -<mz|pq> = -ooov[p,q,m,z]
            self.oovvx = const("oovv_abij", lambda: K.permuted(H.block("oovv"), (2, 3, 0, 1)))
            self.t2x = K.permuted(t2, (2, 3, 0, 1))
        else:
            self.t1 = embed((po, pv), t1)
            self.t2 = embed((po, po, pv, pv), t2)
            self.fov = embed((po, pv), fov)
            far = 1.0e3 + float(ccwfn.eps_v.abs().max()) + float(ccwfn.eps_o.abs().max())
            self.eo = torch.cat((ccwfn.eps_o, torch.full((po - no,), -far, dtype=F64, device=self.dev)))
            self.ev = torch.cat((ccwfn.eps_v, torch.full((pv - nv,), far, dtype=F64, device=self.dev)))
            self.G = const("ovvv_iabe", lambda: embed((po, pv, pv, pv), H.block("ovvv").permute(0, 2, 3, 1)))
            self.Ox = const("ooov_zpqm_neg", lambda: embed((pv, po, po, po), H.block("ooov").permute(3, 0, 1, 2), -1.0))
            self.oovvx = const("oovv_abij", lambda: embed((pv, pv, po, po), H.block("oovv").permute(2, 3, 0, 1)))
            self.t2x = K.permuted(self.t2, (2, 3, 0, 1))
        self.grid = int(grid) if grid else K.NSM
        self.sorted = torch.from_numpy(sorted_occ(self.no)).to(self.dev)
        self.wtile = torch.empty(self.grid * self.no ** 3, dtype=F64, device=self.dev)
        self.partial = torch.empty(self.grid, dtype=F64, device=self.dev)

    @staticmethod
    def applies(ccwfn):
        po, pv = ccwfn.no + (ccwfn.no & 1), ccwfn.nv + (ccwfn.nv & 1)
        return 2 <= po <= K.t_abc_max_no() and pv <= 1023 and not bool(getattr(ccwfn, "mixed", False))

    def abc_all(self):
        """every virtual triple of the (padded) problem"""
        return abc_list(self.nv)

    def energy(self, abc):
        """Sum of the contributions of the packed virtual triples ``abc`` (numpy int32 or device tensor)."""
        et = torch.zeros(1, dtype=F64, device=self.dev)
        if not isinstance(abc, torch.Tensor):
            abc = torch.from_numpy(np.ascontiguousarray(abc, dtype=np.int32)).to(self.dev)
        if abc.numel() == 0:
            return et
        K.t_abc(self.no, self.nv, abc, self.sorted, self.G, self.t2, self.t2x, self.Ox, self.oovvx, self.t1, self.fov,
                self.eo, self.ev, et, self.wtile, self.partial, self.grid, accumulate=True,
                fov_is_zero=self.fov_is_zero)
        return et

    def w_tile(self, a, b, c):
        """Connected numerator W_abc[i,j,k] (= t3c_abc without denominators) of one virtual triple, as left in the CTA's tile."""
        keep = self.grid
        self.grid = 1
        try:
            self.energy(_pack3(np.asarray([a]), np.asarray([b]), np.asarray([c])))
        finally:
            self.grid = keep
        n = self.no_real
        return self.wtile[:self.no ** 3].view(self.no, self.no, self.no)[:n, :n, :n].clone()


def fused_selected(ccwfn):
    """True when t_tjl(ccwfn) runs the fused (a,b,c)-driven kernel (B200CC_T_ALGO / cctriples.ALGO and its shape limits)."""
    return (ALGO == "abc" or (ALGO == "auto" and AUTO_ABC)) and FusedTriples.applies(ccwfn)


def t_flops_abc(no, nv, nabc=None):
    """FP64 flop EXECUTED by the fused form for ``nabc`` virtual triples (default: the whole job): twelve contractions of
    2 o^3 K flop, K = v (particle) or o (hole), per (a,b,c)."""
    n = nv * (nv + 1) * (nv + 2) // 6 - nv if nabc is None else nabc
    return 12.0 * no ** 3 * (nv + no) * n


def t_tjl_abc(ccwfn, abc=None):
    """E(T) through the fused (a,b,c)-driven kernel; same value as :func:`t_tjl`.  Virtual triples are dealt round-robin
    to the ranks of ``ccwfn.comm``; one scalar all-reduce."""
    eng = FusedTriples(ccwfn)
    comm = getattr(ccwfn, "comm", None)
    lst = eng.abc_all() if abc is None else np.asarray(abc, dtype=np.int32)
    if comm is not None and comm.size > 1:
        lst = lst[comm.rank::comm.size]
    et = eng.energy(lst)
    if comm is not None and comm.size > 1:
        comm.all_reduce_sum(et)
    return et[0]


def t_tjl(ccwfn, triples=None):
    """E(T), Lee-Rendell formulation (reference: cctriples.py:177-239).  Returns a 0-d device tensor."""
    if triples is None and fused_selected(ccwfn):
        return t_tjl_abc(ccwfn)
    eng = TriplesEngine(ccwfn, paired=PAIRED)
    comm = getattr(ccwfn, "comm", None)
    trip = triples_list(ccwfn.no) if triples is None else list(triples)
    # i = j = k contributes exactly zero (the bracket vanishes identically): skip those tiles
    trip = [t for t in trip if not (t[0] == t[1] == t[2])]
    if comm is not None and comm.size > 1:
        trip = trip[comm.rank::comm.size]
    try:
        et = eng.energy(trip)
    finally:
        eng.close()
    if comm is not None and comm.size > 1:
        comm.all_reduce_sum(et)
    return et[0]


# ---- per-triple building blocks with the reference's signatures -----------------------------------------
def _eps(F, o, v):
    d = torch.diagonal(F)
    return d[o].contiguous(), d[v].contiguous()


def t3c_ijk(o, v, i, j, k, t2, Wvvvo, Wovoo, F, contract, WithDenom=True):
    """Connected t3 for fixed (i,j,k) from ARBITRARY W blocks (reference: cctriples.py:27-72); the
    twelve contractions go through ``contract`` (DMMA GEMMs), the denominator through the t3 kernel."""
    nv, no = t2.shape[2], t2.shape[0]
    Wv = {x: Wvvvo[:, :, :, x] for x in {i, j, k}}
    t3 = contract('bae,ce->abc', Wv[i], t2[k, j])
    for sub, x, (p, q) in (('cae,be->abc', i, (j, k)), ('ace,be->abc', k, (j, i)), ('bce,ae->abc', k, (i, j)),
                           ('cbe,ae->abc', j, (i, k)), ('abe,ce->abc', j, (k, i))):
        contract(sub, Wv[x], t2[p, q], out=t3, alpha=1.0, beta=1.0)
    for sub, (p, q), x in (('mc,mab->abc', (j, k), i), ('mb,mac->abc', (k, j), i), ('mb,mca->abc', (i, j), k),
                           ('ma,mcb->abc', (j, i), k), ('ma,mbc->abc', (k, i), j), ('mc,mba->abc', (i, k), j)):
        contract(sub, Wovoo[:, :, p, q], t2[x], out=t3, alpha=-1.0, beta=1.0)
    if not WithDenom:
        return t3
    eo, ev = _eps(F, o, v)
    Q = torch.zeros(6 * nv ** 3, dtype=F64, device=t3.device)
    K.strided_axpby(Q[:nv ** 3].view(nv, nv, nv), t3, 1.0, 0.0)
    z1 = torch.zeros((no, nv), dtype=F64, device=t3.device)
    z2 = torch.zeros((no, no, nv, nv), dtype=F64, device=t3.device)
    w3, _ = K.t3_assemble(no, nv, i, j, k, Q, z1, z2, z2, z1, eo, ev, True)
    return w3


def t3d_ijk(o, v, i, j, k, t1, t2, Woovv, F, contract, WithDenom=True):
    """Disconnected t3 for fixed (i,j,k) (reference: cctriples.py:108-147)."""
    no, nv = t1.shape
    eo, ev = _eps(F, o, v)
    Q = torch.zeros(6 * nv ** 3, dtype=F64, device=t2.device)
    Wc = Woovv if Woovv.is_contiguous() else K.permuted(Woovv, (0, 1, 2, 3))
    _, d3 = K.t3_assemble(no, nv, i, j, k, Q, t1.contiguous(), t2.contiguous(), Wc, F[o, v], eo, ev, WithDenom)
    return d3


def _abc_denominator(t3, F, o, v, a, b, c):
    """t3[i,j,k] / (e_i + e_j + e_k - e_a - e_b - e_c) through the d1 kernel on the (ij, k) view."""
    no = t3.shape[0]
    eo, ev = _eps(F, o, v)
    dev = t3.device
    rows = torch.empty((no, no), dtype=F64, device=dev)
    K.strided_axpby(rows, eo.view(-1, 1).expand(no, no), 1.0, 0.0)
    K.strided_axpby(rows, eo.view(1, -1).expand(no, no), 1.0, 1.0)
    shift = float(ev[a] + ev[b] + ev[c])
    ones = torch.ones(no * no, dtype=F64, device=dev)
    rows = rows.view(-1)
    K.axpbyz(1.0, rows, -shift, ones, rows)
    negk = torch.empty(no, dtype=F64, device=dev)
    K.axpbyz(-1.0, eo, 0.0, None, negk)
    return K.div_d1(t3.contiguous().view(no * no, no), rows, negk).view(no, no, no)


def t3c_abc(o, v, a, b, c, t2, Wvvvo, Wovoo, F, contract, WithDenom=True):
    """Connected t3 for fixed (a,b,c), all (i,j,k) (reference: cctriples.py:75-105); generic in the W blocks."""
    t3 = contract('ei,kje->ijk', Wvvvo[b, a], t2[:, :, c])
    for sub, (x, y), z in (('ei,jke->ijk', (c, a), b), ('ek,jie->ijk', (a, c), b), ('ek,ije->ijk', (b, c), a),
                           ('ej,ike->ijk', (c, b), a), ('ej,kie->ijk', (a, b), c)):
        contract(sub, Wvvvo[x, y], t2[:, :, z], out=t3, alpha=1.0, beta=1.0)
    for sub, x, (y, z) in (('mjk,im->ijk', c, (a, b)), ('mkj,im->ijk', b, (a, c)), ('mij,km->ijk', b, (c, a)),
                           ('mji,km->ijk', a, (c, b)), ('mki,jm->ijk', a, (b, c)), ('mik,jm->ijk', c, (b, a))):
        contract(sub, Wovoo[:, x], t2[:, :, y, z], out=t3, alpha=-1.0, beta=1.0)
    return _abc_denominator(t3, F, o, v, a, b, c) if WithDenom else t3


def t3d_abc(o, v, a, b, c, t1, t2, Woovv, F, contract, WithDenom=True):
    """Disconnected t3 for fixed (a,b,c) (reference: cctriples.py:149-173)."""
    Fov = F[o, v]
    t3 = contract('ij,k->ijk', Woovv[:, :, a, b], t1[:, c])
    for sub, X, y in (('ik,j->ijk', Woovv[:, :, a, c], t1[:, b]), ('jk,i->ijk', Woovv[:, :, b, c], t1[:, a]),
                      ('ij,k->ijk', t2[:, :, a, b], Fov[:, c]), ('ik,j->ijk', t2[:, :, a, c], Fov[:, b]),
                      ('jk,i->ijk', t2[:, :, b, c], Fov[:, a])):
        contract(sub, X, y, out=t3, alpha=1.0, beta=1.0)
    return _abc_denominator(t3, F, o, v, a, b, c) if WithDenom else t3


def t3_pert_ijk(o, v, i, j, k, t2, V, F, contract, WithDenom=True):
    """Explicit-field coupling of the connected triples for fixed (i,j,k) (reference: cctriples.py:679-705):
    D t3[a,b,c] = V_ld t2[i,j,a,d] t2[k,l,c,b] -- ONE term, as the reference writes it.  ``cc3_t_residual`` does not call
    this: it folds the term into the numerator GEMMs (TriplesEngine ``pert``); this is the reference's public function."""
    nv, no = t2.shape[2], t2.shape[0]
    tmp = contract('ld,ad->al', V[o, v], t2[i, j])
    t3 = contract('al,lcb->abc', tmp, t2[k])
    if not WithDenom:
        return t3
    eo, ev = _eps(F, o, v)
    Q = torch.zeros(6 * nv ** 3, dtype=F64, device=t3.device)
    K.strided_axpby(Q[:nv ** 3].view(nv, nv, nv), t3, 1.0, 0.0)
    z1 = torch.zeros((no, nv), dtype=F64, device=t3.device)
    z2 = torch.zeros((no, no, nv, nv), dtype=F64, device=t3.device)
    w3, _ = K.t3_assemble(no, nv, i, j, k, Q, z1, z2, z2, z1, eo, ev, True)
    return w3


def t3_pert_abc(o, v, a, b, c, t2, V, F, contract, WithDenom=True):
    """The same term for fixed (a,b,c), all (i,j,k) (reference: cctriples.py:707-720)."""
    tmp = contract('ld,ijd->ijl', V[o, v], t2[:, :, a])
    t3 = contract('ijl,kl->ijk', tmp, t2[:, :, c, b])
    return _abc_denominator(t3, F, o, v, a, b, c) if WithDenom else t3


def t_vikings(ccwfn):
    """E(T), Helgaker-Jorgensen-Olsen full-loop formulation (reference: cctriples.py:243-307).  Cross-check
    only (6x the work of t_tjl): t3 tiles come from the same GEMM + assemble kernels, the X1/X2
    contractions go through the generic contraction backend."""
    ct = ccwfn.contract.engine
    H = ccwfn.H
    no, nv = ccwfn.no, ccwfn.nv
    o, v = ccwfn.o, ccwfn.v
    eng = TriplesEngine(ccwfn)
    t1, t2 = eng.t1, eng.t2
    dev = t2.device
    X1 = torch.zeros_like(t1)
    X2 = torch.zeros_like(t2)
    Loovv = H.derived("Loovv")
    ovvv, ooov = H.block("ovvv"), H.block("ooov")
    Fov = K.permuted(H.F[o, v], (0, 1))
    for i in range(no):
        for j in range(no):
            for k in range(no):
                t3, _ = eng.t3_parts(i, j, k, True)
                u = K.permuted(t3, (0, 1, 2))                           # t3 - t3[c,b,a]
                K.strided_axpby(u, t3.permute(2, 1, 0), -1.0, 1.0)
                w = K.permuted(t3, (0, 1, 2), 2.0)                      # 2 t3 - t3[a,c,b] - t3[c,b,a]
                K.strided_axpby(w, t3.permute(0, 2, 1), -1.0, 1.0)
                K.strided_axpby(w, t3.permute(2, 1, 0), -1.0, 1.0)
                ct('abc,bc->a', u, Loovv[j, k], out=X1[i], alpha=1.0, beta=1.0)
                ct('abc,dcb->ad', w, ovvv[k], out=X2[i, j], alpha=1.0, beta=1.0)   # <dk|bc> = ovvv[k,d,c,b]
                ct('abc,lc->lab', w, ooov[j, k], out=X2[i], alpha=-1.0, beta=1.0)
                ct('abc,c->ab', u, Fov[k], out=X2[i, j], alpha=1.0, beta=1.0)
    s = K.permuted(t2, (0, 1, 2, 3), 4.0)
    K.strided_axpby(s, t2.permute(0, 1, 3, 2), -2.0, 1.0)
    eng.close()
    d1 = K.multi_dot(t1.reshape(-1), [X1.reshape(-1)])
    d2 = K.multi_dot(s.reshape(-1), [X2.reshape(-1)])
    out = torch.empty(1, dtype=F64, device=dev)
    K.axpbyz(2.0, d1, 1.0, d2, out)
    return out[0]


def t_vikings_inverted(ccwfn):
    """E(T), virtual-batched "vikings" form (reference: cctriples.py:311-354): fixed (a,b,c), all (i,j,k).
    Cross-check only; every contraction goes through the generic contraction backend."""
    ct = ccwfn.contract.engine
    H = ccwfn.H
    no, nv = ccwfn.no, ccwfn.nv
    o, v = ccwfn.o, ccwfn.v
    t1, t2 = ccwfn.t1.contiguous(), ccwfn.t2.contiguous()
    dev = t2.device
    ovvv, ooov = H.block("ovvv"), H.block("ooov")
    Loovv = H.derived("Loovv")
    Fov = K.permuted(H.F[o, v], (0, 1))
    X1 = torch.zeros((nv, no), dtype=F64, device=dev)
    X2 = torch.zeros((nv, nv, no, no), dtype=F64, device=dev)
    # Wvvvo[x,y,e,i] = ovvv[i,e,y,x];  Wovoo[m,x,j,k] = ooov[j,k,m,x]   (views, no copies)
    Wvvvo = ovvv.permute(3, 2, 1, 0)
    Wovoo = ooov.permute(2, 3, 0, 1)
    for a in range(nv):
        for b in range(nv):
            for c in range(nv):
                t3 = t3c_abc(o, v, a, b, c, t2, Wvvvo, Wovoo, H.F, ct, True)
                u = K.permuted(t3, (0, 1, 2))
                K.strided_axpby(u, t3.permute(2, 1, 0), -1.0, 1.0)
                w = K.permuted(t3, (0, 1, 2), 2.0)
                K.strided_axpby(w, t3.permute(0, 2, 1), -1.0, 1.0)
                K.strided_axpby(w, t3.permute(2, 1, 0), -1.0, 1.0)
                ct('ijk,jk->i', u, Loovv[:, :, b, c], out=X1[a], alpha=1.0, beta=1.0)
                # ERI[v,o,v,v][d,k,b,c] = ovvv[k,d,c,b]
                ct('ijk,kd->dij', w, ovvv[:, :, c, b], out=X2[a], alpha=1.0, beta=1.0)
                ct('ijk,jkl->il', w, ooov[:, :, :, c], out=X2[a, b], alpha=-1.0, beta=1.0)
                ct('ijk,k->ij', u, Fov[:, c], out=X2[a, b], alpha=1.0, beta=1.0)
    # (4 t2 - 2 t2.swapaxes(2,3)) against X2.T, literally: s[a,d,i,j] = 4 t2[j,i,d,a] - 2 t2[j,i,a,d]
    s = K.permuted(t2, (3, 2, 1, 0), 4.0)
    K.strided_axpby(s, t2.permute(2, 3, 1, 0), -2.0, 1.0)
    t1T = K.permuted(t1, (1, 0))
    d1 = K.multi_dot(t1T.reshape(-1), [X1.reshape(-1)])
    d2 = K.multi_dot(s.reshape(-1), [X2.reshape(-1)])
    out = torch.empty(1, dtype=F64, device=dev)
    K.axpbyz(2.0, d1, 1.0, d2, out)
    return out[0]


# ---- (T) densities and Lambda sources ----------------------------------------------------
T3D_NAMES = ("Doo", "Dvv", "Dov", "Goovv", "Gooov", "Gvvvo", "S1", "S2")


def t3_density(o, v, no, nv, t1, t2, F, ERI, L, contract, comm=None, k_batch=None, work_bytes=None, prof=None,
               pairs=None, mixed=False):
    """(T) energy plus the (T) increments {Doo, Dvv, Dov, Goovv, Gooov, Gvvvo, S1, S2} to the one-/two-particle
    densities and the Lambda residuals; reference: cctriples.py:1063-1157, same signature and return value
    (``ERI`` must be the ``H.ERI`` block view of a :class:`BlockHamiltonian`; ``L`` is derived from it).

    The reference's o^3 loop body is 13 einsums on (v,v,v) tiles.  Here, for a pair i >= j and a run of k:

    1. the t3 numerators of all k are six two-segment GEMMs per triple in ONE launch (TriplesEngine.build_q, the
       kernel of t_tjl); ``b200cc_t3_connected_batch`` assembles M3 = t3c/D (every Q element read once).  Because
       t3c(j,i,k)[b,a,c] = t3c(i,j,k)[a,b,c], the same M3 run serves the loop bodies of (i,j) AND (j,i): the t3 build,
       70 % of the reference's flops, is done for o(o+1)/2 pairs instead of o^2;
    2. ``b200cc_t3_density_forms`` forms sym(t3d/D) on the fly and writes the two GEMM operands W2 = 2 sym(M3) +
       sym(N3) and P = 2 M3 - M3[acb] - M3[cba], each in the two K-major layouts the products below need, and keeps
       every matrix-vector shaped term (Goovv, the Fock part of X2, dvv/doo, Dov, S1) as in-register running sums;
    3. the run index k is folded into the SUMMATION index of six long-K GEMMs (no per-triple launches, no permuted
       copies of <mb|ef>):
         Gvvvo[:,:,:,j] [(a,b),d] += W2ab[(a,b),(k,c)] t2[i,k,d,c]          K = o v          (line 1143)
         S2[i] [(a,b),l]          -= W2ab[(a,b),(k,c)] <jk|lc>              K = o v          (line 1147)
         X2[i] [(a,b),l]          -= Pab [(a,b),(k,c)] <jk|lc>                               (line 1130)
         S2[i,j] [a,d]            += W2n[a,(k,b,c)] <kb|cd>                 K = o v^2        (line 1148)
         X2[i,j] [a,d]            += Pn [a,(k,b,c)] <kb|cd>                                  (line 1129)
         Gooov[j,i] [l,a]         -= t2[l,(k,b,c)] W2n[a,(k,b,c)]                            (line 1142)
       (the last three are issued once per t3 build for BOTH bodies: their left operands are stacked along M)
       (<dk|bc> = <kd|cb> = <kb|cd> = ovvv[k,b,c,d], so the natural layout of <mb|ef> IS the [(k,b,c), d] matrix).

    With a ``comm`` the (i >= j) pairs are dealt round-robin to the ranks and the pieces are summed with all-reduces
    at the end.  Returns ``(ET, dict)``; ET is a 0-d device tensor.  ``prof`` (a dict) collects CUDA-event times per
    phase in ms; ``pairs`` restricts the pair loop (timing probes only: the result is then a partial sum).
    ``mixed`` (precision='MP'): the t3 build and the Gvvvo product run as split-TF32 products on the tcgen05 kernel with
    FP64 accumulation (105 / 90 TFLOP/s FP64-equivalent at o=40,v=300); the stacked <kb|cd> product has too few output
    tiles (4v x v) for that kernel, which has no split-K, and the N = o products are bound by streaming their operand:
    both stay on the FP64 DMMA kernel.
    """
    import types
    H = getattr(ERI, "H", None)
    if H is None:
        raise B200ccError("t3_density needs the block-view ERI of a BlockHamiltonian (ccwfn.H.ERI)")
    ct = contract.engine if hasattr(contract, "engine") else contract
    t1 = t1.contiguous()
    t2 = t2.contiguous()
    dev = t2.device
    eo, ev = _eps(F, o, v)
    shim = types.SimpleNamespace(H=H, no=no, nv=nv, t1=t1, t2=t2, o=o, v=v, eps_o=eo, eps_v=ev, mixed=bool(mixed))
    v2, v3 = nv * nv, nv ** 3
    z = lambda *shape: torch.zeros(shape, dtype=F64, device=dev)
    ovvv, ooov, oovv = H.block("ovvv"), H.block("ooov"), H.block("oovv")
    # <kb|cd> with d slowest, [d,k,b,c] = ovvv[k,b,c,d]: a constant K-major copy (o v^3 doubles, built once per
    # Hamiltonian when it fits) puts the [W2n;Pn] x <kb|cd> products on the TMA kernel -- and, with precision='MP', on
    # the tcgen05 kernel; without it <mb|ef> is streamed in place as an N-major operand (cp.async kernel)
    ovvv_d = H._derived.get("ovvv_dkbc")
    if ovvv_d is None and nv % 2 == 0:
        free = torch.cuda.mem_get_info(dev)[0] if dev.type == "cuda" else 1 << 62
        if free > 6 * 8 * no * v3:
            ovvv_d = H._derived["ovvv_dkbc"] = K.permuted(ovvv, (3, 0, 1, 2))
    t2q = K.permuted(t2, (0, 2, 1, 3))               # [i,d,k,c] = t2[i,k,d,c]
    ooovq = K.permuted(ooov, (0, 2, 1, 3))           # [j,l,k,c] = <jk|lc>
    t2s = K.permuted(t2, (0, 1, 2, 3), 4.0)          # 4 t2 - 2 t2.swapaxes(2,3)
    K.strided_axpby(t2s, t2.permute(0, 1, 3, 2), -2.0, 1.0)
    oovvs = K.permuted(oovv, (0, 1, 2, 3), 4.0)      # 4 <ij|ab> - 2 <ij|ba>
    K.strided_axpby(oovvs, oovv.permute(0, 1, 3, 2), -2.0, 1.0)
    dvv_i, Dov, S1 = z(no, nv), z(no, nv), z(no, nv)
    Goovv, X2, S2 = z(no, no, nv, nv), z(no, no, nv, nv), z(no, no, nv, nv)
    S2T, X2T = z(no, v2, no), z(no, v2, no)         # [i][(a,b)][l]
    Gooov = z(no, no, no, nv)
    Gall = z(no, v2, nv)                            # [j][(a,b)][d]
    # work arrays per k of a run: Q (6 v^3) + M3 + two bodies x (W2ab, W2n, Pab, Pn) = 15 v^3 doubles
    if work_bytes is None:
        work_bytes = 24 << 30
        if dev.type == "cuda":
            free, _ = torch.cuda.mem_get_info(dev)
            work_bytes = int((free - 8 * no * v3) * 0.6)          # Gvvvo itself is still to be allocated
    kb = int(max(1, min(no, work_bytes // (15 * 8 * v3)))) if k_batch is None else int(max(1, min(no, k_batch)))
    kb = max(1, min(kb, (2 ** 32 - 1) // v3))        # the forms kernel indexes a run with 32-bit element offsets
    kb = -(-no // -(-no // kb))                      # equal-sized runs
    eng = TriplesEngine(shim, t1, t2, q_bytes=kb * 6 * v3 * 8)
    eng.fov = F[o, v]
    fov = eng.fov
    M3 = torch.empty(kb * v3, dtype=F64, device=dev)
    # operands of the two loop bodies (i,j), (j,i) served by one t3 build.  The K = (k,b,c) operands of both bodies live
    # in ONE buffer [W2n(0) | W2n(1) | Pn(0) | Pn(1)] so that the products against the common <kb|cd> matrix are a single
    # GEMM with M = 4v (a v x v output tiles raggedly and is too small to amortise its operand traffic)
    W2ab = [torch.empty(kb * v3, dtype=F64, device=dev) for _ in range(2)]
    Pab = [torch.empty(kb * v3, dtype=F64, device=dev) for _ in range(2)]
    NB = torch.empty(4 * kb * v3, dtype=F64, device=dev)
    Ctmp = torch.empty(4 * v2, dtype=F64, device=dev)
    Gtmp = torch.empty(no * 2 * nv, dtype=F64, device=dev)
    size, rank = (comm.size, comm.rank) if comm is not None else (1, 0)
    if pairs is None:
        pairs = [(i, j) for j in range(no) for i in range(j, no)][rank::size]
    marks = []

    def mark(name):
        if prof is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

    try:
        for (i0, j0) in pairs:
            bodies = ((i0, j0, False),) if i0 == j0 else ((i0, j0, False), (j0, i0, True))
            nb = len(bodies)
            for k0 in range(0, no, kb):
                nk = min(kb, no - k0)
                trip = [(i0, j0, k) for k in range(k0, k0 + nk)]
                ijk = torch.tensor(np.asarray(trip, dtype=np.int32), dtype=torch.int32).to(dev)
                mark("start")
                Q = eng.build_q(trip)
                mark("t3_gemm")
                K.t3_connected_batch(no, nv, ijk, Q, eo, ev, M3)
                mark("connected")
                kc, kbc, seg = nk * nv, nk * v2, nk * v3
                for q, (i, j, swap) in enumerate(bodies):
                    mark("start")
                    K.t3_density_forms(no, nv, i, j, k0, nk, M3, t1, t2s, oovvs, fov, eo, ev, W2ab[q],
                                       NB[q * seg:(q + 1) * seg], Pab[q], NB[(nb + q) * seg:(nb + q + 1) * seg],
                                       Goovv[i, j], X2[i, j], dvv_i[i], Dov[i], S1[i], swap_ab=swap)
                    mark("forms")
                    with K.mixed_mode(mixed):
                        K.dgemm(v2, nv, kc, W2ab[q], kc, 0, (t2q, (i * nv * no + k0) * nv), no * nv, 0, Gall[j], nv,
                                1.0, 1.0)
                    mark("gemm_Gvvvo")
                    # N = o columns only: bound by streaming W2ab / Pab, stays on the FP64 kernel in every mode
                    qoff = (j * no * no + k0) * nv
                    K.dgemm(v2, no, kc, W2ab[q], kc, 0, (ooovq, qoff), no * nv, 0, S2T[i], no, -1.0, 1.0)
                    K.dgemm(v2, no, kc, Pab[q], kc, 0, (ooovq, qoff), no * nv, 0, X2T[i], no, -1.0, 1.0)
                    mark("gemm_ooov")
                mark("start")
                # [W2n(q); Pn(q)] (2 nb v rows) x <kb|cd>  ->  S2[i,j] / X2[i,j] increments of every body
                M = 2 * nb * nv
                if ovvv_d is not None:
                    with K.mixed_mode(mixed):
                        K.dgemm(M, nv, kbc, NB, kbc, 0, (ovvv_d, k0 * v2), no * v2, 0, Ctmp, nv, 1.0, 0.0,
                                ksplit=K.balanced_ksplit(M, nv, kbc))
                else:
                    K.dgemm(M, nv, kbc, NB, kbc, 0, (ovvv, k0 * v3), nv, 1, Ctmp, nv, 1.0, 0.0,
                            ksplit=K.balanced_ksplit(M, nv, kbc))
                for q, (i, j, _) in enumerate(bodies):
                    K.axpbyz(1.0, S2[i, j].view(-1), 1.0, Ctmp[q * v2:(q + 1) * v2], S2[i, j].view(-1))
                    K.axpbyz(1.0, X2[i, j].view(-1), 1.0, Ctmp[(nb + q) * v2:(nb + q + 1) * v2], X2[i, j].view(-1))
                mark("gemm_ovvv")
                # t2[l,(k,b,c)] x W2n(q)[a,(k,b,c)]  ->  Gooov[j,i][l,a] increments of every body
                N = nb * nv
                K.dgemm(no, N, kbc, (t2, k0 * v2), no * v2, 0, NB, kbc, 0, Gtmp, N, 1.0, 0.0,
                        ksplit=K.balanced_ksplit(no, N, kbc))
                Gv = Gtmp[:no * N].view(no, nb, nv)
                for q, (i, j, _) in enumerate(bodies):
                    K.strided_axpby(Gooov[j, i], Gv[:, q, :], -1.0, 1.0)
                mark("gemm_Gooov")
    finally:
        eng.close()
    if prof is not None:
        torch.cuda.synchronize()
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            if name != "start":
                prof[name] = prof.get(name, 0.0) + e0.elapsed_time(e1)
    del M3, W2ab, Pab, NB, Ctmp, Gtmp
    _QCACHE.pop(dev, None)                           # release the Q workspace before the o v^3 output is allocated
    # [i][(a,b)][l] -> [i,l,a,b];  [j][a,b,d] -> [a,b,d,j]
    K.strided_axpby(S2, S2T.view(no, nv, nv, no).permute(0, 3, 1, 2), 1.0, 1.0)
    K.strided_axpby(X2, X2T.view(no, nv, nv, no).permute(0, 3, 1, 2), 1.0, 1.0)
    del S2T, X2T
    Gvvvo = K.permuted(Gall.view(no, nv, nv, nv), (1, 2, 3, 0))
    del Gall
    if size > 1:
        for t in (dvv_i, Dov, S1, Goovv, X2, S2, Gooov, Gvvvo):
            comm.all_reduce_sum(t)
    K.symmetrize_r2(S2)                                              # S2 + S2.swapaxes(0,1).swapaxes(2,3)   (1150)
    # dvv = sum_i dvv_i[i];  doo[i] = -sum_a dvv_i[i,a]                                                   (1133-1134)
    Dvv, Doo = z(nv, nv), z(no, no)
    ct('ia,i->a', dvv_i, torch.ones(no, dtype=F64, device=dev), out=torch.diagonal(Dvv), alpha=1.0, beta=0.0)
    ct('ia,a->i', dvv_i, torch.ones(nv, dtype=F64, device=dev), out=torch.diagonal(Doo), alpha=-1.0, beta=0.0)
    # ET = t1.S1 + (4 t2 - 2 t2.swapaxes(2,3)).X2                                                         (1153-1154)
    d1 = K.multi_dot(t1.reshape(-1), [S1.reshape(-1)])
    d2 = K.multi_dot(t2s.reshape(-1), [X2.reshape(-1)])
    et = torch.empty(1, dtype=F64, device=dev)
    K.axpbyz(1.0, d1, 1.0, d2, et)
    return et[0], {'Doo': Doo, 'Dvv': Dvv, 'Dov': Dov, 'Goovv': Goovv, 'Gooov': Gooov, 'Gvvvo': Gvvvo,
                   'S1': S1, 'S2': S2}


# ---- CC3: connected-triples contribution to the T residuals ----------------------------------
def cc3_t_residual(ccwfn, F, t1, t2, Fme, W, k_batch=None, work_bytes=None, V=None, comm=None, reduce=True):
    """(X1, X2) of ``CCwfn._cc3_t_residual`` (reference: ccwfn.py:374-430) from the T1-dressed intermediates ``W`` (dict
    with Wabei, Wmbij, Wmnie, Wamef in the reference's index orders).  ``V`` (the o-v block of F - H.F; real_time = True,
    ccwfn.py:421-423): every t3 is corrected by t3_pert_ijk inside its numerator GEMMs (TriplesEngine ``pert``).  That one
    term is not symmetric under (i,a) <-> (j,b), so the tile of (i,j,k) no longer serves the loop body (j,i): with V the
    loop runs over ALL ordered pairs (i,j), one build each (twice the t3 GEMMs of the field-free case).
    ``comm``: the pairs are dealt round-robin over the ranks (equal cost each); ``reduce`` sums the partial (X1, X2) with
    one all-reduce each, ``reduce=False`` hands the rank's partial sums to a caller that all-reduces them itself.

    Same machinery as :func:`t3_density`, fed with dressed operands: for a pair i >= j and a run of k the t3
    numerators are the batched two-segment GEMMs of the (T) engine on (Wabei, Wmbij); ``b200cc_t3_connected_batch``
    applies the denominators; because t3(j,i,k)[b,a,c] = t3(i,j,k)[a,b,c] holds for ANY W, one build serves the loop
    bodies (i,j) and (j,i).  ``b200cc_t3_density_forms`` then delivers, with its weight arrays pointed at L_jkbc and
    H_me,  P = 2 t3 - t3[acb] - t3[cba]  in both K-major layouts,  X1[i] += (t3 - t3[cba]) L_jkbc  and the H_me part of
    X2[i,j]; the two remaining terms are long-K GEMMs with k folded into the summation index:
        X2[i,j][a,d] += P[a,(k,b,c)] W_amef[d,k,b,c]        X2[i][(a,b),l] -= P[(a,b),(k,c)] W_mnie[j,k,l,c]."""
    import types
    H = ccwfn.H
    no, nv, o, v = ccwfn.no, ccwfn.nv, ccwfn.o, ccwfn.v
    t1, t2 = t1.contiguous(), t2.contiguous()
    dev = t2.device
    eo, ev = _eps(F, o, v)
    shim = types.SimpleNamespace(H=H, no=no, nv=nv, t1=t1, t2=t2, o=o, v=v, eps_o=eo, eps_v=ev, mixed=False)
    v2, v3 = nv * nv, nv ** 3
    z = lambda *shape: torch.zeros(shape, dtype=F64, device=dev)
    if work_bytes is None:
        work_bytes = 24 << 30
        if dev.type == "cuda":
            work_bytes = int(torch.cuda.mem_get_info(dev)[0] * 0.6)
    # per k of a run: Q (6 v^3) + M3 + (unused W2 operands written by the forms kernel: 2) + two bodies x (Pab, Pn)
    kb = int(max(1, min(no, work_bytes // (13 * 8 * v3)))) if k_batch is None else int(max(1, min(no, k_batch)))
    kb = max(1, min(kb, (2 ** 32 - 1) // v3))
    kb = -(-no // -(-no // kb))
    eng = TriplesEngine(shim, t1, t2, q_bytes=kb * 6 * v3 * 8, dressed=(W["Wabei"], W["Wmbij"]), pert=V)
    Wamef = W["Wamef"].contiguous()                               # [d,k,b,c]: K-major in (k,b,c) as stored
    Wq = K.permuted(W["Wmnie"], (0, 2, 1, 3))                      # [j,l,k,c]
    Loovv = H.derived("Loovv")
    Fme = Fme.contiguous()
    X1, X2 = z(no, nv), z(no, no, nv, nv)
    X2T = z(no, v2, no)                                           # [i][(a,b)][l]
    M3 = torch.empty(kb * v3, dtype=F64, device=dev)
    junk = [torch.empty(kb * v3, dtype=F64, device=dev) for _ in range(2)]       # W2 outputs of the forms kernel
    Pab = [torch.empty(kb * v3, dtype=F64, device=dev) for _ in range(2)]
    NB = torch.empty(2 * kb * v3, dtype=F64, device=dev)         # [Pn(0) | Pn(1)]
    Ctmp = torch.empty(2 * v2, dtype=F64, device=dev)
    gij, dv, s1 = z(nv, nv), z(nv), z(nv)                         # accumulators of the forms kernel that CC3 ignores
    try:
        pairs = [(i0, j0) for j0 in range(no) for i0 in range(j0 if V is None else 0, no)]
        if comm is not None:
            pairs = pairs[comm.rank::comm.size]
        for (i0, j0) in pairs:
            if V is not None or i0 == j0:
                bodies = ((i0, j0, False),)
            else:
                bodies = ((i0, j0, False), (j0, i0, True))
            nb = len(bodies)
            for k0 in range(0, no, kb):
                nk = min(kb, no - k0)
                trip = [(i0, j0, k) for k in range(k0, k0 + nk)]
                ijk = torch.tensor(np.asarray(trip, dtype=np.int32), dtype=torch.int32).to(dev)
                Q = eng.build_q(trip)
                K.t3_connected_batch(no, nv, ijk, Q, eo, ev, M3)
                kc, kbc, seg = nk * nv, nk * v2, nk * v3
                for q, (i, j, swap) in enumerate(bodies):
                    # weights: "t2s" -> L_jkbc (its Dov output is the X1[i] increment), "fov" -> H_me
                    K.t3_density_forms(no, nv, i, j, k0, nk, M3, t1, Loovv, Loovv, Fme, eo, ev, junk[0], junk[1],
                                       Pab[q], NB[q * seg:(q + 1) * seg], gij, X2[i, j], dv, X1[i], s1,
                                       swap_ab=swap)
                    K.dgemm(v2, no, kc, Pab[q], kc, 0, (Wq, (j * no * no + k0) * nv), no * nv, 0, X2T[i], no,
                            -1.0, 1.0)
                M = nb * nv
                K.dgemm(M, nv, kbc, NB, kbc, 0, (Wamef, k0 * v2), no * v2, 0, Ctmp, nv, 1.0, 0.0,
                        ksplit=K.balanced_ksplit(M, nv, kbc))
                for q, (i, j, _) in enumerate(bodies):
                    K.axpbyz(1.0, X2[i, j].view(-1), 1.0, Ctmp[q * v2:(q + 1) * v2], X2[i, j].view(-1))
    finally:
        eng.close()
    K.strided_axpby(X2, X2T.view(no, nv, nv, no).permute(0, 3, 1, 2), 1.0, 1.0)
    if comm is not None and reduce:
        comm.all_reduce_sum(X1)
        comm.all_reduce_sum(X2)
    return X1, X2
