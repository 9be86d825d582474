"""Tensor-level wrappers over the C ABI (include/b200cc.h).

Everything here takes torch CUDA float64 tensors (or views of them), turns them into raw
addresses + element strides, and launches the hand-written kernels on torch's current stream.
No arithmetic is done in torch/numpy on this path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import GemmDesc, Gemm3Desc, T3dDesc, TAbcDesc, B200ccError, i64

NSM = 148                   # B200
# tile-config override for experiments (0 = library heuristic); see b200cc_gemm_desc.config
DEFAULT_GEMM_CONFIG = int(os.environ.get("B200CC_GEMM_CONFIG", "0"))
F64 = torch.float64


def _addr(x):
    """tensor | (tensor, element_offset) | int address -> int address"""
    if isinstance(x, int):
        return x
    if isinstance(x, tuple):
        t, off = x
        return _lib.ptr(t) + 8 * int(off)
    return _lib.ptr(x)


def _dev(x):
    if isinstance(x, tuple):
        x = x[0]
    return x.device


SKINNY = os.environ.get("B200CC_GEMM_SKINNY", "1") != "0"


def _gemm_tiles(M, N):
    """output tiles of one batch entry under the library's tile choice (gemm.cu: 40 x 256 / 256 x 40 tiles when one
    extent is <= 40, else 128 x 128)"""
    if SKINNY and M <= 40:
        return (N + 255) // 256
    if SKINNY and N <= 40:
        return (M + 255) // 256
    return ((M + 127) // 128) * ((N + 127) // 128)


def _ragged_rows(M):
    """gemm.cu's test for a last 128-row tile that is much cheaper than a full one (few row tiles, < 0.7 of the cost):
    such products are balanced by a uniform split instead of the split-K of the last wave"""
    rem = M % 128
    if rem == 0 or (M + 127) // 128 > 16:
        return False
    frags_per_warp = -(-(-(-rem // 8)) // 4)          # ceil(ceil(rem / 8) / 4 warps)
    return 10 * frags_per_warp < 7 * 4


def auto_ksplit(M, N, K, batch, tma_like=False):
    """Split-K factor.  The GEMM is persistent (one CTA per SM doing ceil(units/148) rounds), so splitting K
    pays when the output has too few tiles to fill the SMs (Fae, Fmi, r1 terms) or when it fills the last
    round badly (Wmnij: 169 tiles = 2 rounds at 57 %); each split costs an extra M*N partial write + read."""
    tiles = _gemm_tiles(M, N) * batch
    kt = (K + 15) // 16
    if kt >= 2048 and 8 * NSM <= tiles < 16 * NSM:
        # long-K products are scheduled dynamically (one CTA per unit): with 8-16 units per SM the last wave costs up to a
        # whole unit (measured: Z in pair form, 1316 units of 2822 k-tiles, 0.92 of the tensor bound); halves finish closer
        return 2
    if tma_like and kt >= 512 and NSM < tiles <= 16 * NSM and not _ragged_rows(M):
        # K-major long-K products with equal-cost tiles and few waves: the library splits K for the LAST wave only
        # (gemm.cu, "split-K of the last wave"; same conditions there)
        return 1
    if kt < 64 or tiles >= 8 * NSM:
        return 1
    best, best_t = 1, None
    for s in (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 32, 48, 64, 96, 128, 192, 256):
        if s > 1 and kt // s < 32:
            break
        rounds = -(-tiles * s // NSM)
        t = rounds * (kt / s + 6.0)            # k-tiles per round + per-unit prologue/epilogue/partials
        if best_t is None or t < 0.93 * best_t:
            best, best_t = s, t
    return best


def balanced_ksplit(M, N, K, batch=1, rounds=12):
    """Split-K factor for long-K products with FEW, RAGGED output tiles (the o v^2-long contractions of t3_density,
    M = N = v): units are scheduled dynamically, so a ragged 128x128 tiling only balances when every SM gets many
    units -- aim at ``rounds`` units per SM, but keep >= 64 k-tiles per unit."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128) * batch
    kt = (K + 15) // 16
    if tiles >= rounds * NSM or kt < 128:
        return auto_ksplit(M, N, K, batch)
    return int(max(1, min(-(-rounds * NSM // tiles), kt // 64, 4096)))


# ---- mixed precision (precision='MP'): large K-major x K-major products go to the tcgen05 split-TF32 GEMM --------
class _Mixed:
    """Process-wide switch set by DeviceManager(precision='MP').  ``min_flops``: smaller products stay on the FP64
    DMMA kernel (they are a few % of an iteration and keep full precision); ``kchunk``: summation indices accumulated
    in FP32 (TMEM) before the FP64 drain; ``config``: b200cc_gemm3_desc.config."""
    on = False
    min_flops = 1.0e9
    min_dim = 32
    kchunk = int(os.environ.get("B200CC_MP_KCHUNK", "0"))       # 0 = library default (256 for the default kernel)
    config = int(os.environ.get("B200CC_MP_CONFIG", "0"))
    lockstep = int(os.environ.get("B200CC_MP_LOCKSTEP", "0"))   # leader/follower tile schedule: >0 skew in k-blocks, <=0 off
    min_tiles = 74        # fewer 128x128 output tiles than this: the split-K FP64 kernel fills the SMs better
    # Upper bound on the TF32 planes kept per owner (Hamiltonian / (T) engine) for constant operands.  Beyond it a
    # constant is split per use like any other operand (an 8.6 GB block: ~3 ms) instead of doubling its footprint:
    # HBAR + Lambda touch three permuted copies of <ma|ef>, whose cached planes put the o=40,v=300 peak at 179.6 GB.
    cache_bytes = int(float(os.environ.get("B200CC_MP_CACHE_GB", "32")) * (1 << 30))
    stats = {"gemm": 0, "split": 0, "split_cached": 0, "split_uncached_over_budget": 0}


MIXED = _Mixed()


class mixed_mode:
    """``with mixed_mode(flag):`` -- route eligible dgemm calls to the split-TF32 kernel inside the block.
    ``cache=False``: planes of constant operands met inside the block are not added to the cache (entries already
    there are still used).  HBAR / Lambda run this way: measured at o=40,v=300 the cache buys them nothing (0.380 vs
    0.382 s per Lambda iteration) and costs 33 GB (profiles/lambda_probe_r01_o40v300_mp_cache{0,32}.json)."""

    def __init__(self, flag, cache=True):
        self.flag = bool(flag)
        self.cache = bool(cache)

    def __enter__(self):
        self.prev = (MIXED.on, MIXED.cache_bytes)
        MIXED.on = self.flag
        if not self.cache:
            MIXED.cache_bytes = 0
        return self

    def __exit__(self, *exc):
        MIXED.on, MIXED.cache_bytes = self.prev
        return False


_CONST = {}       # storage address of a constant tensor (integral block / derived layout) -> its owner's split cache


def register_constant(t, cache):
    """Mark ``t`` as constant: its TF32 planes are computed once and kept in ``cache`` (owned by the Hamiltonian)."""
    _CONST[t.untyped_storage().data_ptr()] = cache


def unregister_tensor(t):
    """Forget one constant (its storage is about to be released and may be reused by a non-constant tensor)."""
    _CONST.pop(t.untyped_storage().data_ptr(), None)


def unregister_constants(cache):
    for k in [k for k, v in _CONST.items() if v is cache]:
        del _CONST[k]


def _faddr(x):
    """float32 tensor | (tensor, float offset) -> address"""
    if isinstance(x, tuple):
        return _lib.ptr(x[0]) + 4 * int(x[1])
    return _lib.ptr(x)


def split_tf32(src, rows, K, ld, batch=1, stride=0, out=None):
    """FP64 operand (rows x K, pitch ld, batch entries `stride` apart) -> (hi, lo, ldp): FP32 planes
    [batch][rows][ldp] with x ~= hi + lo, both TF32-representable (b200cc_split_tf32).  ``out=(hi, lo, ldp)``
    writes into existing planes (tensors or (tensor, float offset))."""
    nb = int(batch) if (batch > 1 and stride != 0) else 1
    if out is None:
        ldp = (int(K) + 3) // 4 * 4
        dev = _dev(src)
        hi = torch.empty((nb, rows, ldp), dtype=torch.float32, device=dev)
        lo = torch.empty((nb, rows, ldp), dtype=torch.float32, device=dev)
    else:
        hi, lo, ldp = out
    _lib.check(_lib.get().b200cc_split_tf32(_addr(src), int(ld), int(stride), int(rows), int(K), nb, _faddr(hi),
                                            _faddr(lo), int(ldp), _lib.stream()), "b200cc_split_tf32")
    MIXED.stats["split"] += 1
    return hi, lo, ldp


def merge_tf32(hi, lo, ldp, rows, K, out=None):
    """FP64 (rows x K, contiguous) = hi + lo from split-TF32 planes of pitch ldp (b200cc_merge_tf32); plane operands
    may be tensors or (tensor, float offset)."""
    if out is None:
        out = torch.empty((int(rows), int(K)), dtype=torch.float64, device=_dev(hi))
    _lib.check(_lib.get().b200cc_merge_tf32(_faddr(hi), _faddr(lo), int(ldp), int(rows), int(K), _lib.ptr(out), int(K),
                                            _lib.stream()), "b200cc_merge_tf32")
    return out


def _split_operand(X, rows, K, ld, batch, stride):
    t = X[0] if isinstance(X, tuple) else X
    cache = _CONST.get(t.untyped_storage().data_ptr())
    if cache is None:
        return split_tf32(X, rows, K, ld, batch, stride)
    key = (_addr(X), int(rows), int(K), int(ld), int(batch) if stride else 1, int(stride))
    r = cache.get(key)
    if r is None:
        r = split_tf32(X, rows, K, ld, batch, stride)
        held = sum(2 * 4 * v[0].numel() for v in cache.values())
        if held + 2 * 4 * r[0].numel() <= MIXED.cache_bytes:
            cache[key] = r
        else:
            MIXED.stats["split_uncached_over_budget"] += 1
    else:
        MIXED.stats["split_cached"] += 1
    return r


def gemm_tf32x3(M, N, K, Ahi, Alo, lda, Bhi, Blo, ldb, Cmat, ldc, alpha=1.0, beta=0.0, batch=1, sA=0, sB=0, sC=0,
                kchunk=None, config=None, lockstep=None, seg2=None, bcoords=None, nbatch=None):
    """C[b] = alpha * (A[b] B[b]^T [+ A2[b] B2[b]^T]) + beta * C[b] from split-TF32 planes (b200cc_gemm_tf32x3); plane
    operands may be tensors or (tensor, float offset).  seg2 = (A2hi, A2lo, lda2, B2hi, B2lo, ldb2, K2, sA2, sB2);
    bcoords / nbatch as in dgemm."""
    if M == 0 or N == 0 or batch == 0:
        return

    fa = _faddr
    d = Gemm3Desc()
    d.struct_size = C.sizeof(Gemm3Desc)
    d.M, d.N, d.K = int(M), int(N), int(K)
    d.Ahi, d.Alo, d.Bhi, d.Blo = fa(Ahi), fa(Alo), fa(Bhi), fa(Blo)
    d.lda, d.ldb, d.strideA, d.strideB = int(lda), int(ldb), int(sA), int(sB)
    d.C, d.ldc, d.strideC = _addr(Cmat), int(ldc), int(sC)
    d.alpha, d.beta, d.batch = float(alpha), float(beta), int(batch)
    d.kchunk = int(MIXED.kchunk if kchunk is None else kchunk)
    d.config = int(MIXED.config if config is None else config)
    d.lockstep = int(MIXED.lockstep if lockstep is None else lockstep)
    if seg2 is not None:
        A2h, A2l, lda2, B2h, B2l, ldb2, K2, sA2, sB2 = seg2
        d.K2, d.A2hi, d.A2lo, d.B2hi, d.B2lo = int(K2), fa(A2h), fa(A2l), fa(B2h), fa(B2l)
        d.lda2, d.ldb2, d.strideA2, d.strideB2 = int(lda2), int(ldb2), int(sA2), int(sB2)
    if bcoords is not None:
        d.bcoords = _lib.ptr(bcoords)
        d.nbA1, d.nbB1, d.nbA2, d.nbB2 = (int(x) for x in nbatch)
    _lib.check(_lib.get().b200cc_gemm_tf32x3(C.byref(d), _lib.stream()), "b200cc_gemm_tf32x3")
    MIXED.stats["gemm"] += 1


def _mixed_eligible(M, N, K, A, B, transA, transB, batch, seg2, table, bcoords, out_cube_nv):
    if seg2 is not None and (isinstance(seg2[0], int) or isinstance(seg2[2], int)):
        return False
    return (MIXED.on and not transA and not transB and table is None
            and not out_cube_nv and not isinstance(A, int) and not isinstance(B, int)
            and min(M, N, K) >= MIXED.min_dim and 2.0 * M * N * K * batch >= MIXED.min_flops
            and ((M + 127) // 128) * ((N + 127) // 128) * batch >= MIXED.min_tiles)


def dgemm(M, N, K, A, lda, transA, B, ldb, transB, Cmat, ldc, alpha=1.0, beta=0.0,
          batch=1, sA=0, sB=0, sC=0, seg2=None, table=None, table_align16=False, ksplit=None, config=0,
          bcoords=None, nbatch=None, out_cube_nv=0, mp_kchunk=None, seg3=None, seg4=None):
    """C[b] = alpha * (opA[b] opB[b]^T [+ second K segment]) + beta * C[b]; see b200cc_dgemm.
    ``mp_kchunk``: FP32 accumulation run when the product is routed to the mixed-precision kernel.

    A, B, Cmat: tensor, (tensor, element offset) or raw address.  seg2 = (A2, lda2, B2, ldb2, K2, sA2, sB2);
    seg3 / seg4 likewise (TMA kernels only; with bcoords each batch entry then carries 8 slab indices and ``nbatch``
    8 slab counts).
    """
    if M == 0 or N == 0 or batch == 0:
        return
    if seg3 is None and _mixed_eligible(M, N, K, A, B, transA, transB, batch, seg2, table, bcoords, out_cube_nv):
        # with bcoords every operand is a stack of slabs (counts in nbatch) that the batch entries index
        nA1, nB1, nA2, nB2 = (int(x) for x in nbatch) if bcoords is not None else (batch,) * 4
        Ah, Al, lpa = _split_operand(A, M, K, lda, nA1, sA)
        Bh, Bl, lpb = _split_operand(B, N, K, ldb, nB1, sB)
        s2 = None
        if seg2 is not None:
            A2, lda2, B2, ldb2, K2, sA2, sB2 = seg2
            A2h, A2l, lpa2 = _split_operand(A2, M, K2, lda2, nA2, sA2)
            B2h, B2l, lpb2 = _split_operand(B2, N, K2, ldb2, nB2, sB2)
            s2 = (A2h, A2l, lpa2, B2h, B2l, lpb2, K2, M * lpa2 if (nA2 > 1 and sA2) else 0,
                  N * lpb2 if (nB2 > 1 and sB2) else 0)
        return gemm_tf32x3(M, N, K, Ah, Al, lpa, Bh, Bl, lpb, Cmat, ldc, alpha, beta, batch,
                           M * lpa if (nA1 > 1 and sA) else 0, N * lpb if (nB1 > 1 and sB) else 0, sC,
                           seg2=s2, bcoords=bcoords, nbatch=nbatch, kchunk=mp_kchunk)
    d = GemmDesc()
    d.struct_size = C.sizeof(GemmDesc)
    d.M, d.N, d.transA, d.transB = int(M), int(N), int(bool(transA)), int(bool(transB))
    d.K1 = int(K)
    d.A1, d.B1, d.C = _addr(A), _addr(B), _addr(Cmat)
    d.lda1, d.ldb1, d.ldc = int(lda), int(ldb), int(ldc)
    d.strideA1, d.strideB1, d.strideC = int(sA), int(sB), int(sC)
    K2 = 0
    if seg2 is not None:
        A2, lda2, B2, ldb2, K2, sA2, sB2 = seg2
        d.A2, d.B2, d.lda2, d.ldb2, d.K2 = _addr(A2), _addr(B2), int(lda2), int(ldb2), int(K2)
        d.strideA2, d.strideB2 = int(sA2), int(sB2)
    d.alpha, d.beta, d.batch = float(alpha), float(beta), int(batch)
    d.table = _lib.ptr(table) if table is not None else None
    d.table_align16 = int(bool(table_align16))
    K34 = 0
    for n_, seg in ((3, seg3), (4, seg4)):
        if seg is not None:
            As, ldas, Bs, ldbs, Ks, sAs, sBs = seg
            setattr(d, "A%d" % n_, _addr(As)); setattr(d, "B%d" % n_, _addr(Bs))
            setattr(d, "lda%d" % n_, int(ldas)); setattr(d, "ldb%d" % n_, int(ldbs)); setattr(d, "K%d" % n_, int(Ks))
            setattr(d, "strideA%d" % n_, int(sAs)); setattr(d, "strideB%d" % n_, int(sBs))
            K34 += int(Ks)
    if ksplit is None:
        ksplit = auto_ksplit(M, N, K + K2 + K34, batch,
                             tma_like=not transA and not transB and table is None and not out_cube_nv)
    d.ksplit = int(ksplit)
    d.config = int(config) if config else DEFAULT_GEMM_CONFIG
    d.out_cube_nv = int(out_cube_nv)
    if bcoords is not None:
        # per-batch operand indices (int32 [batch,4]) + the extent of each operand's batch dimension
        d.bcoords = _lib.ptr(bcoords)
        nb = [int(x) for x in nbatch]
        d.nbA1, d.nbB1, d.nbA2, d.nbB2 = nb[:4]
        if len(nb) == 8:
            d.nbA3, d.nbB3, d.nbA4, d.nbB4 = nb[4:]
    ws = None
    if ksplit > 1:
        dev = _dev(table if table is not None else Cmat)
        ws = torch.empty(ksplit * batch * M * N, dtype=F64, device=dev)
        d.workspace = _lib.ptr(ws)
    _lib.check(_lib.get().b200cc_dgemm(C.byref(d), _lib.stream()), "b200cc_dgemm")
    del ws


# ---- the ladder in symmetric / antisymmetric pair form (csrc/pairs.cu) ------------------------------------------
def pair_count(n):
    """number of pairs (x >= y) of n indices; pair(x, y) = x(x+1)/2 + y"""
    return int(n) * (int(n) + 1) // 2


def pair_ld(nv):
    """row pitch of packed pair matrices: v(v+1)/2 rounded up to 16 doubles (128-byte rows for TMA boxes)"""
    return (pair_count(nv) + 15) // 16 * 16


def pack_pairs(src, nv, a0, a1, vp, vm, ldq):
    """Rows a in [a0,a1) of <ab|ef> -> V+ / V- rows pair(a,b) - pair(a0,0), b <= a (b200cc_pack_pairs).
    ``src``: a 4-D strided view indexed [a - a0, b, e, f] (any memory order); ``vp``, ``vm``: tensor or
    (tensor, element offset) of the first output row."""
    if src.dim() != 4 or tuple(src.shape) != (a1 - a0, nv, nv, nv):
        raise B200ccError("pack_pairs: src must be a [a1-a0, v, v, v] view (got %s)" % (tuple(src.shape),))
    sa, sb, se, sf = (int(x) for x in src.stride())
    _lib.check(_lib.get().b200cc_pack_pairs(_lib.ptr(src), sa, sb, se, sf, int(nv), int(a0), int(a1), _addr(vp),
                                            _addr(vm), int(ldq), _lib.stream()), "b200cc_pack_pairs")


def unpack_pairs(vp, vm, ldq, nv, npairs, out=None):
    """FP64 slabs <ab|ef>[p, e, f] of ``npairs`` consecutive packed rows starting at ``vp`` / ``vm`` (b200cc_unpack_pairs)."""
    if out is None:
        out = torch.empty((int(npairs), nv, nv), dtype=F64, device=_dev(vp))
    _lib.check(_lib.get().b200cc_unpack_pairs(_addr(vp), _addr(vm), int(ldq), int(nv), int(npairs), _lib.ptr(out),
                                              _lib.stream()), "b200cc_unpack_pairs")
    return out


def pack_tau(tau, tri, out=None):
    """tau (no,no,nv,nv) -> [2, M, ldq]: T+ and T- over the pairs (e >= f); M = o(o+1)/2 rows (i >= j) when ``tri``
    (tau pair-symmetric), else o^2 (b200cc_pack_tau)."""
    no, nv = tau.shape[0], tau.shape[2]
    M = pair_count(no) if tri else no * no
    ldq = pair_ld(nv)
    if out is None:
        out = torch.empty((2, M, ldq), dtype=F64, device=tau.device)
    _lib.check(_lib.get().b200cc_pack_tau(_lib.ptr(_c(tau, "tau")), int(no), int(nv), int(bool(tri)), _lib.ptr(out[0]),
                                          _lib.ptr(out[1]), int(ldq), _lib.stream()), "b200cc_pack_tau")
    return out


def pack_rows(src, nrows, nv, out=None):
    """``nrows`` contiguous (v,v) slabs x[e,f] -> [2, nrows, ldq]: X+ = x_ef + x_fe (x_ee), X- = x_ef - x_fe over the pairs
    (e >= f) (b200cc_pack_rows); ``src``: tensor or (tensor, element offset)."""
    ldq = pair_ld(nv)
    if out is None:
        out = torch.empty((2, int(nrows), ldq), dtype=F64, device=_dev(src))
    _lib.check(_lib.get().b200cc_pack_rows(_addr(src), int(nrows), int(nv), _lib.ptr(out[0]), _lib.ptr(out[1]), int(ldq),
                                           _lib.stream()), "b200cc_pack_rows")
    return out


def pair_rows_unpack(S, A, lds, no, ncols, out, ldo):
    """out[i,j,:ncols] = S[p] + A[p], out[j,i,:ncols] = S[p] - A[p] for p = pair(i,j), i >= j (b200cc_pair_rows_unpack)."""
    _lib.check(_lib.get().b200cc_pair_rows_unpack(_addr(S), _addr(A), int(lds), int(no), int(ncols), _addr(out), int(ldo),
                                                  _lib.stream()), "b200cc_pair_rows_unpack")
    return out


def ring_layouts(t2):
    """(u, tb): u[i,a,m,e] = 2 t2[i,m,a,e] - t2[i,m,e,a], tb[i,a,m,e] = t2[i,m,e,a] in one pass (b200cc_ring_layouts)."""
    no, nv = t2.shape[0], t2.shape[2]
    u = torch.empty((no, nv, no, nv), dtype=F64, device=t2.device)
    tb = torch.empty((no, nv, no, nv), dtype=F64, device=t2.device)
    _lib.check(_lib.get().b200cc_ring_layouts(_lib.ptr(_c(t2, "t2")), int(no), int(nv), _lib.ptr(u), _lib.ptr(tb),
                                              _lib.stream()), "b200cc_ring_layouts")
    return u, tb


def ladder_unpack(S, A, lds, no, nv, tri, a0, a1, alpha, r2):
    """r2 += alpha * (the ladder held as S / A over the pairs of rows a in [a0,a1)) (b200cc_ladder_unpack)."""
    _lib.check(_lib.get().b200cc_ladder_unpack(_addr(S), _addr(A), int(lds), int(no), int(nv), int(bool(tri)), int(a0),
                                               int(a1), float(alpha), _lib.ptr(_c(r2, "r2")), _lib.stream()),
               "b200cc_ladder_unpack")
    return r2


def strided_axpby(out, inp, alpha=1.0, beta=0.0):
    """out = alpha * inp + beta * out for two equally-shaped views with arbitrary strides
    (e.g. ``strided_axpby(buf, t2.permute(0, 1, 3, 2), -1.0, 1.0)``)."""
    if tuple(out.shape) != tuple(inp.shape):
        raise B200ccError("strided_axpby: shape mismatch %s vs %s" % (tuple(out.shape), tuple(inp.shape)))
    r = out.dim()
    if r > 6:
        raise B200ccError("strided_axpby: rank > 6")
    if out.numel() == 0:
        return out
    arr = i64 * max(r, 1)
    shape = arr(*([int(s) for s in out.shape] or [1]))
    si = arr(*([int(s) for s in inp.stride()] or [0]))
    so = arr(*([int(s) for s in out.stride()] or [0]))
    _lib.check(_lib.get().b200cc_permute(max(r, 1) if r else 1, shape, si, so, float(alpha), _lib.ptr(inp),
                                         float(beta), _lib.ptr(out), _lib.stream()), "b200cc_permute")
    return out


def permuted(inp, perm, alpha=1.0):
    """A new contiguous tensor holding alpha * inp.permute(perm)."""
    v = inp.permute(*perm)
    out = torch.empty(tuple(v.shape), dtype=inp.dtype, device=inp.device)
    return strided_axpby(out, v, alpha, 0.0)


def axpbyz(a, x, b, y, z):
    """z = a*x + b*y on contiguous, equally sized tensors (z may alias x or y)."""
    n = z.numel()
    for t in (x, y, z):
        if t is not None and (t.numel() != n or not t.is_contiguous()):
            raise B200ccError("axpbyz: operands must be contiguous and equally sized")
    _lib.check(_lib.get().b200cc_axpbyz(n, float(a), _lib.ptr(x if x is not None else z), float(b),
                                        _lib.ptr(y if y is not None else z), _lib.ptr(z), _lib.stream()),
               "b200cc_axpbyz")
    return z


def _c(t, name):
    if not t.is_contiguous():
        raise B200ccError("%s must be contiguous" % name)
    return t


def build_tau(t1, t2, f1=1.0, f2=1.0, out=None):
    no, nv = t1.shape
    out = torch.empty_like(t2, memory_format=torch.contiguous_format) if out is None else out
    _lib.check(_lib.get().b200cc_build_tau(no, nv, float(f1), float(f2), _lib.ptr(_c(t1, "t1")),
                                           _lib.ptr(_c(t2, "t2")), _lib.ptr(_c(out, "tau")), _lib.stream()),
               "b200cc_build_tau")
    return out


def div_d2(x, eo, ev, out=None):
    no, nv = x.shape[0], x.shape[2]
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.get().b200cc_div_d2(no, nv, _lib.ptr(eo), _lib.ptr(ev), _lib.ptr(_c(x, "x")),
                                        _lib.ptr(_c(out, "out")), _lib.stream()), "b200cc_div_d2")
    return out


def div_d1(x, eo, ev, out=None):
    no, nv = x.shape
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.get().b200cc_div_d1(no, nv, _lib.ptr(eo), _lib.ptr(ev), _lib.ptr(_c(x, "x")),
                                        _lib.ptr(_c(out, "out")), _lib.stream()), "b200cc_div_d1")
    return out


def _scratch(dev, n=16 * 1024):
    return torch.empty(n, dtype=F64, device=dev)


def update_amps(r1, r2, eo, ev, t1, t2, symmetrize=True, write_r2=True):
    """Fused r2 symmetrisation + Jacobi update + sum((r/D)^2).  Returns a 1-element device tensor."""
    no, nv = t1.shape
    out = torch.empty(1, dtype=F64, device=t2.device)
    sc = _scratch(t2.device, 4096)
    _lib.check(_lib.get().b200cc_update_amps(no, nv, _lib.ptr(eo), _lib.ptr(ev), _lib.ptr(_c(r1, "r1")),
                                             _lib.ptr(_c(r2, "r2")), int(symmetrize), int(write_r2),
                                             _lib.ptr(_c(t1, "t1")), _lib.ptr(_c(t2, "t2")), _lib.ptr(out),
                                             _lib.ptr(sc), _lib.stream()), "b200cc_update_amps")
    return out


def update_amps_rows(r1, half, eo, ev, t1, t2, i0, i1):
    """Rows i in [i0,i1) of the fused symmetrise + Jacobi update (b200cc_update_amps_rows); t1 updated in full.
    Returns a 2-element device tensor: sum (r2/D)^2 over the rows, sum (r1/D)^2."""
    no, nv = t1.shape
    out = torch.empty(2, dtype=F64, device=t2.device)
    sc = _scratch(t2.device, 4096)
    _lib.check(_lib.get().b200cc_update_amps_rows(no, nv, int(i0), int(i1), _lib.ptr(eo), _lib.ptr(ev),
                                                  _lib.ptr(_c(r1, "r1")) if r1 is not None else None,
                                                  _lib.ptr(_c(half, "r2")), _lib.ptr(_c(t1, "t1")), _lib.ptr(_c(t2, "t2")),
                                                  _lib.ptr(out), _lib.ptr(sc), _lib.stream()), "b200cc_update_amps_rows")
    return out


def cc_energy_rows(fov, t1, t2, Loovv, i0, i1, with_singles):
    """Partial energy of the rows i in [i0,i1) (b200cc_cc_energy_rows) as a 1-element device tensor."""
    no, nv = t1.shape
    if fov.stride(1) != 1 and nv > 1:
        raise B200ccError("cc_energy: fov must be unit-stride along the virtual index")
    out = torch.empty(1, dtype=F64, device=t2.device)
    sc = _scratch(t2.device, 4096)
    _lib.check(_lib.get().b200cc_cc_energy_rows(no, nv, int(i0), int(i1), int(bool(with_singles)), _lib.ptr(fov),
                                                int(fov.stride(0)), _lib.ptr(_c(t1, "t1")), _lib.ptr(_c(t2, "t2")),
                                                _lib.ptr(_c(Loovv, "Loovv")), _lib.ptr(out), _lib.ptr(sc), _lib.stream()),
               "b200cc_cc_energy_rows")
    return out


def symmetrize_r2(r2):
    no, nv = r2.shape[0], r2.shape[2]
    _lib.check(_lib.get().b200cc_symmetrize_r2(no, nv, _lib.ptr(_c(r2, "r2")), _lib.stream()),
               "b200cc_symmetrize_r2")
    return r2


def cc_energy(fov, t1, t2, Loovv):
    """2 f.t1 + (t2 + t1 t1).L as a 1-element device tensor.  ``fov`` may be a strided (no,nv) view."""
    no, nv = t1.shape
    if fov.stride(1) != 1 and nv > 1:
        raise B200ccError("cc_energy: fov must be unit-stride along the virtual index")
    out = torch.empty(1, dtype=F64, device=t2.device)
    sc = _scratch(t2.device, 4096)
    _lib.check(_lib.get().b200cc_cc_energy(no, nv, _lib.ptr(fov), int(fov.stride(0)), _lib.ptr(_c(t1, "t1")),
                                           _lib.ptr(_c(t2, "t2")), _lib.ptr(_c(Loovv, "Loovv")), _lib.ptr(out),
                                           _lib.ptr(sc), _lib.stream()), "b200cc_cc_energy")
    return out


def multi_dot(x, ys):
    """[x . y for y in ys] as a len(ys) device tensor; one pass over x (len(ys) <= 16)."""
    m = len(ys)
    n = x.numel()
    for y in ys:
        if y.numel() != n or not y.is_contiguous():
            raise B200ccError("multi_dot: operands must be contiguous and equally sized")
    out = torch.empty(m, dtype=F64, device=x.device)
    sc = _scratch(x.device, 16 * 1024)
    arr = (_lib.dptr * m)(*[_lib.ptr(y) for y in ys])
    _lib.check(_lib.get().b200cc_multi_dot(n, _lib.ptr(_c(x, "x")), m, arr, _lib.ptr(out), _lib.ptr(sc),
                                           _lib.stream()), "b200cc_multi_dot")
    return out


def multi_axpy(coeffs, xs, out):
    """out = sum_q coeffs[q] * xs[q]  (len <= 16, out aliases no xs[q])."""
    m = len(xs)
    n = out.numel()
    for x in xs:
        if x.numel() != n or not x.is_contiguous():
            raise B200ccError("multi_axpy: operands must be contiguous and equally sized")
        if x.data_ptr() == out.data_ptr():
            raise B200ccError("multi_axpy: out aliases an input")
    arr = (_lib.dptr * m)(*[_lib.ptr(x) for x in xs])
    cs = (C.c_double * m)(*[float(c) for c in coeffs])
    _lib.check(_lib.get().b200cc_multi_axpy(n, m, cs, arr, _lib.ptr(_c(out, "out")), _lib.stream()),
               "b200cc_multi_axpy")
    return out


def q_size(nv, blocked):
    """doubles per Q array (plain v^3, or padded to whole 8x8x8 cubes when blocked)"""
    return int(_lib.get().b200cc_t_q_size(int(nv), int(blocked) & 1))


def t_energy_batch(no, nv, ijk, Q, t1, t2, oovv, fov, eo, ev, et, accumulate=True, blocked=False):
    """``blocked``: Q layout flags (bit 0 = cube-blocked arrays, bit 1 = paired: three summed arrays per triple)."""
    ntrip = ijk.shape[0]
    n = int(_lib.get().b200cc_t_energy_scratch(nv, ntrip))
    sc = _scratch(Q.device, max(n, 1))
    _lib.check(_lib.get().b200cc_t_energy_batch(no, nv, ntrip, _lib.ptr(ijk), _lib.ptr(Q), int(blocked),
                                                _lib.ptr(_c(t1, "t1")),
                                                _lib.ptr(_c(t2, "t2")), _lib.ptr(_c(oovv, "oovv")), _lib.ptr(fov),
                                                int(fov.stride(0)), _lib.ptr(eo), _lib.ptr(ev), _lib.ptr(et),
                                                int(accumulate), _lib.ptr(sc), _lib.stream()),
               "b200cc_t_energy_batch")
    return et


def t_abc_max_no():
    """largest o the fused (a,b,c)-driven (T) kernel takes (b200cc_t_abc_max_no)"""
    return int(_lib.get().b200cc_t_abc_max_no())


def t_abc(no, nv, abc, sorted_ijk, G, t2, t2x, Ox, oovvx, t1, fov, eo, ev, et, wtile, partial, grid, accumulate=True,
          fov_is_zero=False):
    """E(T) contributions of the packed virtual triples ``abc`` (int32, a | b << 10 | c << 20, a >= b >= c), fused
    (a,b,c)-driven kernel (b200cc_t_abc): ``et[0] (+)= ...``.  ``wtile``: grid * o^3 doubles, ``partial``: grid doubles."""
    d = TAbcDesc()
    d.struct_size = C.sizeof(TAbcDesc)
    d.no, d.nv = int(no), int(nv)
    d.nabc, d.nsorted = int(abc.numel()), int(sorted_ijk.numel())
    d.abc, d.sorted = _lib.ptr(abc), _lib.ptr(sorted_ijk)
    d.G, d.t2, d.t2x = _lib.ptr(_c(G, "G")), _lib.ptr(_c(t2, "t2")), _lib.ptr(_c(t2x, "t2x"))
    d.Ox, d.oovvx, d.t1 = _lib.ptr(_c(Ox, "Ox")), _lib.ptr(_c(oovvx, "oovvx")), _lib.ptr(_c(t1, "t1"))
    d.fov, d.ldf = _lib.ptr(fov), int(fov.stride(0))
    d.eo, d.ev = _lib.ptr(eo), _lib.ptr(ev)
    if wtile.numel() < int(grid) * no ** 3 or partial.numel() < int(grid):
        raise B200ccError("t_abc: scratch too small for grid = %d" % grid)
    d.wtile, d.partial, d.et_out = _lib.ptr(wtile), _lib.ptr(partial), _lib.ptr(et)
    d.accumulate, d.grid = int(bool(accumulate)), int(grid)
    d.fov_is_zero = int(bool(fov_is_zero))
    _lib.check(_lib.get().b200cc_t_abc(C.byref(d), _lib.stream()), "b200cc_t_abc")
    return et


def t3_assemble(no, nv, i, j, k, Q, t1, t2, oovv, fov, eo, ev, with_denom, blocked=False):
    w3 = torch.empty((nv, nv, nv), dtype=F64, device=Q.device)
    d3 = torch.empty((nv, nv, nv), dtype=F64, device=Q.device)
    _lib.check(_lib.get().b200cc_t3_assemble(no, nv, int(i), int(j), int(k), _lib.ptr(Q), int(blocked),
                                             _lib.ptr(_c(t1, "t1")),
                                             _lib.ptr(_c(t2, "t2")), _lib.ptr(_c(oovv, "oovv")), _lib.ptr(fov),
                                             int(fov.stride(0)), _lib.ptr(eo), _lib.ptr(ev), int(bool(with_denom)),
                                             _lib.ptr(w3), _lib.ptr(d3), _lib.stream()), "b200cc_t3_assemble")
    return w3, d3


def t3_connected_batch(no, nv, ijk, Q, eo, ev, out):
    """out[t] = connected t3 of triple t WITH denominators, from the six GEMM outputs Q[t] (b200cc_t3_connected_batch)."""
    ntrip = ijk.shape[0]
    if out.numel() < ntrip * nv ** 3:
        raise B200ccError("t3_connected_batch: output too small")
    _lib.check(_lib.get().b200cc_t3_connected_batch(int(no), int(nv), int(ntrip), _lib.ptr(ijk), _lib.ptr(Q),
                                                    _lib.ptr(eo), _lib.ptr(ev), _lib.ptr(_c(out, "out")),
                                                    _lib.stream()), "b200cc_t3_connected_batch")
    return out


def t3_density_forms(no, nv, i, j, k0, nk, M3, t1, t2s, oovvs, fov, eo, ev, W2ab, W2n, Pab, Pn, Gij, Xij, dvv, Dov, S1,
                     swap_ab=False):
    """The non-GEMM part of the t3_density loop body for fixed (i,j) and k0 <= k < k0+nk (b200cc_t3_density_forms).
    t2s = 4 t2 - 2 t2^T(ab), oovvs = 4<ij|ab> - 2<ij|ba>; swap_ab: M3 is the run of the transposed pair (j,i)."""
    need = nk * nv ** 3
    for t in (M3, W2ab, W2n, Pab, Pn):
        if t.numel() < need or not t.is_contiguous():
            raise B200ccError("t3_density_forms: work arrays must be contiguous with >= nk*nv^3 elements")
    for t, n in ((Gij, nv * nv), (Xij, nv * nv), (dvv, nv), (Dov, nv), (S1, nv)):
        if t.numel() != n or not t.is_contiguous():
            raise B200ccError("t3_density_forms: accumulators must be contiguous (nv,nv) / (nv,) tensors")
    if fov.stride(1) != 1 and nv > 1:
        raise B200ccError("t3_density_forms: fov must be unit-stride along the virtual index")
    sc = _scratch(M3.device, max(1, int(_lib.get().b200cc_t3_density_scratch(int(nv)))))
    d = T3dDesc()
    d.no, d.nv, d.i, d.j, d.k0, d.nk = int(no), int(nv), int(i), int(j), int(k0), int(nk)
    d.swap_ab = int(bool(swap_ab))
    d.M3, d.t1, d.t2s, d.oovvs = _lib.ptr(M3), _lib.ptr(_c(t1, "t1")), _lib.ptr(_c(t2s, "t2s")), _lib.ptr(_c(oovvs, "oovvs"))
    d.fov, d.ldf = _lib.ptr(fov), int(fov.stride(0))
    d.eo, d.ev = _lib.ptr(eo), _lib.ptr(ev)
    d.W2ab, d.W2n, d.Pab, d.Pn = _lib.ptr(W2ab), _lib.ptr(W2n), _lib.ptr(Pab), _lib.ptr(Pn)
    d.Gij, d.Xij = _lib.ptr(Gij), _lib.ptr(Xij)
    d.dvv, d.Dov, d.S1 = _lib.ptr(dvv), _lib.ptr(Dov), _lib.ptr(S1)
    d.scratch = _lib.ptr(sc)
    _lib.check(_lib.get().b200cc_t3_density_forms(C.byref(d), _lib.stream()), "b200cc_t3_density_forms")


# ---- per-phase device timing of an iteration (bench.py "phases"; off unless B200CC_PHASES=1 / PHASES.on = True) ----------
class _Phases:
    """``with PHASES("ladder"):`` brackets a stretch of launches with CUDA events on the current stream; ``collect()``
    synchronises and returns {name: [calls, ms]} accumulated since the last collect.  A no-op when off."""

    def __init__(self):
        self.on = bool(int(os.environ.get("B200CC_PHASES", "0")))
        self._pending = []

    class _Span:
        def __init__(self, owner, name):
            self.owner, self.name = owner, name

        def __enter__(self):
            if self.owner.on and torch.cuda.is_available():
                self.a = torch.cuda.Event(enable_timing=True)
                self.a.record()
            return self

        def __exit__(self, *exc):
            if self.owner.on and torch.cuda.is_available():
                b = torch.cuda.Event(enable_timing=True)
                b.record()
                self.owner._pending.append((self.name, self.a, b))
            return False

    def __call__(self, name):
        return self._Span(self, name)

    def mark(self, name):
        """Sequential form: ends the stretch opened by the previous mark() and, if ``name`` is not None, opens one."""
        if not (self.on and torch.cuda.is_available()):
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        cur = getattr(self, "_open", None)
        if cur is not None:
            self._pending.append((cur[0], cur[1], ev))
        self._open = (name, ev) if name is not None else None

    def collect(self):
        out = {}
        if self._pending:
            torch.cuda.synchronize()
            for name, a, b in self._pending:
                k = out.setdefault(name, [0, 0.0])
                k[0] += 1
                k[1] += a.elapsed_time(b)
            self._pending = []
        return out


PHASES = _Phases()


def launch_count():
    return int(_lib.get().b200cc_launch_count())
