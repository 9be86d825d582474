"""TEST DOUBLE of the libb200cc C ABI (numpy, CPU) -- tests only, never imported by pycc_b200.

There is no GPU in the build container, so the host-side logic of the package (contraction
planning, pointer/stride arithmetic, the residual orchestration, DIIS bookkeeping, triple
batching, rank sharding) is exercised here against an object that exposes the SAME entry points as
``include/b200cc.h`` with the SAME raw-address calling convention, implemented with numpy on host
memory.  ``install()`` swaps it in for the ctypes library and lifts the "CUDA tensors only" check;
the product path itself has no such fallback (``pycc_b200._lib`` raises without the .so / a GPU).
"""
import ctypes as C

import numpy as np
from numpy.lib.stride_tricks import as_strided

from pycc_b200 import _lib


def _arr(addr, shape, strides):
    """float64 view of raw memory: element strides, non-negative."""
    if isinstance(addr, C.c_void_p):
        addr = addr.value
    shape = tuple(int(s) for s in shape)
    strides = tuple(int(s) for s in strides)
    n = 1 + sum((s - 1) * st for s, st in zip(shape, strides))
    if any(s == 0 for s in shape):
        return np.zeros(shape)
    buf = (C.c_double * n).from_address(int(addr))
    base = np.frombuffer(buf, dtype=np.float64)
    return as_strided(base, shape, tuple(8 * st for st in strides))


def _f32(addr, shape):
    """contiguous float32 view of raw memory"""
    if isinstance(addr, C.c_void_p):
        addr = addr.value
    n = int(np.prod(shape))
    buf = (C.c_float * n).from_address(int(addr))
    return np.frombuffer(buf, dtype=np.float32).reshape(shape)


def tf32_rna(f):
    """cvt.rna.tf32.f32: round a float32 array to 10 explicit mantissa bits, ties away from zero"""
    u = np.ascontiguousarray(f, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def _vec(addr, n):
    return _arr(addr, (n,), (1,))


def _ints(addr, n, ctype=C.c_int):
    buf = (ctype * n).from_address(int(addr))
    return np.frombuffer(buf, dtype=np.int32 if ctype is C.c_int else np.int64)


class EmuLib:
    def __init__(self):
        self.launches = 0
        self.err = b""
        self.calls = {}

    def _count(self, name, n=1):
        self.launches += n
        self.calls[name] = self.calls.get(name, 0) + 1

    def b200cc_version(self):
        return 200

    def b200cc_last_error(self):
        return self.err

    def b200cc_launch_count(self):
        return self.launches

    # ---- GEMM -----------------------------------------------------------------------------------
    def b200cc_dgemm(self, dref, stream):
        d = dref._obj
        self._count("dgemm")
        M, N = d.M, d.N
        if M == 0 or N == 0 or d.batch == 0:
            return 0
        tab = None
        if d.table:
            tab = _ints(d.table, 5 * d.batch, C.c_longlong).reshape(d.batch, 5)

        def mat(addr, rows, K, ld, trans):
            if K == 0:
                return np.zeros((rows, 0))
            return _arr(addr, (rows, K), (1, ld) if trans else (ld, 1))

        bc = None
        nbc = 8 if d.K3 > 0 else 4
        if d.bcoords:
            bc = _ints(d.bcoords, nbc * d.batch).reshape(d.batch, nbc)
        for b in range(d.batch):
            if tab is not None:
                a1, b1, a2, b2, c = (int(x) for x in tab[b])
            elif bc is not None:
                a1 = d.A1 + 8 * int(bc[b, 0]) * d.strideA1
                b1 = d.B1 + 8 * int(bc[b, 1]) * d.strideB1
                a2 = (d.A2 or 0) + 8 * int(bc[b, 2]) * d.strideA2
                b2 = (d.B2 or 0) + 8 * int(bc[b, 3]) * d.strideB2
                c = d.C + 8 * b * d.strideC
            else:
                a1 = d.A1 + 8 * b * d.strideA1
                b1 = d.B1 + 8 * b * d.strideB1
                a2 = (d.A2 or 0) + 8 * b * d.strideA2
                b2 = (d.B2 or 0) + 8 * b * d.strideB2
                c = d.C + 8 * b * d.strideC
            acc = mat(a1, M, d.K1, d.lda1, d.transA) @ mat(b1, N, d.K1, d.ldb1, d.transB).T
            if d.K2 > 0:
                acc = acc + mat(a2, M, d.K2, d.lda2, d.transA) @ mat(b2, N, d.K2, d.ldb2, d.transB).T
            if d.K3 > 0:                 # third / fourth K segment (K-major only), slab indices bc[b, 4:8] or b
                ix = [int(x) for x in bc[b, 4:8]] if bc is not None else [b] * 4
                acc = acc + (mat(d.A3 + 8 * ix[0] * d.strideA3, M, d.K3, d.lda3, 0)
                             @ mat(d.B3 + 8 * ix[1] * d.strideB3, N, d.K3, d.ldb3, 0).T)
                if d.K4 > 0:
                    acc = acc + (mat(d.A4 + 8 * ix[2] * d.strideA4, M, d.K4, d.lda4, 0)
                                 @ mat(d.B4 + 8 * ix[3] * d.strideB4, N, d.K4, d.ldb4, 0).T)
            if d.out_cube_nv > 0:        # (T): C written as contiguous 8x8x8 cubes, row = x*nv + y, col = z
                nv = d.out_cube_nv
                nc8 = (nv + 7) // 8
                buf = _vec(c, nc8 ** 3 * 512).reshape(nc8, nc8, nc8, 8, 8, 8)
                full = np.zeros((nc8 * 8, nc8 * 8, nc8 * 8))
                full[:nv, :nv, :N] = (d.alpha * acc).reshape(nv, nv, N)
                buf[...] = full.reshape(nc8, 8, nc8, 8, nc8, 8).transpose(0, 2, 4, 1, 3, 5)
                continue
            Cm = _arr(c, (M, N), (d.ldc, 1))
            if d.beta != 0.0:
                Cm[...] = d.alpha * acc + d.beta * Cm
            else:
                Cm[...] = d.alpha * acc
        return 0

    # ---- mixed precision: TF32 split + 3-product GEMM (FP64 arithmetic on the split values) ---------
    def b200cc_split_tf32(self, src, ld, stride, rows, K, batch, hi, lo, ldp, stream):
        self._count("split_tf32")
        if rows <= 0 or K <= 0 or batch <= 0:
            return 0
        x = _arr(src, (batch, rows, K), (stride, ld, 1))
        H = _f32(hi, (batch, rows, ldp))
        L = _f32(lo, (batch, rows, ldp))
        h = tf32_rna(x.astype(np.float32))
        H[...] = 0.0
        L[...] = 0.0
        H[:, :, :K] = h
        L[:, :, :K] = tf32_rna((x - h.astype(np.float64)).astype(np.float32))
        return 0

    def b200cc_merge_tf32(self, hi, lo, ldp, rows, K, dst, ld, stream):
        self._count("merge_tf32")
        if rows <= 0 or K <= 0:
            return 0
        H = _f32(hi, (rows, ldp))
        L = _f32(lo, (rows, ldp))
        _arr(dst, (rows, K), (ld, 1))[...] = H[:, :K].astype(np.float64) + L[:, :K].astype(np.float64)
        return 0

    def b200cc_gemm_tf32x3(self, dref, stream):
        d = dref._obj
        self._count("gemm_tf32x3")
        if d.M <= 0 or d.N <= 0 or d.batch <= 0:
            return 0
        bc = _ints(d.bcoords, 4 * d.batch).reshape(d.batch, 4) if d.bcoords else None

        def pl(addr, rows, ld, st, b, Kd):
            return _f32(addr + 4 * b * st, (rows, ld))[:, :Kd].astype(np.float64)

        for b in range(d.batch):
            i1, j1, i2, j2 = (int(x) for x in bc[b]) if bc is not None else (b, b, b, b)
            ah, al = pl(d.Ahi, d.M, d.lda, d.strideA, i1, d.K), pl(d.Alo, d.M, d.lda, d.strideA, i1, d.K)
            bh, bl = pl(d.Bhi, d.N, d.ldb, d.strideB, j1, d.K), pl(d.Blo, d.N, d.ldb, d.strideB, j1, d.K)
            acc = ah @ bh.T + ah @ bl.T + al @ bh.T
            if d.K2 > 0:
                ah, al = pl(d.A2hi, d.M, d.lda2, d.strideA2, i2, d.K2), pl(d.A2lo, d.M, d.lda2, d.strideA2, i2, d.K2)
                bh, bl = pl(d.B2hi, d.N, d.ldb2, d.strideB2, j2, d.K2), pl(d.B2lo, d.N, d.ldb2, d.strideB2, j2, d.K2)
                acc = acc + ah @ bh.T + ah @ bl.T + al @ bh.T
            Cm = _arr(d.C + 8 * b * d.strideC, (d.M, d.N), (d.ldc, 1))
            if d.beta != 0.0:
                Cm[...] = d.alpha * acc + d.beta * Cm
            else:
                Cm[...] = d.alpha * acc
        return 0

    # ---- ladder in pair form (csrc/pairs.cu) ------------------------------------------------------------
    @staticmethod
    def _pairs(n):
        """(x, y) index arrays of the pairs x >= y in pair(x,y) = x(x+1)/2 + y order"""
        x, y = np.tril_indices(n)
        return x, y

    def b200cc_pair_count(self, n):
        return n * (n + 1) // 2

    def b200cc_pack_pairs(self, src, sa, sb, se, sf, nv, a0, a1, vp, vm, ldq, stream):
        self._count("pack_pairs")
        na = a1 - a0
        if na <= 0:
            return 0
        X = _arr(src, (na, nv, nv, nv), (sa, sb, se, sf))
        e, f = self._pairs(nv)
        nq = len(e)
        npairs = a1 * (a1 + 1) // 2 - a0 * (a0 + 1) // 2
        P = _arr(vp, (npairs, ldq), (ldq, 1))
        Mn = _arr(vm, (npairs, ldq), (ldq, 1))
        P[...] = 0.0
        Mn[...] = 0.0
        row = 0
        for a in range(a0, a1):
            for b in range(a + 1):
                S = X[a - a0, b]
                x, y = S[e, f], S[f, e]
                P[row, :nq] = np.where(e == f, x, x + y)
                Mn[row, :nq] = 0.0 if a == b else np.where(e == f, 0.0, x - y)
                row += 1
        return 0

    def b200cc_unpack_pairs(self, vp, vm, ldq, nv, npairs, dst, stream):
        self._count("unpack_pairs")
        if npairs <= 0:
            return 0
        P = _arr(vp, (npairs, ldq), (ldq, 1))
        Mn = _arr(vm, (npairs, ldq), (ldq, 1))
        D = _vec(dst, npairs * nv * nv).reshape(npairs, nv, nv)
        e, f = self._pairs(nv)
        nq = len(e)
        lower = 0.5 * (P[:, :nq] + Mn[:, :nq])
        upper = 0.5 * (P[:, :nq] - Mn[:, :nq])
        D[:, f, e] = upper
        D[:, e, f] = lower
        dg = e == f
        D[:, e[dg], f[dg]] = P[:, :nq][:, dg]
        return 0

    def _rows(self, no, tri):
        if tri:
            i, j = self._pairs(no)
        else:
            i, j = np.divmod(np.arange(no * no), no)
        return i, j

    def b200cc_pack_tau(self, tau, no, nv, tri, tp, tm, ldq, stream):
        self._count("pack_tau")
        T = _vec(tau, no * no * nv * nv).reshape(no, no, nv, nv)
        i, j = self._rows(no, tri)
        e, f = self._pairs(nv)
        nq, M = len(e), len(i)
        P = _arr(tp, (M, ldq), (ldq, 1))
        Mn = _arr(tm, (M, ldq), (ldq, 1))
        P[...] = 0.0
        Mn[...] = 0.0
        x = T[i[:, None], j[:, None], e[None, :], f[None, :]]
        y = T[i[:, None], j[:, None], f[None, :], e[None, :]]
        P[:, :nq] = np.where(e == f, x, 0.5 * (x + y))
        Mn[:, :nq] = np.where(e == f, 0.0, 0.5 * (x - y))
        return 0

    def b200cc_pack_rows(self, src, nrows, nv, xp, xm, ldq, stream):
        self._count("pack_rows")
        if nrows <= 0:
            return 0
        X = _vec(src, nrows * nv * nv).reshape(nrows, nv, nv)
        e, f = self._pairs(nv)
        nq = len(e)
        P = _arr(xp, (nrows, ldq), (ldq, 1))
        Mn = _arr(xm, (nrows, ldq), (ldq, 1))
        P[...] = 0.0
        Mn[...] = 0.0
        x, y = X[:, e, f], X[:, f, e]
        P[:, :nq] = np.where(e == f, x, x + y)
        Mn[:, :nq] = np.where(e == f, 0.0, x - y)
        return 0

    def b200cc_pair_rows_unpack(self, S, A, lds, no, ncols, out, ldo, stream):
        self._count("pair_rows_unpack")
        if ncols <= 0:
            return 0
        i, j = self._pairs(no)
        Sm = _arr(S, (len(i), ncols), (lds, 1))
        Am = _arr(A, (len(i), ncols), (lds, 1))
        O = _arr(out, (no * no, ncols), (ldo, 1))
        O[i * no + j] = Sm + Am
        off = i != j
        O[j[off] * no + i[off]] = (Sm - Am)[off]
        return 0

    def b200cc_ring_layouts(self, t2, no, nv, u, tb, stream):
        self._count("ring_layouts")
        n = no * no * nv * nv
        T = _vec(t2, n).reshape(no, no, nv, nv)
        _vec(u, n).reshape(no, nv, no, nv)[...] = (2.0 * T - T.transpose(0, 1, 3, 2)).transpose(0, 2, 1, 3)
        _vec(tb, n).reshape(no, nv, no, nv)[...] = T.transpose(0, 3, 1, 2)
        return 0

    def b200cc_ladder_unpack(self, S, A, lds, no, nv, tri, a0, a1, alpha, r2, stream):
        self._count("ladder_unpack")
        if a1 <= a0:
            return 0
        R = _vec(r2, no * no * nv * nv).reshape(no, no, nv, nv)
        i, j = self._rows(no, tri)
        M = len(i)
        npl = a1 * (a1 + 1) // 2 - a0 * (a0 + 1) // 2
        Sm = _arr(S, (M, npl), (lds, 1))
        Am = _arr(A, (M, npl), (lds, 1))
        a, b = self._pairs(a1)
        keep = a >= a0
        a, b = a[keep], b[keep]
        off = a != b
        for m in range(M):
            u = alpha * (Sm[m] + Am[m])
            w = alpha * (Sm[m] - Am[m])
            R[i[m], j[m], a, b] += u
            R[i[m], j[m], b[off], a[off]] += w[off]
            if tri and i[m] != j[m]:
                R[j[m], i[m], a, b] += w
                R[j[m], i[m], b[off], a[off]] += u[off]
        return 0

    # ---- permute / axpby ---------------------------------------------------------------------------
    def b200cc_permute(self, rank, shape, si, so, alpha, inp, beta, out, stream):
        self._count("permute")
        shp = [int(shape[d]) for d in range(rank)]
        a = _arr(inp, shp, [int(si[d]) for d in range(rank)])
        o = _arr(out, shp, [int(so[d]) for d in range(rank)])
        if beta != 0.0:
            o[...] = alpha * a + beta * o
        else:
            o[...] = alpha * a
        return 0

    def b200cc_axpbyz(self, n, a, x, b, y, z, stream):
        self._count("axpbyz")
        r = 0.0
        if a != 0.0:
            r = r + a * _vec(x, n)
        if b != 0.0:
            r = r + b * _vec(y, n)
        _vec(z, n)[...] = r
        return 0

    # ---- elementwise --------------------------------------------------------------------------------
    def b200cc_build_tau(self, no, nv, f1, f2, t1, t2, tau, stream):
        self._count("tau")
        T1 = _arr(t1, (no, nv), (nv, 1))
        T2 = _vec(t2, no * no * nv * nv).reshape(no, no, nv, nv)
        _vec(tau, no * no * nv * nv).reshape(no, no, nv, nv)[...] = f1 * T2 + f2 * np.einsum("ia,jb->ijab", T1, T1)
        return 0

    @staticmethod
    def _d2(no, nv, eo, ev):
        eo, ev = _vec(eo, no), _vec(ev, nv)
        return eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev

    def b200cc_div_d2(self, no, nv, eo, ev, inp, out, stream):
        self._count("div_d2")
        n = no * no * nv * nv
        _vec(out, n)[...] = (_vec(inp, n).reshape(no, no, nv, nv) / self._d2(no, nv, eo, ev)).ravel()
        return 0

    def b200cc_div_d1(self, no, nv, eo, ev, inp, out, stream):
        self._count("div_d1")
        D = _vec(eo, no)[:, None] - _vec(ev, nv)[None, :]
        _vec(out, no * nv)[...] = (_vec(inp, no * nv).reshape(no, nv) / D).ravel()
        return 0

    def b200cc_update_amps(self, no, nv, eo, ev, r1, r2, symmetrize, write_r2, t1, t2, sumsq, scratch, stream):
        self._count("update_amps", 2)
        n2 = no * no * nv * nv
        R2 = _vec(r2, n2).reshape(no, no, nv, nv)
        full = R2 + R2.transpose(1, 0, 3, 2) if symmetrize else R2.copy()
        if symmetrize and write_r2:
            R2[...] = full
        d2 = full / self._d2(no, nv, eo, ev)
        _vec(t2, n2)[...] += d2.ravel()
        s = float(np.sum(d2 * d2))
        if r1:
            D = _vec(eo, no)[:, None] - _vec(ev, nv)[None, :]
            d1 = _vec(r1, no * nv).reshape(no, nv) / D
            _vec(t1, no * nv)[...] += d1.ravel()
            s += float(np.sum(d1 * d1))
        _vec(sumsq, 1)[0] = s
        return 0

    def b200cc_update_amps_rows(self, no, nv, i0, i1, eo, ev, r1, r2, t1, t2, sumsq2, scratch, stream):
        self._count("update_amps_rows", 2)
        n2 = no * no * nv * nv
        R2 = _vec(r2, n2).reshape(no, no, nv, nv)
        full = (R2 + R2.transpose(1, 0, 3, 2))[i0:i1]
        d2 = full / self._d2(no, nv, eo, ev)[i0:i1]
        _vec(t2, n2).reshape(no, no, nv, nv)[i0:i1] += d2
        out = _vec(sumsq2, 2)
        out[0], out[1] = float(np.sum(d2 * d2)), 0.0
        if r1:
            D = _vec(eo, no)[:, None] - _vec(ev, nv)[None, :]
            d1 = _vec(r1, no * nv).reshape(no, nv) / D
            _vec(t1, no * nv)[...] += d1.ravel()
            out[1] = float(np.sum(d1 * d1))
        return 0

    def b200cc_cc_energy_rows(self, no, nv, i0, i1, with_singles, fov, ldf, t1, t2, L, e_out, scratch, stream):
        self._count("cc_energy_rows", 2)
        f = _arr(fov, (no, nv), (ldf, 1))
        T1 = _arr(t1, (no, nv), (nv, 1))
        n2 = no * no * nv * nv
        tau = _vec(t2, n2).reshape(no, no, nv, nv) + np.einsum("ia,jb->ijab", T1, T1)
        e = np.sum(tau[i0:i1] * _vec(L, n2).reshape(no, no, nv, nv)[i0:i1])
        if with_singles:
            e += 2.0 * np.sum(f * T1)
        _vec(e_out, 1)[0] = e
        return 0

    def b200cc_symmetrize_r2(self, no, nv, r2, stream):
        self._count("symmetrize")
        R2 = _vec(r2, no * no * nv * nv).reshape(no, no, nv, nv)
        R2[...] = R2 + R2.transpose(1, 0, 3, 2)
        return 0

    def b200cc_cc_energy(self, no, nv, fov, ldf, t1, t2, L, e_out, scratch, stream):
        self._count("cc_energy", 2)
        f = _arr(fov, (no, nv), (ldf, 1))
        T1 = _arr(t1, (no, nv), (nv, 1))
        n2 = no * no * nv * nv
        tau = _vec(t2, n2).reshape(no, no, nv, nv) + np.einsum("ia,jb->ijab", T1, T1)
        _vec(e_out, 1)[0] = 2.0 * np.sum(f * T1) + np.sum(tau * _vec(L, n2).reshape(no, no, nv, nv))
        return 0

    def b200cc_multi_dot(self, n, x, m, ys, out, scratch, stream):
        self._count("multi_dot", 2)
        X = _vec(x, n)
        o = _vec(out, m)
        for q in range(m):
            o[q] = float(np.dot(X, _vec(ys[q], n)))
        return 0

    def b200cc_multi_axpy(self, n, m, c, xs, out, stream):
        self._count("multi_axpy")
        r = np.zeros(n)
        for q in range(m):
            r += c[q] * _vec(xs[q], n)
        _vec(out, n)[...] = r
        return 0

    # ---- (T) ------------------------------------------------------------------------------------------
    def b200cc_t_q_size(self, nv, blocked):
        return ((nv + 7) // 8) ** 3 * 512 if (blocked & 1) else nv ** 3

    @staticmethod
    def _unblock(Q6, nv, blocked):
        """[6][qsize] -> [6][nv^3] plain"""
        if not blocked:
            return Q6.reshape(6, nv, nv, nv)
        nc8 = (nv + 7) // 8
        q = Q6.reshape(6, nc8, nc8, nc8, 8, 8, 8).transpose(0, 1, 4, 2, 5, 3, 6).reshape(6, nc8 * 8, nc8 * 8, nc8 * 8)
        return q[:, :nv, :nv, :nv]

    def b200cc_t_energy_scratch(self, nv, ntrip):
        nt = (nv + 7) // 8
        return nt * (nt + 1) * (nt + 2) // 6 * ntrip

    @staticmethod
    def _Wp(R, nv):
        """paired layout: W[a,b,c] = R1[a,b,c] + R2[a,c,b] + R3[c,b,a]"""
        return R[0] + R[1].transpose(0, 2, 1) + R[2].transpose(2, 1, 0)

    def _Wq(self, Q, n, nv, flags):
        """W of triple n from the Q buffer; flags bit 0 = cube-blocked, bit 1 = paired (3 arrays per triple)"""
        blocked, paired = flags & 1, flags & 2
        nq = 3 if paired else 6
        v3 = self.b200cc_t_q_size(nv, blocked)
        raw = _vec(Q + 8 * n * nq * v3, nq * v3)
        if not blocked:
            arr = raw.reshape(nq, nv, nv, nv)
        else:
            nc8 = (nv + 7) // 8
            arr = raw.reshape(nq, nc8, nc8, nc8, 8, 8, 8).transpose(0, 1, 4, 2, 5, 3, 6).reshape(
                nq, nc8 * 8, nc8 * 8, nc8 * 8)[:, :nv, :nv, :nv]
        return self._Wp(arr, nv) if paired else self._W(arr, nv)

    @staticmethod
    def _W(Q, nv):
        return (Q[0] + Q[1].transpose(0, 2, 1) + Q[2].transpose(1, 2, 0) + Q[3].transpose(2, 1, 0)
                + Q[4].transpose(2, 0, 1) + Q[5].transpose(1, 0, 2))

    @staticmethod
    def _disc(no, nv, i, j, k, t1, t2, oovv, fov, ldf):
        T1 = _arr(t1, (no, nv), (nv, 1))
        n2 = no * no * nv * nv
        T2 = _vec(t2, n2).reshape(no, no, nv, nv)
        Kv = _vec(oovv, n2).reshape(no, no, nv, nv)
        f = _arr(fov, (no, nv), (ldf, 1))
        e = np.einsum
        return (e("ab,c->abc", Kv[i, j], T1[k]) + e("ac,b->abc", Kv[i, k], T1[j]) + e("bc,a->abc", Kv[j, k], T1[i])
                + e("ab,c->abc", T2[i, j], f[k]) + e("ac,b->abc", T2[i, k], f[j]) + e("bc,a->abc", T2[j, k], f[i]))

    @staticmethod
    def _den(no, nv, i, j, k, eo, ev):
        eo, ev = _vec(eo, no), _vec(ev, nv)
        return eo[i] + eo[j] + eo[k] - ev[:, None, None] - ev[None, :, None] - ev[None, None, :]

    def b200cc_t_energy_batch(self, no, nv, ntrip, ijk, Q, blocked, t1, t2, oovv, fov, ldf, eo, ev, et, accumulate,
                              scratch, stream):
        self._count("t_energy", 2)
        trip = _ints(ijk, 3 * ntrip).reshape(ntrip, 3)
        a = np.arange(nv)
        eq = ((a[:, None, None] == a[None, :, None]).astype(float) + (a[:, None, None] == a[None, None, :])
              + (a[None, :, None] == a[None, None, :]))
        mask = (a[:, None, None] >= a[None, :, None]) & (a[None, :, None] >= a[None, None, :])
        tot = 0.0
        for n in range(ntrip):
            i, j, k = (int(x) for x in trip[n])
            W = self._Wq(Q, n, nv, blocked)
            V = (W + self._disc(no, nv, i, j, k, t1, t2, oovv, fov, ldf)) / (1.0 + eq)
            p = lambda X, *ax: X.transpose(*ax)
            X3 = (W * V + p(W, 0, 2, 1) * p(V, 0, 2, 1) + p(W, 1, 0, 2) * p(V, 1, 0, 2) + p(W, 1, 2, 0) * p(V, 1, 2, 0)
                  + p(W, 2, 0, 1) * p(V, 2, 0, 1) + p(W, 2, 1, 0) * p(V, 2, 1, 0))
            Y = V + p(V, 1, 2, 0) + p(V, 2, 0, 1)
            Z = p(V, 0, 2, 1) + p(V, 1, 0, 2) + p(V, 2, 1, 0)
            Wc = W + p(W, 1, 2, 0) + p(W, 2, 0, 1)
            Wo = p(W, 0, 2, 1) + p(W, 1, 0, 2) + p(W, 2, 1, 0)
            occ = 2.0 - (float(i == j) + float(i == k) + float(j == k))
            e = ((Y - 2 * Z) * Wc + (Z - 2 * Y) * Wo + 3 * X3) * occ / self._den(no, nv, i, j, k, eo, ev)
            tot += float(np.sum(e[mask]))
        o = _vec(et, 1)
        o[0] = o[0] + tot if accumulate else tot
        return 0

    def b200cc_t3_assemble(self, no, nv, i, j, k, Q, blocked, t1, t2, oovv, fov, ldf, eo, ev, with_denom, w3, d3,
                           stream):
        self._count("t3_assemble")
        v3 = nv ** 3
        W = self._Wq(Q, 0, nv, blocked)
        Dd = self._disc(no, nv, i, j, k, t1, t2, oovv, fov, ldf)
        if with_denom:
            den = self._den(no, nv, i, j, k, eo, ev)
            W, Dd = W / den, Dd / den
        _vec(w3, v3)[...] = W.ravel()
        if d3:
            _vec(d3, v3)[...] = Dd.ravel()
        return 0


    # ---- (T), fused (a,b,c)-driven form ------------------------------------------------------------------
    def b200cc_t_abc_max_no(self):
        return 64

    def b200cc_t_abc(self, dref, stream):
        d = dref._obj
        self._count("t_abc", 2)
        no, nv = d.no, d.nv
        if no % 2 or nv % 2 or no > 64 or nv > 1023:
            self.err = b"b200cc_t_abc: needs even o <= 64 and even v <= 1023"
            return 1
        o2, o3, v2 = no * no, no ** 3, nv * nv
        G = _vec(d.G, no * nv ** 3).reshape(no, nv, nv, nv)            # [l,x,y,e]
        t2 = _vec(d.t2, o2 * v2).reshape(no, no, nv, nv)
        t2x = _vec(d.t2x, o2 * v2).reshape(nv, nv, no, no)              # [x,y,l,m]
        Ox = _vec(d.Ox, nv * o3).reshape(nv, no, no, no)                # [z,p,q,m] = -<mz|pq>
        Kx = _vec(d.oovvx, o2 * v2).reshape(nv, nv, no, no)             # [x,y,i,j]
        t1 = _arr(d.t1, (no, nv), (nv, 1))
        f = _arr(d.fov, (no, nv), (d.ldf, 1))
        eo, ev = _vec(d.eo, no), _vec(d.ev, nv)
        abc = _ints(d.abc, d.nabc)
        srt = _ints(d.sorted, d.nsorted)
        si, sj, sk = srt & 1023, (srt >> 10) & 1023, (srt >> 20) & 1023
        e = np.einsum

        def Gf(x, y, z):                                                # [l,p,q]
            return (e("le,qpe->lpq", G[:, x, y], t2[:, :, z]) + e("le,pqe->lpq", G[:, x, z], t2[:, :, y])
                    + e("lm,pqm->lpq", t2x[x, y], Ox[z]) + e("lm,qpm->lpq", t2x[x, z], Ox[y]))

        grid = min(d.grid, d.nabc)
        tiles = _vec(d.wtile, max(grid, 1) * o3).reshape(max(grid, 1), no, no, no)
        tot = 0.0
        for n in range(d.nabc):
            a, b, c = int(abc[n]) & 1023, (int(abc[n]) >> 10) & 1023, (int(abc[n]) >> 20) & 1023
            W = Gf(a, b, c) + Gf(c, b, a).transpose(2, 1, 0) + Gf(b, c, a).transpose(2, 0, 1)
            tiles[n % grid] = W
            D = (e("ij,k->ijk", Kx[a, b], t1[:, c]) + e("ik,j->ijk", Kx[a, c], t1[:, b]) + e("jk,i->ijk", Kx[b, c], t1[:, a])
                 + e("ij,k->ijk", t2x[a, b], f[:, c]) + e("ik,j->ijk", t2x[a, c], f[:, b]) + e("jk,i->ijk", t2x[b, c], f[:, a]))
            w = lambda X, i, j, k: X[i, j, k]
            sc = 1.0 / (1.0 + (si == sj).astype(float) + (si == sk) + (sj == sk))
            perms = {"ijk": (si, sj, sk), "ikj": (si, sk, sj), "jik": (sj, si, sk), "jki": (sj, sk, si),
                     "kij": (sk, si, sj), "kji": (sk, sj, si)}
            Wp = {key: W[ix] for key, ix in perms.items()}
            Vp = {key: (W[ix] + D[ix]) * sc for key, ix in perms.items()}
            X = sum(Wp[key] * Vp[key] for key in perms)
            Y = Vp["ijk"] + Vp["jki"] + Vp["kij"]
            Z = Vp["ikj"] + Vp["jik"] + Vp["kji"]
            Wc = Wp["ijk"] + Wp["jki"] + Wp["kij"]
            Wo = Wp["ikj"] + Wp["jik"] + Wp["kji"]
            den = eo[si] + eo[sj] + eo[sk] - (ev[a] + ev[b] + ev[c])
            wabc = 2.0 - (float(a == b) + float(a == c) + float(b == c))
            tot += wabc * float(np.sum(((Y - 2 * Z) * Wc + (Z - 2 * Y) * Wo + 3 * X) / den))
        o = _vec(d.et_out, 1)
        o[0] = o[0] + tot if d.accumulate else tot
        return 0


    # ---- (T) densities -------------------------------------------------------------------------------
    def b200cc_t3_connected_batch(self, no, nv, ntrip, ijk, Q, eo, ev, m3, stream):
        self._count("t3_connected")
        trip = _ints(ijk, 3 * ntrip).reshape(ntrip, 3)
        v3 = nv ** 3
        out = _vec(m3, ntrip * v3).reshape(ntrip, nv, nv, nv)
        for n in range(ntrip):
            i, j, k = (int(x) for x in trip[n])
            W = self._W(_vec(Q + 8 * n * 6 * v3, 6 * v3).reshape(6, nv, nv, nv), nv)
            out[n] = W / self._den(no, nv, i, j, k, eo, ev)
        return 0

    def b200cc_t3_density_scratch(self, nv):
        return 3 * nv * ((nv + 7) // 8)

    def b200cc_t3_density_forms(self, dref, stream):
        d = dref._obj
        self._count("t3_density_forms", 4)
        no, nv, i, j, nk = d.no, d.nv, d.i, d.j, d.nk
        v3 = nv ** 3
        M = _vec(d.M3, nk * v3).reshape(nk, nv, nv, nv)
        T1 = _arr(d.t1, (no, nv), (nv, 1))
        Ts = _vec(d.t2s, no * no * nv * nv).reshape(no, no, nv, nv)
        Ks = _vec(d.oovvs, no * no * nv * nv).reshape(no, no, nv, nv)
        f = _arr(d.fov, (no, nv), (d.ldf, 1))
        W2ab = _vec(d.W2ab, nk * v3).reshape(nv, nv, nk, nv)
        W2n = _vec(d.W2n, nk * v3).reshape(nv, nk, nv, nv)
        Pab = _vec(d.Pab, nk * v3).reshape(nv, nv, nk, nv)
        Pn = _vec(d.Pn, nk * v3).reshape(nv, nk, nv, nv)
        Gij, Xij = _vec(d.Gij, nv * nv).reshape(nv, nv), _vec(d.Xij, nv * nv).reshape(nv, nv)
        dvv, Dov, S1 = _vec(d.dvv, nv), _vec(d.Dov, nv), _vec(d.S1, nv)

        def sym(A):
            return (8.0 * A - 4.0 * A.transpose(1, 0, 2) - 4.0 * A.transpose(0, 2, 1) - 4.0 * A.transpose(2, 1, 0)
                    + 2.0 * A.transpose(2, 0, 1) + 2.0 * A.transpose(1, 2, 0))

        def symd(Aij, Aik, Ajk, wi, wj, wk):
            """sym() of  Aij[x,y] wk[z] + Aik[x,z] wj[y] + Ajk[y,z] wi[x]  from the A~ = 4A - 2A^T combinations"""
            e = np.einsum
            return (2.0 * e("ab,c->abc", Aij, wk) - e("ac,b->abc", Aij, wk) - e("cb,a->abc", Aij, wk)
                    + 2.0 * e("ac,b->abc", Aik, wj) - e("bc,a->abc", Aik, wj) - e("ab,c->abc", Aik, wj)
                    + 2.0 * e("bc,a->abc", Ajk, wi) - e("ac,b->abc", Ajk, wi) - e("ba,c->abc", Ajk, wi))
        e = np.einsum
        for kk in range(nk):
            k = d.k0 + kk
            M3 = M[kk].transpose(1, 0, 2) if d.swap_ab else M[kk]
            Y3 = (symd(Ks[i, j], Ks[i, k], Ks[j, k], T1[i], T1[j], T1[k])
                  + symd(Ts[i, j], Ts[i, k], Ts[j, k], f[i], f[j], f[k])) / self._den(no, nv, i, j, k, d.eo, d.ev)
            X3 = sym(M3)
            W2 = 2.0 * X3 + Y3
            P = 2.0 * M3 - M3.transpose(0, 2, 1) - M3.transpose(2, 1, 0)
            U = M3 - M3.transpose(2, 1, 0)
            Z3 = 2.0 * (M3 - M3.transpose(0, 2, 1)) - (M3.transpose(1, 0, 2) - M3.transpose(2, 0, 1))
            W2ab[:, :, kk, :] = W2
            W2n[:, kk] = W2
            Pab[:, :, kk, :] = P
            Pn[:, kk] = P
            Gij += 4.0 * e("c,abc->ab", T1[k], Z3)
            Xij += e("abc,c->ab", U, f[k])
            dvv += 0.5 * e("abc,abc->a", M3, X3 + Y3)
            Dov += e("abc,bc->a", U, Ts[j, k])
            S1 += e("abc,bc->a", M3 - M3.transpose(1, 0, 2), Ks[j, k])
        return 0


class install:
    """Context manager / fixture helper: route pycc_b200 through the numpy double on CPU tensors."""

    def __enter__(self):
        self.saved = (_lib._LIB, _lib.REQUIRE_CUDA)
        self.lib = EmuLib()
        _lib._LIB = self.lib
        _lib.REQUIRE_CUDA = False
        return self.lib

    def __exit__(self, *exc):
        _lib._LIB, _lib.REQUIRE_CUDA = self.saved
        return False
