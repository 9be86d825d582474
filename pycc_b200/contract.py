"""Two-operand tensor contraction on the DMMA GEMM (transpose-transpose-GEMM-transpose).

``contract('ijef,abef->ijab', tau, vvvv)`` classifies the indices into batch / M / N / K groups,
looks at the operands' *strides* to see whether each can be handed to ``b200cc_dgemm`` as a
(batched) matrix in place -- either orientation, any leading dimension, sliced views included --
and only when it cannot does it make a permuted copy with ``b200cc_permute``.  The output is
written in place when its layout is a (batched) row-major matrix in (M,N) or (N,M) order
(the operand roles are swapped for the latter); otherwise the product goes to a temporary and one
permute pass folds ``alpha``/``beta`` into the destination.

This is the replacement for the reference's ``opt_einsum.contract`` -> ``torch.tensordot`` ->
cuBLAS path (pycc/device.py:64-86); plans are cached per (subscripts, shapes, strides).
"""
from __future__ import annotations

import torch

from . import kernels as K
from ._lib import B200ccError


def _parse(sub, nops):
    sub = sub.replace(" ", "")
    if "->" in sub:
        lhs, out = sub.split("->")
    else:
        lhs = sub
        seen = {}
        for ch in lhs.replace(",", ""):
            seen[ch] = seen.get(ch, 0) + 1
        out = "".join(sorted(ch for ch, n in seen.items() if n == 1))
    ins = lhs.split(",")
    if len(ins) != nops:
        raise B200ccError("contract: %d operands for subscripts %r" % (nops, sub))
    return ins, out


def _merge(idx, shape, stride, group):
    """Can the indices in ``group`` (in this order) be fused into ONE strided dimension of a tensor
    whose index string is ``idx``?  Returns (extent, stride) or None.  Extent-1 indices are free."""
    ext, st = 1, None
    for ch in reversed(group):
        p = idx.index(ch)
        n, s = shape[p], stride[p]
        if n == 1:
            continue
        if st is None:
            st = s
            ext = n
        else:
            if s != st * ext:
                return None
            ext *= n
    return ext, (st if st is not None else 1)


class _Mat:
    """A tensor seen as batch x rows x cols with element strides (None if not expressible)."""
    __slots__ = ("ok", "bs", "rs", "cs")

    def __init__(self, idx, shape, stride, bgrp, rgrp, cgrp):
        b = _merge(idx, shape, stride, bgrp)
        r = _merge(idx, shape, stride, rgrp)
        c = _merge(idx, shape, stride, cgrp)
        self.ok = b is not None and r is not None and c is not None
        if self.ok:
            self.bs, self.rs, self.cs = b[1], r[1], c[1]


def _orient(m, nrows, ncols):
    """(trans, ld) if the rows x cols view can feed the GEMM: trans=0 -> cols contiguous
    ("K-major"), trans=1 -> rows contiguous.  None otherwise."""
    if not m.ok:
        return None
    if ncols == 1 or m.cs == 1:
        ld = m.rs if nrows > 1 else max(ncols, 1)
        if ld >= 1:
            return 0, ld
    if nrows == 1 or m.rs == 1:
        ld = m.cs if ncols > 1 else max(nrows, 1)
        if ld >= 1:
            return 1, ld
    return None


def _prod(shape, idx, group):
    p = 1
    for ch in group:
        p *= shape[idx.index(ch)]
    return p


class Contractor:
    """Callable with the signature of the reference's ``ContractionBackend.__call__``
    (pycc/device.py:64): ``contract(subscripts, *operands)`` -> new tensor; plus the in-place form
    ``contract(sub, A, B, out=C, alpha=..., beta=...)`` used by the fused residual."""

    def __init__(self):
        self._plans = {}
        self.stats = {"gemm": 0, "permute": 0}

    # -- public -------------------------------------------------------------------------------
    def __call__(self, sub, *ops, out=None, alpha=1.0, beta=0.0):
        if len(ops) == 1:
            return self._unary(sub, ops[0], out, alpha, beta)
        if len(ops) > 2:
            # left-to-right pairwise; intermediate keeps every index still needed later
            ins, o = _parse(sub, len(ops))
            cur, cur_idx = ops[0], ins[0]
            for n in range(1, len(ops)):
                later = "".join(ins[n + 1:]) + o
                keep = "".join(ch for ch in dict.fromkeys(cur_idx + ins[n]) if ch in later)
                last = n == len(ops) - 1
                tgt = o if last else keep
                cur = self._binary("%s,%s->%s" % (cur_idx, ins[n], tgt), cur, ops[n],
                                   out if last else None, alpha if last else 1.0, beta if last else 0.0)
                cur_idx = tgt
            return cur
        return self._binary(sub, ops[0], ops[1], out, alpha, beta)

    # -- one operand: pure permutation / trace-free reindexing ---------------------------------------
    def _unary(self, sub, A, out, alpha, beta):
        (ia,), io = _parse(sub, 1)
        if sorted(ia) != sorted(io) or len(set(ia)) != len(ia):
            raise B200ccError("contract: unary %r is not a pure permutation" % sub)
        v = A.permute(*[ia.index(ch) for ch in io])
        if out is None:
            out = torch.empty(tuple(v.shape), dtype=A.dtype, device=A.device)
            beta = 0.0
        self.stats["permute"] += 1
        return K.strided_axpby(out, v, alpha, beta)

    # -- two operands ----------------------------------------------------------------------------
    def _binary(self, sub, A, B, out, alpha, beta):
        (ia, ib), io = _parse(sub, 2)
        for s in (ia, ib, io):
            if len(set(s)) != len(s):
                raise B200ccError("contract: repeated index inside one operand (%r) is not supported" % sub)
        if A.dtype != torch.float64 or B.dtype != torch.float64:
            raise B200ccError("contract: float64 operands required (got %s, %s)" % (A.dtype, B.dtype))
        dims = {}
        for idx, t in ((ia, A), (ib, B)):
            if len(idx) != t.dim():
                raise B200ccError("contract: %r does not match a rank-%d operand" % (idx, t.dim()))
            for ch, n in zip(idx, t.shape):
                if dims.setdefault(ch, n) != n:
                    raise B200ccError("contract: extent mismatch on index %r in %r" % (ch, sub))
        for ch in io:
            if ch not in dims:
                raise B200ccError("contract: output index %r not in any operand" % ch)
        # indices living in one operand only and not in the output: sum them out first is not
        # needed on this path -> refuse clearly
        for ch in ia:
            if ch not in ib and ch not in io:
                raise B200ccError("contract: index %r is summed inside one operand only (%r)" % (ch, sub))
        for ch in ib:
            if ch not in ia and ch not in io:
                raise B200ccError("contract: index %r is summed inside one operand only (%r)" % (ch, sub))
        oshape = tuple(dims[ch] for ch in io)
        fresh = out is None
        if fresh:
            out = torch.empty(oshape, dtype=torch.float64, device=A.device)
            beta = 0.0
        elif tuple(out.shape) != oshape:
            raise B200ccError("contract: out has shape %s, expected %s" % (tuple(out.shape), oshape))
        if out.numel() == 0:
            return out
        key = (ia, ib, io, tuple(A.shape), tuple(A.stride()), tuple(B.shape), tuple(B.stride()),
               tuple(out.stride()))
        plan = self._plans.get(key)
        if plan is None:
            plan = self._make_plan(ia, ib, io, A, B, out)
            self._plans[key] = plan
        self._run(plan, A, B, out, alpha, beta)
        return out

    def _make_plan(self, ia, ib, io, A, B, out):
        bt = [ch for ch in io if ch in ia and ch in ib]
        Mg = [ch for ch in io if ch in ia and ch not in ib]
        Ng = [ch for ch in io if ch in ib and ch not in ia]
        Kg = [ch for ch in ia if ch in ib and ch not in io]
        sa, sb, so = tuple(A.stride()), tuple(B.stride()), tuple(out.stride())
        sha, shb, sho = tuple(A.shape), tuple(B.shape), tuple(out.shape)
        nb = _prod(sho, io, bt)
        M = _prod(sha, ia, Mg)
        N = _prod(shb, ib, Ng)
        Kd = _prod(sha, ia, Kg)

        # --- output: in place as [b][M][N] (or [b][N][M] with swapped roles), else via a temporary
        c_direct = None
        for swap in (False, True):
            rg, cg = (Ng, Mg) if swap else (Mg, Ng)
            nr, nc = (N, M) if swap else (M, N)
            mc = _Mat(io, sho, so, bt, rg, cg)
            if mc.ok and (nc == 1 or mc.cs == 1) and (nb == 1 or mc.bs >= 0):
                ldc = mc.rs if nr > 1 else max(nc, 1)
                if nr == 1 or ldc >= nc:
                    c_direct = (swap, ldc, mc.bs if nb > 1 else 0)
                    break
        # group orders: M/N follow the output when written in place, else the operand's own order
        if c_direct is None:
            Mg = [ch for ch in ia if ch in Mg]
            Ng = [ch for ch in ib if ch in Ng]
        plan = {"nb": nb, "M": M, "N": N, "K": Kd, "c_direct": c_direct, "io": io}

        # --- K order: try A's, then B's
        def operand(idx, shape, stride, rowg, kg):
            m = _Mat(idx, shape, stride, bt, rowg, kg)
            nrows = _prod(shape, idx, rowg)
            o = _orient(m, nrows, Kd)
            if o is None:
                return None
            return (o[0], o[1], m.bs if nb > 1 else 0)

        Ka = [ch for ch in ia if ch in Kg]
        Kb = [ch for ch in ib if ch in Kg]
        best = None
        for korder in (Ka, Kb):
            oa = operand(ia, sha, sa, Mg, korder)
            ob = operand(ib, shb, sb, Ng, korder)
            # cost = elements that must be copied to make this K order work
            cost = (0 if oa is not None else A.numel()) + (0 if ob is not None else B.numel())
            if best is None or cost < best[0]:
                best = (cost, korder, oa, ob)
        _, korder, oa, ob = best
        plan["A"] = oa if oa is not None else ("copy", [ia.index(ch) for ch in bt + Mg + korder])
        plan["B"] = ob if ob is not None else ("copy", [ib.index(ch) for ch in bt + Ng + korder])
        if c_direct is None:
            # temp is [bt..., Mg..., Ng...]; remember how to view it in the output's index order
            tidx = "".join(bt + Mg + Ng)
            plan["tshape"] = tuple(sho[io.index(ch)] for ch in tidx)
            plan["tperm"] = [tidx.index(ch) for ch in io]
        return plan

    def _run(self, plan, A, B, out, alpha, beta):
        nb, M, N, Kd = plan["nb"], plan["M"], plan["N"], plan["K"]

        def prep(t, spec, rows):
            if spec[0] == "copy":
                self.stats["permute"] += 1
                c = K.permuted(t, spec[1])
                # contiguous [b][rows][K]
                return c, 0, max(Kd, 1), rows * Kd
            return t, spec[0], spec[1], spec[2]

        At, ta, lda, sA = prep(A, plan["A"], M)
        Bt, tb, ldb, sB = prep(B, plan["B"], N)
        self.stats["gemm"] += 1
        cd = plan["c_direct"]
        if cd is not None:
            swap, ldc, sC = cd
            if swap:
                K.dgemm(N, M, Kd, Bt, ldb, tb, At, lda, ta, out, ldc, alpha, beta, nb, sB, sA, sC)
            else:
                K.dgemm(M, N, Kd, At, lda, ta, Bt, ldb, tb, out, ldc, alpha, beta, nb, sA, sB, sC)
            return
        tmp = torch.empty(plan["tshape"], dtype=torch.float64, device=out.device)
        K.dgemm(M, N, Kd, At, lda, ta, Bt, ldb, tb, tmp, max(N, 1), 1.0, 0.0, nb, sA, sB, M * N)
        self.stats["permute"] += 1
        K.strided_axpby(out, tmp.permute(*plan["tperm"]), alpha, beta)
