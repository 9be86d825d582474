"""Complex tensors as pairs of real planes, and the contraction of such pairs on the REAL kernels.

The library has no complex kernel.  Where a complex quantity has to be carried through the term tables of HBAR and of the
Lambda residual (the Lambda half of the RT-CC right-hand side, rt/rtcc.py:143-147; cclambda.py:202-256 with complex
t, lambda and a Hermitian F), it is a :class:`Planes` -- two float64 device tensors -- and every table term
``alpha * contract(sub, A, B)`` accumulated into an output pair is evaluated by :func:`cterm`:

* real x real              one GEMM into the real plane
* complex x real integral  two GEMMs (one per plane) -- the integrals are real, so every "amplitude x integral block" term,
                           which is where the o^2v^4 / o^3v^3 flops of HBAR are, costs 2 real products (5 when the whole
                           residual is sampled at five real points, utils.complex_from_real_samples)
* complex x complex        four GEMMs accumulated in place, or three with temporaries (``use3m``: the o^3v^3 products of
                           lambda2 with H_ovvo / H_ovov)
"""
from __future__ import annotations

import torch

from . import kernels as K

F64 = torch.float64
# complex x complex terms with at least this many output elements take the three-product form (tests lower it)
THREE_M_MIN_OUT = 1 << 22


class Planes(object):
    """(re, im) float64 tensors of equal shape; ``im`` is None for a real quantity."""
    __slots__ = ("re", "im")

    def __init__(self, re, im=None):
        self.re, self.im = re, im

    @property
    def shape(self):
        return self.re.shape

    def view(self, fn):
        """The same strided view of both planes, e.g. ``F.view(lambda x: x[o, v])``."""
        return Planes(fn(self.re), None if self.im is None else fn(self.im))

    def full(self):
        """Both planes present (zeros for a missing imaginary plane)."""
        if self.im is None:
            self.im = torch.zeros_like(self.re, memory_format=torch.contiguous_format)
        return self


def parts(X):
    return (X.re, X.im) if isinstance(X, Planes) else (X, None)


def _sum(a, b):
    s = K.permuted(a, tuple(range(a.dim())))
    return K.strided_axpby(s, b, 1.0, 1.0)


def cterm(ct, alpha, sub, A, B, out, use3m=None):
    """out += alpha * contract(sub, A, B) for real tensors / Planes ``A``, ``B`` and a full Planes ``out``."""
    (ar, ai), (br, bi) = parts(A), parts(B)
    if ai is None and bi is None:
        ct(sub, ar, br, out=out.re, alpha=alpha, beta=1.0)
    elif bi is None:
        ct(sub, ar, br, out=out.re, alpha=alpha, beta=1.0)
        ct(sub, ai, br, out=out.im, alpha=alpha, beta=1.0)
    elif ai is None:
        ct(sub, ar, br, out=out.re, alpha=alpha, beta=1.0)
        ct(sub, ar, bi, out=out.im, alpha=alpha, beta=1.0)
    else:
        if use3m is None:
            use3m = out.re.numel() >= THREE_M_MIN_OUT
        if use3m:
            # Re = P1 - P2, Im = (Ar + Ai)(Br + Bi) - P1 - P2
            p1 = ct(sub, ar, br)
            p2 = ct(sub, ai, bi)
            p3 = ct(sub, _sum(ar, ai), _sum(br, bi))
            K.strided_axpby(out.re, p1, alpha, 1.0)
            K.strided_axpby(out.re, p2, -alpha, 1.0)
            K.strided_axpby(out.im, p3, alpha, 1.0)
            K.strided_axpby(out.im, p1, -alpha, 1.0)
            K.strided_axpby(out.im, p2, -alpha, 1.0)
        else:
            ct(sub, ar, br, out=out.re, alpha=alpha, beta=1.0)
            ct(sub, ai, bi, out=out.re, alpha=-alpha, beta=1.0)
            ct(sub, ar, bi, out=out.im, alpha=alpha, beta=1.0)
            ct(sub, ai, br, out=out.im, alpha=alpha, beta=1.0)
    return out


def cprod(ct, alpha, sub, A, B):
    """alpha * contract(sub, A, B) as a new Planes (real result: im None)."""
    (ar, ai), (br, bi) = parts(A), parts(B)
    re = ct(sub, ar, br, alpha=alpha)
    if ai is None and bi is None:
        return Planes(re, None)
    out = Planes(re, torch.zeros_like(re))
    if ai is not None and bi is not None:
        ct(sub, ai, bi, out=out.re, alpha=-alpha, beta=1.0)
        ct(sub, ar, bi, out=out.im, alpha=alpha, beta=1.0)
        ct(sub, ai, br, out=out.im, alpha=alpha, beta=1.0)
    elif bi is None:
        ct(sub, ai, br, out=out.im, alpha=alpha, beta=1.0)
    else:
        ct(sub, ar, bi, out=out.im, alpha=alpha, beta=1.0)
    return out


def copy_real(x, alpha=1.0):
    """Planes(alpha * x, 0) from a real tensor / view."""
    re = K.permuted(x, tuple(range(x.dim())), alpha)
    return Planes(re, torch.zeros_like(re))


def copy_planes(P, alpha=1.0):
    """contiguous, full copy of alpha * P"""
    ident = tuple(range(P.re.dim()))
    re = K.permuted(P.re, ident, alpha)
    return Planes(re, torch.zeros_like(re) if P.im is None else K.permuted(P.im, ident, alpha))


def complex_tau(t1, t2):
    """tau = t2 + t1 t1 for complex amplitudes given as Planes (ccwfn.py:455):
    Re = x2 + x1 x1 - y1 y1,  Im = y2 + x1 y1 + y1 x1 = y2 + (x1+y1)(x1+y1) - x1 x1 - y1 y1."""
    (x1, y1), (x2, y2) = parts(t1), parts(t2)
    tre = K.build_tau(x1, x2, 1.0, 1.0)
    if y1 is None and y2 is None:
        return Planes(tre, None)
    z2 = torch.zeros_like(x2) if y2 is None else y2
    if y1 is None:
        return Planes(tre, K.permuted(z2, (0, 1, 2, 3)))
    yy = K.build_tau(y1, x2, 0.0, 1.0)                                   # y1 y1
    xx = K.build_tau(x1, x2, 0.0, 1.0)                                   # x1 x1
    s1 = K.axpbyz(1.0, x1, 1.0, y1, torch.empty_like(x1))
    tim = K.build_tau(s1, z2, 1.0, 1.0)                                  # y2 + (x1+y1)(x1+y1)
    K.axpbyz(1.0, tim.view(-1), -1.0, xx.view(-1), tim.view(-1))
    K.axpbyz(1.0, tim.view(-1), -1.0, yy.view(-1), tim.view(-1))
    K.axpbyz(1.0, tre.view(-1), -1.0, yy.view(-1), tre.view(-1))
    return Planes(tre, tim)
