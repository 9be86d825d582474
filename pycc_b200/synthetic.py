"""Synthetic closed-shell MO integrals of a named (o, v) shape.

psi4 is not available offline, so every size beyond the recorded goldens is fed
with synthetic integrals built from a seeded, factorised (density-fitting-like)
tensor.  The recipe is the one fixed in SURVEY.md section 8(d):

    B[P,p,q]  = N(0,1), symmetrised in (p,q), naux = 2n
    (pq|rs)   = sum_P B[P,p,q] B[P,r,s]           chemist, 8-fold symmetric, PSD
    <pq|rs>   = (pr|qs)                           Dirac, as pycc's Hamiltonian.ERI
                                                  (reference hamiltonian.py:67-68)
    eps       = linspace(-2,-0.5,o) ++ linspace(0.5,3,v);  F = diag(eps) (+ noise)
    <pq|rs>  *= 0.15 / || <ij|ab> / D_ijab ||_F

Only the *factor* B and the scalars are produced here (host numpy, seeded); the
integral blocks themselves are contracted from B slices, on the host for the
small parity cases (:func:`full_eri`, :func:`blocks_from_factor`) and on the
device with the package's own GEMM for the bench sizes
(:func:`pycc_b200.hamiltonian.BlockHamiltonian.from_factor`), so the n^4 array
is never needed.
"""
from __future__ import annotations

import numpy as np

BLOCK_NAMES = ("oooo", "ooov", "oovv", "ovov", "ovvv", "vvvv")


class Synthetic:
    """Factorised synthetic Hamiltonian: B (naux,n,n), F (n,n), eps (n), scale."""

    def __init__(self, no, nv, B, F, scale, seed, eref=0.0):
        self.no, self.nv = int(no), int(nv)
        self.n = self.no + self.nv
        self.B = B
        self.F = F
        self.eps = np.diagonal(F).copy()
        self.scale = float(scale)      # multiplies the *integrals* (B B^T), not B
        self.seed = seed
        self.eref = eref

    @property
    def o(self):
        return slice(0, self.no)

    @property
    def v(self):
        return slice(self.no, self.n)


def make_synthetic(no, nv, seed=0, fock_noise=0.0, naux=None, target=0.15, device=None):
    """Seeded synthetic problem of shape (no, nv).  ``fock_noise`` > 0 adds a
    symmetric off-diagonal Fock perturbation |f_pq| <= fock_noise (variant B of
    SURVEY 8(d)) so the non-canonical F_me terms are exercised.  With ``device``
    the MP1-norm that fixes the scale is evaluated on that device with the
    package's own kernels (bench sizes); otherwise on the host with numpy."""
    rng = np.random.default_rng(seed)
    n = no + nv
    naux = 2 * n if naux is None else naux
    B = rng.standard_normal((naux, n, n))
    B += B.transpose(0, 2, 1)
    eps = np.concatenate((np.linspace(-2.0, -0.5, no), np.linspace(0.5, 3.0, nv)))
    F = np.diag(eps)
    if fock_noise:
        N = rng.uniform(-fock_noise, fock_noise, (n, n))
        N = 0.5 * (N + N.T)
        np.fill_diagonal(N, 0.0)
        F = F + N
    if device is not None:
        return Synthetic(no, nv, B, F, _device_scale(B, eps, no, nv, target, device), seed)
    # scale from the MP1 doubles norm:  <ij|ab> = sum_P B[P,i,a] B[P,j,b]
    Bov = B[:, :no, no:]
    oovv = np.einsum('Pia,Pjb->ijab', Bov, Bov, optimize=True)
    eo, ev = eps[:no], eps[no:]
    D = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev
    scale = target / np.linalg.norm(oovv / D)
    return Synthetic(no, nv, B, F, scale, seed)


def full_eri(syn):
    """The full n^4 Dirac array <pq|rs> (small cases only)."""
    chem = np.einsum('Ppq,Prs->pqrs', syn.B, syn.B, optimize=True)
    return np.ascontiguousarray(chem.swapaxes(1, 2)) * syn.scale


def block_from_factor(syn, name):
    """One Dirac block, e.g. 'ovvv' -> <mb|ef>[m,b,e,f] = sum_P B[P,m,e] B[P,b,f]."""
    sl = {'o': syn.o, 'v': syn.v}
    p, q, r, s = (sl[c] for c in name)
    return np.einsum('Ppr,Pqs->pqrs', syn.B[:, p, r], syn.B[:, q, s], optimize=True) * syn.scale


def blocks_from_factor(syn, names=BLOCK_NAMES):
    return {k: block_from_factor(syn, k) for k in names}


def _device_scale(B, eps, no, nv, target, device):
    """target / || <ij|ab> / D_ijab ||_F with the GEMM / denominator / dot kernels of the package."""
    import torch
    from . import kernels as K
    from .contract import Contractor
    Bov = torch.from_numpy(np.ascontiguousarray(B[:, :no, no:])).to(device)
    oovv = Contractor()("Pia,Pjb->ijab", Bov, Bov)
    eo = torch.from_numpy(eps[:no].copy()).to(device)
    ev = torch.from_numpy(eps[no:].copy()).to(device)
    t = K.div_d2(oovv, eo, ev).view(-1)
    return target / float(K.multi_dot(t, [t])[0]) ** 0.5
