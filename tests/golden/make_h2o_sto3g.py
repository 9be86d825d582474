#!/usr/bin/env python
"""Known-answer fixture: H2O / STO-3G (the molecule of the reference's hot-path tests).

    python tests/golden/make_h2o_sto3g.py        # rewrites tests/golden/h2o_sto3g.npz

The reference's own tests pin the path on this molecule with hard-coded numbers
(pycc/tests/test_002_ccsd_energy.py:22-31, test_005_ccsd_t_energy.py:21-36,
test_044_ccsd_t_gpu.py:21-38; geometry ``moldict["H2O"]`` = pycc/data/molecules.py:42-46,
frozen core, SCF converged to 1e-12 — pycc/tests/conftest.py:28-36):

    E_corr(CCSD)      = -0.070616830152761
    E(T)              = -0.000099957499645
    E_corr(CCSD(T))   = -0.0707167876524093

psi4 (which supplies the integrals to the reference, hamiltonian.py:58-68) is not installable
offline, so this script computes them itself: the STO-3G basis holds only s and p Gaussians, for
which a McMurchie-Davidson scheme in plain numpy is a page of code.  It then runs RHF and

  * stores the AO quantities (S, Hcore, (pq|rs), C, eps, E_nuc, E_SCF) as the fixture,
  * runs the UNMODIFIED reference (loader of make_golden.py) on the MO integrals and records
    its E(CCSD) / E(T), next to the three hard-coded numbers above.

The integrals are therefore NOT psi4's; what anchors them is that the reference's code, fed with
them, reproduces the reference's published numbers (asserted below to 1e-9; the achieved
difference is stored in the fixture as ``dev_*``).

Basis: STO-3G as distributed with psi4 (share/basis/sto-3g.gbs = the EMSL table), contracted
functions normalised to unit self-overlap.  Geometry: Z-matrix O / H 1 1.1 / H 1 1.1 2 104 in
Angstrom; 1 bohr = 0.52917721067 Angstrom (CODATA 2014, the value of psi4 >= 1.2 / qcelemental).
"""
import contextlib
import io
import math
import os
import sys

import numpy as np
from scipy.special import gammainc, gamma

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

BOHR = 0.52917721067

STO3G = {
    "H": [("s", [3.42525091, 0.62391373, 0.16885540], [0.15432897, 0.53532814, 0.44463454])],
    "O": [("s", [130.7093200, 23.8088610, 6.4436083], [0.15432897, 0.53532814, 0.44463454]),
          ("s", [5.0331513, 1.1695961, 0.3803890], [-0.09996723, 0.39951283, 0.70011547]),
          ("p", [5.0331513, 1.1695961, 0.3803890], [0.15591627, 0.60768372, 0.39195739])],
}
CHARGE = {"H": 1.0, "O": 8.0}


def geometry():
    r = 1.1 / BOHR
    th = math.radians(104.0)
    return [("O", np.zeros(3)),
            ("H", np.array([0.0, 0.0, r])),
            ("H", np.array([r * math.sin(th), 0.0, r * math.cos(th)]))]


def boys(n, T):
    if T < 1e-13:
        return 1.0 / (2 * n + 1) - T / (2 * n + 3)
    a = n + 0.5
    return gamma(a) * gammainc(a, T) / (2.0 * T ** a)


def hermite_E(i, j, t, Q, a, b):
    """Expansion coefficient of the Cartesian overlap distribution in Hermite Gaussians."""
    p = a + b
    if t < 0 or t > i + j:
        return 0.0
    if i == j == t == 0:
        return math.exp(-a * b / p * Q * Q)
    if j == 0:
        return (hermite_E(i - 1, j, t - 1, Q, a, b) / (2 * p)
                - a * b / p * Q / a * hermite_E(i - 1, j, t, Q, a, b)
                + (t + 1) * hermite_E(i - 1, j, t + 1, Q, a, b))
    return (hermite_E(i, j - 1, t - 1, Q, a, b) / (2 * p)
            + a * b / p * Q / b * hermite_E(i, j - 1, t, Q, a, b)
            + (t + 1) * hermite_E(i, j - 1, t + 1, Q, a, b))


def hermite_R_table(L, p, PC):
    """R[t,u,v] (order n = 0) for t+u+v <= L: Hermite Coulomb integrals."""
    T = p * float(PC @ PC)
    memo = {}

    def R(t, u, v, n):
        key = (t, u, v, n)
        if key in memo:
            return memo[key]
        if t == u == v == 0:
            val = (-2.0 * p) ** n * boys(n, T)
        elif t == u == 0:
            val = PC[2] * R(t, u, v - 1, n + 1)
            if v > 1:
                val += (v - 1) * R(t, u, v - 2, n + 1)
        elif t == 0:
            val = PC[1] * R(t, u - 1, v, n + 1)
            if u > 1:
                val += (u - 1) * R(t, u - 2, v, n + 1)
        else:
            val = PC[0] * R(t - 1, u, v, n + 1)
            if t > 1:
                val += (t - 1) * R(t - 2, u, v, n + 1)
        memo[key] = val
        return val

    tab = np.zeros((L + 1, L + 1, L + 1))
    for t in range(L + 1):
        for u in range(L + 1 - t):
            for v in range(L + 1 - t - u):
                tab[t, u, v] = R(t, u, v, 0)
    return tab


def dfact(n):
    return 1.0 if n <= 0 else n * dfact(n - 2)


class Fn:
    """One contracted Cartesian Gaussian."""

    def __init__(self, center, lmn, exps, coefs):
        self.A = center
        self.lmn = lmn
        self.exps = np.array(exps)
        l, m, n = lmn
        L = l + m + n
        norm = np.array([(2 * a / math.pi) ** 0.75 * (4 * a) ** (L / 2.0)
                         / math.sqrt(dfact(2 * l - 1) * dfact(2 * m - 1) * dfact(2 * n - 1)) for a in exps])
        self.c = np.array(coefs) * norm
        s = sum(ci * cj * prim_overlap(a, lmn, center, b, lmn, center)
                for a, ci in zip(self.exps, self.c) for b, cj in zip(self.exps, self.c))
        self.c = self.c / math.sqrt(s)


def prim_overlap(a, lmn1, A, b, lmn2, B):
    p = a + b
    val = (math.pi / p) ** 1.5
    for x in range(3):
        val *= hermite_E(lmn1[x], lmn2[x], 0, A[x] - B[x], a, b)
    return val


def prim_kinetic(a, lmn1, A, b, lmn2, B):
    l2, m2, n2 = lmn2

    def S(d):
        lmn = (l2 + d[0], m2 + d[1], n2 + d[2])
        if min(lmn) < 0:
            return 0.0
        return prim_overlap(a, lmn1, A, b, lmn, B)

    t0 = b * (2 * (l2 + m2 + n2) + 3) * S((0, 0, 0))
    t1 = -2 * b * b * (S((2, 0, 0)) + S((0, 2, 0)) + S((0, 0, 2)))
    t2 = -0.5 * (l2 * (l2 - 1) * S((-2, 0, 0)) + m2 * (m2 - 1) * S((0, -2, 0)) + n2 * (n2 - 1) * S((0, 0, -2)))
    return t0 + t1 + t2


def pair_E(a, lmn1, A, b, lmn2, B):
    """E[t,u,v] of a primitive pair and its (exponent, centre)."""
    p = a + b
    P = (a * A + b * B) / p
    Ls = [lmn1[x] + lmn2[x] for x in range(3)]
    Ex = [np.array([hermite_E(lmn1[x], lmn2[x], t, A[x] - B[x], a, b) for t in range(Ls[x] + 1)]) for x in range(3)]
    return p, P, np.einsum("t,u,v->tuv", *Ex)


def prim_nuclear(a, lmn1, A, b, lmn2, B, C):
    p, P, E = pair_E(a, lmn1, A, b, lmn2, B)
    L = sum(lmn1) + sum(lmn2)
    R = hermite_R_table(L, p, P - C)
    nt, nu, nv = E.shape
    return 2 * math.pi / p * float(np.sum(E * R[:nt, :nu, :nv]))


def contracted(f1, f2, prim, *args):
    return sum(ci * cj * prim(a, f1.lmn, f1.A, b, f2.lmn, f2.A, *args)
               for a, ci in zip(f1.exps, f1.c) for b, cj in zip(f2.exps, f2.c))


def eri_contracted(pairs, ij, kl):
    """(ij|kl) from the per-pair lists of primitive (coef, p, P, E)."""
    val = 0.0
    for (c1, p, P, E1) in pairs[ij]:
        for (c2, q, Q, E2) in pairs[kl]:
            alpha = p * q / (p + q)
            n1, n2 = E1.shape, E2.shape
            L = sum(n1) + sum(n2) - 6
            R = hermite_R_table(L, alpha, P - Q)
            s = 0.0
            for t in range(n2[0]):
                for u in range(n2[1]):
                    for v in range(n2[2]):
                        e2 = E2[t, u, v]
                        if e2 != 0.0:
                            s += (-1) ** (t + u + v) * e2 * float(
                                np.sum(E1 * R[t:t + n1[0], u:u + n1[1], v:v + n1[2]]))
            val += c1 * c2 * s * 2 * math.pi ** 2.5 / (p * q * math.sqrt(p + q))
    return val


def integrals():
    atoms = geometry()
    basis = []
    for sym, xyz in atoms:
        for kind, exps, coefs in STO3G[sym]:
            for lmn in ([(0, 0, 0)] if kind == "s" else [(1, 0, 0), (0, 1, 0), (0, 0, 1)]):
                basis.append(Fn(xyz, lmn, exps, coefs))
    n = len(basis)
    S = np.zeros((n, n)); T = np.zeros((n, n)); V = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            S[i, j] = S[j, i] = contracted(basis[i], basis[j], prim_overlap)
            T[i, j] = T[j, i] = contracted(basis[i], basis[j], prim_kinetic)
            V[i, j] = V[j, i] = sum(-CHARGE[s] * contracted(basis[i], basis[j], prim_nuclear, xyz) for s, xyz in atoms)
    pairs = {}
    for i in range(n):
        for j in range(i + 1):
            f1, f2 = basis[i], basis[j]
            pairs[(i, j)] = [(ci * cj,) + pair_E(a, f1.lmn, f1.A, b, f2.lmn, f2.A)
                             for a, ci in zip(f1.exps, f1.c) for b, cj in zip(f2.exps, f2.c)]
    eri = np.zeros((n, n, n, n))          # chemists' (ij|kl)
    keys = sorted(pairs)
    for x, (i, j) in enumerate(keys):
        for (k, l) in keys[:x + 1]:
            val = eri_contracted(pairs, (i, j), (k, l))
            for (a, b) in ((i, j), (j, i)):
                for (c, d) in ((k, l), (l, k)):
                    eri[a, b, c, d] = eri[c, d, a, b] = val
    enuc = sum(CHARGE[atoms[a][0]] * CHARGE[atoms[b][0]] / np.linalg.norm(atoms[a][1] - atoms[b][1])
               for a in range(len(atoms)) for b in range(a))
    return S, T + V, eri, enuc


def rhf(S, H, eri, ndocc, tol=1e-13, maxiter=200):
    """Plain RHF with Pulay DIIS on FDS - SDF."""
    w, U = np.linalg.eigh(S)
    X = U @ np.diag(w ** -0.5) @ U.T
    def diag(F):
        e, Cp = np.linalg.eigh(X @ F @ X)
        return e, X @ Cp
    eps, C = diag(H)
    D = C[:, :ndocc] @ C[:, :ndocc].T
    Fs, Es = [], []
    for it in range(maxiter):
        J = np.einsum("pqrs,rs->pq", eri, D)
        K = np.einsum("prqs,rs->pq", eri, D)
        F = H + 2 * J - K
        E = float(np.sum(D * (H + F)))
        err = X @ (F @ D @ S - S @ D @ F) @ X
        Fs.append(F); Es.append(err)
        Fs, Es = Fs[-8:], Es[-8:]
        if np.max(np.abs(err)) < tol:
            break
        m = len(Fs)
        Bm = -np.ones((m + 1, m + 1)); Bm[-1, -1] = 0.0
        for a in range(m):
            for b in range(m):
                Bm[a, b] = np.sum(Es[a] * Es[b])
        rhs = np.zeros(m + 1); rhs[-1] = -1.0
        c = np.linalg.lstsq(Bm, rhs, rcond=None)[0]
        eps, C = diag(sum(ci * Fi for ci, Fi in zip(c[:m], Fs)))
        D = C[:, :ndocc] @ C[:, :ndocc].T
    else:
        raise RuntimeError("RHF did not converge")
    eps, C = diag(F)
    return E, eps, C, F


REF_ECCSD = -0.070616830152761       # pycc/tests/test_002_ccsd_energy.py:31
REF_ET = -0.000099957499645          # pycc/tests/test_005_ccsd_t_energy.py:33
REF_ECCSD_T = -0.0707167876524093    # pycc/tests/test_044_ccsd_t_gpu.py:37


def main():
    import types
    from make_golden import load_reference, reference_wfn
    S, H, eri, enuc = integrals()
    escf_el, eps, C, F_ao = rhf(S, H, eri, ndocc=5)
    escf = escf_el + enuc
    print("n = %d  E_nuc = %.12f  E_SCF = %.12f" % (S.shape[0], enuc, escf))
    print("eps =", eps)

    nfzc, no = 1, 4
    n = S.shape[0]
    nv = n - nfzc - no
    mo = np.einsum("pqrs,pi,qj,rk,sl->ijkl", eri, C, C, C, C, optimize=True)
    ERI = mo.swapaxes(1, 2).copy()                      # Dirac <pq|rs>, hamiltonian.py:67
    F = C.T @ F_ao @ C

    ccwfn_mod, cctriples, utils, device_mod = load_reference()
    syn = types.SimpleNamespace(no=no, nv=nv, n=n, F=F, eps=np.diag(F).copy(),
                                o=slice(nfzc, nfzc + no), v=slice(nfzc + no, n))
    w = reference_wfn(ccwfn_mod, device_mod, syn, ERI, model="CCSD")
    w.nfzc = nfzc
    with contextlib.redirect_stdout(io.StringIO()):
        eccsd = float(w.solve_cc(1e-12, 1e-12, 75))
        et_tjl = float(cctriples.t_tjl(w))
        et_vik = float(cctriples.t_vikings(w))
        et_inv = float(cctriples.t_vikings_inverted(w))
    print("reference code on these integrals: E(CCSD) = %.15f   (hard-coded %.15f, diff %.2e)"
          % (eccsd, REF_ECCSD, eccsd - REF_ECCSD))
    print("                                   E(T)    = %.15f   (hard-coded %.15f, diff %.2e)"
          % (et_tjl, REF_ET, et_tjl - REF_ET))
    print("t_vikings %.15f  t_vikings_inverted %.15f" % (et_vik, et_inv))
    assert abs(eccsd - REF_ECCSD) < 1e-9 and abs(et_tjl - REF_ET) < 1e-9
    assert abs(et_vik - et_tjl) < 1e-13 and abs(et_inv - et_tjl) < 1e-13

    np.savez_compressed(
        os.path.join(HERE, "h2o_sto3g.npz"),
        S=S, Hcore=H, eri_ao=eri, C=C, eps=eps, F_ao=F_ao, enuc=enuc, escf=escf,
        no=no, nv=nv, nfzc=nfzc,
        ref_eccsd_hardcoded=REF_ECCSD, ref_et_hardcoded=REF_ET, ref_eccsd_t_hardcoded=REF_ECCSD_T,
        ref_eccsd=eccsd, ref_et=et_tjl, ref_t1=w.t1, ref_t2=w.t2,
        dev_eccsd=eccsd - REF_ECCSD, dev_et=et_tjl - REF_ET)
    print("wrote h2o_sto3g.npz")


if __name__ == "__main__":
    main()
