"""Gaussian one- and two-electron integrals (McMurchie-Davidson, any angular momentum) + RHF in numpy.

Test infrastructure for the known-answer fixtures (tests/golden/make_h2o.py): psi4, which supplies the integrals to
the reference (hamiltonian.py:58-68, wavefunction.py:304-324), is not installable offline.  Vectorised over the
primitive pairs of a contracted quartet; nothing here is on the product path.
"""
import math

import numpy as np
from scipy.special import gammainc, gamma

BOHR = 0.52917721067          # Angstrom per bohr, CODATA 2014 (psi4 >= 1.2 / qcelemental)


def boys(n, T):
    T = np.asarray(T, dtype=float)
    a = n + 0.5
    small = T < 1e-13
    Ts = np.where(small, 1.0, T)
    return np.where(small, 1.0 / (2 * n + 1) - T / (2 * n + 3), gamma(a) * gammainc(a, Ts) / (2.0 * Ts ** a))


def hermite_E(i, j, t, Q, a, b):
    """Expansion coefficient E_t^{ij} of a 1-D Cartesian overlap distribution in Hermite Gaussians (arrays a, b, Q)."""
    p = a + b
    if t < 0 or t > i + j:
        return np.zeros_like(p)
    if i == j == t == 0:
        return np.exp(-a * b / p * Q * Q)
    if j == 0:
        return (hermite_E(i - 1, j, t - 1, Q, a, b) / (2 * p) - b / p * Q * hermite_E(i - 1, j, t, Q, a, b)
                + (t + 1) * hermite_E(i - 1, j, t + 1, Q, a, b))
    return (hermite_E(i, j - 1, t - 1, Q, a, b) / (2 * p) + a / p * Q * hermite_E(i, j - 1, t, Q, a, b)
            + (t + 1) * hermite_E(i, j - 1, t + 1, Q, a, b))


def hermite_R(L, p, PC):
    """R[t,u,v,...] for t+u+v <= L (zero elsewhere): Hermite Coulomb integrals; p, PC[...,3] arrays."""
    T = p * np.sum(PC * PC, axis=-1)
    F = [boys(n, T) for n in range(L + 1)]
    memo = {}

    def R(t, u, v, n):
        key = (t, u, v, n)
        if key not in memo:
            if t == u == v == 0:
                val = (-2.0 * p) ** n * F[n]
            elif t == u == 0:
                val = PC[..., 2] * R(t, u, v - 1, n + 1) + ((v - 1) * R(t, u, v - 2, n + 1) if v > 1 else 0.0)
            elif t == 0:
                val = PC[..., 1] * R(t, u - 1, v, n + 1) + ((u - 1) * R(t, u - 2, v, n + 1) if u > 1 else 0.0)
            else:
                val = PC[..., 0] * R(t - 1, u, v, n + 1) + ((t - 1) * R(t - 2, u, v, n + 1) if t > 1 else 0.0)
            memo[key] = val
        return memo[key]

    tab = np.zeros((L + 1, L + 1, L + 1) + T.shape)
    for t in range(L + 1):
        for u in range(L + 1 - t):
            for v in range(L + 1 - t - u):
                tab[t, u, v] = R(t, u, v, 0)
    return tab


def dfact(n):
    return 1.0 if n <= 0 else n * dfact(n - 2)


class Fn:
    """One contracted Cartesian Gaussian: centre, (l,m,n), exponents, coefficients of NORMALISED primitives."""

    def __init__(self, center, lmn, exps, coefs):
        self.A = np.asarray(center, dtype=float)
        self.lmn = tuple(lmn)
        self.exps = np.asarray(exps, dtype=float)
        l, m, n = lmn
        L = l + m + n
        norm = ((2 * self.exps / math.pi) ** 0.75 * (4 * self.exps) ** (L / 2.0)
                / math.sqrt(dfact(2 * l - 1) * dfact(2 * m - 1) * dfact(2 * n - 1)))
        self.c = np.asarray(coefs, dtype=float) * norm


class Pair:
    """All primitive pairs of two contracted functions, flattened."""

    def __init__(self, f1, f2):
        self.f1, self.f2 = f1, f2
        self.a = np.repeat(f1.exps, len(f2.exps))
        self.b = np.tile(f2.exps, len(f1.exps))
        self.c = np.repeat(f1.c, len(f2.c)) * np.tile(f2.c, len(f1.c))
        self.p = self.a + self.b
        self.P = (self.a[:, None] * f1.A + self.b[:, None] * f2.A) / self.p[:, None]
        self.L = sum(f1.lmn) + sum(f2.lmn)

    def E1d(self, x, l1, l2):
        Q = self.f1.A[x] - self.f2.A[x]
        return np.stack([hermite_E(l1, l2, t, Q, self.a, self.b) for t in range(l1 + l2 + 1)], axis=1)   # [pp, t]

    def E(self, lmn2=None):
        l1, l2 = self.f1.lmn, (self.f2.lmn if lmn2 is None else lmn2)
        return np.einsum("at,au,av->atuv", *[self.E1d(x, l1[x], l2[x]) for x in range(3)])

    def overlap(self, lmn2=None):
        l2 = self.f2.lmn if lmn2 is None else lmn2
        if min(l2) < 0:
            return np.zeros_like(self.p)
        val = (math.pi / self.p) ** 1.5
        for x in range(3):
            val = val * hermite_E(self.f1.lmn[x], l2[x], 0, self.f1.A[x] - self.f2.A[x], self.a, self.b)
        return val

    def S(self):
        return float(np.sum(self.c * self.overlap()))

    def T(self):
        l2, m2, n2 = self.f2.lmn
        b = self.b
        sh = lambda dx, dy, dz: self.overlap((l2 + dx, m2 + dy, n2 + dz))
        val = (b * (2 * (l2 + m2 + n2) + 3) * sh(0, 0, 0) - 2 * b * b * (sh(2, 0, 0) + sh(0, 2, 0) + sh(0, 0, 2))
               - 0.5 * (l2 * (l2 - 1) * sh(-2, 0, 0) + m2 * (m2 - 1) * sh(0, -2, 0) + n2 * (n2 - 1) * sh(0, 0, -2)))
        return float(np.sum(self.c * val))

    def V(self, C):
        """<f1| 1/|r-C| |f2>"""
        E = self.E()
        R = hermite_R(self.L, self.p, self.P - C)                  # [t,u,v,pp]
        nt, nu, nv = E.shape[1:]
        return float(np.sum(self.c * 2 * math.pi / self.p * np.einsum("atuv,tuva->a", E, R[:nt, :nu, :nv])))


def eri_pair(P1, E1, P2, E2):
    """(f1 f2|f3 f4) for two Pair objects and their E tensors."""
    p, q = P1.p[:, None], P2.p[None, :]
    alpha = p * q / (p + q)
    R = hermite_R(P1.L + P2.L, alpha, P1.P[:, None, :] - P2.P[None, :, :])       # [T,U,V,a,b]
    pref = 2 * math.pi ** 2.5 / (p * q * np.sqrt(p + q)) * P1.c[:, None] * P2.c[None, :]
    n1, n2 = E1.shape[1:], E2.shape[1:]
    val = 0.0
    for t in range(n2[0]):
        for u in range(n2[1]):
            for v in range(n2[2]):
                e2 = E2[:, t, u, v]
                if np.any(e2):
                    val += (-1) ** (t + u + v) * np.einsum(
                        "atuv,tuvab,ab,b->", E1, R[t:t + n1[0], u:u + n1[1], v:v + n1[2]], pref, e2, optimize=True)
    return float(val)


CART = {0: [(0, 0, 0)], 1: [(1, 0, 0), (0, 1, 0), (0, 0, 1)],
        2: [(2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)]}
# columns: pure combinations of the Cartesian components above (any basis of the pure subspace gives the same energies)
PURE = {0: np.eye(1), 1: np.eye(3),
        2: np.array([[0, 1, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0], [0, 0, 0, 0, 1, 0],
                     [1, 0, 0, -1, 0, 0], [-1, 0, 0, -1, 0, 2]], dtype=float).T}


def build_basis(atoms, shells):
    """atoms: [(symbol, xyz)], shells: {symbol: [(l, exps, coefs)]} -> (Cartesian functions, Cartesian->pure matrix).
    Within a d shell the primitive norm of x^2 is used for every component (the pure combinations need a common one)."""
    fns, blocks = [], []
    for sym, xyz in atoms:
        for l, exps, coefs in shells[sym]:
            first = Fn(xyz, CART[l][0], exps, coefs)
            for lmn in CART[l]:
                f = Fn(xyz, lmn, exps, coefs)
                f.c = first.c.copy()
                fns.append(f)
            blocks.append(PURE[l])
    nc = sum(b.shape[0] for b in blocks)
    npure = sum(b.shape[1] for b in blocks)
    U = np.zeros((nc, npure))
    r = c = 0
    for b in blocks:
        U[r:r + b.shape[0], c:c + b.shape[1]] = b
        r += b.shape[0]
        c += b.shape[1]
    return fns, U


def integrals(atoms, shells, charge):
    """(S, Hcore, chemists' (pq|rs), E_nuc) over the pure functions, each normalised to unit self-overlap."""
    fns, U = build_basis(atoms, shells)
    n = len(fns)
    S = np.zeros((n, n)); H = np.zeros((n, n))
    pairs, Es = {}, {}
    for i in range(n):
        for j in range(i + 1):
            P = pairs[(i, j)] = Pair(fns[i], fns[j])
            Es[(i, j)] = P.E()
            S[i, j] = S[j, i] = P.S()
            H[i, j] = H[j, i] = P.T() - sum(charge[s] * P.V(xyz) for s, xyz in atoms)
    eri = np.zeros((n, n, n, n))
    keys = sorted(pairs)
    for x, (i, j) in enumerate(keys):
        for (k, l) in keys[:x + 1]:
            val = eri_pair(pairs[(i, j)], Es[(i, j)], pairs[(k, l)], Es[(k, l)])
            for (a, b) in ((i, j), (j, i)):
                for (c, d) in ((k, l), (l, k)):
                    eri[a, b, c, d] = eri[c, d, a, b] = val
    S = U.T @ S @ U
    H = U.T @ H @ U
    eri = np.einsum("pqrs,pi,qj,rk,sl->ijkl", eri, U, U, U, U, optimize=True)
    d = 1.0 / np.sqrt(np.diag(S))
    S = S * d[:, None] * d[None, :]
    H = H * d[:, None] * d[None, :]
    eri = eri * d[:, None, None, None] * d[None, :, None, None] * d[None, None, :, None] * d[None, None, None, :]
    enuc = sum(charge[atoms[a][0]] * charge[atoms[b][0]] / np.linalg.norm(atoms[a][1] - atoms[b][1])
               for a in range(len(atoms)) for b in range(a))
    return S, H, eri, enuc


def rhf(S, H, eri, ndocc, tol=1e-13, maxiter=200):
    """Plain RHF with Pulay DIIS on the orthogonalised commutator FDS - SDF.  Returns (E_el, eps, C, F_ao)."""
    w, U = np.linalg.eigh(S)
    X = U @ np.diag(w ** -0.5) @ U.T

    def diag(F):
        e, Cp = np.linalg.eigh(X @ F @ X)
        return e, X @ Cp

    eps, C = diag(H)
    D = C[:, :ndocc] @ C[:, :ndocc].T
    Fs, Es = [], []
    for it in range(maxiter):
        F = H + 2 * np.einsum("pqrs,rs->pq", eri, D) - np.einsum("prqs,rs->pq", eri, D)
        E = float(np.sum(D * (H + F)))
        err = X @ (F @ D @ S - S @ D @ F) @ X
        if np.max(np.abs(err)) < tol:
            break
        Fs, Es = (Fs + [F])[-8:], (Es + [err])[-8:]
        m = len(Fs)
        Bm = -np.ones((m + 1, m + 1)); Bm[-1, -1] = 0.0
        Bm[:m, :m] = [[np.sum(ea * eb) for eb in Es] for ea in Es]
        rhs = np.zeros(m + 1); rhs[-1] = -1.0
        c = np.linalg.lstsq(Bm, rhs, rcond=None)[0]
        eps, C = diag(sum(ci * Fi for ci, Fi in zip(c[:m], Fs)))
        D = C[:, :ndocc] @ C[:, :ndocc].T
    else:
        raise RuntimeError("RHF did not converge")
    eps, C = diag(F)
    return E, eps, C, F


def pack_eri(eri):
    """Unique elements of an 8-fold symmetric (pq|rs): [pair(p>=q) >= pair(r>=s)]."""
    n = eri.shape[0]
    iu = np.tril_indices(n)
    M = eri[iu[0], iu[1]][:, iu[0], iu[1]]
    ju = np.tril_indices(M.shape[0])
    return M[ju]


def unpack_eri(packed, n):
    iu = np.tril_indices(n)
    npair = len(iu[0])
    ju = np.tril_indices(npair)
    M = np.zeros((npair, npair))
    M[ju] = packed
    M = M + M.T - np.diag(np.diag(M))
    idx = np.zeros((n, n), dtype=int)
    idx[iu] = np.arange(npair)
    idx = np.maximum(idx, idx.T)
    return M[idx[:, :, None, None], idx[None, None, :, :]]
