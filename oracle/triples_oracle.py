"""CPU ORACLE (test infrastructure, NOT a product path) for the (T) correction.

Plain-numpy restatement of CrawfordGroup/pycc ``cctriples.py`` for closed-shell
RHF, on the stored Dirac blocks (``ovvv``, ``ooov``, ``oovv``):

  t3c_ijk   cctriples.py:27-72      connected numerator (6 particle + 6 hole terms)
  t3d_ijk   cctriples.py:108-147    disconnected numerator
  t_tjl     cctriples.py:177-239    Lee-Rendell energy; the reference's two
                                    interpreted a,b,c loops (210-213, 231-237) are
                                    restated with masks/weights over the full
                                    (a,b,c) cube -- same sum, vectorised
  t_vikings cctriples.py:243-307    full-loop cross-check (same E(T))
  t3c_abc   cctriples.py:75-105     connected numerator for fixed (a,b,c): an o^3 tile
  t3d_abc   cctriples.py:149-173    disconnected numerator for fixed (a,b,c)
  abc_energy                        the Lee-Rendell bracket of cctriples.py:208-237 applied to one
                                    (a,b,c) tile with the roles of the occupied and virtual indices
                                    exchanged (the form the fused CUDA kernel evaluates); the sum over
                                    a >= b >= c equals t_tjl -- checked against the reference's e_t_tjl

Block identities (8-fold symmetry):  ERI[v,v,v,o][x,y,e,i] = ovvv[i,e,y,x],
ERI[o,v,o,o][m,c,j,k] = ooov[j,k,m,c],  ERI[v,o,v,v][d,k,b,c] = ovvv[k,d,c,b].

PARITY PINNED: ``tests/test_oracle_golden.py`` checks W3, V3, t3c/t3d with
denominators and E(T) against the reference's own outputs
(``tests/golden/ref_*.npz``), including ``t_tjl == t_vikings ==
t_vikings_inverted`` as in the reference's tests/test_005_ccsd_t_energy.py:30-36.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import numpy as np


def es(sub, *ops):
    return np.einsum(sub, *ops, optimize=True)


def _denom(F, no, i, j, k, nfzc=0):
    eps = np.diagonal(F)
    eo, ev = eps[nfzc:nfzc + no], eps[nfzc + no:]
    return (eo[i] + eo[j] + eo[k]) - (ev[:, None, None] + ev[None, :, None] + ev[None, None, :])


def t3c_ijk(i, j, k, t2, ovvv, ooov, F=None, with_denom=False, nfzc=0):
    """Connected t3 numerator for one (i,j,k)   (cctriples.py:50-62).

    Wvvvo[:,:,:,i][x,y,e] = ovvv[i,e,y,x];  Wovoo[:,:,j,k][m,c] = ooov[j,k,m,c].
    """
    def Wv(p):                       # [x,y,e]
        return ovvv[p].transpose(2, 1, 0)
    W = es("bae,ce->abc", Wv(i), t2[k, j])
    W += es("cae,be->abc", Wv(i), t2[j, k])
    W += es("ace,be->abc", Wv(k), t2[j, i])
    W += es("bce,ae->abc", Wv(k), t2[i, j])
    W += es("cbe,ae->abc", Wv(j), t2[i, k])
    W += es("abe,ce->abc", Wv(j), t2[k, i])
    W -= es("mc,mab->abc", ooov[j, k], t2[i])
    W -= es("mb,mac->abc", ooov[k, j], t2[i])
    W -= es("mb,mca->abc", ooov[i, j], t2[k])
    W -= es("ma,mcb->abc", ooov[j, i], t2[k])
    W -= es("ma,mbc->abc", ooov[k, i], t2[j])
    W -= es("mc,mba->abc", ooov[i, k], t2[j])
    if with_denom:
        return W / _denom(F, t2.shape[0], i, j, k, nfzc)
    return W


def t3d_ijk(i, j, k, t1, t2, oovv, F, with_denom=False, nfzc=0):
    """Disconnected t3 numerator for one (i,j,k)   (cctriples.py:131-137)."""
    no = t2.shape[0]
    Fov = F[nfzc:nfzc + no, nfzc + no:]
    V = es("ab,c->abc", oovv[i, j], t1[k])
    V += es("ac,b->abc", oovv[i, k], t1[j])
    V += es("bc,a->abc", oovv[j, k], t1[i])
    V += es("ab,c->abc", t2[i, j], Fov[k])
    V += es("ac,b->abc", t2[i, k], Fov[j])
    V += es("bc,a->abc", t2[j, k], Fov[i])
    if with_denom:
        return V / _denom(F, no, i, j, k, nfzc)
    return V


def triple_energy(W, V, D, occ_weight):
    """Lee-Rendell contribution of one (i>=j>=k) batch   (cctriples.py:210-237).

    ``V`` must already hold W + disconnected, NOT yet divided by (1+delta_ab+...).
    The a>=b>=c loop becomes a 0/1 mask over the cube; the 1/(1+d_ab+d_ac+d_bc)
    scaling is applied to V first, exactly as the reference does.
    """
    nv = W.shape[0]
    a = np.arange(nv)
    eq = ((a[:, None, None] == a[None, :, None]).astype(float)
          + (a[:, None, None] == a[None, None, :]).astype(float)
          + (a[None, :, None] == a[None, None, :]).astype(float))
    V = V / (1.0 + eq)
    p = lambda X, *ax: X.transpose(*ax)
    # X3[a,b,c] = sum over the six permutations P of W[P(abc)] V[P(abc)]   (215-220)
    X = (W * V + p(W, 0, 2, 1) * p(V, 0, 2, 1) + p(W, 1, 0, 2) * p(V, 1, 0, 2)
         + p(W, 1, 2, 0) * p(V, 1, 2, 0) + p(W, 2, 0, 1) * p(V, 2, 0, 1) + p(W, 2, 1, 0) * p(V, 2, 1, 0))
    # Y = V[abc]+V[bca]+V[cab]; Z = V[acb]+V[bac]+V[cba]                   (222-223)
    Y = V + p(V, 1, 2, 0) + p(V, 2, 0, 1)
    Z = p(V, 0, 2, 1) + p(V, 1, 0, 2) + p(V, 2, 1, 0)
    Wc = W + p(W, 1, 2, 0) + p(W, 2, 0, 1)
    Wo = p(W, 0, 2, 1) + p(W, 1, 0, 2) + p(W, 2, 1, 0)
    mask = (a[:, None, None] >= a[None, :, None]) & (a[None, :, None] >= a[None, None, :])
    e = ((Y - 2.0 * Z) * Wc + (Z - 2.0 * Y) * Wo + 3.0 * X) / D
    return occ_weight * np.sum(e[mask])


def t_tjl(t1, t2, F, ovvv, ooov, oovv, nfzc=0, triples=None):
    """E(T), Lee-Rendell form (cctriples.py:177-239).  ``triples`` restricts the
    sum to a list of (i,j,k) (used for sampled CPU baselines / sharding checks)."""
    no = t2.shape[0]
    if triples is None:
        triples = [(i, j, k) for i in range(no) for j in range(i + 1) for k in range(j + 1)]
    et = 0.0
    for (i, j, k) in triples:
        W = t3c_ijk(i, j, k, t2, ovvv, ooov)
        V = W + t3d_ijk(i, j, k, t1, t2, oovv, F, nfzc=nfzc)
        D = _denom(F, no, i, j, k, nfzc)
        w = 2.0 - (float(i == j) + float(i == k) + float(j == k))
        et += triple_energy(W, V, D, w)
    return et


def t3c_abc(a, b, c, t2, ovvv, ooov, F=None, with_denom=False, nfzc=0):
    """Connected t3 numerator for one (a,b,c) as an (o,o,o) tile   (cctriples.py:83-95), term by term in the
    reference's order.  Wvvvo[x,y][e,i] = ovvv[i,e,y,x];  Wovoo[:,c,:,:][m,j,k] = ooov[j,k,m,c].
    ``ovvv`` may be a dict {i: slab} or an array; only slices [:, :, y, x] of it are taken."""
    no = t2.shape[0]

    def Wv(x, y):                    # [e,i] = <xy|ei> = ovvv[i,e,y,x]
        return np.stack([np.asarray(ovvv[i])[:, y, x] for i in range(no)], axis=1)

    def Wo(x):                       # [m,j,k] = <mx|jk> = ooov[j,k,m,x]
        return ooov[:, :, :, x].transpose(2, 0, 1)
    W = es("ei,kje->ijk", Wv(b, a), t2[:, :, c])
    W += es("ei,jke->ijk", Wv(c, a), t2[:, :, b])
    W += es("ek,jie->ijk", Wv(a, c), t2[:, :, b])
    W += es("ek,ije->ijk", Wv(b, c), t2[:, :, a])
    W += es("ej,ike->ijk", Wv(c, b), t2[:, :, a])
    W += es("ej,kie->ijk", Wv(a, b), t2[:, :, c])
    W -= es("mjk,im->ijk", Wo(c), t2[:, :, a, b])
    W -= es("mkj,im->ijk", Wo(b), t2[:, :, a, c])
    W -= es("mij,km->ijk", Wo(b), t2[:, :, c, a])
    W -= es("mji,km->ijk", Wo(a), t2[:, :, c, b])
    W -= es("mki,jm->ijk", Wo(a), t2[:, :, b, c])
    W -= es("mik,jm->ijk", Wo(c), t2[:, :, b, a])
    if with_denom:
        return W / _denom_abc(F, no, a, b, c, nfzc)
    return W


def _denom_abc(F, no, a, b, c, nfzc=0):
    eps = np.diagonal(F)
    eo, ev = eps[nfzc:nfzc + no], eps[nfzc + no:]
    return (eo[:, None, None] + eo[None, :, None] + eo[None, None, :]) - (ev[a] + ev[b] + ev[c])


def t3d_abc(a, b, c, t1, t2, oovv, F, with_denom=False, nfzc=0):
    """Disconnected t3 numerator for one (a,b,c)   (cctriples.py:155-163)."""
    no = t2.shape[0]
    Fov = F[nfzc:nfzc + no, nfzc + no:]
    V = es("ij,k->ijk", oovv[:, :, a, b], t1[:, c])
    V += es("ik,j->ijk", oovv[:, :, a, c], t1[:, b])
    V += es("jk,i->ijk", oovv[:, :, b, c], t1[:, a])
    V += es("ij,k->ijk", t2[:, :, a, b], Fov[:, c])
    V += es("ik,j->ijk", t2[:, :, a, c], Fov[:, b])
    V += es("jk,i->ijk", t2[:, :, b, c], Fov[:, a])
    if with_denom:
        return V / _denom_abc(F, no, a, b, c, nfzc)
    return V


def abc_energy(a, b, c, t1, t2, F, ovvv, ooov, oovv, nfzc=0):
    """Contribution of the virtual triple a >= b >= c to E(T): ``triple_energy`` (cctriples.py:210-237) on the (o,o,o)
    tile, i.e. with the roles of (i,j,k) and (a,b,c) exchanged -- 1/(1+delta) over the occupied indices, the i >= j >= k
    mask, the weight 2 - d_ab - d_ac - d_bc.  Summed over all a >= b >= c this is t_tjl (W and V are symmetric under
    simultaneous permutations of the occupied and the virtual triple)."""
    no = t2.shape[0]
    W = t3c_abc(a, b, c, t2, ovvv, ooov)
    V = W + t3d_abc(a, b, c, t1, t2, oovv, F, nfzc=nfzc)
    w = 2.0 - (float(a == b) + float(a == c) + float(b == c))
    return triple_energy(W, V, _denom_abc(F, no, a, b, c, nfzc), w)


def t_tjl_abc(t1, t2, F, ovvv, ooov, oovv, nfzc=0):
    """E(T) as the sum of abc_energy over a >= b >= c."""
    nv = t2.shape[2]
    return sum(abc_energy(a, b, c, t1, t2, F, ovvv, ooov, oovv, nfzc)
               for a in range(nv) for b in range(a + 1) for c in range(b + 1))


def t_vikings(t1, t2, F, ovvv, ooov, oovv, nfzc=0):
    """E(T), Helgaker-Jorgensen-Olsen full-loop form (cctriples.py:243-307)."""
    no, nv = t1.shape
    Fov = F[nfzc:nfzc + no, nfzc + no:]
    Loovv = 2.0 * oovv - oovv.transpose(0, 1, 3, 2)
    X1 = np.zeros_like(t1)
    X2 = np.zeros_like(t2)
    for i in range(no):
        for j in range(no):
            for k in range(no):
                t3 = t3c_ijk(i, j, k, t2, ovvv, ooov, F, True, nfzc)
                u = t3 - t3.transpose(2, 1, 0)
                w = 2.0 * t3 - t3.transpose(0, 2, 1) - t3.transpose(2, 1, 0)
                X1[i] += es("abc,bc->a", u, Loovv[j, k])
                # ERI[v,o,v,v][d,k,b,c] = ovvv[k,d,c,b]
                X2[i, j] += es("abc,dcb->ad", w, ovvv[k])
                X2[i] -= es("abc,lc->lab", w, ooov[j, k])
                X2[i, j] += es("abc,c->ab", u, Fov[k])
    return 2.0 * np.sum(t1 * X1) + np.sum((4.0 * t2 - 2.0 * t2.transpose(0, 1, 3, 2)) * X2)
