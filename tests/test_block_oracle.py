"""oracle/block_oracle.py (r1 in full and single (i,j) blocks of r2 straight from the factor of the synthetic integrals:
the checker of the FULL-SIZE residual test, tests/test_fullsize.py) against the pinned oracle at sizes where both run."""
import numpy as np
import pytest

from pycc_b200.synthetic import blocks_from_factor, make_synthetic
from oracle import ccsd_oracle as co
from oracle.block_oracle import FactorProblem


@pytest.mark.parametrize("no,nv,seed,noise", [(4, 10, 0, 0.0), (3, 7, 2, 0.01), (5, 6, 1, 0.02)])
def test_blocks_match_the_pinned_oracle(no, nv, seed, noise):
    syn = make_synthetic(no, nv, seed=seed, fock_noise=noise)
    P = co.Problem(blocks_from_factor(syn), syn.F, no)
    rng = np.random.default_rng(seed)
    t1 = 0.1 * rng.standard_normal((no, nv))
    t2 = 0.1 * rng.standard_normal((no, no, nv, nv))                # generic: no pair symmetry
    m = rng.standard_normal(syn.F.shape)
    F = syn.F + 0.02 * (m + m.T)
    r1, r2, inter, _ = P.residuals(F, t1, t2, parts=True)
    Q = FactorProblem(syn).bind(F, t1, t2)
    assert np.abs(Q.Fae - inter["Fae"]).max() < 1e-13 and np.abs(Q.Fmi - inter["Fmi"]).max() < 1e-13
    assert np.abs(Q.Fme - inter["Fme"]).max() < 1e-13
    for q in range(no):
        assert np.abs(Q.Wmbej_col(q) - inter["Wmbej"][:, :, :, q]).max() < 1e-13
        assert np.abs(Q.Wmbje_col(q) - inter["Wmbje"][:, :, q, :]).max() < 1e-13
    assert np.abs(Q.r1() - r1).max() < 1e-13
    for i in range(no):
        for j in range(no):
            assert np.abs(Q.r2_block(i, j) - r2[i, j]).max() < 1e-13, (i, j)


# ---- the product against the block oracle: small sizes on the numpy double (CPU lane), BASELINE configs[2] on the GPU ----
def _residual_blocks(no, nv, dev, pairs, tol):
    """r1 in full and the (i,j) blocks ``pairs`` of r2 of the fused residual -- general mode (``residuals``) and the
    (i >= j) pair-symmetric mode that ``solve_cc`` / bench.py iterate (``_residuals_half(symmetric=True)``) -- against
    the host evaluation from the factor, at pair-symmetric amplitudes and a non-canonical Fock matrix."""
    import torch
    import pycc_b200
    from pycc_b200 import kernels as K
    syn = make_synthetic(no, nv, seed=0, device=dev if dev.type == "cuda" else None)
    rng = np.random.default_rng(11)
    t1 = 0.05 * rng.standard_normal((no, nv))
    t2 = 0.02 * rng.standard_normal((no, no, nv, nv))
    t2 = t2 + t2.transpose(1, 0, 3, 2)                              # t2[i,j,a,b] = t2[j,i,b,a], as solve_cc's iterates
    m = rng.uniform(-0.01, 0.01, syn.F.shape)
    F = syn.F + 0.5 * (m + m.T)
    cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    Fd, a1, a2 = T(F), T(t1), T(t2)
    r1g, r2g = cc.residuals(Fd, a1, a2)
    r1g, r2g_blocks = r1g.cpu().numpy(), {p: r2g[p].cpu().numpy() for p in pairs}
    del r2g
    r1s, half = cc._residuals_half(Fd, a1, a2, symmetric=True)
    K.symmetrize_r2(half)
    r1s, r2s_blocks = r1s.cpu().numpy(), {p: half[p].cpu().numpy() for p in pairs}
    del half, cc
    Q = FactorProblem(syn).bind(F, t1, t2)
    want1 = Q.r1()
    scale = max(1.0, float(np.abs(want1).max()))
    assert np.abs(r1g - want1).max() < tol * scale and np.abs(r1s - want1).max() < tol * scale
    for p in pairs:
        want = Q.r2_block(*p)
        scale = max(1.0, float(np.abs(want).max()))
        assert np.abs(want).max() > 1e-3                                # a non-trivial block
        assert np.abs(r2g_blocks[p] - want).max() < tol * scale, ("general", p)
        assert np.abs(r2s_blocks[p] - want).max() < tol * scale, ("symmetric", p)


def test_product_blocks_small():
    import torch
    from tests import emu
    with emu.install():
        _residual_blocks(6, 14, torch.device("cpu"), [(4, 1), (2, 2), (0, 5)], 1e-12)


@pytest.mark.gpu
def test_config2_residual_blocks_match_oracle_at_full_size():
    """BASELINE configs[2], o=40 v=300: every term of the fused residual (ladder in pair form, the ring-fused o^3v^3
    products, Wmnij, Z, the t1 terms) enters r1 and the sampled r2[i,j,:,:] blocks compared here (north_star: 1e-9)."""
    import gc
    import torch
    gc.collect()
    torch.cuda.empty_cache()
    try:
        _residual_blocks(40, 300, torch.device("cuda:0"), [(7, 3), (5, 5)], 1e-10)
    finally:
        gc.collect()
        torch.cuda.empty_cache()
