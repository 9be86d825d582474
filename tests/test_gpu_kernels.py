"""Unit tests of the CUDA kernels through the C ABI (libb200cc.so) against plain torch float64 ops on
the same device.  All `-m gpu`.  FP64 tolerance: 1e-11 relative to the magnitude of the result
(different summation order only)."""
import itertools

import numpy as np
import pytest
import torch

from pycc_b200 import kernels as K

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def rnd(*shape, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return torch.randn(*shape, dtype=torch.float64, generator=g).to(DEV)


def relerr(got, ref):
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-300))


SHAPES = [(1, 1, 1), (8, 8, 4), (37, 19, 53), (129, 130, 17), (128, 128, 64), (300, 257, 1000),
          (400, 150, 333), (1, 300, 77), (257, 1, 40), (64, 513, 16), (130, 131, 1)]


@pytest.mark.parametrize("M,N,Kd", SHAPES)
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("misalign", [0, 1])
def test_dgemm_orientations(M, N, Kd, ta, tb, misalign):
    # operands live at an odd element offset when misalign=1 -> exercises the 8-byte cp.async path
    Abuf = rnd(M * Kd + 3, seed=1)
    Bbuf = rnd(N * Kd + 3, seed=2)
    A = Abuf[misalign:misalign + M * Kd].view((Kd, M) if ta else (M, Kd))
    B = Bbuf[misalign:misalign + N * Kd].view((Kd, N) if tb else (N, Kd))
    C0 = rnd(M, N, seed=3)
    C = C0.clone()
    K.dgemm(M, N, Kd, A, A.stride(0), ta, B, B.stride(0), tb, C, N, alpha=0.75, beta=-0.5, ksplit=1)
    Am = A.t() if ta else A
    Bm = B.t() if tb else B
    ref = 0.75 * (Am @ Bm.t()) - 0.5 * C0
    assert relerr(C, ref) < 1e-12


@pytest.mark.parametrize("ksplit", [2, 5, 64])
def test_dgemm_splitk_and_ldc(ksplit):
    M, N, Kd = 45, 70, 5000
    A, B = rnd(M, Kd), rnd(N, Kd)
    Cbig = rnd(M, N + 9)
    ref = 2.0 * (A @ B.t()) + 1.5 * Cbig[:, :N]
    keep = Cbig[:, N:].clone()
    K.dgemm(M, N, Kd, A, Kd, 0, B, Kd, 0, Cbig, N + 9, alpha=2.0, beta=1.5, ksplit=ksplit)
    assert relerr(Cbig[:, :N], ref) < 1e-12
    assert torch.equal(Cbig[:, N:], keep)          # nothing written past column N


def test_dgemm_auto_splitk_long_k():
    M, N, Kd = 40, 300, 90000
    A, B = rnd(M, Kd), rnd(N, Kd)
    C = torch.empty(M, N, dtype=torch.float64, device=DEV)
    K.dgemm(M, N, Kd, A, Kd, 0, B, Kd, 0, C, N)
    assert relerr(C, A @ B.t()) < 1e-12


def test_dgemm_strided_batch_and_broadcast_operand():
    nb, M, N, Kd = 7, 33, 65, 47
    A, B = rnd(M, Kd), rnd(nb, Kd, N)             # A shared (stride 0), B N-major per batch
    C = torch.zeros(nb, M, N, dtype=torch.float64, device=DEV)
    K.dgemm(M, N, Kd, A, Kd, 0, B, N, 1, C, N, batch=nb, sA=0, sB=Kd * N, sC=M * N)
    ref = torch.einsum("mk,bkn->bmn", A, B)
    assert relerr(C, ref) < 1e-12


def test_dgemm_two_segments_and_table():
    nb, M, N, K1, K2 = 5, 150, 41, 61, 13
    A1, B1 = rnd(nb, K1, M), rnd(nb, N, K1)        # A M-major, B K-major (the (T) orientation)
    A2, B2 = rnd(nb, K2, M, seed=5), rnd(nb, N, K2, seed=6)
    C = torch.zeros(nb, M, N, dtype=torch.float64, device=DEV)
    ref = torch.einsum("bkm,bnk->bmn", A1, B1) + torch.einsum("bkm,bnk->bmn", A2, B2)
    # strided form
    K.dgemm(M, N, K1, A1, M, 1, B1, K1, 0, C, N, batch=nb, sA=K1 * M, sB=N * K1, sC=M * N,
            seg2=(A2, M, B2, K2, K2, K2 * M, N * K2))
    assert relerr(C, ref) < 1e-12
    # address-table form, batches in scrambled order
    order = [3, 0, 4, 1, 2]
    tab = np.array([[A1[b].data_ptr(), B1[b].data_ptr(), A2[b].data_ptr(), B2[b].data_ptr(), C[b].data_ptr()]
                    for b in order], dtype=np.int64)
    C.zero_()
    for aligned in (False, bool(np.all(tab % 16 == 0))):
        K.dgemm(M, N, K1, A1, M, 1, B1, K1, 0, C, N, batch=nb, seg2=(A2, M, B2, K2, K2, 0, 0),
                table=torch.from_numpy(tab).to(DEV), table_align16=aligned)
        assert relerr(C, ref) < 1e-12


def test_dgemm_large_batch_chunking():
    nb, M, N, Kd = 70000, 3, 5, 4
    A, B = rnd(nb, M, Kd), rnd(nb, N, Kd)
    C = torch.empty(nb, M, N, dtype=torch.float64, device=DEV)
    K.dgemm(M, N, Kd, A, Kd, 0, B, Kd, 0, C, N, batch=nb, sA=M * Kd, sB=N * Kd, sC=M * N)
    assert relerr(C, torch.einsum("bmk,bnk->bmn", A, B)) < 1e-12


@pytest.mark.parametrize("shape", [(5,), (7, 9), (33, 65), (3, 4, 5, 6), (6, 6, 40, 41), (2, 3, 4, 5, 6, 7)])
def test_permute_all_orders(shape):
    x = rnd(*shape)
    perms = list(itertools.permutations(range(len(shape))))
    if len(perms) > 30:
        rng = np.random.default_rng(0)
        perms = [perms[i] for i in rng.choice(len(perms), 30, replace=False)]
    for p in perms:
        got = K.permuted(x, p, 1.5)
        assert torch.equal(got, 1.5 * x.permute(*p).contiguous()), p
    # accumulate into a strided destination
    p = perms[-1]
    dst = rnd(*[2 * s for s in x.permute(*p).shape], seed=9)
    view = dst[tuple(slice(0, s) for s in x.permute(*p).shape)]
    ref = dst.clone()
    ref[tuple(slice(0, s) for s in x.permute(*p).shape)] = -2.0 * x.permute(*p) + 0.5 * view
    K.strided_axpby(view, x.permute(*p), -2.0, 0.5)
    assert relerr(dst, ref) < 1e-15


def test_elementwise_kernels():
    no, nv = 5, 37
    t1, t2 = rnd(no, nv), rnd(no, no, nv, nv)
    eo = -torch.rand(no, dtype=torch.float64, device=DEV) - 0.5
    ev = torch.rand(nv, dtype=torch.float64, device=DEV) + 0.5
    D1 = eo[:, None] - ev
    D2 = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev
    tau = K.build_tau(t1, t2, 0.5, 2.0)
    assert relerr(tau, 0.5 * t2 + 2.0 * torch.einsum("ia,jb->ijab", t1, t1)) < 1e-15
    assert relerr(K.div_d2(t2, eo, ev), t2 / D2) < 1e-15
    assert relerr(K.div_d1(t1, eo, ev), t1 / D1) < 1e-15
    # fused symmetrise + update + rms
    r1, half = rnd(no, nv, seed=4), rnd(no, no, nv, nv, seed=5)
    full = half + half.permute(1, 0, 3, 2)
    a1, a2 = t1.clone(), t2.clone()
    h = half.clone()
    ssq = K.update_amps(r1, h, eo, ev, a1, a2, symmetrize=True, write_r2=True)
    assert relerr(h, full) < 1e-15
    assert relerr(a1, t1 + r1 / D1) < 1e-15
    assert relerr(a2, t2 + full / D2) < 1e-15
    ref = ((r1 / D1) ** 2).sum() + ((full / D2) ** 2).sum()
    assert abs(float(ssq[0]) - float(ref)) < 1e-12 * float(ref)
    a1, a2 = t1.clone(), t2.clone()
    h = half.clone()
    K.update_amps(r1, h, eo, ev, a1, a2, symmetrize=True, write_r2=False)
    assert torch.equal(h, half) and relerr(a2, t2 + full / D2) < 1e-15
    a1, a2 = t1.clone(), t2.clone()
    ssq2 = K.update_amps(r1, full.contiguous(), eo, ev, a1, a2, symmetrize=False)
    assert relerr(a2, t2 + full / D2) < 1e-15 and abs(float(ssq2[0]) - float(ref)) < 1e-12 * float(ref)
    h = half.clone()
    assert relerr(K.symmetrize_r2(h), full) < 1e-15
    # energy
    F = rnd(no + nv + 2, no + nv + 2, seed=7)
    fov = F[1:1 + no, 1 + no:1 + no + nv]
    L = rnd(no, no, nv, nv, seed=8)
    e = K.cc_energy(fov, t1, t2, L)
    ref = 2.0 * (fov * t1).sum() + ((t2 + torch.einsum("ia,jb->ijab", t1, t1)) * L).sum()
    assert abs(float(e[0]) - float(ref)) < 1e-11 * abs(float(ref))


def test_diis_kernels():
    n = 100003
    xs = [rnd(n, seed=s) for s in range(9)]
    d = K.multi_dot(xs[0], xs)
    ref = torch.stack([torch.dot(xs[0], x) for x in xs])
    assert relerr(d, ref) < 1e-12
    c = [0.3, -1.2, 0.0, 2.5, 1.0, -0.1, 0.7, 0.9, -3.0]
    out = torch.empty(n, dtype=torch.float64, device=DEV)
    K.multi_axpy(c, xs, out)
    assert relerr(out, sum(ci * x for ci, x in zip(c, xs))) < 1e-14
    z = torch.empty(n, dtype=torch.float64, device=DEV)
    K.axpbyz(2.0, xs[0], -3.0, xs[1], z)
    assert relerr(z, 2.0 * xs[0] - 3.0 * xs[1]) < 1e-15
    # determinism: identical bits on repeat
    assert torch.equal(K.multi_dot(xs[0], xs), d)


@pytest.mark.parametrize("M,N,Kd", [(1, 1, 1), (37, 19, 53), (129, 130, 18), (300, 257, 1000), (400, 150, 334), (64, 513, 16)])
def test_dgemm_tma_config(M, N, Kd):
    """config 6: cp.async.bulk.tensor (TMA, 128B swizzle) operand staging; K-major operands with even pitches."""
    lda, ldb = Kd + (Kd & 1) + 2, Kd + (Kd & 1)
    A = rnd(M, lda, seed=1)[:, :Kd]
    B = rnd(N, ldb, seed=2)[:, :Kd]
    C0 = rnd(M, N + 3, seed=3)
    C = C0.clone()
    K.dgemm(M, N, Kd, A, lda, 0, B, ldb, 0, C, N + 3, alpha=0.75, beta=-0.5, ksplit=1, config=6)
    ref = 0.75 * (A @ B.t()) - 0.5 * C0[:, :N]
    assert relerr(C[:, :N], ref) < 1e-12
    assert torch.equal(C[:, N:], C0[:, N:])


def test_dgemm_tma_batch_segments_splitk():
    nb, M, N, K1, K2 = 5, 150, 41, 64, 14
    A1, B1 = rnd(nb, M, K1), rnd(N, K1)                      # B shared by all batches (stride 0)
    A2, B2 = rnd(nb, M, K2, seed=5), rnd(nb, N, K2, seed=6)
    C = torch.zeros(nb, M, N, dtype=torch.float64, device=DEV)
    ref = torch.einsum("bmk,nk->bmn", A1, B1) + torch.einsum("bmk,bnk->bmn", A2, B2)
    K.dgemm(M, N, K1, A1, K1, 0, B1, K1, 0, C, N, batch=nb, sA=M * K1, sB=0, sC=M * N,
            seg2=(A2, K2, B2, K2, K2, M * K2, N * K2), ksplit=1, config=6)
    assert relerr(C, ref) < 1e-12
    M, N, Kd = 45, 70, 5000
    A, B = rnd(M, Kd), rnd(N, Kd)
    C = torch.empty(M, N, dtype=torch.float64, device=DEV)
    K.dgemm(M, N, Kd, A, Kd, 0, B, Kd, 0, C, N, ksplit=7, config=6)
    assert relerr(C, A @ B.t()) < 1e-12


SKINNY = [(40, 300, 300), (8, 700, 64), (30, 257, 1000), (40, 40, 5000), (3600, 40, 300), (1000, 24, 90), (257, 38, 17),
          (40, 1300, 40)]


@pytest.mark.parametrize("M,N,Kd", SKINNY)
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("batch", [1, 5])
def test_dgemm_skinny_tiles(M, N, Kd, ta, tb, batch):
    """one extent <= 40 takes the 40 x 256 / 256 x 40 tile configurations (gemm.cu configs 8-11: cp.async producers for
    transposed / unaligned operands, TMA for K-major ones), with split-K chosen by kernels.auto_ksplit"""
    A = rnd(batch, Kd, M, seed=1) if ta else rnd(batch, M, Kd, seed=1)
    B = rnd(batch, Kd, N, seed=2) if tb else rnd(batch, N, Kd, seed=2)
    C0 = rnd(batch, M, N, seed=3)
    C = C0.clone()
    K.dgemm(M, N, Kd, A, A.stride(1), ta, B, B.stride(1), tb, C, N, alpha=1.25, beta=0.5, batch=batch,
            sA=M * Kd, sB=N * Kd, sC=M * N)
    Am = A.transpose(1, 2) if ta else A
    Bm = B.transpose(1, 2) if tb else B
    ref = 1.25 * torch.einsum("bmk,bnk->bmn", Am, Bm) + 0.5 * C0
    assert relerr(C, ref) < 1e-12


@pytest.mark.parametrize("cfg", [8, 9, 10, 11])
def test_dgemm_skinny_configs_forced_on_general_shapes(cfg):
    """the skinny tile configurations are ordinary tilings: forced onto a shape with several ragged tiles in both
    directions they give the same product"""
    M, N, Kd = 300, 530, 777 + (cfg % 2)
    Kd += Kd % 2
    A, B = rnd(M, Kd, seed=4), rnd(N, Kd, seed=5)
    C = torch.zeros(M, N, dtype=torch.float64, device=DEV)
    K.dgemm(M, N, Kd, A, Kd, 0, B, Kd, 0, C, N, config=cfg, ksplit=1)
    assert relerr(C, A @ B.t()) < 1e-12


@pytest.mark.parametrize("M,N,Kd,batch", [(2000, 1500, 9000, 1), (1290, 1100, 8200, 2), (820, 3000, 8192, 2)])
@pytest.mark.parametrize("beta", [0.0, -0.5])
def test_dgemm_split_k_of_the_last_wave(M, N, Kd, batch, beta):
    """long-K K-major products with more units than SMs: the full waves run unsplit, the tiles of the last partial wave get
    their own launch with K split (compact partial tiles + reduction kernel) -- same product, bit-identical between runs"""
    A, B = rnd(batch, M, Kd, seed=6), rnd(batch, N, Kd, seed=7)
    C0 = rnd(batch, M, N, seed=8)
    outs = []
    for _ in range(2):
        C = C0.clone()
        K.dgemm(M, N, Kd, A, Kd, 0, B, Kd, 0, C, N, alpha=0.5, beta=beta, batch=batch, sA=M * Kd, sB=N * Kd, sC=M * N)
        outs.append(C)
    ref = 0.5 * torch.einsum("bmk,bnk->bmn", A, B) + beta * C0
    assert relerr(outs[0], ref) < 1e-12
    assert torch.equal(outs[0], outs[1])
