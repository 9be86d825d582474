"""Mixed-precision (split-TF32 on tcgen05, FP64 accumulation) kernels through the C ABI against torch float64 on the
same device.  All `-m gpu`.

Tolerance (written out in ``tol``): the tensor core's FP32 accumulate truncates, so an accumulator that receives
n MMAs is off by at most n ulp = n * 1.2e-7 of its magnitude (measured mean: n * 1.6e-8, a systematic shrink);
n = MMAs per K chunk into the large accumulator (3 per 8 summation indices for the shared-accumulator kernels,
configs 1-4; 1 per 8 for the register-accumulating kernel, configs 0/5/6).  On top: the 2^-22 split error and one
FP32 rounding per chunk -> 1e-6.  Everything relative to the largest |result|.
"""
import numpy as np
import pytest
import torch

from pycc_b200 import kernels as K

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def rnd(*shape, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return torch.randn(*shape, dtype=torch.float64, generator=g).to(DEV)


def tf32_rna_torch(x32):
    u = x32.view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)


def test_split_tf32_planes():
    rows, Kd, ld = 37, 53, 61
    buf = rnd(2, rows, ld) * 3.0
    hi, lo, ldp = K.split_tf32(buf, rows, Kd, ld, batch=2, stride=rows * ld)
    assert ldp == 56 and tuple(hi.shape) == (2, rows, ldp)
    x = buf[:, :, :Kd]
    h = tf32_rna_torch(x.to(torch.float32))
    l = tf32_rna_torch((x - h.double()).to(torch.float32))
    assert torch.equal(hi[:, :, :Kd], h) and torch.equal(lo[:, :, :Kd], l)
    assert float(hi[:, :, Kd:].abs().max()) == 0.0 and float(lo[:, :, Kd:].abs().max()) == 0.0
    # hi + lo reproduces the double to ~2^-22
    assert float(((hi.double() + lo.double())[:, :, :Kd] - x).abs().max() / x.abs().max()) < 2.0 ** -21
    # low 13 mantissa bits are clear: the tensor core reads exactly these values
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0 and int((lo.view(torch.int32) & 0x1FFF).abs().max()) == 0


def test_merge_tf32_is_the_inverse_of_split():
    """b200cc_merge_tf32: dst = hi + lo, exactly (both planes are FP32, their sum is exact in FP64), with a row offset into
    the planes and an odd K (scalar tail), untouched columns beyond K."""
    rows, Kd = 45, 361                                        # 19^2: the cc-pVDZ water case
    x = rnd(rows, Kd) * 2.0
    hi, lo, ldp = K.split_tf32(x, rows, Kd, Kd)
    hi, lo = hi[0], lo[0]
    want = hi.double() + lo.double()
    got = K.merge_tf32(hi, lo, ldp, rows, Kd)
    assert tuple(got.shape) == (rows, Kd) and torch.equal(got, want[:, :Kd])
    assert float((got - x).abs().max() / x.abs().max()) < 2.0 ** -21
    r0 = 7
    part = K.merge_tf32((hi, r0 * ldp), (lo, r0 * ldp), ldp, rows - r0, Kd)
    assert torch.equal(part, want[r0:, :Kd])
    out = torch.full((rows, Kd), -7.0, dtype=torch.float64, device=x.device)
    K.merge_tf32(hi, lo, ldp, rows, Kd - 3, out=out[:, :Kd - 3].contiguous())      # K < ldp - 3: whole float4 groups skipped
    assert float(out.min()) == -7.0


SHAPES = [(128, 256, 32), (128, 256, 64), (128, 128, 256), (256, 512, 2048), (37, 19, 53), (129, 300, 1000),
          (400, 513, 333), (1, 300, 77), (257, 1, 40), (1600, 1200, 4096)]


def tol(config, kchunk, Kd, ref):
    per8 = 1 if config in (0, 5, 6, 7) else 3
    nops = per8 * (min(kchunk, Kd) + 7) // 8
    return (nops * 1.2e-7 + 1e-6) * float(ref.abs().max())


@pytest.mark.parametrize("config", [7, 5, 6, 1, 2, 3, 4])
@pytest.mark.parametrize("M,N,Kd", SHAPES)
def test_gemm_tf32x3(M, N, Kd, config):
    A, B = rnd(M, Kd, seed=1), rnd(N, Kd, seed=2)
    C0 = rnd(M, N + 3, seed=3)
    C = C0.clone()
    Ah, Al, lpa = K.split_tf32(A, M, Kd, Kd)
    Bh, Bl, lpb = K.split_tf32(B, N, Kd, Kd)
    K.gemm_tf32x3(M, N, Kd, Ah, Al, lpa, Bh, Bl, lpb, C, N + 3, alpha=0.75, beta=-0.5, kchunk=256, config=config)
    ref = 0.75 * (A @ B.t()) - 0.5 * C0[:, :N]
    assert float((C[:, :N] - ref).abs().max()) < tol(config, 256, Kd, ref)
    assert torch.equal(C[:, N:], C0[:, N:])            # nothing written past column N


@pytest.mark.parametrize("config", [0, 1, 7])
@pytest.mark.parametrize("kchunk", [32, 256, 100000])
def test_gemm_tf32x3_chunked_accumulation(kchunk, config):
    """Many FP64 drains (kchunk=32: one per k-block), few, and none: same answer; beta applied exactly once."""
    M, N, Kd = 300, 520, 6000
    A, B = rnd(M, Kd, seed=4), rnd(N, Kd, seed=5)
    C0 = rnd(M, N, seed=6)
    C = C0.clone()
    Ah, Al, lpa = K.split_tf32(A, M, Kd, Kd)
    Bh, Bl, lpb = K.split_tf32(B, N, Kd, Kd)
    K.gemm_tf32x3(M, N, Kd, Ah, Al, lpa, Bh, Bl, lpb, C, N, alpha=-1.25, beta=2.0, kchunk=kchunk, config=config)
    ref = -1.25 * (A @ B.t()) + 2.0 * C0
    assert float((C - ref).abs().max()) < tol(config, kchunk, Kd, ref)


@pytest.mark.parametrize("config", [0, 7])
def test_gemm_tf32x3_batched_and_shared_operand(config):
    nb, M, N, Kd = 5, 150, 260, 700
    A, B = rnd(nb, M, Kd, seed=7), rnd(N, Kd, seed=8)
    C = torch.zeros(nb, M, N, dtype=torch.float64, device=DEV)
    Ah, Al, lpa = K.split_tf32(A, M, Kd, Kd, batch=nb, stride=M * Kd)
    Bh, Bl, lpb = K.split_tf32(B, N, Kd, Kd)
    K.gemm_tf32x3(M, N, Kd, Ah, Al, lpa, Bh, Bl, lpb, C, N, batch=nb, sA=M * lpa, sB=0, sC=M * N, config=config)
    ref = torch.einsum("bmk,nk->bmn", A, B)
    assert float((C - ref).abs().max()) < tol(0, 256, Kd, ref)


@pytest.mark.parametrize("lockstep", [-1, 0, 1, 4, 7])
@pytest.mark.parametrize("M,N,Kd,nb", [(19 * 128 - 50, 11 * 128 + 5, 1536, 1), (13 * 128 - 64, 30 * 128, 800, 1),
                                       (5 * 128, 3 * 128 + 1, 640, 3)])
def test_gemm_tf32x3_leader_follower_schedule(M, N, Kd, nb, lockstep):
    """Plain unit schedule (-1 / auto at these sizes) vs the leader/follower block schedule (skew 1, 4, 7): ragged bands of
    m-tiles (19 -> 12 + 7 with idle CTAs in the second band), several rounds of n-blocks, batches."""
    A, B = rnd(nb, M, Kd, seed=11), rnd(N, Kd, seed=12)
    C = torch.zeros(nb, M, N, dtype=torch.float64, device=DEV)
    Ah, Al, lpa = K.split_tf32(A, M, Kd, Kd, batch=nb, stride=M * Kd)
    Bh, Bl, lpb = K.split_tf32(B, N, Kd, Kd)
    K.gemm_tf32x3(M, N, Kd, Ah, Al, lpa, Bh, Bl, lpb, C, N, batch=nb, sA=M * lpa if nb > 1 else 0, sB=0, sC=M * N,
                  lockstep=lockstep)
    ref = torch.einsum("bmk,nk->bmn", A, B)
    assert float((C - ref).abs().max()) < tol(0, 256, Kd, ref)


def test_gemm_tf32x3_pairs_many_tiles():
    """CTA-pair kernel: several 256x128 tiles per pair, ragged M (last pair tile half empty) and N."""
    M, N, Kd = 9 * 256 + 100, 21 * 128 + 70, 1100
    A, B = rnd(M, Kd, seed=13), rnd(N, Kd, seed=14)
    C0 = rnd(M, N, seed=15)
    C = C0.clone()
    Ah, Al, lpa = K.split_tf32(A, M, Kd, Kd)
    Bh, Bl, lpb = K.split_tf32(B, N, Kd, Kd)
    K.gemm_tf32x3(M, N, Kd, Ah, Al, lpa, Bh, Bl, lpb, C, N, alpha=0.5, beta=1.0, config=7)
    ref = 0.5 * (A @ B.t()) + C0
    assert float((C - ref).abs().max()) < tol(7, 256, Kd, ref)


def test_two_segment_slab_indexed_gemm_matches_fp64():
    """The (T) launch shape: batch entries pick operand slabs by index (bcoords) and chain a second K segment into the
    same tile; mixed kernel vs the FP64 DMMA kernel on identical arguments."""
    nsA, nsB, M, N, K1, K2, nb = 4, 6, 700, 136, 136, 12, 9
    A1, B1 = rnd(nsA, M, K1, seed=21), rnd(nsB, N, K1, seed=22)
    A2, B2 = rnd(nsA, M, K2, seed=23), rnd(nsB, N, K2, seed=24)
    g = torch.Generator().manual_seed(5)
    co = torch.stack([torch.randint(0, nsA, (nb,), generator=g), torch.randint(0, nsB, (nb,), generator=g),
                      torch.randint(0, nsA, (nb,), generator=g), torch.randint(0, nsB, (nb,), generator=g)], 1)
    co = co.to(torch.int32).to(DEV)
    out = {}
    for mixed in (False, True):
        C = torch.zeros(nb, M, N, dtype=torch.float64, device=DEV)
        keep = (K.MIXED.min_flops, K.MIXED.min_tiles)
        K.MIXED.min_flops, K.MIXED.min_tiles = 0.0, 1
        try:
            with K.mixed_mode(mixed):
                K.dgemm(M, N, K1, A1, K1, 0, B1, K1, 0, C, N, 1.0, 0.0, batch=nb, sA=M * K1, sB=N * K1, sC=M * N,
                        seg2=(A2, K2, B2, K2, K2, M * K2, N * K2), bcoords=co, nbatch=(nsA, nsB, nsA, nsB), ksplit=1)
        finally:
            K.MIXED.min_flops, K.MIXED.min_tiles = keep
        out[mixed] = C
    c = co.cpu().numpy()
    ref = torch.stack([A1[c[b, 0]] @ B1[c[b, 1]].t() + A2[c[b, 2]] @ B2[c[b, 3]].t() for b in range(nb)])
    assert float((out[False] - ref).abs().max()) < 1e-11 * float(ref.abs().max())
    err = float((out[True] - ref).abs().max())
    assert 0.0 < err < tol(0, 256, K1 + K2, ref)


def test_dgemm_routes_to_mixed_when_enabled():
    M, N, Kd = 1280, 1024, 2048
    A, B = rnd(M, Kd, seed=9), rnd(N, Kd, seed=10)
    C = torch.zeros(M, N, dtype=torch.float64, device=DEV)
    before = dict(K.MIXED.stats)
    K.MIXED.on = True
    try:
        K.dgemm(M, N, Kd, A, Kd, 0, B, Kd, 0, C, N)
    finally:
        K.MIXED.on = False
    assert K.MIXED.stats["gemm"] == before["gemm"] + 1
    ref = A @ B.t()
    err = float((C - ref).abs().max())
    assert 0.0 < err < tol(0, 256, Kd, ref)      # TF32-split, not bitwise FP64
