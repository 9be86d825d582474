"""(T) densities / Lambda sources (reference cctriples.py:1063-1157, ccwfn.py:300-304,1819-1829):
the numpy oracle against the reference's golden vectors, and the product path (pycc_b200.cctriples.t3_density,
CCwfn.t3_density, solve_cc with make_t3_density=True) against both.  `emu`: host logic through the numpy double of the
C ABI; `cuda` (-m gpu): the same assertions through libb200cc.so, plus medium sizes against the oracle.
FP64 tolerances (north_star): 1e-10 Eh on energies, 1e-9 max-abs on tensors; the golden checks are far tighter."""
import glob
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import pycc_b200
from pycc_b200 import cctriples
from pycc_b200.synthetic import Synthetic, blocks_from_factor, make_synthetic
from oracle import ccsd_oracle as co, t3density_oracle as do, triples_oracle as to
from tests import emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T3D = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "t3d_*.npz")))
DEV = [torch.device("cpu")]


def load(path):
    g = dict(np.load(path))
    tag = os.path.basename(path)[4:-4]
    r = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % tag)))
    syn = Synthetic(int(r["no"]), int(r["nv"]), r["B"], r["F"], float(r["scale"]), int(r["seed"]))
    return g, r, syn


@pytest.fixture(params=T3D, ids=[os.path.basename(p)[4:-4] for p in T3D])
def t3d(request):
    return load(request.param)


@pytest.fixture(params=[pytest.param("emu"), pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request):
    if request.param == "emu":
        DEV[0] = torch.device("cpu")
        with emu.install():
            yield DEV[0]
    else:
        assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
        DEV[0] = torch.device("cuda:0")
        yield DEV[0]
        DEV[0] = torch.device("cpu")


def T(x):
    return torch.from_numpy(np.array(x, dtype=np.float64, order="C", copy=True)).to(DEV[0])


def worst(dens, want):
    return max(float(np.abs(dens[k].cpu().numpy() - want[k]).max()) for k in do.NAMES)


# ---------------------------------------------------------------------------------------------------------
# oracle vs the reference's own outputs
# ---------------------------------------------------------------------------------------------------------
def test_oracle_t3_density(t3d):
    g, r, syn = t3d
    b = blocks_from_factor(syn)
    et, d = do.t3_density(g["t1"], g["t2"], syn.F, b["ovvv"], b["ooov"], b["oovv"])
    assert abs(et - float(g["et"])) < 1e-14
    assert abs(et - float(r["e_t_tjl"])) < 1e-13          # same E(T) as the Lee-Rendell driver
    for k in do.NAMES:
        assert np.abs(d[k] - g[k]).max() < 1e-13, k


def test_oracle_j_partition_sums_to_whole(t3d):
    g, r, syn = t3d
    b = blocks_from_factor(syn)
    args = (g["t1"], g["t2"], syn.F, b["ovvv"], b["ooov"], b["oovv"])
    _, p0 = do.t3_density(*args, js=range(0, syn.no, 2))
    _, p1 = do.t3_density(*args, js=range(1, syn.no, 2))
    S2 = p0["S2"] + p1["S2"]
    assert np.abs(S2 + S2.transpose(1, 0, 3, 2) - g["S2"]).max() < 1e-13
    for k in ("Dov", "Goovv", "Gooov", "Gvvvo", "S1", "Doo", "Dvv"):
        assert np.abs(p0[k] + p1[k] - g[k]).max() < 1e-13, k


# ---------------------------------------------------------------------------------------------------------
# product vs the reference's own outputs
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k_batch", [None, 1, 3])
def test_t3_density_matches_reference(t3d, dev, k_batch):
    g, r, syn = t3d
    cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True)
    cc.t1, cc.t2 = T(g["t1"]), T(g["t2"])
    et, dens = cctriples.t3_density(cc.o, cc.v, cc.no, cc.nv, cc.t1, cc.t2, cc.H.F, cc.H.ERI, cc.H.L, cc.contract,
                                    k_batch=k_batch)
    assert abs(float(et) - float(g["et"])) < 1e-13
    for k in do.NAMES:
        assert np.abs(dens[k].cpu().numpy() - g[k]).max() < 1e-12, k
    assert set(dens) == set(do.NAMES)


def test_solve_cc_with_make_t3_density(t3d, dev):
    """ccwfn.py:300-304: with make_t3_density the (T) step of solve_cc is t3_density, which also caches the pieces."""
    g, r, syn = t3d
    cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True, make_t3_density=True)
    e = cc.solve_cc(1e-12, 1e-12)
    assert abs(float(e) - float(r["e_total_ccsd_t"])) < 1e-10
    for k in do.NAMES:
        assert np.abs(getattr(cc, k).cpu().numpy() - g[k]).max() < 1e-9, k
    # and the plain driver still agrees
    assert abs(float(cctriples.t_tjl(cc)) - float(g["et"])) < 1e-12


def test_make_t3_density_keyword():
    from pycc_b200.exceptions import InvalidKeywordError
    with pytest.raises(InvalidKeywordError):
        pycc_b200.ccwfn(make_synthetic(2, 3), model="CCSD(T)", make_t3_density="yes")


def test_odd_sizes_and_fock_noise(dev):
    """odd o / v take the non-TMA operand path of the t3 GEMMs; ragged 8-cubes; non-canonical f_ov terms"""
    for (no, nv, noise) in ((3, 5, 0.02), (2, 9, 0.0), (5, 6, 0.01)):
        syn = make_synthetic(no, nv, seed=5, fock_noise=noise)
        b = blocks_from_factor(syn)
        rng = np.random.default_rng(7)
        t1 = 0.05 * rng.standard_normal((no, nv))
        t2 = 0.05 * rng.standard_normal((no, no, nv, nv))
        t2 = t2 + t2.transpose(1, 0, 3, 2)
        et, want = do.t3_density(t1, t2, syn.F, b["ovvv"], b["ooov"], b["oovv"])
        cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True)
        cc.t1, cc.t2 = T(t1), T(t2)
        e = float(cc.t3_density())
        assert abs(e - et) < 1e-12, (no, nv)
        assert worst({k: getattr(cc, k) for k in do.NAMES}, want) < 1e-12, (no, nv)


def test_mixed_precision_t3_density(t3d, dev):
    """precision='MP': t3 build and the K-major density products as split-TF32 GEMMs with FP64 accumulation; E(T) to
    1e-6 Eh (north_star), the pieces to 1e-6 relative to their largest element"""
    from pycc_b200 import kernels as K
    g, r, syn = t3d
    keep = (K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles)
    K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles = 0.0, 1, 1
    try:
        cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True, precision="MP")
        cc.t1, cc.t2 = T(g["t1"]), T(g["t2"])
        g0 = K.MIXED.stats["gemm"]
        e = float(cc.t3_density())
        assert K.MIXED.stats["gemm"] > g0
    finally:
        K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles = keep
    assert abs(e - float(g["et"])) < 1e-6
    for k in do.NAMES:
        assert np.abs(getattr(cc, k).cpu().numpy() - g[k]).max() < 1e-6 * max(1.0, np.abs(g[k]).max()), k


def test_frozen_core_offsets(dev):
    """nfzc > 0: o / v are offset slices of the full MO space (wavefunction.py:304-315); eps and f_ov must follow"""
    from pycc_b200.synthetic import full_eri
    from pycc_b200.wavefunction import IntegralReference
    nf, no, nv = 2, 3, 6
    syn = make_synthetic(nf + no, nv, seed=9, fock_noise=0.01)
    ERI = full_eri(syn)
    n = nf + no + nv
    o, v = slice(nf, nf + no), slice(nf + no, n)
    rng = np.random.default_rng(3)
    t1 = 0.05 * rng.standard_normal((no, nv))
    t2 = 0.05 * rng.standard_normal((no, no, nv, nv))
    t2 = t2 + t2.transpose(1, 0, 3, 2)
    et, want = do.t3_density(t1, t2, syn.F, ERI[o, v, v, v], ERI[o, o, o, v], ERI[o, o, v, v], nfzc=nf)
    cc = pycc_b200.ccwfn(IntegralReference.from_arrays(syn.F, ERI, no, nf), model="CCSD(T)", device="GPU", quiet=True)
    cc.t1, cc.t2 = T(t1), T(t2)
    assert abs(float(cc.t3_density()) - et) < 1e-12
    assert worst({k: getattr(cc, k) for k in do.NAMES}, want) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("no,nv,seed,noise,kb", [(6, 26, 0, 0.01, None), (8, 40, 1, 0.0, 3), (5, 33, 2, 0.01, 2)])
def test_medium_size_vs_oracle(no, nv, seed, noise, kb):
    syn = make_synthetic(no, nv, seed=seed, fock_noise=noise)
    b = blocks_from_factor(syn)
    P = co.Problem(b, syn.F, no)
    _, t1, t2, _ = co.solve_cc(P, 1e-11, 1e-11, 100)
    et, want = do.t3_density(t1, t2, syn.F, b["ovvv"], b["ooov"], b["oovv"])
    DEV[0] = torch.device("cuda:0")
    try:
        cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True)
        cc.t1, cc.t2 = T(t1), T(t2)
        e, dens = cctriples.t3_density(cc.o, cc.v, cc.no, cc.nv, cc.t1, cc.t2, cc.H.F, cc.H.ERI, cc.H.L, cc.contract,
                                       k_batch=kb)
        assert abs(float(e) - et) < 1e-10
        assert abs(float(e) - to.t_tjl(t1, t2, syn.F, b["ovvv"], b["ooov"], b["oovv"])) < 1e-10
        assert worst(dens, want) < 1e-9
    finally:
        DEV[0] = torch.device("cpu")


# ---------------------------------------------------------------------------------------------------------
# N > 1: j dealt round-robin to the ranks, pieces all-reduced (gloo, numpy double of the C ABI)
# ---------------------------------------------------------------------------------------------------------
def _worker(rank, world, port, path, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pycc_b200.parallel import Comm
        g, r, syn = load(path)
        with emu.install():
            cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True, comm=Comm())
            cc.t1, cc.t2 = torch.from_numpy(g["t1"].copy()), torch.from_numpy(g["t2"].copy())
            e = float(cc.t3_density())
            q.put((rank, abs(e - float(g["et"])), worst({k: getattr(cc, k) for k in do.NAMES}, g)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_match_reference():
    path = [p for p in T3D if "o4v10_s1" in p][0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 11
    procs = [ctx.Process(target=_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, de, dd in res:
        assert de < 1e-13 and dd < 1e-12, (rank, de, dd)
