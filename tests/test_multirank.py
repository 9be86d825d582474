"""N > 1 host logic on CPU: two gloo ranks, each driving the numpy double of the C ABI.  Checks the
rank partition of the residual (a-sharded ladder, occupied-sliced ring/W/Z terms, one all-reduce of r2)
and the round-robin (T) against the reference goldens."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, golden_path, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pycc_b200
        from pycc_b200 import cctriples
        from pycc_b200.parallel import Comm
        from tests import emu
        from tests.conftest import load_golden
        g, syn = load_golden(golden_path)
        with emu.install():
            comm = Comm()
            cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True, comm=comm)
            assert cc.H.block("vvvv").shape[0] == comm.a_range(syn.nv)[1] - comm.a_range(syn.nv)[0]
            t1 = torch.from_numpy(g["rand_t1"].copy())
            t2 = torch.from_numpy(g["rand_t2"].copy())
            r1, r2 = cc.residuals(cc.H.F, t1, t2)
            e1 = float(np.abs(r1.numpy() - g["rand_r1"]).max())
            e2 = float(np.abs(r2.numpy() - g["rand_r2"]).max())
            ecc = cc.solve_cc(1e-11, 1e-11)
            ee = abs(float(ecc) - float(g["e_total_ccsd_t"]))
            cc.t1 = torch.from_numpy(g["conv_t1"].copy())
            cc.t2 = torch.from_numpy(g["conv_t2"].copy())
            et = abs(float(cctriples.t_tjl(cc)) - float(g["e_t_tjl"]))
            q.put((rank, e1, e2, ee, et, len(cc.trace)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_ranks_match_reference(world):
    from tests.conftest import GOLDEN
    path = [p for p in GOLDEN if "o4v10_s1" in p][0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    iters = set()
    for rank, e1, e2, ee, et, n in res:
        assert e1 < 1e-12 and e2 < 1e-12, (rank, e1, e2)
        assert ee < 1e-10 and et < 1e-12, (rank, ee, et)
        iters.add(n)
    assert len(iters) == 1


def _worker_mp(rank, world, port, golden_path, q):
    """precision='MP' with two ranks: <ab|ef> planes a-sharded, mixed GEMMs on the rank-local slices."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pycc_b200
        from pycc_b200 import kernels as K
        from pycc_b200.parallel import Comm
        from tests import emu
        from tests.conftest import load_golden
        g, syn = load_golden(golden_path)
        K.MIXED.min_flops, K.MIXED.min_dim, K.MIXED.min_tiles = 0.0, 1, 1
        with emu.install():
            comm = Comm()
            cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True, comm=comm, precision="MP")
            a0, a1 = comm.a_range(syn.nv)
            # only the TF32 planes of this rank's pair-packed rows (a >= b, a in its range) are resident
            assert not cc.H.has("vvvv") and cc.H.vvvv_packed is None
            assert cc.H.vvvv_planes[0].shape[:2] == (2, a1 * (a1 + 1) // 2 - a0 * (a0 + 1) // 2)
            g0 = K.MIXED.stats["gemm"]
            ecc = cc.solve_cc(1e-7, 1e-7)
            ee = abs(float(ecc) - float(g["e_total_ccsd_t"]))
            dt = float(np.abs(cc.t2.numpy() - g["conv_t2"]).max())
            q.put((rank, ee, dt, len(cc.trace), K.MIXED.stats["gemm"] - g0, float(ecc)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_mixed_precision():
    from tests.conftest import GOLDEN
    path = [p for p in GOLDEN if "o4v10_s1" in p][0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 7
    procs = [ctx.Process(target=_worker_mp, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ee, dt, n, ngemm, e in res:
        assert ee < 1e-6 and dt < 1e-6, (rank, ee, dt)
        assert ngemm > 0
    assert res[0][3] == res[1][3] and res[0][5] == res[1][5]        # replicated state stays identical


def test_split_is_a_partition():
    from pycc_b200.parallel import split
    for n in (0, 1, 7, 40, 300):
        for size in (1, 2, 3, 8, 11):
            parts = [split(n, size, r) for r in range(size)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[r][1] == parts[r + 1][0] for r in range(size - 1))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1


def _worker_lambda(rank, world, port, lam_path, q, precision="DP"):
    """HBAR + Lambda with an a-sharded <ab|ef>: the t1.<ab|ef> term of Hvvvo and the Lambda ladder are rank-local pieces
    summed by all-reduces, everything else is replicated."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pycc_b200
        from pycc_b200.parallel import Comm
        from tests import emu
        from tests.test_lambda import load
        g, syn, model = load(lam_path)
        with emu.install():
            cc = pycc_b200.ccwfn(syn, model=model, device="GPU", quiet=True, comm=Comm(), precision=precision)
            assert cc._vvvv_released() == (precision == "MP")
            cc.t1, cc.t2 = torch.from_numpy(g["t1"].copy()), torch.from_numpy(g["t2"].copy())
            hb = pycc_b200.cchbar(cc)
            dh = float(np.abs(hb.Hvvvo.numpy() - g["Hvvvo"]).max())
            lm = pycc_b200.cclambda(cc, hb)
            lecc = lm.solve_lambda(*((1e-12, 1e-12, 100) if precision == "DP" else (1e-8, 1e-8, 100)))
            q.put((rank, dh, abs(float(lecc) - float(g["lecc"])), float(np.abs(lm.l2.numpy() - g["conv_l2"]).max()),
                   len(lm.trace), len(g["trace_lecc_rms"])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_lambda_with_sharded_vvvv(world):
    import glob
    path = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "lam_o4v10_s1_noise_ccsd.npz")))[0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 20 + world
    procs = [ctx.Process(target=_worker_lambda, args=(r, world, port, path, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, dh, de, dl, n, nref in res:
        assert dh < 1e-12 and de < 1e-11 and dl < 1e-10 and n == nref, (rank, dh, de, dl, n, nref)


def test_lambda_with_sharded_vvvv_mixed_precision():
    """precision='MP' on two ranks: the a-sharded <ab|ef> exists only as TF32 planes; the t1.<ab|ef> piece of Hvvvo is
    rebuilt from them (b200cc_merge_tf32), the Lambda ladder uses them directly.  Within 1e-6 of the FP64 reference."""
    import glob
    path = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "lam_o4v10_s1_noise_ccsd.npz")))[0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 31
    procs = [ctx.Process(target=_worker_lambda, args=(r, 2, port, path, q, "MP")) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, dh, de, dl, n, nref in res:
        assert dh < 1e-6 and de < 1e-6 and dl < 1e-6, (rank, dh, de, dl)


def _worker_cc3(rank, world, port, tag, q):
    """model='CC3' on several ranks: t1.<ab|ef> of W_abei from the rank's pairs (one all-reduce), the (i,j) pairs of the
    triples loop dealt round-robin, their partial (X1, X2) summed by the residual's own all-reduce."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pycc_b200
        from pycc_b200.parallel import Comm
        from tests import emu
        from tests.test_cc3 import load, load_rt
        g, r, syn = load(os.path.join(ROOT, "tests", "golden", "cc3_%s.npz" % tag))
        grt, _, _ = load_rt(os.path.join(ROOT, "tests", "golden", "rtcc3_%s.npz" % tag))
        with emu.install():
            comm = Comm()
            cc = pycc_b200.ccwfn(syn, model="CC3", device="GPU", quiet=True, comm=comm)
            assert cc.H.a_range == tuple(comm.a_range(syn.nv))
            o, v, H = cc.o, cc.v, cc.H
            t1, t2 = torch.from_numpy(g["t1"].copy()), torch.from_numpy(g["t2"].copy())
            dw = float(np.abs(cc.build_cc3_Wabei(o, v, H.ERI, t1).numpy() - g["Wabei"]).max())
            X1, X2 = cc._cc3_t_residual(o, v, H.F, H.ERI, H.L, t1, t2, cc.build_Fme(o, v, H.F, H.L, t1))
            dx = max(float(np.abs(X1.numpy() - g["X1"]).max()), float(np.abs(X2.numpy() - g["X2"]).max()))
            r1, r2 = cc.residuals(H.F, t1, t2)
            dr = max(float(np.abs(r1.numpy() - g["r1"]).max()), float(np.abs(r2.numpy() - g["r2"]).max()))
            F_el = torch.from_numpy(grt["F_el"].copy())
            q1, q2 = cc.residuals(F_el, t1, t2, real_time=True)
            drt = max(float(np.abs(q1.numpy() - grt["r1_el"]).max()), float(np.abs(q2.numpy() - grt["r2_el"]).max()))
            ecc = cc.solve_cc(1e-12, 1e-12)
            q.put((rank, dw, dx, dr, drt, abs(float(ecc) - float(g["ecc"])), len(cc.trace), len(g["trace_ecc_rms"]),
                   float(np.abs(cc.t2.numpy() - g["conv_t2"]).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_cc3_on_several_ranks(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 40 + world
    procs = [ctx.Process(target=_worker_cc3, args=(r, world, port, "o4v10_s1_noise", q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, dw, dx, dr, drt, de, n, nref, dt in res:
        assert dw < 1e-12 and dx < 1e-12 and dr < 1e-12 and drt < 1e-12, (rank, dw, dx, dr, drt)
        assert de < 1e-11 and n == nref and dt < 1e-10, (rank, de, n, nref, dt)


def _worker_complex(rank, world, port, tag, q):
    """the RT-CC right-hand side on several ranks: sampled residuals with the sharded formulation, the two ladder planes
    all-reduced separately; generic and pair-symmetric complex amplitudes against the reference's golden / the oracle"""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pycc_b200
        from pycc_b200.parallel import Comm
        from pycc_b200.synthetic import blocks_from_factor
        from oracle import ccsd_oracle as co
        from tests import emu
        from tests.test_complex import load
        g, r, syn = load(os.path.join(ROOT, "tests", "golden", "cplx_%s.npz" % tag))
        P = co.Problem(blocks_from_factor(syn), syn.F, syn.no)
        T = lambda x: torch.from_numpy(np.array(x, order="C", copy=True))
        with emu.install():
            cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True, comm=Comm())
            r1, r2 = cc.residuals(T(g["F_mag"]), T(g["t1"]), T(g["t2"]), real_time=True)
            d_gen = max(float(np.abs(r1.numpy() - g["r1_mag"]).max()), float(np.abs(r2.numpy() - g["r2_mag"]).max()))
            t2s = 0.5 * (g["t2"] + g["t2"].transpose(1, 0, 3, 2))
            w1, w2 = P.residuals(g["F_mag"], g["t1"], t2s)
            r1, r2 = cc.residuals(T(g["F_mag"]), T(g["t1"]), T(t2s), real_time=True)
            d_sym = max(float(np.abs(r1.numpy() - w1).max()), float(np.abs(r2.numpy() - w2).max()))
            # the Lambda half on planes with the sharded <ab|ef>: t1.<ab|ef> of H_abei and the lambda2 ladder per plane
            cc.t1, cc.t2 = T(r["conv_t1"]), T(r["conv_t2"])
            lm = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc))
            q1, q2 = lm.residuals(T(g["F_mag"]), T(g["t1"]), T(g["t2"]), T(g["l1"]), T(g["l2"]))
            d_lam = max(float(np.abs(q1.numpy() - g["rl1_mag"]).max()), float(np.abs(q2.numpy() - g["rl2_mag"]).max()))
            q.put((rank, d_gen, max(d_sym, d_lam / 10.0)))
    finally:
        dist.destroy_process_group()


def test_complex_residual_on_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 55
    procs = [ctx.Process(target=_worker_complex, args=(r, 2, port, "o4v10_s1_noise", q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, d_gen, d_sym in res:
        assert d_gen < 1e-11 and d_sym < 1e-11, (rank, d_gen, d_sym)
