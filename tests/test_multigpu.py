"""N > 1 on real GPUs (NCCL): skipped unless at least two CUDA devices are visible.  Same checks as
tests/test_multirank.py (gloo + numpy double), but through libb200cc.so on each rank's own GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import pycc_b200
        from pycc_b200.parallel import Comm
        from pycc_b200.synthetic import make_synthetic, blocks_from_factor
        from oracle import ccsd_oracle as co, triples_oracle as to
        no, nv = 6, 26
        syn = make_synthetic(no, nv, seed=3, fock_noise=0.01)
        comm = Comm()
        cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True, comm=comm)
        e = float(cc.solve_cc(1e-11, 1e-11))
        b = blocks_from_factor(syn)
        P = co.Problem(b, syn.F, no)
        e_ref, t1, t2, trace = co.solve_cc(P, 1e-11, 1e-11)
        et = to.t_tjl(t1, t2, syn.F, b["ovvv"], b["ooov"], b["oovv"])
        # (T) densities: (i >= j) pairs dealt over the ranks, pieces all-reduced (NCCL)
        from oracle import t3density_oracle as do
        et_d, want = do.t3_density(t1, t2, syn.F, b["ovvv"], b["ooov"], b["oovv"])
        e_d = float(cc.t3_density())
        dd = max(float(np.abs(getattr(cc, k).cpu().numpy() - want[k]).max()) for k in do.NAMES)
        # HBAR + Lambda with the (T) sources, <ab|ef> a-sharded (rank-local ladder pieces + all-reduce)
        from oracle import lambda_oracle as lo
        lecc_ref, l1_ref, l2_ref, ltrace = lo.solve_lambda(P, t1, t2, 1e-11, 1e-11, 100, model="CCSD(T)",
                                                           s1=want["S1"], s2=want["S2"])
        lm = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc))
        lecc = float(lm.solve_lambda(1e-11, 1e-11, 100))
        dl = float(np.abs(lm.l2.cpu().numpy() - l2_ref).max())
        q.put((rank, abs(e - (e_ref + et)), float(np.abs(cc.t2.cpu().numpy() - t2).max()), len(cc.trace), len(trace),
               abs(e_d - et_d), dd, abs(lecc - lecc_ref), dl))
    finally:
        dist.destroy_process_group()


def test_nccl_ranks_match_oracle():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, de, dt, n, nref, ded, dd, dle, dl in res:
        assert de < 1e-10 and dt < 1e-9 and n == nref, (rank, de, dt, n, nref)
        assert ded < 1e-10 and dd < 1e-9, (rank, ded, dd)
        assert dle < 1e-10 and dl < 1e-9, (rank, dle, dl)


def _worker_cc3(rank, world, port, q):
    """model='CC3' (and its real-time residual) and precision='MP' CCSD on NCCL ranks, against the numpy oracle."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import pycc_b200
        from pycc_b200.parallel import Comm
        from pycc_b200.synthetic import make_synthetic, blocks_from_factor
        from oracle import ccsd_oracle as co, cc3_oracle as c3
        no, nv = 6, 26
        syn = make_synthetic(no, nv, seed=4, fock_noise=0.01)
        P = co.Problem(blocks_from_factor(syn), syn.F, no)
        comm = Comm()
        dev = torch.device("cuda", rank)
        cc = pycc_b200.ccwfn(syn, model="CC3", device="GPU", quiet=True, comm=comm)
        rng = np.random.default_rng(8)
        t1 = 0.05 * rng.standard_normal((no, nv))
        t2 = 0.05 * rng.standard_normal((no, no, nv, nv))
        m = rng.standard_normal(syn.F.shape)
        F = syn.F + 0.02 * (m + m.T)
        want1, want2 = c3.residuals(P, F, t1, t2, real_time=True)
        T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        r1, r2 = cc.residuals(T(F), T(t1), T(t2), real_time=True)
        d_rt = max(float(np.abs(r1.cpu().numpy() - want1).max()), float(np.abs(r2.cpu().numpy() - want2).max()))
        e_ref, _, t2_ref, trace = c3.solve_cc(P, 1e-11, 1e-11)
        e = float(cc.solve_cc(1e-11, 1e-11))
        d_t2 = float(np.abs(cc.t2.cpu().numpy() - t2_ref).max())
        # mixed precision on NCCL ranks: a-sharded TF32 planes, within 1e-6 Eh of the FP64 oracle
        e64, _, _, _ = co.solve_cc(P, 1e-9, 1e-9)
        mp_ = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True, comm=comm, precision="MP")
        e_mp = float(mp_.solve_cc(1e-8, 1e-8))
        q.put((rank, d_rt, abs(e - e_ref), d_t2, len(cc.trace), len(trace), abs(e_mp - e64)))
    finally:
        dist.destroy_process_group()


def test_nccl_cc3_and_mixed_precision():
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_worker_cc3, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, d_rt, de, dt, n, nref, dmp in res:
        assert d_rt < 1e-10 and de < 1e-10 and dt < 1e-9 and n == nref, (rank, d_rt, de, dt, n, nref)
        assert dmp < 1e-6, (rank, dmp)
