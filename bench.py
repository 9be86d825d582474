#!/usr/bin/env python
"""bench.py -- CCSD seconds/iteration and (T) FP64 TFLOP/s at o=40, v=300 on N B200s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this package (CUDA, FP64 DMMA)
    python bench.py --impl reference [...]                          # the reference's algorithm on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N    # N ranks, one per GPU, NCCL

A step = ONE full CCSD iteration of solve_cc (reference ccwfn.py:268-319): residuals (all intermediates,
r1, r2 incl. the v^4 ladder), the fused symmetrise/Jacobi-update/rms pass, the energy, and the DIIS
add + extrapolate -- on synthetic integrals of the named shape (SURVEY.md 8d recipe, seed 0) held in HBM.
Work is fixed as N grows (strong scaling): <ab|ef> is a-sharded, the other r2 terms are split over an
occupied index, one all-reduce of r2 per iteration (pycc_b200/parallel.py).

One JSON line on stdout (rank 0).  value = seconds per iteration (lower is better), timed with CUDA
events around exactly K steps after W warm-up steps, max over ranks.  Extra keys: roofline (the ladder
GEMM, the dominant kernel, timed live), cpu_baseline, e2e, t (the (T) rate), clocks, gpu_launches.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CCSD s/iter + (T) FP64 TFLOP/s at o=40 v=300, 1/2/4/8 B200 vs host CPU"
FP64_PEAK_FALLBACK = 36.18     # TFLOP/s, cuBLAS DGEMM 8192^3 measured on this pool's B200 (profiles/probe_r01_first.json)


def ccsd_flops(o, v, factorised=True):
    """Algorithmic flop of one CCSD iteration: ladder + o^3v^3 terms (7 when the two t1 x t1
    terms are factorised, 9 as written in the reference) + 2 x o^4v^2 + 8 x o^2v^3."""
    n33 = 7 if factorised else 9
    return 2 * o**2 * v**4 + n33 * 2 * o**3 * v**3 + 2 * 2 * o**4 * v**2 + 8 * 2 * o**2 * v**3


def t_flops_per_triple(o, v):
    return 12 * v**4 + 12 * o * v**3


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower() == "active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
def cpu_sample(o_s, v_s, o, v, threads, reps=1):
    """The oracle (numpy restatement of the reference algorithm) timed on the host cores for one full
    CCSD iteration at the reduced size (o_s, v_s), scaled to (o, v) by the reference's algorithmic flops."""
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    from oracle import ccsd_oracle as co
    from pycc_b200.synthetic import make_synthetic, blocks_from_factor
    syn = make_synthetic(o_s, v_s, seed=0)
    P = co.Problem(blocks_from_factor(syn), syn.F, o_s)
    t1, t2 = P.guess()
    diis = co.Diis(t1, t2, 8)
    times = []
    for _ in range(reps + 1):                       # first pass warms BLAS / einsum paths
        t0 = time.perf_counter()
        r1, r2 = P.residuals(syn.F, t1, t2)
        t1 = t1 + r1 / P.Dia
        t2 = t2 + r2 / P.Dijab
        P.cc_energy(syn.F, t1, t2)
        diis.add_error_vector(t1, t2)
        t1, t2 = diis.extrapolate(t1, t2)
        times.append(time.perf_counter() - t0)
    t_s = min(times[1:])
    scale = ccsd_flops(o, v, False) / ccsd_flops(o_s, v_s, False)
    sample = ("oracle (numpy port of ccwfn.py:321-372 + update/energy/DIIS) full CCSD iteration at o=%d,v=%d: %.3f s; "
              "scaled to o=%d,v=%d by reference algorithmic flops x%.1f (estimate)" % (o_s, v_s, t_s, o, v, scale))
    return t_s * scale, t_s, sample


def cpu_sample_fullsize(o, v, threads, na=None, budget_s=8.0):
    """Bounded sample of the SAME workload on the host cores: the reference's two dominant contraction shapes evaluated
    with the reference's own backend call (an exact einsum -> tensordot -> BLAS, device.py:84) on arrays of the real
    (o, v) shape:
      * the ladder 'ijef,abef->ijab' (ccwfn.py:931) on ``na`` of the v rows a of <ab|ef> (the full block is 64.8 GB);
      * ONE of the nine o^3v^3 contractions, 'imae,mbej->ijab' (ccwfn.py:933), at full size.
    The iteration is then  ladder_time * v/na + ring_time * (9 + the o^4v^2 / o^2v^3 terms at the ring's flop rate);
    HBM-bound passes (tau, update, DIIS) are not counted, which favours the CPU."""
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    rng = np.random.default_rng(0)
    # size both slices from a quick BLAS rate probe: ~budget_s/3 for the ladder rows, ~2 budget_s/3 for the ring columns
    a = rng.standard_normal((1536, 1536))
    a @ a
    t0 = time.perf_counter()
    a @ a
    rate = 2 * 1536**3 / max(time.perf_counter() - t0, 1e-4)
    if na is None:
        na = int(max(1, min(v, (budget_s / 3.0) * rate / (2.0 * o * o * v**3))))
    nj = int(max(1, min(o, (2.0 * budget_s / 3.0) * rate / (2.0 * o * o * v**3))))
    tau = rng.standard_normal((o, o, v, v))
    vslice = rng.standard_normal((na, v, v, v))
    t0 = time.perf_counter()
    np.einsum("ijef,abef->ijab", tau, vslice, optimize=True)
    t_lad = time.perf_counter() - t0
    del vslice
    W = rng.standard_normal((o, v, v, nj))
    t0 = time.perf_counter()
    np.einsum("imae,mbej->ijab", tau, W, optimize=True)
    t_ring = time.perf_counter() - t0
    del W, tau
    ring_fl = 2.0 * o**3 * v**3
    other = (2 * 2 * o**4 * v**2 + 8 * 2 * o**2 * v**3) / ring_fl
    est = t_lad * v / na + t_ring * (o / nj) * (9.0 + other)
    sample = ("same shapes as the workload, reference backend call (einsum->tensordot->BLAS): ladder ijef,abef->ijab on "
              "%d of %d rows a of <ab|ef>: %.2f s (%.0f GFLOP/s); one of the nine o^3v^3 terms imae,mbej->ijab on %d of "
              "%d columns j: %.2f s (%.0f GFLOP/s); iteration = ladder*v/na + ring*(o/nj)*(9 + %.2f for the "
              "o^4v^2/o^2v^3 terms); HBM-bound passes not counted"
              % (na, v, t_lad, 2.0 * o * o * na * v**3 / t_lad / 1e9, nj, o, t_ring,
                 ring_fl * nj / o / t_ring / 1e9, other))
    return est, t_lad + t_ring, sample


def run_reference(args):
    """--impl reference: the reference's own algorithm on the host cores (oracle port; the reference is
    pure Python + psi4 and cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    sample = ""
    nwarm = min(1, max(0, args.warmup))              # one untimed sample warms BLAS threads / page cache
    for _ in range(max(1, args.steps) + nwarm):
        est, t_s, sample = cpu_sample_fullsize(args.o, args.v, cores)
        vals.append(est)
    vals = vals[nwarm:]
    v = float(np.median(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "s/iter", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "RHF-CCSD iteration o=%d v=%d FP64 (synthetic integrals, seed 0); inputs >> L2 "
                                   "(75 GB of integrals streamed per step)" % (args.o, args.v),
                       "parallelism": "host cores, all BLAS threads"},
            "cpu_baseline": {"value": v, "unit": "s/iter", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "s/iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200cc", choices=["b200cc", "reference"])
    ap.add_argument("--o", type=int, default=40)
    ap.add_argument("--v", type=int, default=300)
    ap.add_argument("--cpu-o", type=int, default=16)
    ap.add_argument("--cpu-v", type=int, default=120)
    ap.add_argument("--t-triples", type=int, default=48, help="(T) sample: triples timed per rank-set")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-mp", action="store_true", help="skip the mixed-precision (precision='MP') leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import pycc_b200
    from pycc_b200 import kernels as K, cctriples
    from pycc_b200.synthetic import make_synthetic
    from pycc_b200.parallel import Comm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the host baseline")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the single JSON line: NCCL writes its banner ("NCCL version ...", printed from VERSION up,
        # i.e. also at WARN) and debug lines to stdout unless NCCL_DEBUG_FILE points elsewhere
        os.environ["NCCL_DEBUG"] = os.environ.get("B200CC_NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        comm = Comm()
    o, v = args.o, args.v

    def sync():
        if comm is not None:
            comm.barrier()
        torch.cuda.synchronize()

    # ---- problem: synthetic factor on the host (seeded), integral blocks contracted on the device
    t_setup = time.time()
    syn = make_synthetic(o, v, seed=0, device=dev)
    cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True, comm=comm)
    diis = pycc_b200.helper_diis(cc.t1, cc.t2, 8)
    sync()
    t_setup = time.time() - t_setup

    energies = []                                   # E after every iteration (warm-up included): the MP leg repeats them

    def step():
        ecc, rms = cc.iterate()
        cc.diis_step(diis, True)
        energies.append(ecc)
        return ecc, rms

    for _ in range(max(3, args.warmup)):
        step()
    # ---- timed region: exactly K steps; inputs (integral blocks, amplitudes) resident in HBM.
    # The working set of one step (>= 75 GB of integrals streamed) is far larger than the 126 MB L2.
    sampler = ClockSampler(local)
    sync()
    sampler.start()
    l0 = K.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        ecc, rms = step()
    ev1.record()
    sync()
    clocks = sampler.stop()
    launches = K.launch_count() - l0
    sec = ev0.elapsed_time(ev1) * 1e-3
    if comm is not None:
        sec = comm.all_reduce_max_scalar(sec)
    s_iter = sec / args.steps

    # ---- roofline of the dominant kernel, timed live: the ladder GEMM (this rank's a-slice)
    tau = K.build_tau(cc.t1, cc.t2)
    r2 = torch.zeros_like(cc.t2)
    a_lo, a_hi = cc.part.a_range(v)
    lad_flops = 2.0 * o * o * (a_hi - a_lo) * v * v * v
    cc._ladder(tau, r2)
    torch.cuda.synchronize()
    la, lb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nl = 3
    la.record()
    for _ in range(nl):
        cc._ladder(tau, r2)
    lb.record()
    torch.cuda.synchronize()
    t_lad = la.elapsed_time(lb) * 1e-3 / nl
    del tau, r2
    peak = FP64_PEAK_FALLBACK
    peak_src = "cuBLAS DGEMM 8192^3 measured on this pool (profiles/probe_r01_first.json); MEASURED_PEAKS.json has no FP64 entry"
    try:   # live re-measurement of the denominator on this very GPU (library GEMM, not on the product path)
        A = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        C = torch.empty_like(A)
        for _ in range(2):
            torch.matmul(A, A, out=C)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(5):
            pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            pa.record()
            torch.matmul(A, A, out=C)
            pb.record()
            torch.cuda.synchronize()
            best = max(best, 2.0 * 8192**3 / (pa.elapsed_time(pb) * 1e-3) / 1e12)
        peak = best
        peak_src = "cuBLAS DGEMM 8192^3, best of 5, measured live in this run (MEASURED_PEAKS.json has no FP64 entry)"
        del A, C
    except Exception:
        pass
    # DRAM bytes of the ladder launch from the committed ncu --set full capture (profiles/), if it is this shape
    traffic = None
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ladder_traffic.json")))
        key = "o%dv%d_n%d" % (o, v, world)
        if key in rec:
            traffic = rec[key]["dram_bytes"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "dgemm_kernel (ladder, ccwfn.py:931)", "achieved": lad_flops / t_lad / 1e12,
                "peak": peak, "unit": "TFLOP/s", "frac": lad_flops / t_lad / 1e12 / peak, "traffic": traffic,
                "algorithmic_bytes": 8.0 * ((a_hi - a_lo) * v ** 3 + 2 * o * o * v * v),
                "peak_source": peak_src, "launch_ms": t_lad * 1e3,
                "share_of_step": t_lad / s_iter,
                "whole_step_tflops_per_gpu": ccsd_flops(o, v) / world / s_iter / 1e12}

    # ---- (T): FP64 TFLOP/s on a bounded sample of (i>=j>=k) triples, sharded round-robin over the ranks
    trip = [t for t in cctriples.triples_list(o) if not (t[0] == t[1] == t[2])]
    nt = min(len(trip), args.t_triples * world)
    sample_trip = trip[:: max(1, len(trip) // nt)][:nt]
    cctriples.t_tjl(cc, sample_trip)                       # warm-up (allocates the Q workspace once)
    sync()
    ta, tb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ta.record()
    et = cctriples.t_tjl(cc, sample_trip)
    tb.record()
    sync()
    t_t = ta.elapsed_time(tb) * 1e-3
    if comm is not None:
        t_t = comm.all_reduce_max_scalar(t_t)
    t_rate = t_flops_per_triple(o, v) * len(sample_trip) / t_t / 1e12
    t_info = {"tflops": t_rate, "unit": "TFLOP/s (FP64, whole job)", "triples_timed": len(sample_trip),
              "triples_total": len(trip), "seconds": t_t, "full_t_seconds_est": t_t * len(trip) / len(sample_trip),
              "frac_of_fp64_peak_per_gpu": t_rate / world / peak, "e_t_sample": float(et)}

    # ---- e2e: the same step through the public API with HOST amplitudes: every step uploads t1, t2 (and F)
    # from pinned host memory, runs the iteration, and reads (ecc, rms) back
    h_t1 = cc.t1.cpu().pin_memory()
    h_t2 = cc.t2.cpu().pin_memory()
    h_F = cc.H.F.cpu().pin_memory()
    d_F = torch.empty_like(cc.H.F)
    e2e_steps = max(2, min(args.steps, 3))
    sync()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for _ in range(e2e_steps):
        cc.t1.copy_(h_t1, non_blocking=True)
        cc.t2.copy_(h_t2, non_blocking=True)
        d_F.copy_(h_F, non_blocking=True)
        ecc2, rms2 = cc.iterate(d_F)                     # returns host floats (D2H of 2 doubles)
        cc.diis_step(diis, True)
    eb.record()
    sync()
    t_e2e = ea.elapsed_time(eb) * 1e-3 / e2e_steps
    if comm is not None:
        t_e2e = comm.all_reduce_max_scalar(t_e2e)
    e2e = {"value": t_e2e, "unit": "s/iter", "h2d_bytes_per_step": int((h_t1.numel() + h_t2.numel() + h_F.numel()) * 8),
           "d2h_bytes_per_step": 16}

    # ---- BASELINE configs[4]: the same iterations with precision='MP' (split-TF32 contractions on tcgen05, FP64
    # accumulation) from the same starting guess: s/iter, |E_MP - E_FP64| after each iteration, and the ladder GEMM
    # against the TF32 tensor peak (cuBLAS TF32 SGEMM 8192^3 measured live).  The FP64 objects are released first
    # (<ab|ef> FP64 + its TF32 planes do not fit together at v=300).
    mp = None
    if not args.no_mp:
        n_dp = len(energies)
        del cc, diis, h_t1, h_t2, h_F, d_F
        cctriples._QCACHE.clear()
        import gc
        gc.collect()                                # the Hamiltonian <-> its ERI/L views form reference cycles
        torch.cuda.empty_cache()
        ccm = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", precision="MP", quiet=True, comm=comm)
        diism = pycc_b200.helper_diis(ccm.t1, ccm.t2, 8)
        e_mp = []

        def mstep():
            e, r = ccm.iterate()
            ccm.diis_step(diism, True)
            e_mp.append(e)

        for _ in range(max(3, args.warmup)):
            mstep()
        sync()
        l0 = K.launch_count()
        ma, mb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ma.record()
        for _ in range(args.steps):
            mstep()
        mb.record()
        sync()
        mp_launches = K.launch_count() - l0
        t_mp = ma.elapsed_time(mb) * 1e-3 / args.steps
        if comm is not None:
            t_mp = comm.all_reduce_max_scalar(t_mp)
        nmp = min(n_dp, len(e_mp))
        de = max(abs(a - b) for a, b in zip(energies[:nmp], e_mp[:nmp]))
        tau = K.build_tau(ccm.t1, ccm.t2)
        r2 = torch.zeros_like(ccm.t2)
        with K.mixed_mode(True):
            ccm._ladder(tau, r2)
            torch.cuda.synchronize()
            la, lb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            la.record()
            for _ in range(3):
                ccm._ladder(tau, r2)
            lb.record()
            torch.cuda.synchronize()
        t_lad_mp = la.elapsed_time(lb) * 1e-3 / 3
        del tau, r2
        # (T) in MP mode: the same sample of triples, t3 build on the split-TF32 kernel (two K segments per GEMM)
        cctriples.t_tjl(ccm, sample_trip)
        sync()
        tma_, tmb_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tma_.record()
        et_mp = cctriples.t_tjl(ccm, sample_trip)
        tmb_.record()
        sync()
        t_t_mp = tma_.elapsed_time(tmb_) * 1e-3
        if comm is not None:
            t_t_mp = comm.all_reduce_max_scalar(t_t_mp)
        tf32_peak, tf32_src = 1130.0, "nominal dense TF32 (no measured entry in MEASURED_PEAKS.json)"
        try:
            torch.backends.cuda.matmul.allow_tf32 = True
            A = torch.randn(8192, 8192, dtype=torch.float32, device=dev)
            C = torch.empty_like(A)
            for _ in range(2):
                torch.matmul(A, A, out=C)
            torch.cuda.synchronize()
            best = 0.0
            for _ in range(5):
                pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                pa.record()
                torch.matmul(A, A, out=C)
                pb.record()
                torch.cuda.synchronize()
                best = max(best, 2.0 * 8192**3 / (pa.elapsed_time(pb) * 1e-3) / 1e12)
            tf32_peak, tf32_src = best, "cuBLAS TF32 SGEMM 8192^3, best of 5, measured live in this run"
            torch.backends.cuda.matmul.allow_tf32 = False
            del A, C
        except Exception:
            pass
        mp = {"value": t_mp, "unit": "s/iter", "speedup_vs_fp64": s_iter / t_mp, "dtype": "tf32x3 products, f64 accumulate",
              "max_abs_dE_vs_fp64_same_iteration": de, "iterations_compared": nmp, "ecc_last": e_mp[-1],
              "gpu_launches": int(mp_launches), "stats": dict(K.MIXED.stats),
              "t": {"tflops_fp64_equivalent": t_flops_per_triple(o, v) * len(sample_trip) / t_t_mp / 1e12,
                    "seconds": t_t_mp, "speedup_vs_fp64": t_t / t_t_mp, "triples_timed": len(sample_trip),
                    "full_t_seconds_est": t_t_mp * len(trip) / len(sample_trip),
                    "e_t_sample": float(et_mp), "abs_dE_vs_fp64": abs(float(et_mp) - float(et))},
              "roofline": {"bound": "tensor", "kernel": "tf32x3_gemm_r_kernel (ladder, ccwfn.py:931)",
                           "achieved": 3.0 * lad_flops / t_lad_mp / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                           "frac": 3.0 * lad_flops / t_lad_mp / 1e12 / tf32_peak, "peak_source": tf32_src,
                           "fp64_equivalent_tflops": lad_flops / t_lad_mp / 1e12, "launch_ms": t_lad_mp * 1e3,
                           "note": "achieved counts the three TF32 products actually executed per FP64-equivalent flop"}}
        del ccm, diism

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:          # reported at N=1 only (rank 0)
            cores = os.cpu_count() or 1
            est, t_s, sample = cpu_sample_fullsize(o, v, cores)
            cpu = {"value": est, "unit": "s/iter", "cores": cores, "kind": "port", "sample": sample}
            # cross-check: the whole oracle iteration (residuals + update + energy + DIIS) at a reduced size, flop-scaled
            est_s, t_small, sample_s = cpu_sample(args.cpu_o, args.cpu_v, o, v, cores)
            cpu["cross_check"] = {"value": est_s, "sample": sample_s}
        line = {"metric": METRIC, "value": s_iter, "unit": "s/iter", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": s_iter * 1e3, "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "RHF-CCSD iteration o=%d v=%d FP64 (synthetic integrals, seed 0); inputs >> L2 "
                                       "(75 GB of integrals streamed per step)" % (o, v),
                           "parallelism": "a-sharded ladder + occupied-sliced ring terms, 1 all-reduce/iter" if world > 1 else "single GPU",
                           "diis": 8, "setup_s": t_setup},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "t": t_info, "mp": mp, "clocks": clocks,
                "gpu_launches": int(launches), "ecc_last": ecc, "rms_last": rms}
        print(json.dumps(line), flush=True)
    if comm is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
