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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# BLAS thread count for the host legs (--impl reference, cpu_baseline) -- fixed BEFORE numpy / torch are imported, and
# regardless of torchrun, which exports OMP_NUM_THREADS=1 to every rank: the reference arm always gets all host cores.
_REFERENCE_ARM = any(a == "reference" or a == "--impl=reference" for a in sys.argv[1:])
if _REFERENCE_ARM or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(host_cores())

import numpy as np  # noqa: E402

METRIC = "CCSD s/iter + (T) FP64 TFLOP/s at o=40 v=300, 1/2/4/8 B200 vs host CPU"
FP64_PEAK_FALLBACK = 36.18     # TFLOP/s, cuBLAS DGEMM 8192^3 measured on this pool's B200 (profiles/probe_r01_first.json)


def ccsd_flops(o, v, factorised=True):
    """Algorithmic flop of one CCSD iteration: ladder + o^3v^3 terms (7 when the two t1 x t1
    terms are factorised, 9 as written in the reference) + 2 x o^4v^2 + 8 x o^2v^3."""
    n33 = 7 if factorised else 9
    return 2 * o**2 * v**4 + n33 * 2 * o**3 * v**3 + 2 * 2 * o**4 * v**2 + 8 * 2 * o**2 * v**3


def ccsd_flops_executed(o, v):
    """Flop one iteration of THIS package executes (solve_cc path): the ladder in pair form on rows (i >= j) --
    2 x 2 x o(o+1)/2 x (v(v+1)/2)^2 -- plus seven o^3v^3 terms, two o^4v^2 and eight o^2v^3."""
    nq, m = v * (v + 1) // 2, o * (o + 1) // 2
    return 2 * 2 * m * nq * nq + 7 * 2 * o**3 * v**3 + 2 * 2 * o**4 * v**2 + 8 * 2 * o**2 * v**3


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
def blas_threads(n):
    """Pin the BLAS pool to n threads at run time as well (threadpoolctl) and report what the pools say."""
    info = []
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=int(n))
        info = [{"api": d.get("internal_api"), "threads": d.get("num_threads")} for d in threadpoolctl.threadpool_info()]
    except Exception as exc:                              # the env variables above still apply
        info = [{"api": "threadpoolctl unavailable: %s" % exc, "threads": None}]
    return info


FULLSIZE_RECORD = os.path.join(ROOT, "profiles", "reference_fullsize_r02.json")


def reference_iterations(o_s, v_s, n_warm, n_timed, log=None):
    """Real CCSD iterations of the reference's own CPU path on the host cores at (o_s, v_s).

    With the reference available (baseline/_ref, see baseline/install_ref.py): its UNMODIFIED ``CCwfn.solve_cc`` loop --
    ``residuals`` (ccwfn.py:321-372), Jacobi update + rms (281-284), ``cc_energy`` (286), ``helper_diis`` (317-319) -- on
    integrals held in the memory layout its Hamiltonian uses, timed per iteration from its own prints.  Without it
    (kind 'port'): the numpy oracle, which restates the same equations.  Returns (seconds per timed iteration, kind,
    energies)."""
    from pycc_b200.synthetic import make_synthetic
    syn = make_synthetic(o_s, v_s, seed=0)
    n = n_warm + n_timed
    try:
        from baseline import refload
        ref = refload.load_reference()
    except ImportError:
        ref = None
    if ref is not None:
        w = refload.reference_wfn(ref, syn.F, syn.no, None, blocks=refload.host_blocks(syn, log=log))
        secs, en = refload.timed_solve_cc(w, n, echo=log)
        return secs[n_warm:], "reference", en
    from oracle import ccsd_oracle as co
    from pycc_b200.synthetic import blocks_from_factor
    P = co.Problem(blocks_from_factor(syn), syn.F, o_s)
    t1, t2 = P.guess()
    diis = co.Diis(t1, t2, 8)
    secs, en = [], []
    for _ in range(n):
        t0 = time.perf_counter()
        r1, r2 = P.residuals(syn.F, t1, t2)
        t1 = t1 + r1 / P.Dia
        t2 = t2 + r2 / P.Dijab
        en.append(float(P.cc_energy(syn.F, t1, t2)))
        diis.add_error_vector(t1, t2)
        t1, t2 = diis.extrapolate(t1, t2)
        secs.append(time.perf_counter() - t0)
    return secs[n_warm:], "port", en


def reference_scale(o_s, v_s, o, v):
    """Factor from a measured (o_s, v_s) iteration of the reference to one at (o, v), and where it comes from.
    Preferred: the ratio MEASURED on this pool's host by running the reference itself once at both sizes
    (profiles/reference_fullsize_r02.json, scripts/reference_fullsize.py).  Fallback: the reference's algorithmic flop
    count (9 o^3v^3 terms as written), which ignores that its transposition copies grow more slowly than its flops."""
    flop = ccsd_flops(o, v, False) / ccsd_flops(o_s, v_s, False)
    try:
        rec = json.load(open(FULLSIZE_RECORD))
        if (rec["o"], rec["v"], rec["o_s"], rec["v_s"]) == (o, v, o_s, v_s):
            return float(rec["seconds_full"]) / float(rec["seconds_small"]), (
                "ratio measured with the reference itself on this pool's host: one real iteration at o=%d,v=%d took %.1f s, "
                "at o=%d,v=%d %.2f s (%d cores, %s)" % (o, v, rec["seconds_full"], o_s, v_s, rec["seconds_small"],
                                                        rec["cores"], os.path.relpath(FULLSIZE_RECORD, ROOT)))
    except Exception:
        pass
    return flop, "the reference's algorithmic flop ratio (9 o^3v^3 terms), no full-size measurement on record"


def cpu_leg(args, n_warm, n_timed, log=None):
    """The host-CPU measurement shared by --impl reference and the cpu_baseline key: real iterations of the reference
    at BASELINE configs[1] (o=20, v=150 -- the largest config it finishes in seconds per iteration), scaled to the
    bench workload (o, v).  A full (o=40, v=300) iteration of the reference needs ~165 GB of host memory and minutes of
    strided copies per step, so it cannot be repeated --steps times; it is measured once by scripts/reference_fullsize.py
    and its record calibrates the scale factor."""
    cores = host_cores()
    pools = blas_threads(cores)
    o_s, v_s = args.cpu_o, args.cpu_v
    secs, kind, en = reference_iterations(o_s, v_s, n_warm, n_timed, log)
    t_small = float(np.median(secs))
    scale, how = reference_scale(o_s, v_s, args.o, args.v)
    sample = ("%d real CCSD iterations (%d warm-up) of %s at o=%d,v=%d (BASELINE configs[1]) on %d host cores: median "
              "%.3f s/iter (min %.3f, max %.3f); value = median x %.2f -- %s [estimate for o=%d,v=%d]"
              % (n_timed, n_warm,
                 "the unmodified reference (baseline/_ref: CCwfn.solve_cc loop = residuals + update + cc_energy + helper_diis)"
                 if kind == "reference" else "the numpy oracle port of the reference (baseline/_ref absent)",
                 o_s, v_s, cores, t_small, min(secs), max(secs), scale, how, args.o, args.v))
    return {"value": t_small * scale, "unit": "s/iter", "cores": cores, "blas_threads": pools, "kind": kind,
            "estimate": True, "measured_small": {"o": o_s, "v": v_s, "s_per_iter": t_small, "iterations": n_timed,
                                                 "all": [round(x, 4) for x in secs], "ecc_last": en[-1]},
            "scale": scale, "sample": sample}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores, all BLAS threads.
    Under torchrun rank 0 alone runs it; the other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    log = (lambda m: print(m, file=sys.stderr, flush=True)) if args.verbose else None
    t0 = time.time()
    cpu = cpu_leg(args, max(1, args.warmup), max(1, args.steps), log)
    v = cpu["value"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "s/iter", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.o, args.v, "host cores, all BLAS threads"),
            "cpu_baseline": cpu, "wall_s": time.time() - t0,
            "e2e": {"value": v, "unit": "s/iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gpu_library_leg(o, v, dev, log=None):
    """The same-box LIBRARY bar ('GPU-side comparison bar'): the reference's own, unmodified CCSD iteration
    (baseline/_ref: residuals as written -- nine o^3v^3 and the dense o^2v^4 ladder --, update, energy, helper_diis) on
    torch tensors, i.e. torch.einsum -> cuBLAS DGEMM + ATen elementwise kernels, with every integral block ALREADY
    RESIDENT on the GPU -- the reference's device='GPU' path minus the per-call host->device copies of the ERI slices it
    would pay (device.py:70-74: 64.8 GB for <ab|ef> per iteration over PCIe).  What a library formulation of the step
    reaches on this very GPU.  None of this package's kernels run in the timed part (the blocks are generated by them
    beforehand)."""
    import gc
    import torch
    from baseline import refload
    from pycc_b200.hamiltonian import BlockHamiltonian
    from pycc_b200.synthetic import make_synthetic
    ref = refload.load_reference()
    syn = make_synthetic(o, v, seed=0, device=dev)
    keep = BlockHamiltonian.keep_vvvv_bytes
    BlockHamiltonian.keep_vvvv_bytes = 1 << 50            # the library path needs the full FP64 <ab|ef> block
    try:
        H = BlockHamiltonian.from_factor(syn, dev)
    finally:
        BlockHamiltonian.keep_vvvv_bytes = keep
    blocks = {k: H.block(k) for k in ("oooo", "ooov", "oovv", "ovov", "ovvv", "vvvv")}
    del H
    gc.collect()
    torch.cuda.empty_cache()
    w = refload.reference_wfn(ref, syn.F, syn.no, None, blocks=blocks, device="GPU")
    assert w.t2.is_cuda and w.device1.type == "cuda"
    torch.cuda.synchronize()

    def echo(line):                                        # device work is asynchronous: settle it at every print
        torch.cuda.synchronize()
        if log is not None:
            log(line)
    secs, en = refload.timed_solve_cc(w, 3, echo=echo)
    peak = torch.cuda.max_memory_allocated(dev) / 2**30
    return {"value": float(min(secs[1:])), "unit": "s/iter", "all": [round(x, 4) for x in secs], "energies": en,
            "peak_gb": peak,
            "what": "unmodified reference CCSD iteration (9 o^3v^3 terms + dense 2 o^2 v^4 ladder) on torch CUDA tensors: "
                    "torch.einsum -> cuBLAS DGEMM, integral blocks resident in HBM (no per-call H2D copies)"}


def workload_config(o, v, parallelism):
    return {"workload": "RHF-CCSD iteration o=%d v=%d FP64 (synthetic integrals, seed 0); inputs >> L2 "
                        "(integral blocks streamed from HBM every step)" % (o, v), "parallelism": parallelism}


# ------------------------------------------------------------------------------------------------
def cuda_time(fn, reps=1, sync=None):
    """Seconds per call of fn(), CUDA events on the current stream, synchronised on both sides."""
    import torch
    (sync or torch.cuda.synchronize)()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    (sync or torch.cuda.synchronize)()
    return a.elapsed_time(b) * 1e-3 / reps, out


def library_peak(dev, dtype, allow_tf32=False):
    """cuBLAS GEMM 8192^3, best of 5: the denominator of the tensor rooflines (not on the product path)."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
    try:
        A = torch.randn(8192, 8192, dtype=dtype, device=dev)
        C = torch.empty_like(A)
        for _ in range(2):
            torch.matmul(A, A, out=C)
        best = 0.0
        for _ in range(5):
            t, _ = cuda_time(lambda: torch.matmul(A, A, out=C))
            best = max(best, 2.0 * 8192**3 / t / 1e12)
        return best
    finally:
        torch.backends.cuda.matmul.allow_tf32 = False


PARITY_RECORD = os.path.join(ROOT, "profiles", "parity_r02.json")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200cc", choices=["b200cc", "reference"])
    ap.add_argument("--o", type=int, default=40)
    ap.add_argument("--v", type=int, default=300)
    ap.add_argument("--cpu-o", type=int, default=20, help="size of the real reference iterations (configs[1])")
    ap.add_argument("--cpu-v", type=int, default=150)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--t-triples", type=int, default=0, help="(T): time only this many triples per rank (0 = the full job)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-mp", action="store_true", help="skip the mixed-precision (precision='MP') leg")
    ap.add_argument("--t-compare", action="store_true", help="(T): also time the (i,j,k)-driven formulation of the same job")
    ap.add_argument("--no-c4", action="store_true", help="skip BASELINE configs[3]: the full (T) at o=30, v=280")
    ap.add_argument("--no-lib", action="store_true", help="skip the same-box library bar (reference code on torch/cuBLAS)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import gc
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
        # NCCL_DEBUG / NCCL_DEBUG_FILE are the caller's (the driver reads NCCL's own log to count ranks).  Only when
        # nobody set them: keep stdout to the single JSON line -- NCCL prints its banner and debug lines to stdout
        # unless NCCL_DEBUG_FILE points elsewhere.
        if "NCCL_DEBUG" not in os.environ:
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        comm = Comm()
    o, v = args.o, args.v
    nwarm = max(3, args.warmup)

    def sync():
        if comm is not None:
            comm.barrier()
        torch.cuda.synchronize()

    def rmax(x):
        return comm.all_reduce_max_scalar(x) if comm is not None else x

    def release():
        cctriples._QCACHE.clear()
        gc.collect()                                # the Hamiltonian <-> its ERI/L views form reference cycles
        torch.cuda.empty_cache()

    # ---- problem: synthetic factor on the host (seeded), integral blocks contracted on the device
    t_setup = time.time()
    syn = make_synthetic(o, v, seed=0, device=dev)
    cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True, comm=comm)
    diis = cc.make_diis(8)
    e_mp2 = float(cc.cc_energy(cc.o, cc.v, cc.H.F, cc.H.L, cc.t1, cc.t2))
    sync()
    t_setup = time.time() - t_setup

    energies = []                                   # E after every iteration (warm-up included): the MP leg repeats them

    def step():
        ecc, rms = cc.iterate()
        cc.diis_step(diis, True)
        energies.append(ecc)
        return ecc, rms

    for _ in range(nwarm):
        step()
    # ---- timed region: exactly K steps; inputs (integral blocks, amplitudes) resident in HBM.
    # The working set of one step (> 40 GB of integrals streamed) is far larger than the 126 MB L2.
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
    s_iter = rmax(ev0.elapsed_time(ev1) * 1e-3) / args.steps

    # ---- where the step goes on this rank (rank 0): two more steps bracketed phase by phase with CUDA events
    K.PHASES.on = True
    K.PHASES.collect()
    for _ in range(2):
        step()
    phases = {k.strip(): {"calls_per_step": c / 2.0, "ms_per_step": ms / 2.0, "nested": k.startswith("  ")}
              for k, (c, ms) in K.PHASES.collect().items()}
    K.PHASES.on = False

    # ---- parity at the FULL size, in this very run: the first iterations against the energies the unmodified
    # reference printed when it was run once at this size on this pool's host (profiles/reference_fullsize_r02.json)
    parity = {"e_mp2": e_mp2, "ecc_iter1": energies[0], "ecc_iter2": energies[1]}
    try:
        rec = json.load(open(FULLSIZE_RECORD))
        if (rec["o"], rec["v"]) == (o, v):
            want = rec["energies_full"]
            devs = [abs(e_mp2 - want[0])] + [abs(a - b) for a, b in zip(energies, want[1:])]
            parity.update({"reference_energies": want, "max_abs_dE_vs_reference": max(devs), "tolerance": 1e-10,
                           "ok": bool(max(devs) < 1e-10),
                           "what": "MP2 + first %d CCSD iteration energies of the unmodified reference at this size" % (len(want) - 1)})
            if not parity["ok"]:
                raise SystemExit("PARITY FAILURE vs the reference at o=%d,v=%d: %r" % (o, v, parity))
    except (OSError, KeyError, ValueError):
        pass

    # ---- roofline of the dominant kernel, timed live: the ladder GEMM (this rank's share of the pair rows)
    tau = K.build_tau(cc.t1, cc.t2)
    r2 = torch.zeros_like(cc.t2)
    cc._ladder(tau, r2, symmetric=True)
    t_lad, _ = cuda_time(lambda: cc._ladder(tau, r2, symmetric=True), 3)
    lad_exec = cc.ladder_flops
    a_lo, a_hi = cc.part.a_range(v)
    npl, nq = K.pair_count(a_hi) - K.pair_count(a_lo), K.pair_count(v)
    M_tri = K.pair_count(o)
    # the GEMM alone (same operands, same launch as inside _ladder)
    T = K.pack_tau(tau, True)
    V, ldv = cc.H.packed()
    lds = (npl + 1) // 2 * 2
    SA = torch.empty((2, M_tri, lds), dtype=torch.float64, device=dev)
    row0 = K.pair_count(a_lo) - K.pair_count(cc.H.a_range[0])

    def ladder_gemm():
        K.dgemm(M_tri, npl, nq, T, T.shape[2], 0, (V, row0 * ldv), ldv, 0, SA, lds, 1.0, 0.0, batch=2,
                sA=M_tri * T.shape[2], sB=cc.H.npairs_local * ldv, sC=M_tri * lds)
    ladder_gemm()
    t_gemm, _ = cuda_time(ladder_gemm, 3)
    del tau, r2, T, SA
    try:
        peak = library_peak(dev, torch.float64)
        peak_src = "cuBLAS DGEMM 8192^3, best of 5, measured live in this run (MEASURED_PEAKS.json has no FP64 entry)"
    except Exception:
        peak, peak_src = FP64_PEAK_FALLBACK, "cuBLAS DGEMM 8192^3 measured on this pool (profiles/probe_r01_first.json)"
    traffic = None
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ladder_traffic.json")))
        key = "r02_o%dv%d_n%d" % (o, v, world)
        if key in rec:
            traffic = rec[key]["dram_bytes"]
    except Exception:
        pass
    lad_dense = 2.0 * o * o * v * v * (2.0 * npl)          # 2 o^2 v^2 x (this rank's share of the v^2 (a,b) columns)
    roofline = {"bound": "tensor", "kernel": "dgemm_tma_kernel: the ladder as S = T+ V+^T, A = T- V-^T (ccwfn.py:931 in pair form)",
                "achieved": lad_exec / t_gemm / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": lad_exec / t_gemm / 1e12 / peak,
                "traffic": traffic, "algorithmic_bytes": 8.0 * 2 * (npl * nq + 2 * M_tri * nq),
                "executed_flop": lad_exec, "dense_equivalent_flop": lad_dense,
                "dense_equivalent_tflops": lad_dense / t_lad / 1e12,
                "peak_source": peak_src, "launch_ms": t_gemm * 1e3, "ladder_ms_with_pack_unpack": t_lad * 1e3,
                "share_of_step": t_gemm / s_iter,
                "whole_step_tflops_per_gpu": ccsd_flops_executed(o, v) / world / s_iter / 1e12,
                "note": "achieved counts EXECUTED flop (pair-packed: 1/4 of the dense 2 o^2 v^4); dense_equivalent_* "
                        "is the reference's flop count for the same term over the time of the whole ladder"}

    # ---- (T): the FULL job at this size, all (i>=j>=k) triples dealt round-robin to the ranks
    def t_job(wfn, nsample, force_ijk=False):
        trip = [t for t in cctriples.triples_list(wfn.no) if not (t[0] == t[1] == t[2])]
        if cctriples.fused_selected(wfn) and not force_ijk:
            # fused (a,b,c)-driven kernel (csrc/triples_abc.cu): the units are the virtual triples a >= b >= c, dealt
            # round-robin to the ranks; --t-triples R times R rounds of one (a,b,c) per SM and rank from the middle of the list
            lst = cctriples.abc_list(wfn.nv)
            n = K.NSM * world
            run = lst if not nsample else lst[lst.size // 2: lst.size // 2 + nsample * n]
            cctriples.t_tjl_abc(wfn, run[:n])                      # warm-up
            t_t, et = cuda_time(lambda: cctriples.t_tjl_abc(wfn, run), 1, sync)
            t_t = rmax(t_t)
            rate = cctriples.t_flops_abc(wfn.no, wfn.nv, len(run)) / t_t / 1e12
            full = len(run) == lst.size
            return {"o": wfn.no, "v": wfn.nv, "tflops": rate, "unit": "TFLOP/s (FP64 executed, whole job over all GPUs)",
                    "algorithm": "fused (a,b,c)-driven: t3 tile kept on chip (cctriples.py:75-105 + 208-237 in exchanged roles)",
                    "triples_timed": len(trip) if full else 0, "triples_total": len(trip),
                    "abc_timed": int(len(run)), "abc_total": int(lst.size), "seconds": t_t,
                    "frac_of_fp64_peak_per_gpu": rate / world / peak, "e_t": float(et)}, run
        run = trip if not nsample else trip[:: max(1, len(trip) // (nsample * world))][:nsample * world]
        cctriples.t_tjl(wfn, run[:2 * world])                      # warm-up (allocates the Q workspace once)
        t_t, et = cuda_time(lambda: cctriples.t_tjl(wfn, run), 1, sync)
        t_t = rmax(t_t)
        rate = t_flops_per_triple(wfn.no, wfn.nv) * len(run) / t_t / 1e12
        return {"o": wfn.no, "v": wfn.nv, "tflops": rate, "unit": "TFLOP/s (FP64, whole job over all GPUs)",
                "algorithm": "(i,j,k)-driven: paired GEMMs + energy kernel, t3 through HBM once",
                "triples_timed": len(run), "triples_total": len(trip), "seconds": t_t,
                "frac_of_fp64_peak_per_gpu": rate / world / peak, "e_t": float(et)}, run

    t_info, t_run = t_job(cc, args.t_triples)
    if args.t_compare and "abc_total" in t_info:
        t_info["ijk_driven"] = t_job(cc, args.t_triples, force_ijk=True)[0]
    t_info["amplitudes"] = "after %d CCSD iterations" % len(energies)

    # ---- e2e: the same step through the public API with HOST amplitudes: every step uploads t1, t2 (and F)
    # from pinned host memory, runs the iteration, and reads (ecc, rms) back
    h_t1 = cc.t1.cpu().pin_memory()
    h_t2 = cc.t2.cpu().pin_memory()
    h_F = cc.H.F.cpu().pin_memory()
    d_F = torch.empty_like(cc.H.F)
    e2e_steps = max(2, min(args.steps, 3))

    i0, i1 = cc.part.occ_range(o)

    def e2e_step():
        cc._gather_rows()
        cc.t1.copy_(h_t1, non_blocking=True)
        if world > 1:       # each rank uploads ITS rows of t2 over PCIe, the rest arrives over NVLink
            cc.t2[i0:i1].copy_(h_t2[i0:i1], non_blocking=True)
            cc.part.all_gather_rows(cc.t2)
        else:
            cc.t2.copy_(h_t2, non_blocking=True)
        d_F.copy_(h_F, non_blocking=True)
        out = cc.iterate(d_F)                              # returns host floats (D2H of 2 doubles)
        cc.diis_step(diis, True)
        return out
    t_e2e, _ = cuda_time(e2e_step, e2e_steps, sync)
    t_e2e = rmax(t_e2e)
    e2e = {"value": t_e2e, "unit": "s/iter",
           "h2d_bytes_per_step": int((h_t1.numel() + h_t2[i0:i1].numel() + h_F.numel()) * 8), "d2h_bytes_per_step": 16,
           "note": "per rank: t1, F and this rank's rows of t2 from pinned host memory%s; (ecc, rms) read back"
                   % (" + all-gather of the rows over NVLink" if world > 1 else "")}
    n_dp = len(energies)
    del cc, diis, h_t1, h_t2, h_F, d_F, V
    release()

    # ---- BASELINE configs[3]: RHF-CCSD(T) o=30 v=280, the full (T) job sharded over (i,j,k) on all ranks; the
    # amplitudes are those after exactly three CCSD iterations (the (T) cost does not depend on convergence), so E(T)
    # is a fixed number that every N must reproduce
    c4 = None
    if not args.no_c4:
        syn4 = make_synthetic(30, 280, seed=0, device=dev)
        cc4 = pycc_b200.ccwfn(syn4, model="CCSD(T)", device="GPU", quiet=True, comm=comm)
        d4 = cc4.make_diis(8)
        for _ in range(3):
            e4, _ = cc4.iterate()
            cc4.diis_step(d4, True)
        del d4
        c4, _ = t_job(cc4, args.t_triples)
        c4["ecc_after_3_iterations"] = e4
        del cc4, syn4
        release()

    # ---- driver-visible multi-GPU parity: energies that every N must reproduce, recorded at N = 1
    # (profiles/parity_r02.json, written by `bench.py --record-parity` runs at N=1; compared here at any N)
    checks = {}
    try:
        rec = json.load(open(PARITY_RECORD))
        key = "o%dv%d" % (o, v)
        if key in rec:
            for k_it, e_ref in rec[key].get("ecc_by_iteration", {}).items():
                if int(k_it) <= len(energies):
                    checks["ecc_iter%s" % k_it] = abs(energies[int(k_it) - 1] - e_ref)
            e_t_ref = rec[key].get("e_t_by_iteration", {}).get(str(n_dp))
            if e_t_ref is not None and t_info["triples_timed"] == t_info["triples_total"]:
                checks["e_t_full_o%dv%d" % (o, v)] = abs(t_info["e_t"] - e_t_ref)
        if c4 is not None and "o30v280" in rec and c4["triples_timed"] == c4["triples_total"]:
            checks["c4_ecc_after_3_iterations"] = abs(c4["ecc_after_3_iterations"] - rec["o30v280"]["ecc_after_3_iterations"])
            checks["c4_e_t_full"] = abs(c4["e_t"] - rec["o30v280"]["e_t"])
    except (OSError, ValueError, KeyError):
        pass
    parity["vs_n1_record"] = {"max_abs_dE": max(checks.values()) if checks else None, "checks": checks, "tolerance": 1e-10}
    if checks and max(checks.values()) >= 1e-10:
        raise SystemExit("PARITY FAILURE vs the N=1 record (profiles/parity_r02.json): %r" % checks)

    # ---- BASELINE configs[4]: the same iterations with precision='MP' (split-TF32 contractions on tcgen05, FP64
    # accumulation) from the same starting guess: s/iter, |E_MP - E_FP64| after each iteration, and the ladder GEMM
    # against the TF32 tensor peak (cuBLAS TF32 SGEMM 8192^3 measured live).
    mp = None
    if not args.no_mp:
        ccm = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", precision="MP", quiet=True, comm=comm)
        diism = ccm.make_diis(8)
        e_mp = []

        def mstep():
            e, r = ccm.iterate()
            ccm.diis_step(diism, True)
            e_mp.append(e)

        # warm up until the DIIS history is full: while it grows every iteration allocates two more t2-sized vectors, and in
        # this mode (FP64 blocks + cached TF32 planes) the caching allocator has to release and re-map memory for them
        for _ in range(max(nwarm, 9)):
            mstep()
        l0 = K.launch_count()
        t_mp, _ = cuda_time(mstep, args.steps, sync)
        t_mp = rmax(t_mp)
        mp_launches = K.launch_count() - l0
        nmp = min(n_dp, len(e_mp))
        de = max(abs(a - b) for a, b in zip(energies[:nmp], e_mp[:nmp]))
        tau = K.build_tau(ccm.t1, ccm.t2)
        r2 = torch.zeros_like(ccm.t2)
        with K.mixed_mode(True):
            ccm._ladder(tau, r2, symmetric=True)
            t_lad_mp, _ = cuda_time(lambda: ccm._ladder(tau, r2, symmetric=True), 3)
        mp_exec = ccm.ladder_flops
        del tau, r2
        # (T) in MP mode: a bounded sample of the triples, t3 build on the split-TF32 kernel (two K segments per GEMM)
        mt, _ = t_job(ccm, args.t_triples or 48)
        try:
            tf32_peak, tf32_src = library_peak(dev, torch.float32, True), "cuBLAS TF32 SGEMM 8192^3, best of 5, measured live in this run"
        except Exception:
            tf32_peak, tf32_src = 1130.0, "nominal dense TF32 (no measured entry in MEASURED_PEAKS.json)"
        mp = {"value": t_mp, "unit": "s/iter", "speedup_vs_fp64": s_iter / t_mp, "dtype": "tf32x3 products, f64 accumulate",
              "max_abs_dE_vs_fp64_same_iteration": de, "iterations_compared": nmp, "ecc_last": e_mp[-1],
              "gpu_launches": int(mp_launches), "stats": dict(K.MIXED.stats),
              "t": {"tflops_fp64_equivalent": mt["tflops"], "seconds": mt["seconds"], "triples_timed": mt["triples_timed"],
                    "triples_total": mt["triples_total"], "e_t_sample": mt["e_t"]},
              "roofline": {"bound": "tensor", "kernel": "tf32x3_gemm_r_kernel (ladder in pair form, ccwfn.py:931)",
                           "achieved": 3.0 * mp_exec / t_lad_mp / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                           "frac": 3.0 * mp_exec / t_lad_mp / 1e12 / tf32_peak, "peak_source": tf32_src,
                           "fp64_equivalent_tflops": mp_exec / t_lad_mp / 1e12,
                           "dense_equivalent_tflops": lad_dense / t_lad_mp / 1e12, "launch_ms": t_lad_mp * 1e3,
                           "note": "achieved counts the three TF32 products actually executed per FP64-equivalent flop; "
                                   "launch_ms includes the pack / split / unpack passes around the GEMM"}}
        del ccm, diism
        release()

    lib = None
    if world == 1 and not args.no_lib:
        try:
            lib = gpu_library_leg(o, v, dev)
            lib["speedup_of_this_package"] = lib["value"] / s_iter
        except Exception as exc:                    # e.g. baseline/_ref absent, or out of memory
            lib = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
        release()

    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:          # reported at N=1 only (rank 0)
            cpu = cpu_leg(args, 1, 2)               # ~3 real reference iterations at configs[1]: 10-30 s of CPU work
        cfg = workload_config(o, v, "pair-packed ladder a-sharded by pair count + occupied-sliced ring terms, 1 all-reduce/iter"
                              if world > 1 else "single GPU")
        cfg.update({"diis": 8, "setup_s": t_setup})
        line = {"metric": METRIC, "value": s_iter, "unit": "s/iter", "n_gpus": world, "steps": args.steps,
                "warmup": nwarm, "ms_per_step": s_iter * 1e3, "higher_is_better": False,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg, "roofline": roofline, "cpu_baseline": cpu, "gpu_library_baseline": lib, "e2e": e2e, "t": t_info, "t_c4": c4, "mp": mp,
                "parity": parity, "phases": phases, "clocks": clocks, "gpu_launches": int(launches), "ecc_last": ecc, "rms_last": rms,
                "energies": energies}
        print(json.dumps(line), flush=True)
    if comm is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
