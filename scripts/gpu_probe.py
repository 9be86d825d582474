#!/usr/bin/env python
"""One-shot B200 probe (run under gpurun): measured FP64 peaks (cuBLAS DGEMM via torch.matmul, the
roofline denominator MEASURED_PEAKS.json lacks), this package's DMMA GEMM on the same and on the
path's own shapes, and the HBM-bound kernels.  Writes gpurun_out/probe.json."""
import json
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycc_b200 import kernels as K   # noqa: E402

DEV = torch.device("cuda:0")
OUT = {}
CONFIGS = tuple(int(x) for x in os.environ.get("PROBE_CONFIGS", "2,4").split(","))


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    tot = 0.0
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = min(best, ms)
        tot += ms
    return best * 1e-3, tot / reps * 1e-3


def clocks():
    try:
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active"
        return subprocess.check_output(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader"], text=True).strip()
    except Exception as e:  # noqa
        return str(e)


def gemm_case(name, M, N, Kd, ta=0, tb=0, batch=1, cublas=True, reps=5):
    A = torch.randn((batch, Kd, M) if ta else (batch, M, Kd), dtype=torch.float64, device=DEV)
    B = torch.randn((batch, Kd, N) if tb else (batch, N, Kd), dtype=torch.float64, device=DEV)
    C = torch.empty((batch, M, N), dtype=torch.float64, device=DEV)
    flops = 2.0 * M * N * Kd * batch
    lda = M if ta else Kd
    ldb = N if tb else Kd
    r = {"M": M, "N": N, "K": Kd, "batch": batch, "ta": ta, "tb": tb}
    for cfg in CONFIGS:
        try:
            best, avg = timeit(lambda: K.dgemm(M, N, Kd, A, lda, ta, B, ldb, tb, C, N, batch=batch, sA=M * Kd,
                                               sB=N * Kd, sC=M * N, ksplit=1, config=cfg), reps=reps)
            r["b200cc_cfg%d_tflops" % cfg] = flops / best / 1e12
        except Exception as e:   # config not applicable to this operand layout
            r["b200cc_cfg%d_tflops" % cfg] = float("nan")
    r["clocks_after"] = clocks()
    if cublas:
        Am = A.transpose(1, 2) if ta else A
        Bm = B if tb else B.transpose(1, 2)
        ref = torch.matmul(Am, Bm)
        err = float((ref - C).abs().max() / ref.abs().max())
        best, avg = timeit(lambda: torch.matmul(Am, Bm, out=ref), reps=reps)
        r.update({"cublas_tflops_best": flops / best / 1e12, "cublas_tflops_avg": flops / avg / 1e12, "relerr": err})
    OUT[name] = r
    print(name, json.dumps(r), flush=True)
    del A, B, C


def main():
    OUT["device"] = torch.cuda.get_device_name(0)
    OUT["clocks_idle"] = clocks()
    print(OUT["device"], OUT["clocks_idle"], flush=True)
    gemm_case("square_4096", 4096, 4096, 4096)
    gemm_case("square_8192", 8192, 8192, 8192, reps=3)
    gemm_case("ladder_o20", 400, 22500, 22500, reps=3)                 # o=20,v=150 ladder, full
    gemm_case("ladder_o40_slice", 1600, 9000, 90000, reps=2)           # o=40,v=300 ladder, 30 of 300 a-rows
    gemm_case("ring_o20", 3000, 3000, 3000)
    gemm_case("ring_o40", 12000, 12000, 12000, reps=2)
    gemm_case("t_o30v280", 78400, 280, 310, ta=1, tb=0, batch=6, reps=3)
    gemm_case("wmnij_o40", 1600, 1600, 90000, reps=3)
    # sustained: back-to-back 8192^3 for ~3 s
    A = torch.randn(8192, 8192, dtype=torch.float64, device=DEV)
    B = torch.randn(8192, 8192, dtype=torch.float64, device=DEV)
    C = torch.empty(8192, 8192, dtype=torch.float64, device=DEV)
    for label, fn in (("b200cc", lambda: K.dgemm(8192, 8192, 8192, A, 8192, 0, B, 8192, 0, C, 8192, ksplit=1)),
                      ("cublas", lambda: torch.matmul(A, B.t(), out=C))):
        fn()
        torch.cuda.synchronize()
        t0 = time.time()
        n = 0
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        while time.time() - t0 < 3.0:
            fn()
            n += 1
            if n % 4 == 0:
                torch.cuda.synchronize()
        b.record()
        torch.cuda.synchronize()
        OUT["sustained_%s_tflops" % label] = 2.0 * 8192 ** 3 * n / (a.elapsed_time(b) * 1e-3) / 1e12
        OUT["sustained_%s_clocks" % label] = clocks()
        print(label, "sustained", OUT["sustained_%s_tflops" % label], OUT["sustained_%s_clocks" % label], flush=True)
    del A, B, C
    # HBM-bound kernels at o=40, v=300 (t2-sized tensors, 1.152 GB)
    no, nv = 40, 300
    t2 = torch.randn(no, no, nv, nv, dtype=torch.float64, device=DEV)
    t1 = torch.randn(no, nv, dtype=torch.float64, device=DEV)
    out = torch.empty_like(t2)
    nbytes = t2.numel() * 8
    hb = {}
    for name, perm in (("copy", (0, 1, 2, 3)), ("swap_ab", (0, 1, 3, 2)), ("iajb", (0, 2, 1, 3)), ("jbnf", (1, 3, 0, 2))):
        v = t2.permute(*perm)
        o2 = torch.empty(tuple(v.shape), dtype=torch.float64, device=DEV)
        best, _ = timeit(lambda: K.strided_axpby(o2, v, 1.0, 0.0))
        hb["permute_" + name] = 2 * nbytes / best / 1e9
    best, _ = timeit(lambda: K.build_tau(t1, t2, 1.0, 1.0, out=out))
    hb["tau"] = 2 * nbytes / best / 1e9
    eo = -torch.rand(no, dtype=torch.float64, device=DEV) - 0.5
    ev = torch.rand(nv, dtype=torch.float64, device=DEV) + 0.5
    r2 = torch.randn_like(t2) * 1e-6
    best, _ = timeit(lambda: K.update_amps(t1, r2, eo, ev, t1, t2, True, False))
    hb["sym_update"] = 3 * nbytes / best / 1e9
    best, _ = timeit(lambda: K.cc_energy(t1, t1, t2, out))
    hb["energy"] = 2 * nbytes / best / 1e9
    xs = [t2.view(-1), out.view(-1), r2.view(-1)]
    best, _ = timeit(lambda: K.multi_dot(xs[0], xs))
    hb["multi_dot3"] = 3 * nbytes / best / 1e9
    best, _ = timeit(lambda: torch.add(t2, r2, out=out))
    hb["torch_add_ref"] = 3 * nbytes / best / 1e9
    best, _ = timeit(lambda: out.copy_(t2))
    hb["torch_copy_ref"] = 2 * nbytes / best / 1e9
    OUT["hbm_gbs"] = hb
    print(json.dumps(hb), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(OUT, open("gpurun_out/probe.json", "w"), indent=1)


if __name__ == "__main__":
    main()
