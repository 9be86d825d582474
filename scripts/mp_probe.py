"""GPU probe of the split-TF32 tcgen05 GEMM: throughput per config and accuracy vs kchunk on ladder / ring shapes.
Writes gpurun_out/mp_probe.json.  Run: python scripts/mp_probe.py [quick]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycc_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
out = {}


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def run(name, M, N, Kd, configs, kchunks, locksteps=(0,)):
    g = torch.Generator(device="cpu").manual_seed(1)
    A = (torch.randn(M, Kd, dtype=torch.float64, generator=g) * 0.05).to(dev)
    B = torch.randn(N, Kd, dtype=torch.float64, device=dev) * 0.05
    ref = torch.empty(M, N, dtype=torch.float64, device=dev)
    t64 = timed(lambda: K.dgemm(M, N, Kd, A, Kd, 0, B, Kd, 0, ref, N), 1)
    ts = timed(lambda: K.split_tf32(B, N, Kd, Kd), 2)
    Ah, Al, lpa = K.split_tf32(A, M, Kd, Kd)
    Bh, Bl, lpb = K.split_tf32(B, N, Kd, Kd)
    fl = 2.0 * M * N * Kd
    rec = {"M": M, "N": N, "K": Kd, "fp64_s": t64, "fp64_tflops": fl / t64 / 1e12,
           "split_B_s": ts, "split_GBps": 16.0 * N * Kd / ts / 1e9, "runs": []}
    C = torch.empty(M, N, dtype=torch.float64, device=dev)
    scale = float(ref.abs().max())
    for cfg in configs:
      for ls in locksteps:
        for kc in kchunks:
            t = timed(lambda: K.gemm_tf32x3(M, N, Kd, Ah, Al, lpa, Bh, Bl, lpb, C, N, kchunk=kc, config=cfg, lockstep=ls))
            d = C - ref
            r = {"config": cfg, "kchunk": kc, "lockstep": ls, "s": t, "eff_tflops": fl / t / 1e12, "tf32_tflops": 3 * fl / t / 1e12,
                 "max_err_rel": float(d.abs().max()) / scale, "rms_err_rel": float(d.pow(2).mean().sqrt()) / scale,
                 "mean_err_rel": float(d.mean()) / scale,
                 "bias_vs_abs": float((d * ref.sign()).mean()) / float(ref.abs().mean())}
            rec["runs"].append(r)
            print(name, json.dumps(r), flush=True)
    out[name] = rec
    print(name, json.dumps({k: v for k, v in rec.items() if k != "runs"}), flush=True)


if quick:
    run("small", 1600, 2048, 8192, [5, 6, 1, 3], [256, 2048])
else:
    run("ladder_slice", 1600, 36000, 90000, [5, 7], [256])
    run("ladder_slice_T", 36000, 1600, 90000, [5, 7], [256])
    run("ring", 12000, 12000, 12000, [5, 7], [256])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/mp_probe.json", "w"), indent=1)
