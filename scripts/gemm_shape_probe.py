"""Tile-config / split-K sweep for the long-K, small-output GEMMs of t3_density (M = N = v, K = o v^2, B N-major).
python scripts/gemm_shape_probe.py [V] [O]  -> gpurun_out/gemm_shape_probe.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycc_b200 import kernels as K  # noqa: E402

v = int(sys.argv[1]) if len(sys.argv) > 1 else 300
o = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dev = torch.device("cuda:0")
Kd = o * v * v
A = torch.randn(v * Kd, dtype=torch.float64, device=dev)
B = torch.randn(Kd * v, dtype=torch.float64, device=dev)
Cm = torch.zeros(v * v, dtype=torch.float64, device=dev)
res = []


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


fl = 2.0 * v * v * Kd
for transB in (1, 0):
    for cfg in (4, 5) if transB else (6, 7):
        for ks in (16, 64, 128, 198, 256, 400, 800):
            try:
                ms = timeit(lambda: K.dgemm(v, v, Kd, A, Kd, 0, B, v if transB else Kd, transB, Cm, v, 1.0, 1.0,
                                            ksplit=ks, config=cfg))
            except Exception as exc:           # noqa: BLE001
                res.append({"transB": transB, "config": cfg, "ksplit": ks, "error": str(exc)[:80]})
                continue
            res.append({"transB": transB, "config": cfg, "ksplit": ks, "ms": ms, "tflops": fl / ms / 1e9})
            print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"v": v, "o": o, "results": res}, open("gpurun_out/gemm_shape_probe.json", "w"), indent=1)
