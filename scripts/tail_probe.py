#!/usr/bin/env python
"""Times the long-K GEMM shapes one rank sees at N = 1, 2, 4, 8 (o=40, v=300), with the library's split-K of the last wave
on or off (B200CC_GEMM_TAIL, read once per process):   B200CC_GEMM_TAIL=0 python scripts/tail_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pycc_b200 import kernels as K

dev = torch.device("cuda:0")
out = {}
for n in (1, 2, 4, 8):
    shapes = {"o3v3": (12000, 12000 // n, 12000, 1), "o3v3_T": (12000 // n, 12000, 12000, 1), "ladder": (820, (45150 + n - 1) // n, 45150, 2),
              "Z": (820, 12000 // n, 45150, 2)}
    for name, (M, N, Kd, b) in shapes.items():
        ld = (Kd + 15) // 16 * 16
        A = torch.randn(b, M, ld, dtype=torch.float64, device=dev)
        B = torch.randn(b, N, ld, dtype=torch.float64, device=dev)
        C = torch.empty(b, M, N, dtype=torch.float64, device=dev)
        f = lambda: K.dgemm(M, N, Kd, A, ld, 0, B, ld, 0, C, N, batch=b, sA=M * ld, sB=N * ld, sC=M * N)
        f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        fl = 2.0 * M * N * Kd * b
        out["%s_n%d" % (name, n)] = {"ms": min(ts), "tflops": fl / min(ts) / 1e9}
        print(name, n, (M, N, Kd, b), "%.3f ms  %.2f TFLOP/s" % (min(ts), fl / min(ts) / 1e9), flush=True)
        del A, B, C
json.dump(out, open("gpurun_out/tail_probe_%s.json" % os.environ.get("B200CC_GEMM_TAIL", "1"), "w"), indent=1)
