"""Full o=40,v=300 ladder shape on the split-TF32 kernel: one launch vs several launches over slices of the
<ab|ef> rows.  python scripts/mp_ladder_slices.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycc_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
o, v = 40, 300
M, N, Kd = o * o, v * v, v * v
Ah = torch.randn(M, Kd, dtype=torch.float32, device=dev) * 0.05
Al = Ah * 1e-4
Bh = torch.empty(N, Kd, dtype=torch.float32, device=dev)
for r in range(0, N, 9000):
    Bh[r:r + 9000] = torch.randn(min(9000, N - r), Kd, dtype=torch.float32, device=dev) * 0.05
Bl = Bh * 1e-4
C = torch.zeros(M, N, dtype=torch.float64, device=dev)
out = {}


def run(nsl, cfg=5):
    rows = (N + nsl - 1) // nsl
    rows = (rows + 127) // 128 * 128
    for r0 in range(0, N, rows):
        n = min(rows, N - r0)
        K.gemm_tf32x3(M, n, Kd, Ah, Al, Kd, (Bh, r0 * Kd), (Bl, r0 * Kd), Kd, (C, r0), N, 0.5, 1.0, config=cfg)


for cfg in (5,):
    for nsl in (8, 11, 16, 22, 32, 64):
        run(nsl, cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            run(nsl, cfg)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / 2
        r = {"config": cfg, "slices": nsl, "s": t, "eff_tflops": 2.0 * M * N * Kd / t / 1e12}
        out["c%d_s%d" % (cfg, nsl)] = r
        print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/mp_ladder_slices.json", "w"), indent=1)
