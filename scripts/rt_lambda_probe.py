"""The Lambda half of the RT-CC right-hand side on the GPU: cclambda.residuals(F, t1, t2, l1, l2) with complex amplitudes
(HBAR rebuilt from (F, t1, t2) in every call, cclambda.py:202-256) on pairs of real planes (default) and from five real
samples of the whole residual (round 1), next to one real call.
python scripts/rt_lambda_probe.py O V -> gpurun_out/rt_lambda_probe_o<O>v<V>.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycc_b200  # noqa: E402
from pycc_b200.synthetic import make_synthetic  # noqa: E402

o, v = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
syn = make_synthetic(o, v, seed=0, device=dev)
cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
g = torch.Generator(device=dev).manual_seed(1)
cc.t1 = cc.t1 + 0.01 * torch.randn(o, v, dtype=torch.float64, device=dev, generator=g)
lm = pycc_b200.cclambda(cc, pycc_b200.cchbar(cc))
t1, t2, l1, l2 = cc.t1, cc.t2, lm.l1, lm.l2
z1 = torch.complex(t1, 0.01 * torch.randn(o, v, dtype=torch.float64, device=dev, generator=g))
z2 = torch.complex(t2, 0.1 * t2)
y1 = torch.complex(l1, 0.01 * torch.randn(o, v, dtype=torch.float64, device=dev, generator=g))
y2 = torch.complex(l2, 0.1 * l2)
m = torch.randn(cc.H.F.shape, dtype=torch.float64, device=dev, generator=g)
F = cc.H.F + 0.01 * (m + m.T)


QUICK = "--quick" in sys.argv          # large sizes: no warm-up call, no sampled evaluation


def timeit(fn, reps=1):
    if not QUICK:
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


ms_real, _ = timeit(lambda: lm.residuals(F, t1, t2, l1, l2))
lm.complex_on_planes = True
ms_planes, a = timeit(lambda: lm.residuals(F, z1, z2, y1, y2))
ms_samples, diff = None, None
if not QUICK:
    lm.complex_on_planes = False
    ms_samples, b = timeit(lambda: lm.residuals(F, z1, z2, y1, y2))
    lm.complex_on_planes = True
    diff = max(float((a[0] - b[0]).abs().max()), float((a[1] - b[1]).abs().max()))
out = {"o": o, "v": v, "real_call_ms": ms_real, "complex_on_planes_ms": ms_planes, "complex_five_samples_ms": ms_samples,
       "planes_over_real": ms_planes / ms_real, "samples_over_real": ms_samples / ms_real if ms_samples else None,
       "speedup": ms_samples / ms_planes if ms_samples else None, "max_abs_diff_between_the_two": diff,
       "quick": QUICK,
       "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/rt_lambda_probe_o%dv%d.json" % (o, v), "w"), indent=1)
