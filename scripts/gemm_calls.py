#!/usr/bin/env python
"""Every dgemm call of ONE CCSD iteration at a bench shape, with its shape, device time (CUDA events around the call) and
the fraction of the FP64 tensor / HBM bound it reaches -- the list that shows which small products are worth a tile config.

    python scripts/gemm_calls.py [o v]  ->  gpurun_out/gemm_calls.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pycc_b200
from pycc_b200 import kernels as K
from pycc_b200.synthetic import make_synthetic

o, v = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 300)
dev = torch.device("cuda:0")
syn = make_synthetic(o, v, seed=0, device=dev)
cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
diis = cc.make_diis(8)
for _ in range(2):
    cc.iterate()
    cc.diis_step(diis, True)
calls = []
orig = K.dgemm


import inspect
SIG = inspect.signature(orig)


def timed(*args, **kwargs):
    ba = SIG.bind(*args, **kwargs)
    ba.apply_defaults()
    kw = ba.arguments
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    r = orig(*args, **kwargs)
    e1.record()
    torch.cuda.synchronize()
    K2 = sum(int(kw[s][4]) for s in ("seg2", "seg3", "seg4") if kw.get(s) is not None)
    calls.append({"M": int(kw["M"]), "N": int(kw["N"]), "K": int(kw["K"]) + K2, "batch": int(kw["batch"]),
                  "tA": int(bool(kw["transA"])), "tB": int(bool(kw["transB"])), "ksplit": kw.get("ksplit"),
                  "ms": e0.elapsed_time(e1)})
    return r


K.dgemm = timed
cc.iterate()
K.dgemm = orig
tot = sum(c["ms"] for c in calls)
for c in calls:
    fl = 2.0 * c["M"] * c["N"] * c["K"] * c["batch"]
    by = 8.0 * c["batch"] * (c["M"] * c["K"] + c["N"] * c["K"] + c["M"] * c["N"])
    c["tensor_bound_ms"] = fl / 35.4e12 * 1e3
    c["hbm_bound_ms"] = by / 6.5e12 * 1e3
    c["frac_of_bound"] = max(c["tensor_bound_ms"], c["hbm_bound_ms"]) / c["ms"]
calls.sort(key=lambda c: -c["ms"])
print("total %.1f ms in %d calls" % (tot, len(calls)))
for c in calls[:40]:
    print("%8.3f ms  bound %7.3f / %6.3f  frac %.2f   M=%-8d N=%-7d K=%-8d b=%-6d tA=%d tB=%d ks=%s"
          % (c["ms"], c["tensor_bound_ms"], c["hbm_bound_ms"], c["frac_of_bound"], c["M"], c["N"], c["K"], c["batch"],
             c["tA"], c["tB"], c["ksplit"]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(calls, open("gpurun_out/gemm_calls.json", "w"), indent=1)
