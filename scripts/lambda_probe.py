"""HBAR build + Lambda iteration on the GPU: wall times and FP64 rate.
    python scripts/lambda_probe.py O V [conv] [DP|MP]
Writes gpurun_out/lambda_probe_o<O>v<V>[_mp].json"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycc_b200  # noqa: E402
from pycc_b200 import kernels as K  # noqa: E402
from pycc_b200.synthetic import make_synthetic  # noqa: E402

o, v = int(sys.argv[1]), int(sys.argv[2])
conv = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
prec = sys.argv[4] if len(sys.argv) > 4 else "DP"
dev = torch.device("cuda:0")
syn = make_synthetic(o, v, seed=0, device=dev)
cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True, precision=prec)
torch.cuda.synchronize()
t0 = time.time()
ecc = cc.solve_cc(conv, conv, 60)
torch.cuda.synchronize()
t_cc = time.time() - t0
peaks = {"setup+ccsd": torch.cuda.max_memory_allocated() / 1e9, "held_after_ccsd": torch.cuda.memory_allocated() / 1e9}
torch.cuda.reset_peak_memory_stats()
l0 = K.launch_count()
t0 = time.time()
hb = pycc_b200.cchbar(cc)
torch.cuda.synchronize()
t_hbar = time.time() - t0
peaks["hbar"] = torch.cuda.max_memory_allocated() / 1e9
peaks["held_after_hbar"] = torch.cuda.memory_allocated() / 1e9
torch.cuda.reset_peak_memory_stats()
n_hbar = K.launch_count() - l0
lm = pycc_b200.cclambda(cc, hb)
l0 = K.launch_count()
t0 = time.time()
lecc = lm.solve_lambda(conv, conv, 60)
torch.cuda.synchronize()
t_lam = time.time() - t0
peaks["lambda"] = torch.cuda.max_memory_allocated() / 1e9
iters = len(lm.trace)
# dominant terms of one Lambda iteration: Hvvvv ladder, Hoooo, three o^3v^3 ring terms, l2.Hvvvo / l2.Hovoo, Goo/Gvv
fl = 2 * o**2 * v**4 + 2 * o**4 * v**2 + 3 * 2 * o**3 * v**3 + 2 * o**2 * v**3 * 2 + 2 * o**3 * v**2 * 2 + 2 * o**2 * v**3 + 2 * o**3 * v**2
out = {"o": o, "v": v, "conv": conv, "precision": prec, "mixed_stats": dict(K.MIXED.stats), "hvvvv_materialised": hb._Hvvvv is not None,
       "peak_mem_gb": max(peaks["setup+ccsd"], peaks["hbar"], peaks["lambda"]), "peaks_gb": peaks, "ecc": float(ecc), "ccsd_s_per_iter": t_cc / len(cc.trace), "hbar_s": t_hbar, "hbar_launches": n_hbar,
       "lambda_pseudoE": float(lecc) if lecc is not None else None, "lambda_iters": iters,
       "lambda_s_per_iter": t_lam / iters, "lambda_launches_per_iter": (K.launch_count() - l0) / iters,
       "lambda_tflops": fl / (t_lam / iters) / 1e12}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/lambda_probe_o%dv%d%s.json" % (o, v, "" if prec == "DP" else "_" + prec.lower()), "w"), indent=1)
