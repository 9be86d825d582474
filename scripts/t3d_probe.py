"""(T) densities (cctriples.t3_density) on one GPU: wall time, FP64 rate and per-phase CUDA-event breakdown.
python scripts/t3d_probe.py O V [NP]   -- NP = number of (i > j) pairs timed (default all pairs i >= j).
Writes gpurun_out/t3d_probe_o<O>v<V>.json"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycc_b200  # noqa: E402
from pycc_b200 import kernels as K, cctriples  # noqa: E402
from pycc_b200.hamiltonian import BlockHamiltonian  # noqa: E402
from pycc_b200.synthetic import make_synthetic  # noqa: E402
import types  # noqa: E402

o, v = int(sys.argv[1]), int(sys.argv[2])
npair = int(sys.argv[3]) if len(sys.argv) > 3 else 0
mixed = len(sys.argv) > 4 and sys.argv[4].upper() == "MP"
dev = torch.device("cuda:0")
syn = make_synthetic(o, v, seed=0, device=dev)
H = BlockHamiltonian.from_factor(syn, dev, names=("ooov", "oovv", "ovvv"))
eo, ev = H.eps[H.o].contiguous(), H.eps[H.v].contiguous()
t1 = 0.01 * torch.randn(o, v, dtype=torch.float64, device=dev)
t2 = K.div_d2(H.block("oovv"), eo, ev)
ct = pycc_b200.device.DeviceManager(device="GPU", precision="DP").contract
allp = [(i, j) for j in range(o) for i in range(j, o)]
pairs = allp if not npair else [(o - 1 - n, n) for n in range(npair)]          # i > j: both loop bodies per t3 build
# warm-up on one pair (allocations, derived layouts), then the timed run
cctriples.t3_density(H.o, H.v, o, v, t1, t2, H.F, H.ERI, H.L, ct, pairs=pairs[:1], mixed=mixed)
torch.cuda.synchronize()
prof = {}
l0 = K.launch_count()
t0 = time.time()
et, dens = cctriples.t3_density(H.o, H.v, o, v, t1, t2, H.F, H.ERI, H.L, ct, pairs=pairs, prof=prof, mixed=mixed)
torch.cuda.synchronize()
wall = time.time() - t0
nbuild = len(pairs) * o                                   # t3 tiles built
ntrip = sum(1 if i == j else 2 for i, j in pairs) * o     # loop bodies of the reference served
# executed flops: t3 build 12 v^4 + 12 o v^3 per built tile; per loop body Gvvvo 2 v^4, S2/X2 <kb|cd> 2 x 2 v^4,
# S2/X2 <jk|lc> 2 x 2 o v^3, Gooov 2 o v^3
fl_build = nbuild * (12 * v**4 + 12 * o * v**3)
fl_dens = ntrip * (6 * v**4 + 6 * o * v**3)
dens_ms = sum(prof[k] for k in prof if k.startswith("gemm_"))
out = {"o": o, "v": v, "precision": "MP" if mixed else "DP", "pairs": len(pairs), "t3_tiles_built": nbuild, "loop_bodies": ntrip, "wall_s": wall,
       "tflops_executed": (fl_build + fl_dens) / wall / 1e12,
       "tflops_reference_formulation": ntrip * (18 * v**4 + 18 * o * v**3) / wall / 1e12,
       "full_o3_s_est": wall * o * o * o / ntrip, "launches": K.launch_count() - l0, "phase_ms": prof,
       "phase_tflops": {"t3_gemm": fl_build / (prof["t3_gemm"] * 1e-3) / 1e12,
                        "density_gemms": fl_dens / (dens_ms * 1e-3) / 1e12,
                        "gemm_Gvvvo": ntrip * 2 * v**4 / (prof["gemm_Gvvvo"] * 1e-3) / 1e12,
                        "gemm_ovvv": ntrip * 4 * v**4 / (prof["gemm_ovvv"] * 1e-3) / 1e12},
       "phase_gbs": {"connected": nbuild * 7 * v**3 * 8 / (prof["connected"] * 1e-3) / 1e9,
                     "forms": ntrip * 5 * v**3 * 8 / (prof["forms"] * 1e-3) / 1e9},
       "et_partial": float(et)}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/t3d_probe_o%dv%d%s.json" % (o, v, "_mp" if mixed else ""), "w"), indent=1)
