"""(T) densities (cctriples.t3_density) on one GPU: wall time, FP64 rate and per-phase CUDA-event breakdown.
python scripts/t3d_probe.py O V [NJ]   -- NJ = number of j values timed (default all).  Writes gpurun_out/t3d_probe_o<O>v<V>.json"""
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
nj = int(sys.argv[3]) if len(sys.argv) > 3 else o
dev = torch.device("cuda:0")
syn = make_synthetic(o, v, seed=0, device=dev)
H = BlockHamiltonian.from_factor(syn, dev, names=("ooov", "oovv", "ovvv"))
eo, ev = H.eps[H.o].contiguous(), H.eps[H.v].contiguous()
t1 = 0.01 * torch.randn(o, v, dtype=torch.float64, device=dev)
t2 = K.div_d2(H.block("oovv"), eo, ev)
ct = pycc_b200.device.DeviceManager(device="GPU", precision="DP").contract
js = list(range(nj))
# warm-up on one j (allocations, derived layouts), then the timed run
cctriples.t3_density(H.o, H.v, o, v, t1, t2, H.F, H.ERI, H.L, ct, js=js[:1])
torch.cuda.synchronize()
prof = {}
l0 = K.launch_count()
t0 = time.time()
et, dens = cctriples.t3_density(H.o, H.v, o, v, t1, t2, H.F, H.ERI, H.L, ct, js=js, prof=prof)
torch.cuda.synchronize()
wall = time.time() - t0
ntrip = nj * o * o
# executed flops per triple: t3 build 12 v^4 + 12 o v^3; Gvvvo 2 v^4; S2/X2 <kb|cd> 2 x 2 v^4; S2/X2 <jk|lc> 2 x 2 o v^3; Gooov 2 o v^3
fl = ntrip * (12 * v**4 + 12 * o * v**3 + 2 * v**4 + 4 * v**4 + 4 * o * v**3 + 2 * o * v**3)
out = {"o": o, "v": v, "j_values": nj, "triples": ntrip, "wall_s": wall, "tflops": fl / wall / 1e12,
       "full_o3_s_est": wall * o / nj, "launches": K.launch_count() - l0, "phase_ms": prof,
       "phase_tflops": {"t3_gemm": ntrip * (12 * v**4 + 12 * o * v**3) / (prof["t3_gemm"] * 1e-3) / 1e12,
                        "density_gemm": ntrip * (6 * v**4 + 6 * o * v**3) / (prof["density_gemm"] * 1e-3) / 1e12},
       "phase_gbs": {"connected": ntrip * 7 * v**3 * 8 / (prof["connected"] * 1e-3) / 1e9,
                     "forms": ntrip * 5 * v**3 * 8 / (prof["forms"] * 1e-3) / 1e9},
       "et_partial": float(et)}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/t3d_probe_o%dv%d.json" % (o, v), "w"), indent=1)
