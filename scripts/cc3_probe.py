"""CC3 iteration on one GPU: time of the dressed intermediates and of the connected-triples step next to the CCSD part.
python scripts/cc3_probe.py O V -> gpurun_out/cc3_probe_o<O>v<V>.json"""
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
dev = torch.device("cuda:0")
syn = make_synthetic(o, v, seed=0, device=dev)
cc = pycc_b200.ccwfn(syn, model="CC3", device="GPU", quiet=True)


def clock():
    torch.cuda.synchronize()
    return time.time()


cc.iterate()                                   # warm-up (derived layouts, allocations)
t0 = clock()
l0 = K.launch_count()
ecc, rms = cc.iterate()
t1 = clock()
launches = K.launch_count() - l0
cc.model = "CCSD"
t2 = clock()
cc.iterate()
t3 = clock()
cc.model = "CC3"
H = cc.H
ta = clock()
Wmnij = cc.build_cc3_Wmnij(cc.o, cc.v, H.ERI, cc.t1)
W = {"Wmbij": cc.build_cc3_Wmbij(cc.o, cc.v, H.ERI, cc.t1, Wmnij), "Wmnie": cc.build_cc3_Wmnie(cc.o, cc.v, H.ERI, cc.t1),
     "Wamef": cc.build_cc3_Wamef(cc.o, cc.v, H.ERI, cc.t1), "Wabei": cc.build_cc3_Wabei(cc.o, cc.v, H.ERI, cc.t1)}
tb = clock()
npairs = o * (o + 1) // 2
# executed flops of the triples step: t3 build 12 v^4 + 12 o v^3 per built tile (o per pair); per loop body (o^3 of them)
# the W_amef product 2 v^4 and the W_mnie product 2 o v^3
fl = npairs * o * (12 * v**4 + 12 * o * v**3) + o**3 * (2 * v**4 + 2 * o * v**3)
t_trip = (t1 - t0) - (t3 - t2) - (tb - ta)
out = {"o": o, "v": v, "cc3_iter_s": t1 - t0, "ccsd_iter_s": t3 - t2, "intermediates_s": tb - ta,
       "triples_s_est": t_trip, "triples_tflops_executed": fl / t_trip / 1e12,
       "triples_tflops_reference_formulation": o**3 * (14 * v**4 + 14 * o * v**3) / t_trip / 1e12,
       "launches_per_iter": launches, "ecc": ecc, "rms": rms, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
# real-time CC3 residual with real amplitudes (explicit-field triples: one t3 build per ORDERED pair) next to the
# field-free residual of the same amplitudes
n = cc.no + cc.nv
g = torch.Generator(device=dev).manual_seed(1)
M = 0.01 * torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
Ff = (H.F + M + M.T).contiguous()
cc.residuals(H.F, cc.t1, cc.t2)
tc = clock()
cc.residuals(H.F, cc.t1, cc.t2)
td = clock()
cc.residuals(Ff, cc.t1, cc.t2, real_time=True)
te = clock()
out["residual_s"] = td - tc
out["residual_real_time_s"] = te - td
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/cc3_probe_o%dv%d.json" % (o, v), "w"), indent=1)
