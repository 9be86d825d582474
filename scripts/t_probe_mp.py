"""(T) on the GPU, precision 'DP' vs 'MP': t3-build GEMM time, energy-kernel time, E(T) of the sample.
Writes gpurun_out/t_probe_mp.json.   python scripts/t_probe_mp.py"""
import gc
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycc_b200  # noqa: E402
from pycc_b200 import kernels as K, cctriples  # noqa: E402
from pycc_b200.synthetic import make_synthetic  # noqa: E402

dev = torch.device("cuda:0")
out = {}
for (o, v, nb) in [(20, 150, 48), (30, 280, 12), (40, 300, 12)]:
    syn = make_synthetic(o, v, seed=0, device=dev)
    for prec in ("DP", "MP"):
        cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", precision=prec, quiet=True)
        cc.t1 = 0.01 * torch.randn(o, v, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
        eng = cctriples.TriplesEngine(cc)
        trip = [t for t in cctriples.triples_list(o) if not (t[0] == t[1] == t[2])][-nb:]
        ijk = torch.tensor(trip, dtype=torch.int32).to(dev)
        et = torch.zeros(1, dtype=torch.float64, device=dev)
        Q = eng.build_q(trip)
        K.t_energy_batch(o, v, ijk, Q, eng.t1, eng.t2, eng.oovv, eng.fov, cc.eps_o, cc.eps_v, et)
        torch.cuda.synchronize()
        e_sample = float(et)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps, tg, te = 3, 0.0, 0.0
        for _ in range(reps):
            ev[0].record()
            Q = eng.build_q(trip)
            ev[1].record()
            K.t_energy_batch(o, v, ijk, Q, eng.t1, eng.t2, eng.oovv, eng.fov, cc.eps_o, cc.eps_v, et)
            ev[2].record()
            torch.cuda.synchronize()
            tg += ev[0].elapsed_time(ev[1]) * 1e-3 / reps
            te += ev[1].elapsed_time(ev[2]) * 1e-3 / reps
        fl = (12 * v ** 4 + 12 * o * v ** 3) * nb
        r = {"o": o, "v": v, "precision": prec, "triples": nb, "gemm_s": tg, "energy_s": te,
             "gemm_tflops_fp64_equiv": fl / tg / 1e12, "total_tflops_fp64_equiv": fl / (tg + te) / 1e12,
             "q_write_GBps": 48.0 * v ** 3 * nb / tg / 1e9, "energy_GBps": 48.0 * v ** 3 * nb / te / 1e9,
             "e_t_sample": e_sample}
        out["o%dv%d_%s" % (o, v, prec)] = r
        print(json.dumps(r), flush=True)
        eng.close()
        del eng, Q, cc
        gc.collect()
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/t_probe_mp.json", "w"), indent=1)
