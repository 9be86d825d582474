#!/usr/bin/env python
"""(T) timing breakdown on one GPU: the batched two-segment GEMM vs the fused energy kernel."""
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycc_b200 import kernels as K, cctriples              # noqa: E402
from pycc_b200.hamiltonian import BlockHamiltonian          # noqa: E402
from pycc_b200.synthetic import make_synthetic              # noqa: E402

dev = torch.device("cuda:0")
out = {}
for (o, v, nb) in [(20, 150, 48), (30, 280, 12), (40, 300, 12)]:
    syn = make_synthetic(o, v, seed=0, device=dev)
    H = BlockHamiltonian.from_factor(syn, dev, names=("ooov", "oovv", "ovvv"))
    w = types.SimpleNamespace(H=H, no=o, nv=v, o=H.o, v=H.v, comm=None)
    w.eps_o, w.eps_v = H.eps[H.o].contiguous(), H.eps[H.v].contiguous()
    w.t1 = 0.01 * torch.randn(o, v, dtype=torch.float64, device=dev)
    w.t2 = K.div_d2(H.block("oovv"), w.eps_o, w.eps_v)
    for cube in (False, True):
        eng = cctriples.TriplesEngine(w, cube_q=cube)
        trip = [t for t in cctriples.triples_list(o) if not (t[0] == t[1] == t[2])][-nb:]
        ijk = torch.tensor(trip, dtype=torch.int32).to(dev)
        et = torch.zeros(1, dtype=torch.float64, device=dev)
        Q = eng.build_q(trip)
        K.t_energy_batch(o, v, ijk, Q, eng.t1, eng.t2, eng.oovv, eng.fov, w.eps_o, w.eps_v, et, blocked=eng.cube)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 3
        tg = te = 0.0
        for _ in range(reps):
            ev[0].record()
            Q = eng.build_q(trip)
            ev[1].record()
            K.t_energy_batch(o, v, ijk, Q, eng.t1, eng.t2, eng.oovv, eng.fov, w.eps_o, w.eps_v, et, blocked=eng.cube)
            ev[2].record()
            torch.cuda.synchronize()
            tg += ev[0].elapsed_time(ev[1]) * 1e-3 / reps
            te += ev[1].elapsed_time(ev[2]) * 1e-3 / reps
        fl = (12 * v ** 4 + 12 * o * v ** 3) * nb
        qbytes = 6 * v ** 3 * 8 * nb
        r = {"o": o, "v": v, "cube_q": cube, "triples": nb, "gemm_s": tg, "energy_s": te,
             "gemm_tflops": fl / tg / 1e12, "total_tflops": fl / (tg + te) / 1e12,
             "energy_GBps_algorithmic": qbytes / te / 1e9}
        out["o%dv%d_cube%d" % (o, v, int(cube))] = r
        print(json.dumps(r), flush=True)
        del eng, Q
    del H, w
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/t_probe.json", "w"), indent=1)
