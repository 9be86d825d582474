#!/usr/bin/env python
"""Fused (a,b,c)-driven (T) kernel on one GPU: parity of sampled tiles against the (i,j,k)-driven path, timing of a
sample of virtual triples, and (with --full) the whole job next to the (i,j,k)-driven one.

    python scripts/t_abc_probe.py [--full] [--shapes 30x280,40x300] [--rounds 16]
"""
import argparse
import json
import os
import sys
import time
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycc_b200 import kernels as K, cctriples              # noqa: E402
from pycc_b200.hamiltonian import BlockHamiltonian          # noqa: E402
from pycc_b200.synthetic import make_synthetic              # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--full", action="store_true")
ap.add_argument("--shapes", default="30x280,40x300")
ap.add_argument("--rounds", type=int, default=16)
ap.add_argument("--out", default="gpurun_out/t_abc_probe.json")
args = ap.parse_args()

dev = torch.device("cuda:0")
out = {}


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3, r


for shape in args.shapes.split(","):
    o, v = (int(x) for x in shape.split("x"))
    syn = make_synthetic(o, v, seed=0, device=dev)
    H = BlockHamiltonian.from_factor(syn, dev, names=("ooov", "oovv", "ovvv"))
    w = types.SimpleNamespace(H=H, no=o, nv=v, o=H.o, v=H.v, comm=None, mixed=False)
    w.eps_o, w.eps_v = H.eps[H.o].contiguous(), H.eps[H.v].contiguous()
    w.t1 = 0.01 * torch.randn(o, v, dtype=torch.float64, device=dev)
    w.t2 = K.div_d2(H.block("oovv"), w.eps_o, w.eps_v)
    r = {"o": o, "v": v}
    t0 = time.time()
    eng = cctriples.FusedTriples(w)
    torch.cuda.synchronize()
    r["setup_s"] = time.time() - t0
    lst = cctriples.abc_list(v)
    r["nabc"] = int(lst.size)
    # ---- parity of one tile against the (i,j,k)-driven numerators
    ijk = cctriples.TriplesEngine(w, paired=True)
    a, b, c = v - 3, v // 2, 5
    W = eng.w_tile(a, b, c)
    err = 0.0
    for (i, j, k) in ((o - 1, o // 2, 0), (3, 3, 1), (o - 2, 1, 1)):
        w3, _ = ijk.t3_parts(i, j, k, False)
        for (I, J, Kk), (A, B, Cc) in (((i, j, k), (a, b, c)), ((j, i, k), (b, a, c)), ((k, j, i), (c, b, a)),
                                        ((i, k, j), (a, c, b))):
            err = max(err, abs(float(W[I, J, Kk]) - float(w3[A, B, Cc])))
    r["tile_max_abs_err_vs_ijk"] = err
    # ---- timing of a contiguous sample from the middle of the list
    n = K.NSM * args.rounds
    mid = lst.size // 2
    sample = torch.from_numpy(lst[mid:mid + n].copy()).to(dev)
    eng.energy(sample[:K.NSM])
    ts = min(timed(lambda: eng.energy(sample))[0] for _ in range(3))
    kp = (v + 15) // 16 * 16
    ko = (o + 15) // 16 * 16
    fl_exec = 3 * 2.0 * o ** 3 * (2 * v + 2 * o)           # per (a,b,c), unpadded
    r.update(sample=n, sample_s=ts, ms_per_abc_per_cta=ts / args.rounds * 1e3,
             tflops=fl_exec * n / ts / 1e12, projected_job_s=ts * lst.size / n,
             k_padding=(2 * kp + 2 * ko) / (2.0 * v + 2 * o))
    print(json.dumps(r), flush=True)
    if args.full:
        trip = [t for t in cctriples.triples_list(o) if not (t[0] == t[1] == t[2])]
        ijk.energy(trip[:4])
        t_ijk, e_ijk = timed(lambda: float(ijk.energy(trip)[0]))
        dl = torch.from_numpy(lst).to(dev)
        t_abc, e_abc = timed(lambda: float(eng.energy(dl)[0]))
        r.update(full_ijk_s=t_ijk, full_abc_s=t_abc, e_t_ijk=e_ijk, e_t_abc=e_abc, abs_dE=abs(e_ijk - e_abc),
                 full_abc_tflops=fl_exec * lst.size / t_abc / 1e12,
                 full_ijk_tflops=len(trip) * 12.0 * v ** 3 * (v + o) / t_ijk / 1e12)
        print(json.dumps(r), flush=True)
    out[shape] = r
    ijk.close()
    del eng, ijk, H, w
    cctriples._QCACHE.clear()
    torch.cuda.empty_cache()
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump(out, open(args.out, "w"), indent=1)
