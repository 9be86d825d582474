#!/usr/bin/env python
"""Full (T) job, paired (three summed arrays per triple, four-segment GEMMs) vs six-array form, on one GPU.
    python scripts/t_compare.py [o v]"""
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pycc_b200
from pycc_b200 import cctriples, kernels as K
from pycc_b200.synthetic import make_synthetic

o, v = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (30, 280)
dev = torch.device("cuda:0")
syn = make_synthetic(o, v, seed=0, device=dev)
cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True)
d = cc.make_diis(8)
for _ in range(3):
    cc.iterate()
    cc.diis_step(d, True)
del d
trip = [t for t in cctriples.triples_list(o) if not (t[0] == t[1] == t[2])]
flops = (12 * v**4 + 12 * o * v**3) * len(trip)
out = {"o": o, "v": v, "triples": len(trip)}
for name, paired in (("six_arrays", False), ("paired", True), ("six_arrays_again", False), ("paired_again", True)):
    cctriples.PAIRED = paired
    cctriples.t_tjl(cc, trip[:4])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = K.launch_count()
    a.record()
    et = cctriples.t_tjl(cc)
    b.record()
    torch.cuda.synchronize()
    t = a.elapsed_time(b) * 1e-3
    out[name] = {"seconds": t, "tflops": flops / t / 1e12, "e_t": float(et), "launches": K.launch_count() - l0}
print(json.dumps(out))
