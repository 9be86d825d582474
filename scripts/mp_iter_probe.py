#!/usr/bin/env python
"""precision='MP' CCSD iterations at o=40,v=300 with per-phase device times (B200CC_GEMM_SKINNY / _TAIL toggles are read
from the environment):   python scripts/mp_iter_probe.py [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pycc_b200
from pycc_b200 import kernels as K
from pycc_b200.synthetic import make_synthetic

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
syn = make_synthetic(40, 300, seed=0, device=dev)
cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", precision="MP", quiet=True)
diis = cc.make_diis(8)
for _ in range(4):
    cc.iterate(); cc.diis_step(diis, True)
torch.cuda.synchronize()
K.PHASES.on = True
K.PHASES.collect()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    cc.iterate(); cc.diis_step(diis, True)
e1.record()
torch.cuda.synchronize()
ph = K.PHASES.collect()
print("s/iter %.4f" % (e0.elapsed_time(e1) * 1e-3 / steps))
for k, (n, ms) in ph.items():
    print("  %-60s %8.2f" % (k, ms / steps))
