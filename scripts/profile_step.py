#!/usr/bin/env python
"""Profiling harness (run under ncu with --profile-from-start off): one CCSD iteration + a few (T)
triples, or only the ladder GEMM / only the (T) kernels, inside a cudaProfilerStart/Stop range.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python scripts/profile_step.py --o 20 --v 150
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dgemm -c 1 \
      -o gpurun_out/ladder python scripts/profile_step.py --o 20 --v 150 --only ladder
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycc_b200                                   # noqa: E402
from pycc_b200 import kernels as K, cctriples      # noqa: E402
from pycc_b200.synthetic import make_synthetic     # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--o", type=int, default=20)
ap.add_argument("--v", type=int, default=150)
ap.add_argument("--only", default="step", choices=["step", "ladder", "t"])
ap.add_argument("--triples", type=int, default=4)
ap.add_argument("--precision", default="DP", choices=["DP", "MP"])
args = ap.parse_args()
dev = torch.device("cuda:0")
syn = make_synthetic(args.o, args.v, seed=0, device=dev)
cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True, precision=args.precision)
mixed = K.mixed_mode(args.precision == "MP")
mixed.__enter__()
diis = pycc_b200.helper_diis(cc.t1, cc.t2, 8)
for _ in range(2):
    cc.iterate()
    cc.diis_step(diis)
trip = [t for t in cctriples.triples_list(args.o) if not (t[0] == t[1] == t[2])][-args.triples:]
cctriples.t_tjl(cc, trip[:1])
tau = K.build_tau(cc.t1, cc.t2)
r2 = torch.zeros_like(cc.t2)
cc._ladder(tau, r2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
if args.only == "ladder":
    cc._ladder(tau, r2)
elif args.only == "t":
    cctriples.t_tjl(cc, trip)
else:
    cc.iterate()
    cc.diis_step(diis)
    cctriples.t_tjl(cc, trip)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled", args.only, "launches so far", K.launch_count())
