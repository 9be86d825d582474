#!/usr/bin/env python
"""Ladder GEMM (o=40, v=300 shape, NA rows of a): time (and, under ncu, DRAM traffic) of the default kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycc_b200 import kernels as K
dev = torch.device("cuda:0")
o, v = 40, 300
na = int(os.environ.get("NA", "300"))
tau = torch.randn(o * o, v * v, dtype=torch.float64, device=dev)
vvvv = torch.randn(na * v, v * v, dtype=torch.float64, device=dev)
r2 = torch.zeros(o * o, v * v, dtype=torch.float64, device=dev)
fl = 2.0 * o * o * na * v * v * v
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    K.dgemm(o * o, na * v, v * v, tau, v * v, 0, vvvv, v * v, 0, r2, v * v, 0.5, 1.0, ksplit=1)
    b.record()
    torch.cuda.synchronize()
    print("persistent=%s" % os.environ.get("B200CC_GEMM_PERSISTENT", "auto"), "na", na, "ms", a.elapsed_time(b), "TFLOP/s", fl / a.elapsed_time(b) / 1e9, flush=True)
