"""One split-TF32 GEMM launch for ncu: python scripts/mp_one.py M N K lockstep [config]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycc_b200 import kernels as K  # noqa: E402

M, N, Kd, ls = (int(x) for x in sys.argv[1:5])
cfg = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dev = torch.device("cuda:0")
A = torch.randn(M, Kd, dtype=torch.float64, device=dev) * 0.05
B = torch.randn(N, Kd, dtype=torch.float64, device=dev) * 0.05
C = torch.empty(M, N, dtype=torch.float64, device=dev)
Ah, Al, lpa = K.split_tf32(A, M, Kd, Kd)
Bh, Bl, lpb = K.split_tf32(B, N, Kd, Kd)
del A, B
K.gemm_tf32x3(M, N, Kd, Ah, Al, lpa, Bh, Bl, lpb, C, N, config=cfg, lockstep=ls)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
K.gemm_tf32x3(M, N, Kd, Ah, Al, lpa, Bh, Bl, lpb, C, N, config=cfg, lockstep=ls)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
