import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycc_b200 import kernels as K
dev = torch.device("cuda:0")
for (M, N, Kd) in [(128, 128, 64), (300, 257, 1000), (4096, 4096, 512)]:
    A = torch.randn(M, Kd, dtype=torch.float64, device=dev)
    B = torch.randn(N, Kd, dtype=torch.float64, device=dev)
    C = torch.zeros(M, N, dtype=torch.float64, device=dev)
    ref = A @ B.t()
    for cfg in (4, 6):
        C.zero_()
        K.dgemm(M, N, Kd, A, Kd, 0, B, Kd, 0, C, N, ksplit=1, config=cfg)
        torch.cuda.synchronize()
        print("cfg", cfg, M, N, Kd, float((C - ref).abs().max() / ref.abs().max()), flush=True)
