import torch
dev = torch.device("cuda:0")
o, v = 40, 300
tau = torch.randn(o * o, v * v, dtype=torch.float64, device=dev)
vvvv = torch.randn(v * v, v * v, dtype=torch.float64, device=dev)
out = torch.empty(o * o, v * v, dtype=torch.float64, device=dev)
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); torch.matmul(tau, vvvv.t(), out=out); b.record(); torch.cuda.synchronize()
    print("cublas ladder ms", a.elapsed_time(b), 2.0 * o * o * v ** 4 / a.elapsed_time(b) / 1e9)
