#!/usr/bin/env python
"""One launch of the fused (a,b,c)-driven (T) kernel at a bench shape for an ncu capture:

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:t_abc_kernel \
        -o gpurun_out/t_abc python scripts/ncu_t_abc.py 40 300 4
"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycc_b200 import kernels as K, cctriples              # noqa: E402
from pycc_b200.hamiltonian import BlockHamiltonian          # noqa: E402
from pycc_b200.synthetic import make_synthetic              # noqa: E402

o, v, rounds = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (40, 300, 4)
dev = torch.device("cuda:0")
syn = make_synthetic(o, v, seed=0, device=dev)
H = BlockHamiltonian.from_factor(syn, dev, names=("ooov", "oovv", "ovvv"))
w = types.SimpleNamespace(H=H, no=o, nv=v, o=H.o, v=H.v, comm=None, mixed=False)
w.eps_o, w.eps_v = H.eps[H.o].contiguous(), H.eps[H.v].contiguous()
w.t1 = 0.01 * torch.randn(o, v, dtype=torch.float64, device=dev)
w.t2 = K.div_d2(H.block("oovv"), w.eps_o, w.eps_v)
eng = cctriples.FusedTriples(w)
lst = cctriples.abc_list(v)
mid = lst.size // 2
sample = torch.from_numpy(lst[mid:mid + K.NSM * rounds].copy()).to(dev)
eng.energy(sample[:K.NSM])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.energy(sample)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
