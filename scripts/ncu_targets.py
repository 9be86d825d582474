#!/usr/bin/env python
"""Launch exactly the kernels worth a full ncu capture, once each, at the bench shapes (o=40, v=300):
the pair-packed ladder in FP64 (pack_tau, dgemm_tma_kernel, ladder_unpack) and in precision='MP' (tf32x3_gemm_r_kernel),
and the paired (T) energy kernel on a small batch of triples.

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:'dgemm_tma|tf32x3_gemm|t_energy_cp|pack_tau|ladder_unpack' -o gpurun_out/r02_targets python scripts/ncu_targets.py
(the set-up -- integral generation, guesses -- runs outside the cudaProfilerStart / Stop bracket)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pycc_b200
from pycc_b200 import cctriples, kernels as K
from pycc_b200.hamiltonian import BlockHamiltonian
from pycc_b200.synthetic import make_synthetic

o, v = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 300)
which = sys.argv[3] if len(sys.argv) > 3 else "all"
dev = torch.device("cuda:0")
syn = make_synthetic(o, v, seed=0, device=dev)
if which in ("all", "fp64"):
    cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True)
    tau = K.build_tau(cc.t1, cc.t2)
    r2 = torch.zeros_like(cc.t2)
    eng = cctriples.TriplesEngine(cc, paired=True)      # constant transposed copies built outside the bracket
    trip = [(5, 3, 1), (7, 7, 2), (9, 4, 4), (11, 6, 0), (20, 11, 3), (31, 30, 2), (38, 12, 12), (39, 39, 0)]
    eng.qbuf(len(trip))
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    cc._ladder(tau, r2, symmetric=True)                 # pack_tau, dgemm_tma_kernel (batch 2), ladder_unpack
    if which == "all":
        eng.energy(trip)                                # dgemm_tma_kernel (4 K segments), t_energy_cp_kernel<3>
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    eng.close()
    del eng
    del cc, tau, r2
    cctriples._QCACHE.clear()
    import gc
    gc.collect()
    torch.cuda.empty_cache()
if which in ("all", "mp"):
    ccm = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", precision="MP", quiet=True)
    tau = K.build_tau(ccm.t1, ccm.t2)
    r2 = torch.zeros_like(ccm.t2)
    with K.mixed_mode(True):
        ccm._ladder(tau, r2, symmetric=True)            # warm-up: nothing to build, but the first call sizes buffers
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        ccm._ladder(tau, r2, symmetric=True)            # split_tf32, tf32x3_gemm_r_kernel x slices
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
print("done")
