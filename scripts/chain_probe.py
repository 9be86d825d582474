"""The whole CCSD(T) chain of the reference on one GPU, at a size the reference cannot reach: amplitudes with
make_t3_density=True (ccwfn.py:216-319, 300-304), HBAR (cchbar.py:54-99), Lambda with the (T) sources
(cclambda.py:69-200).  python scripts/chain_probe.py O V [CONV] -> gpurun_out/chain_probe_o<O>v<V>.json"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycc_b200  # noqa: E402
from pycc_b200 import kernels as K, cctriples  # noqa: E402
from pycc_b200.synthetic import make_synthetic  # noqa: E402

o, v = int(sys.argv[1]), int(sys.argv[2])
conv = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
dev = torch.device("cuda:0")


def clock():
    torch.cuda.synchronize()
    return time.time()


t0 = clock()
syn = make_synthetic(o, v, seed=0, device=dev)
cc = pycc_b200.ccwfn(syn, model="CCSD(T)", device="GPU", quiet=True, make_t3_density=True)
t1 = clock()
l0 = K.launch_count()
# amplitudes first (timed alone), then the density-producing (T) step that solve_cc would run
cc.model = "CCSD"
ecc = float(cc.solve_cc(conv, conv, 100))
cc.model = "CCSD(T)"
t2 = clock()
et = float(cc.t3_density())
t3 = clock()
et_tjl = float(cctriples.t_tjl(cc))
t4 = clock()
hb = pycc_b200.cchbar(cc)
t5 = clock()
lm = pycc_b200.cclambda(cc, hb)
lecc = lm.solve_lambda(conv, conv, 100)
t6 = clock()
out = {"o": o, "v": v, "conv": conv, "setup_s": t1 - t0, "ccsd_s": t2 - t1, "ccsd_iters": len(cc.trace),
       "e_ccsd": ecc, "t3_density_s": t3 - t2, "e_t_from_t3_density": et, "t_tjl_s": t4 - t3, "e_t_from_t_tjl": et_tjl,
       "abs_dE_t": abs(et - et_tjl), "hbar_s": t5 - t4, "lambda_s": t6 - t5, "lambda_iters": len(lm.trace),
       "lambda_pseudoE": None if lecc is None else float(lecc), "launches": K.launch_count() - l0,
       "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/chain_probe_o%dv%d.json" % (o, v), "w"), indent=1)
