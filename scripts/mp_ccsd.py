"""precision='MP' vs 'DP' CCSD on the GPU at a given size: per-iteration energies, converged energy difference,
s/iter.  Writes gpurun_out/mp_ccsd_o<o>v<v>.json.   python scripts/mp_ccsd.py O V [kchunk ...]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycc_b200  # noqa: E402
from pycc_b200 import kernels as K  # noqa: E402
from pycc_b200.synthetic import make_synthetic  # noqa: E402

o, v = int(sys.argv[1]), int(sys.argv[2])
kchunks = [int(x) for x in sys.argv[3:]] or [0]
dev = torch.device("cuda:0")
syn = make_synthetic(o, v, seed=0, device=dev)
out = {"o": o, "v": v}


def solve(precision, e_conv, r_conv):
    cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", precision=precision, quiet=True)
    torch.cuda.synchronize()
    t0 = time.time()
    e = cc.solve_cc(e_conv, r_conv, 60)
    torch.cuda.synchronize()
    dt = time.time() - t0
    return (float(e) if e is not None else None), cc.trace, dt


e_dp, tr_dp, t_dp = solve("DP", 1e-10, 1e-10)
out["DP"] = {"e": e_dp, "iters": len(tr_dp), "s_per_iter": t_dp / len(tr_dp)}
print("DP", json.dumps(out["DP"]), flush=True)
for kc in kchunks:
    K.MIXED.kchunk = kc
    e_mp, tr_mp, t_mp = solve("MP", 1e-8, 1e-7)
    n = min(len(tr_dp), len(tr_mp))
    rec = {"kchunk": kc, "e": e_mp, "iters": len(tr_mp), "s_per_iter": t_mp / len(tr_mp),
           "dE_converged": None if e_mp is None else e_mp - e_dp,
           "max_dE_same_iteration": max(abs(tr_dp[i][0] - tr_mp[i][0]) for i in range(n)),
           "rms_last": tr_mp[-1][1], "stats": dict(K.MIXED.stats)}
    out["MP_kc%d" % kc] = rec
    print("MP", json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/mp_ccsd_o%dv%d.json" % (o, v), "w"), indent=1)
