"""RT-CC right-hand side on the GPU: CCwfn.residuals with complex amplitudes and a field-dressed Fock matrix (five fused
FP64 residual evaluations), next to one real residual.  python scripts/rt_probe.py O V -> gpurun_out/rt_probe_o<O>v<V>.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pycc_b200  # noqa: E402
from pycc_b200 import kernels as K  # noqa: E402
from pycc_b200.synthetic import make_synthetic  # noqa: E402

o, v = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
syn = make_synthetic(o, v, seed=0, device=dev)
cc = pycc_b200.ccwfn(syn, model="CCSD", device="GPU", quiet=True)
g = torch.Generator(device=dev).manual_seed(1)
t1 = cc.t1 + 0.01 * torch.randn(o, v, dtype=torch.float64, device=dev, generator=g)
t2 = cc.t2.clone()
z1 = torch.complex(t1, 0.01 * torch.randn(o, v, dtype=torch.float64, device=dev, generator=g))
z2 = torch.complex(t2, 0.1 * t2)
m = torch.randn(cc.H.F.shape, dtype=torch.float64, device=dev, generator=g)
F = cc.H.F + 0.01 * (m + m.T)


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


REPS = 3 if o * v <= 6000 else 1
fl = 2 * o**2 * v**4 + 7 * 2 * o**3 * v**3 + 2 * 2 * o**4 * v**2 + 8 * 2 * o**2 * v**3      # reference count, one real residual


def sym_residual():
    r1, half = cc._residuals_half(F, t1, t2, symmetric=True)
    K.symmetrize_r2(half)


ms_real = timeit(lambda: cc.residuals(F, t1, t2), REPS)                  # general mode (what residuals() runs)
ms_sym = timeit(sym_residual, REPS)                                      # (i >= j) mode (what iterate() runs)
out = {"o": o, "v": v, "real_residual_ms": ms_real, "real_residual_pair_mode_ms": ms_sym, "complex_ms": {},
       "amplitudes": "pair-symmetric in both planes (MP1 doubles), Hermitian field"}
for pair, native, heavy in ((True, True, True), (True, True, False), (True, False, False), (False, True, True),
                            (False, False, False)):
        cc.complex_pair_mode, cc.complex_native_ladder, cc.complex_native_heavy = pair, native, heavy
        l0 = K.launch_count()
        ms = timeit(lambda: cc.residuals(F, z1, z2, real_time=True), REPS)
        out["complex_ms"]["pair_mode=%d,native_ladder=%d,native_heavy=%d" % (pair, native, heavy)] = {
            "ms": ms, "ratio_to_real_general": ms / ms_real, "ratio_to_real_pair_mode": ms / ms_sym,
            "launches": (K.launch_count() - l0) // (REPS + 1)}
cc.complex_pair_mode = cc.complex_native_ladder = cc.complex_native_heavy = True
best = out["complex_ms"]["pair_mode=1,native_ladder=1,native_heavy=1"]["ms"]
out["complex_residual_ms"] = best
out["ratio"] = best / ms_real
out["round1_formulation_ms"] = out["complex_ms"]["pair_mode=0,native_ladder=0,native_heavy=0"]["ms"]
out["complex_equivalent_tflops"] = 4 * fl / (best * 1e-3) / 1e12
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/rt_probe_o%dv%d.json" % (o, v), "w"), indent=1)
