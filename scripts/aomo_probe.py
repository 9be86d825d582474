"""AO -> MO staging on the GPU (BlockHamiltonian.from_ao): wall time and FP64 rate of the quarter transformations,
AO tensor already resident in HBM; with --stream also with the AO tensor on the HOST (numpy array swept in slabs through
pinned staging buffers, one sweep per first quarter transformation).
python scripts/aomo_probe.py NO NV [NBF] [--stream]  -> gpurun_out/aomo_probe_o<NO>v<NV>.json"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycc_b200 import kernels as K  # noqa: E402
from pycc_b200.hamiltonian import BlockHamiltonian  # noqa: E402

STREAM = "--stream" in sys.argv
argv = [a for a in sys.argv if not a.startswith("--")]
no, nv = int(argv[1]), int(argv[2])
n = int(argv[3]) if len(argv) > 3 else no + nv
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
# timing only (parity is covered by tests/test_aomo.py): any dense AO tensor will do
AO = torch.randn((n,) * 4, dtype=torch.float64, device=dev, generator=g)
C = torch.linalg.qr(torch.randn((n, n), dtype=torch.float64, device=dev, generator=g))[0][:, :no + nv].contiguous()
F_ao = torch.randn((n, n), dtype=torch.float64, device=dev, generator=g)
out = {"no": no, "nv": nv, "nbf": n, "ao_bytes": 8 * n**4}
for rep in range(2):
    torch.cuda.synchronize()
    l0 = K.launch_count()
    t0 = time.time()
    H = BlockHamiltonian.from_ao(F_ao, AO, C, no, 0, dev)
    torch.cuda.synchronize()
    out["wall_s"] = time.time() - t0
    out["launches"] = K.launch_count() - l0
    del H
# executed flops of the quarter transformations (2 x rows x cols x summed index each)
o, v = no, nv
fl = 2 * n**4 * o + 2 * n**3 * o * (o + v)                         # first index -> o; second -> o / v
fl += 3 * 2 * n**2 * o * o * (o + v) - 2 * n**2 * o * o * o          # (oo|nu sig): third index o, o(again, ooov), v  [three finishes]
fl += 2 * n * o * o * (o * o + o * v + v * v)                      # fourth index of oooo, ooov, ovov
fl += 2 * n**2 * o * v * (o + v) + 2 * n * o * v * (o * v + v * v)  # (ov|..): oovv, ovvv
fl += 2 * n**4 * v + 2 * n**3 * v * v + 2 * n**2 * v**3 + 2 * n * v**4   # vvvv, a-chunked
out["flops"] = fl
out["tflops"] = fl / out["wall_s"] / 1e12
if STREAM:
    AO_h = AO.cpu().numpy()
    del AO
    torch.cuda.empty_cache()
    for chunk_gb in (4, 16):
        torch.cuda.synchronize()
        t0 = time.time()
        H = BlockHamiltonian.from_ao(F_ao, AO_h, C, no, 0, dev, stream_ao=True, chunk_bytes=chunk_gb << 30)
        torch.cuda.synchronize()
        w = time.time() - t0
        rows = max(1, min(nv, (chunk_gb << 30) // (8 * n ** 3)))
        sweeps = 1 + -(-nv // rows)
        out["streamed_chunk%dGB" % chunk_gb] = {"wall_s": w, "sweeps_over_host_ao": sweeps,
                                               "host_to_device_GBps": sweeps * 8 * n ** 4 / w / 1e9}
        del H
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/aomo_probe_o%dv%d.json" % (no, nv), "w"), indent=1)
