#!/usr/bin/env python
"""ONE real iteration of the unmodified reference at the bench workload (o=40, v=300) on the host cores, next to the
same at BASELINE configs[1] (o=20, v=150) -- the measurement that calibrates ``bench.py``'s reference-arm scale factor.

    python scripts/reference_fullsize.py [--o 40 --v 300] > profiles/reference_fullsize_r02.json

Needs ~165 GB of host memory: <ab|ef> (64.8 GB, chemist layout as the reference's Hamiltonian holds it) plus the
transposed copy numpy.tensordot makes of it for 'ijef,abef->ijab' (ccwfn.py:931), plus <ma|ef>, L and o^2v^2 temporaries.
The address space is capped (RLIMIT_AS) so that running out raises MemoryError instead of taking the box down.
"""
import argparse
import json
import os
import resource
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = str(len(os.sched_getaffinity(0)))

import numpy as np  # noqa: E402


def mem_available_gb():
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable"):
            return int(line.split()[1]) / 1048576.0
    return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--o", type=int, default=40)
    ap.add_argument("--v", type=int, default=300)
    ap.add_argument("--o-s", type=int, default=20)
    ap.add_argument("--v-s", type=int, default=150)
    ap.add_argument("--iterations", type=int, default=2)
    ap.add_argument("--mem-cap-gb", type=float, default=0.0, help="RLIMIT_AS; 0 = MemAvailable - 6 GB")
    args = ap.parse_args()
    from baseline import refload
    from pycc_b200.synthetic import make_synthetic
    import bench
    log = lambda m: print(m, file=sys.stderr, flush=True)   # noqa: E731
    cores = len(os.sched_getaffinity(0))
    pools = bench.blas_threads(cores)
    avail = mem_available_gb()
    cap = args.mem_cap_gb or max(8.0, avail - 6.0)
    resource.setrlimit(resource.RLIMIT_AS, (int(cap * 2**30), int(cap * 2**30)))
    ref = refload.load_reference()
    out = {"o": args.o, "v": args.v, "o_s": args.o_s, "v_s": args.v_s, "cores": cores, "blas_threads": pools,
           "mem_available_gb": avail, "reference_root": os.path.relpath(ref.root, ROOT) if ref.root.startswith(ROOT) else ref.root}
    # small size first (also warms the BLAS pool)
    syn = make_synthetic(args.o_s, args.v_s, seed=0)
    w = refload.reference_wfn(ref, syn.F, syn.no, None, blocks=refload.host_blocks(syn))
    secs, en = refload.timed_solve_cc(w, 4, echo=log)
    out["seconds_small"] = float(np.median(secs[1:]))
    out["small_all"] = secs
    del w
    t0 = time.time()
    syn = make_synthetic(args.o, args.v, seed=0)
    log("factor ready %.1f s" % (time.time() - t0))
    blocks = refload.host_blocks(syn, log=log)
    out["setup_s"] = time.time() - t0
    log("blocks ready %.1f s" % out["setup_s"])
    w = refload.reference_wfn(ref, syn.F, syn.no, None, blocks=blocks)
    log("wavefunction ready %.1f s" % (time.time() - t0))
    secs, en = refload.timed_solve_cc(w, args.iterations, echo=log)
    out["full_all"] = secs
    out["seconds_full"] = float(min(secs))
    out["energies_full"] = en
    out["ratio"] = out["seconds_full"] / out["seconds_small"]
    out["flop_ratio"] = bench.ccsd_flops(args.o, args.v, False) / bench.ccsd_flops(args.o_s, args.v_s, False)
    out["max_rss_gb"] = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1048576.0
    print(json.dumps(out))


if __name__ == "__main__":
    main()
