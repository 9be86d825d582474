#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals, and -- for bench.py runs --
the launches of ONE CCSD step (between two consecutive energy_kernel launches of the timed region).

    python scripts/launch_summary.py gpurun_out/launches_r02a.csv [step_index]
"""
import csv
import collections
import re
import sys


def rows(path):
    lines = [l for l in open(path) if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            unit = r["Metric Unit"]
            t = float(r["Metric Value"].replace(",", ""))
            t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
            name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("b200cc::", "")
            yield name, t, r["Grid Size"]


def main():
    L = list(rows(sys.argv[1]))
    which = int(sys.argv[2]) if len(sys.argv) > 2 else -2
    marks = [i for i, (n, _, _) in enumerate(L) if n.startswith("energy_kernel")]
    print("%d launches, %.1f ms total; %d energy_kernel launches" % (len(L), sum(t for _, t, _ in L), len(marks)))
    if len(marks) >= 3:
        a, b = marks[which - 1] + 1, marks[which] + 1
        # include the trailing final_reduce of the energy
        seg = L[a:b]
        tot = sum(t for _, t, _ in seg)
        print("one step = launches %d..%d: %d launches, %.2f ms of kernel time" % (a, b, len(seg), tot))
        agg = collections.OrderedDict()
        for n, t, g in seg:
            k = agg.setdefault(n, [0, 0.0, 0.0])
            k[0] += 1
            k[1] += t
            k[2] = max(k[2], t)
        for n, (c, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("  %-70s x%-4d %9.3f ms  %5.1f %%   (largest %.3f)" % (n[:70], c, t, 100 * t / tot, mx))


if __name__ == "__main__":
    main()
