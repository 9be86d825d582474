#!/usr/bin/env python
"""Write profiles/parity_r02.json from N=1 bench.py lines: the energies every N > 1 run must reproduce to 1e-10 Eh.

    python scripts/record_parity.py gpurun_out/bench_n1.json [more N=1 lines ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "parity_r02.json")


def main():
    rec = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for path in sys.argv[1:]:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        assert d["n_gpus"] == 1, "the record is taken at N = 1"
        o, v = d["t"]["o"], d["t"]["v"]
        key = "o%dv%d" % (o, v)
        r = rec.setdefault(key, {"ecc_by_iteration": {}, "e_t_by_iteration": {}})
        for n, e in enumerate(d["energies"], 1):
            r["ecc_by_iteration"][str(n)] = e
        if d["t"]["triples_timed"] == d["t"]["triples_total"]:
            r["e_t_by_iteration"][str(len(d["energies"]))] = d["t"]["e_t"]
        c4 = d.get("t_c4")
        if c4 and c4["triples_timed"] == c4["triples_total"]:
            rec["o30v280"] = {"ecc_after_3_iterations": c4["ecc_after_3_iterations"], "e_t": c4["e_t"]}
        rec.setdefault("sources", []).append(os.path.basename(path))
    json.dump(rec, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
