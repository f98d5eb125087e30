"""Collect bench.py lines of several runs (gpurun_out/<tag>_bench_<workload>_<N>.json) into one scaling table.

  python tools/collect_scaling.py r2m r2i > profiles/r02_scaling.json     # later tags fill what earlier ones lack
"""
import glob
import json
import os
import re
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def last_json(path):
    try:
        lines = [l for l in open(path) if l.startswith("{")]
        return json.loads(lines[-1]) if lines else None
    except Exception:
        return None


def main(tags):
    out = {}
    for tag in tags:
        for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_bench_*_*.json"))):
            m = re.match(rf"{tag}_bench_(\w+)_(\d+)\.json", os.path.basename(f))
            b = last_json(f)
            if not m or not b:
                continue
            wl, n = m.group(1), int(m.group(2))
            rec = {"source": os.path.basename(f), "ms_per_step": b["ms_per_step"], "cells_per_s": b["value"],
                   "stage_ms": {k: round(v, 3) for k, v in b["stage_ms"].items() if v},
                   "parity_checksum": b.get("parity", {}).get("checksum")}
            if "e2e" in b:
                rec["e2e_ms_per_step"] = b["e2e"]["ms_per_step"]
                rec["e2e_cells_per_s"] = b["e2e"]["value"]
            out.setdefault(wl, {}).setdefault(str(n), rec)
    for wl, rows in out.items():
        if "1" in rows:
            t1 = rows["1"]["ms_per_step"]
            for n, r in rows.items():
                r["speedup_vs_1"] = round(t1 / r["ms_per_step"], 3)
                r["efficiency"] = round(t1 / r["ms_per_step"] / int(n), 3)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1:] or ["r2"])
