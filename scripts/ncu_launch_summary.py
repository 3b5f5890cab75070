#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel launch count, total time, share of the run.
usage: ncu_launch_summary.py launches.csv [steps_in_run]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
col = {h: i for i, h in enumerate(rows[hdr])}
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    v = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    tot[name][0] += 1
    tot[name][1] += ms
allms = sum(v[1] for v in tot.values())
steps = int(sys.argv[2]) if len(sys.argv) > 2 else None
print("kernel,launches,total_ms,share" + (",launches_per_step" if steps else ""))
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{n},{ms:.4f},{ms / allms:.4f}" + (f",{n / steps:.1f}" if steps else ""))
print(f"TOTAL,{sum(v[0] for v in tot.values())},{allms:.4f},1.0" + (f",{sum(v[0] for v in tot.values()) / steps:.1f}" if steps else ""))
