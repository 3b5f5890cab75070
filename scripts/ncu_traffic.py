#!/usr/bin/env python
"""One contract_plans call of a workload on a single stream lane (deterministic launch order), with the engine's
per-launch records (algorithmic ops / bytes of every launch) dumped to a CSV.  Run it under
    ncu --set full --clock-control none --import-source on -k regex:k_gemm2 -c N -o <rep> python scripts/ncu_traffic.py ...
and join the two with scripts/ncu_traffic_join.py: DRAM bytes and algorithmic bytes of THE SAME launches."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TB_LANES"] = "1"

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--max-branches", type=int, default=None)
ap.add_argument("--first", type=int, default=0, help="first branch of the workload to use")
ap.add_argument("--count", type=int, default=None)
ap.add_argument("--dump", required=True)
ap.add_argument("--weights", default="unit", choices=["unit", "f32"])
args = ap.parse_args()

import numpy as np  # noqa: E402

import bench  # noqa: E402
import tbcuda  # noqa: E402

brs = bench.make_workload(args.workload, args.max_branches)
brs = brs[args.first:(args.first + args.count) if args.count else None]
rng = np.random.default_rng(15)
sliced = []
for b in brs:
    w = None if args.weights == "unit" or b.nv == 0 else (1.0 + rng.random(b.nv)).astype(np.float32)
    sliced.append(tbcuda.SlicedBranch.from_parts(b.nv, b.edges, w, b.ixs, b.tree, b.r))
eng = tbcuda.Engine(0)
plans = [tbcuda.Plan(s, np.float32, engine=eng) if s.code is not None else None for s in sliced]
if os.path.exists(args.dump):
    os.remove(args.dump)
os.environ["TB_DUMP_LAUNCHES"] = args.dump
eng.profile(2)
vals, status, mx = eng.contract_plans(plans, np.array([b.r for b in brs], dtype=np.float64))
print("branches", len(brs), "max", mx, "launches", eng.last_timing()[1], "device ms", eng.last_timing()[0],
      "value type", plans[0].info().value_type if plans and plans[0] else None)
