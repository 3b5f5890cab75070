"""Timeline of one end-to-end tb_contract_networks call (TB_TRACE_CALL=1): per pipeline batch, when the host had it
compiled / enqueued and when the device finished it.  usage: python scripts/trace_call.py [workload] [calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TB_TRACE_CALL"] = "1"
import numpy as np  # noqa: E402

import bench  # noqa: E402
import tbcuda  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 4
branches = bench.make_workload(wl)
sliced = [tbcuda.SlicedBranch.from_parts(b.nv, b.edges, b.weights, b.ixs, b.tree, b.r) for b in branches]
eng = tbcuda.Engine(0)
for c in range(calls):
    sys.stderr.write(f"---- call {c}\n")
    sys.stderr.flush()
    t0 = time.perf_counter()
    vals = tbcuda.contract_slices(sliced, np.float32, True, engine=eng)
    dt = (time.perf_counter() - t0) * 1e3
    sys.stderr.write(f"---- call {c}: {dt:.3f} ms wall (incl. the trace's own synchronisation), max {vals.max()}\n")
eng.close()
