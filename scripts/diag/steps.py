"""Diagnostics: per-step wall and device time of the resident path.  usage: python scripts/diag/steps.py [workload] [i32|i16] [steps]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch, tbcuda, bench
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
vt = sys.argv[2] if len(sys.argv) > 2 else "i16"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 12
branches = bench.make_workload(wl)
sliced = [tbcuda.SlicedBranch.from_parts(b.nv, b.edges, b.weights, b.ixs, b.tree, b.r) for b in branches]
eng = tbcuda.Engine(0, plan_flags=(tbcuda.TB_PLAN_NO_I16 if vt == "i32" else 0))
eng.set_stream(torch.cuda.current_stream().cuda_stream)
plans = [tbcuda.Plan(s, np.float32, engine=eng) for s in sliced if s.code is not None]
r = np.zeros(len(plans))
for i in range(n):
    torch.cuda.synchronize(); t = time.perf_counter()
    vals, status, _ = eng.contract_plans(plans, r)
    torch.cuda.synchronize()
    print(f"step {i}: wall {(time.perf_counter()-t)*1e3:8.2f} ms  device {eng.last_timing()[0]:8.2f} ms  launches {eng.last_timing()[1]}  max {vals.max()}", flush=True)
free, tot = torch.cuda.mem_get_info()
print(f"device memory in use {(tot-free)/2**30:.1f} GiB")
eng.close()
