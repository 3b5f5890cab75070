#!/bin/bash
# batched branching tables: tests, sanitizers, timing (single call vs a batch of 256 regions)
O=gpurun_out/r2z3; mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests/test_table_configs.py tests/test_wide_value_types.py tests/test_abi_layout.py -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_table.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_table_configs.py -m gpu -x -q -k "region_table or (test_gpu_table_configs and not large)" 2>&1 | tail -8 | tee $O/memcheck_table.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_table_configs.py -m gpu -x -q -k "region_table" 2>&1 | tail -8 | tee $O/racecheck_table.txt
python - <<'PY' 2>&1 | tee $O/region_table_timing.txt
import sys, time
sys.path.insert(0, "tests")
import numpy as np
import tbcuda as tb
from helpers import regular_root
eng = tb.Engine(0)
def region(n, n_open, seed):
    root = regular_root(n, seed)
    rng = np.random.default_rng(seed + 7)
    ol = sorted(int(v) for v in rng.choice(root.nv, size=n_open, replace=False))
    return tb.SlicedBranch(tb.MISProblem(root.nv, root.edges, root.weights), tb.CompressedEinsum(root.ixs, ol, root.tree), 0), ol
print("one region per call: n open rows configs device_ms launches host_wall_ms (Python mirror)")
for n, n_open in [(12, 4), (16, 5), (20, 6), (20, 8), (24, 8), (28, 8)]:
    br, ol = region(n, n_open, 5)
    eng.region_table(br, ol)
    best = 1e9; bw = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        sizes, keep, rows = eng.region_table(br, ol)
        bw = min(bw, (time.perf_counter() - t0) * 1e3)
        best = min(best, eng.last_timing()[0])
    print(n, n_open, len(rows), sum(len(r[2]) for r in rows), round(best, 3), eng.last_timing()[1], round(bw, 3))
print("batches: regions n open device_ms host_wall_ms  -> per region device_us / wall_us")
for cnt, n, n_open in [(1, 20, 6), (16, 20, 6), (256, 20, 6), (256, 16, 5), (1024, 14, 4)]:
    regs = [region(n, n_open, 100 + i) for i in range(cnt)]
    brs, ols = [r[0] for r in regs], [r[1] for r in regs]
    eng.region_tables(brs, ols)
    best = 1e9; bw = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        eng.region_tables(brs, ols)
        bw = min(bw, (time.perf_counter() - t0) * 1e3)
        best = min(best, eng.last_timing()[0])
    print(cnt, n, n_open, round(best, 3), round(bw, 3), " -> ", round(best * 1e3 / cnt, 2), round(bw * 1e3 / cnt, 2))
PY
