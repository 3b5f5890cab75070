#!/bin/bash
# single-call branching tables: tests, racecheck / memcheck of the region kernels, timing of the one-call table
O=gpurun_out/r2z2; mkdir -p $O; rm -f $O/*
timeout 600 python -m pytest tests/test_table_configs.py tests/test_wide_value_types.py tests/test_abi_layout.py -m gpu -x -q 2>&1 | tail -15 | tee $O/pytest_table.txt
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_table_configs.py -m gpu -x -q -k "region_table or (test_gpu_table_configs and not large)" 2>&1 | tail -8 | tee $O/racecheck_table.txt
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_table_configs.py -m gpu -x -q -k "region_table" 2>&1 | tail -8 | tee $O/memcheck_region_table.txt
python - <<'PY' 2>&1 | tee $O/region_table_timing.txt
import sys, time
sys.path.insert(0, "tests")
import numpy as np
import tbcuda as tb
from helpers import regular_root
eng = tb.Engine(0)
print("n open rows configs device_ms launches host_wall_ms(one call, Python mirror)")
for n, n_open in [(12, 4), (16, 5), (20, 6), (20, 8), (24, 8), (28, 8)]:
    root = regular_root(n, 5)
    rng = np.random.default_rng(7)
    ol = sorted(int(v) for v in rng.choice(root.nv, size=n_open, replace=False))
    br = tb.SlicedBranch(tb.MISProblem(root.nv, root.edges, root.weights), tb.CompressedEinsum(root.ixs, ol, root.tree), 0)
    eng.region_table(br, ol)
    best = 1e9; bw = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        sizes, keep, rows = eng.region_table(br, ol)
        bw = min(bw, (time.perf_counter() - t0) * 1e3)
        best = min(best, eng.last_timing()[0])
    print(n, n_open, len(rows), sum(len(r[2]) for r in rows), round(best, 3), eng.last_timing()[1], round(bw, 3))
PY
