"""Device time of tb_table_configs (all optimal configurations of a branching table) for regions of growing size.
usage: python scripts/table/time_table_configs.py > profiles/..."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import tbcuda as tb  # noqa: E402
from helpers import regular_root  # noqa: E402

eng = tb.Engine(0)
print("n  open  rows_kept  configs  table_configs_ms(count+write)  vertex_sets/s  contract_table_ms  compactify_ms")
for n, n_open in [(12, 4), (16, 5), (20, 6), (20, 8), (24, 8), (28, 8), (30, 8), (32, 10)]:
    root = regular_root(n, 5)
    rng = np.random.default_rng(7)
    ol = sorted(int(v) for v in rng.choice(root.nv, size=n_open, replace=False))
    br = tb.SlicedBranch(tb.MISProblem(root.nv, root.edges, root.weights), tb.CompressedEinsum(root.ixs, ol, root.tree), 0)
    p = tb.Plan(br, value_type=tb.TB_VALUE_SIZE_CONFIG, engine=eng)
    labels, sizes, _ = eng.contract_table(p)
    t_con = eng.last_timing()[0]
    keep = eng.compactify_table(sizes)
    t_cmp = eng.last_timing()[0]
    best = None
    for _ in range(3):
        own, off, cfgs = eng.table_configs(br, labels, keep)
        ms = eng.last_timing()[0]
        best = ms if best is None else min(best, ms)
    assert np.array_equal(own, sizes)
    print(f"{n:2d} {n_open:3d} {int(keep.sum()):6d} {len(cfgs):8d} {best:10.3f} {3 * 2.0 ** n / (best * 1e-3):12.3e} {t_con:8.3f} {t_cmp:8.3f}")
