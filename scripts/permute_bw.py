"""Bandwidth of k_permute_bits (tb_permute_bits): bytes read + written / kernel time, for a few ranks and permutations.
usage: python scripts/permute_bw.py  (needs a GPU)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import tbcuda  # noqa: E402

eng = tbcuda.Engine(0)
rng = np.random.default_rng(0)
for rank in (20, 24, 26, 28):
    x = rng.integers(0, 1 << 30, size=1 << rank, dtype=np.int32)
    for name, perm in (("identity", list(range(rank))), ("reverse", list(range(rank))[::-1]),
                       ("rotate by 7", [(i + 7) % rank for i in range(rank)]), ("random", list(rng.permutation(rank)))):
        y = eng.permute_bits(x, perm)
        best = 1e30
        for _ in range(3):
            eng.permute_bits(x, perm)
            best = min(best, eng.last_timing()[0])
        print(f"rank {rank:2d} ({x.nbytes >> 20:5d} MiB) {name:12s}: {best:8.3f} ms  {2 * x.nbytes / best * 1e-6:8.1f} GB/s")
        if name == "identity":
            assert np.array_equal(x, y)
eng.close()
