"""Plan-compiler microbenchmark.  usage: python scripts/plan_bench/run.py [workload] [max_plans]
Builds scripts/plan_bench/plan_bench.cpp + csrc/plan.cpp with g++ (twice: plain, and with the per-pass timers of
-DTB_PLAN_PROFILE), then prints the single-thread time per plan, the per-pass profile and the multi-thread throughput."""
import ctypes
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import tbcuda  # noqa: E402
from tbcuda import _lib as L  # noqa: E402
from tbcuda import contract as Cn  # noqa: E402

CSRC = os.path.join(ROOT, "tensorbranching.jl_b200", "csrc")
HERE = os.path.dirname(os.path.abspath(__file__))


def build(tmp, name, extra):
    out = os.path.join(tmp, name)
    subprocess.check_call(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
                           *extra, "-o", out, os.path.join(HERE, "plan_bench.cpp"), os.path.join(CSRC, "plan.cpp")])
    lib = ctypes.CDLL(out)
    lib.pb_single.restype = ctypes.c_double
    lib.pb_single.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.pb_threads.restype = ctypes.c_double
    lib.pb_threads.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    return lib


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    cap = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 30
    branches = [b for b in bench.make_workload(wl) if b.nv][:cap]
    sliced = [tbcuda.SlicedBranch.from_parts(b.nv, b.edges, b.weights, b.ixs, b.tree, b.r) for b in branches]
    arr = (L.tb_network * len(sliced))()
    keep = []
    for i, s in enumerate(sliced):
        net, w = Cn._network_of(s, np.float32)
        arr[i] = net
        keep.append((net, w))
    n = len(sliced)
    leaves = float(np.mean([len(s.code.ixs) for s in sliced]))
    with tempfile.TemporaryDirectory() as tmp:
        lib = build(tmp, "libpb.so", [])
        prof = build(tmp, "libpb_prof.so", ["-DTB_PLAN_PROFILE"])
        reps = 8
        print(f"{wl}: {n} plans, {leaves:.0f} leaves each on average; {os.cpu_count()} logical CPUs")
        print(f"single thread: {lib.pb_single(arr, n, reps):.1f} us per plan (best of {reps} passes)")
        prof.pb_single(arr, n, reps)
        print("per pass (us per plan, mean over all passes, timers included):")
        sys.stdout.flush()
        prof.pb_profile_dump(n * reps)
        sys.stdout.flush()
        th = 1
        while th <= (os.cpu_count() or 1):
            us = lib.pb_threads(arr, n, th, 5)
            print(f"{th:3d} threads: {us:7.2f} us per plan wall, {us * th:7.1f} us thread time per plan")
            th *= 2


if __name__ == "__main__":
    main()
