"""Diagnostics (library built with -DTB_KPROF): per-CTA timeline of one resident step, dumped by tb_shutdown.
usage: TBCUDA_LIB=.../libtbcuda_kprof.so TB_TL_DUMP=out.bin python scripts/diag/timeline.py [workload]
then   python scripts/diag/timeline.py --analyze out.bin"""
import sys, time
import numpy as np
sys.path.insert(0, ".")

REC = np.dtype([("key", "<u8"), ("t0", "<u8"), ("t1", "<u8"), ("smid", "<u4"), ("kind", "<u4")])


def analyze(path):
    r = np.fromfile(path, dtype=REC)
    t_lo, t_hi = r["t0"].min(), r["t1"].max()
    span = (t_hi - t_lo) * 1e-6
    print(f"{len(r)} CTA records, span {span:.3f} ms")
    names = {0: "fused", 1: "generic", 2: "gemm"}
    for k in (0, 1, 2):
        q = r[r["kind"] == k]
        if not len(q):
            continue
        dur = (q["t1"] - q["t0"]).astype(np.float64)
        # launches = distinct keys; per launch wall interval
        keys, inv = np.unique(q["key"], return_inverse=True)
        l0 = np.full(len(keys), np.iinfo(np.uint64).max, dtype=np.uint64); np.minimum.at(l0, inv, q["t0"])
        l1 = np.zeros(len(keys), dtype=np.uint64); np.maximum.at(l1, inv, q["t1"])
        print(f"{names[k]:8s} CTAs {len(q):7d} launches {len(keys):4d} CTA-time {dur.sum()*1e-6:9.2f} ms  "
              f"mean CTA {dur.mean()*1e-3:7.1f} us  sum of launch walls {(l1-l0).sum()*1e-6:7.2f} ms")
    # SM occupancy by GEMM CTAs over time (2 slots per SM): sample on a 2 us grid
    g = r[r["kind"] == 2]
    step = 2000
    nb = int((t_hi - t_lo) // step) + 2
    occ = np.zeros(nb)
    a = ((g["t0"] - t_lo) // step).astype(np.int64); b = ((g["t1"] - t_lo) // step).astype(np.int64)
    d = np.zeros(nb + 1); np.add.at(d, a, 1); np.add.at(d, b + 1, -1)
    occ = np.cumsum(d)[:nb]
    n_sm = int(r["smid"].max()) + 1
    slots = 2 * n_sm
    print(f"SMs {n_sm}; GEMM CTA slots busy: mean {occ.mean()/slots:.3f} of {slots}")
    hist, edges = np.histogram(occ / slots, bins=[0, 0.02, 0.25, 0.5, 0.75, 0.9, 1.01])
    for h, e0, e1 in zip(hist, edges[:-1], edges[1:]):
        print(f"  GEMM slot occupancy {e0:4.2f}-{e1:4.2f}: {h*step*1e-6:7.2f} ms")
    # time with no GEMM CTA anywhere but other kernels running
    o = r[r["kind"] != 2]
    d2 = np.zeros(nb + 1)
    a = ((o["t0"] - t_lo) // step).astype(np.int64); b = ((o["t1"] - t_lo) // step).astype(np.int64)
    np.add.at(d2, a, 1); np.add.at(d2, b + 1, -1)
    occ2 = np.cumsum(d2)[:nb]
    print(f"  grid cells with GEMM occupancy < 25%: {(occ/slots < 0.25).sum()*step*1e-6:.2f} ms, of which other kernels active: "
          f"{((occ/slots < 0.25) & (occ2 > 0)).sum()*step*1e-6:.2f} ms, nothing at all: {((occ == 0) & (occ2 == 0)).sum()*step*1e-6:.2f} ms")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--analyze":
        analyze(sys.argv[2]); sys.exit(0)
    import torch, tbcuda, bench
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    branches = bench.make_workload(wl)
    sliced = [tbcuda.SlicedBranch.from_parts(b.nv, b.edges, b.weights, b.ixs, b.tree, b.r) for b in branches]
    eng = tbcuda.Engine(0)
    plans = [tbcuda.Plan(s, np.float32, engine=eng) for s in sliced if s.code is not None]
    r = np.zeros(len(plans))
    for i in range(4):
        torch.cuda.synchronize(); t = time.perf_counter()
        vals, status, _ = eng.contract_plans(plans, r)
        torch.cuda.synchronize(); print(f"step {i}: {(time.perf_counter()-t)*1e3:.2f} ms, max {vals.max()}")
    eng.close()
