"""Join an ncu launch list (gpu__time_duration per launch) with the plans' per-level work to get a
per-launch roofline: ops, algorithmic bytes, achieved Gop/s and GB/s, which bound applies.
usage: python scripts/launch_roofline.py gpurun_out/r01/launches.csv cfg2 [skip_launches]"""
import csv
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import tbcuda  # noqa: E402

P_DPX = 17.95e12
BW = 6547.5e9


def main():
    path, wl = sys.argv[1], sys.argv[2]
    branches = bench.make_workload(wl)
    sliced = [tbcuda.SlicedBranch.from_parts(b.nv, b.edges, b.weights, b.ixs, b.tree, b.r) for b in branches]
    plans = [tbcuda.Plan(s) if s.code is not None else None for s in sliced]
    half = any(p is not None and p.info().value_type == 3 for p in plans)
    eb = 2 if half else 4
    global P_DPX
    if half:
        P_DPX = 35.4e12
    steps = [p.steps() if p else [] for p in plans]
    # waves of 256 plans in order (the engine's default), per level per kind
    expected = []
    W = 128
    live = [i for i, p in enumerate(plans) if p is not None]
    for w0 in range(0, len(live), W):
        mem = live[w0:w0 + W]
        nl = max(plans[i].info().n_levels for i in mem)
        f_ops = sum(s_.fused_ops for s_ in (plans[i].info() for i in mem))
        expected.append(("k_fused", f_ops, 0.0, 0))
        for lv in range(1, nl + 1):
            for kind, nm in ((1, "k_generic"), (2, "k_gemm")):
                ops = byts = 0.0
                cnt = 0
                desc = {}
                for i in mem:
                    for s in steps[i]:
                        if s.level == lv and s.kind == kind:
                            tc = s.rank_c + s.n_k + s.n_ka + s.n_kb
                            ops += 2.0 ** tc
                            byts += eb * (2.0 ** s.rank_a + 2.0 ** s.rank_b + 2.0 ** s.rank_c)
                            cnt += 1
                            key = (s.n_m, s.n_n, s.n_b, s.n_k, s.n_ka + s.n_kb)
                            desc[key] = desc.get(key, 0) + 2.0 ** tc
                if cnt:
                    top = sorted(desc.items(), key=lambda kv: -kv[1])[:2]
                    expected.append((nm, ops, byts, cnt, top))
        expected.append(("k_finalize", 0, 0, 0))
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    L = [(r[ki], float(r[vi].replace(",", "")) * 1e-9) for r in rows[hi + 1:] if len(r) > vi and "tb::" in r[ki]]
    if len(sys.argv) > 3:  # launches of earlier (warm-up) calls; the first call grows the arena and has more, smaller waves
        L = L[int(sys.argv[3]):]
    print(f"{'kernel':10s} {'us':>9s} {'Gop':>9s} {'MB':>8s} {'Gop/s':>8s} {'%dpx':>6s} {'GB/s':>7s} {'%hbm':>6s} {'t_roof_us':>9s} {'eff':>5s}  nodes top(m,n,b,k)")
    tot_t = tot_roof = 0
    for (name, t), e in zip(L, expected):
        assert e[0] in name, (name, e[0])
        ops, byts = e[1], e[2]
        roof = max(ops / P_DPX, byts / BW)
        tot_t += t
        tot_roof += roof
        extra = f"{e[3]:5d} {e[4]}" if len(e) > 4 else ""
        print(f"{e[0]:10s} {t*1e6:9.1f} {ops*1e-9:9.3f} {byts*1e-6:8.1f} {ops/t*1e-9:8.0f} {ops/t/P_DPX*100:6.1f} {byts/t*1e-9:7.0f} {byts/t/BW*100:6.1f} {roof*1e6:9.1f} {roof/t:5.2f}  {extra}")
    print(f"total {tot_t*1e3:.2f} ms, roofline {tot_roof*1e3:.2f} ms, efficiency {tot_roof/tot_t:.3f}")


if __name__ == "__main__":
    main()
