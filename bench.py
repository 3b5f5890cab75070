#!/usr/bin/env python
"""bench.py -- tropical-contraction throughput of the hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps K --warmup W            this engine (libtbcuda.so)
    python bench.py --impl reference ...                     CPU arm: the oracle's C/OpenMP port of the
                                                             reference's algorithm on the host cores
    torchrun ... bench.py --gpus N ...                       one rank per GPU.  Default --scaling strong: ONE unit list,
                                                             dealt longest-first (LPT) by tropical ops, every rank compiles
                                                             and contracts only its shard, one all-reduce(max) over the
                                                             result vector.  --scaling weak: every rank contracts its own copy.
    --workload cfg1|cfg2|cfg3|cfg4|cfg5                      BASELINE.json configs[0..4]; cfg3 is index-sliced (--slice-k)

Default workload = cfg4 = BASELINE.json configs[3], the north-star target (3-regular n=500, sc_target=28; first 32
finished branches of the stand-in host's depth-first slicer, 2^45.5 tropical ops per step).  At N=1 the line also carries
`other_configs`: short runs of cfg2 / cfg3 / cfg5 (value, ms, roofline fraction, e2e, agreement with the CPU port).

A "step" = one pass of contract_slices over the whole branch list of the workload.
value  = tropical Gop/s with plans resident in HBM (ops = sum over nodes 2^(m+n+k+b))
e2e    = the same metric through contract_slices(branches) from host objects: cost estimation, LPT, plan compilation,
         descriptor upload (H2D), contraction, result read-back (D2H) all inside the timed region.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, params, sc_target, max_branches)
    "cfg1": ("regular", dict(n=100, d=3, seed=1), 10, None),
    "cfg2": ("regular", dict(n=200, d=3, seed=2), 20, None),
    "cfg2s": ("regular", dict(n=160, d=3, seed=2), 18, None),
    "cfg3": ("ksg", dict(m=30, n=30, rho=0.8, seed=3), 24, None),
    # the full branch set of n=500 at sc 28 is far beyond one bench step (the stand-in host's greedy tree starts at
    # sc 86): a step contracts the first 32 branches the depth-first slicer finishes, 2^45.5 tropical ops
    "cfg4": ("regular", dict(n=500, d=3, seed=4), 28, 32),
    "cfg5": ("regular_many", dict(n=150, d=3, seed0=1000, count=1024), 16, None),
}
# index slicing: every branch is cut into 2^k independent slices (k labels fixed).  cfg3's stand-in
# branch list is ONE branch (the kernelised KSG already has sc 23 <= 24), so slices are the only units to shard.
DEFAULT_SLICE_K = {"cfg3": 3}
# branches too heavy for the CPU legs (one cfg4 branch is minutes of CPU time): the CPU sample is made of index
# slices of one branch, cut with this many labels
CPU_SLICE_K = {"cfg4": 10}
OTHER_CONFIGS = ("cfg2", "cfg3", "cfg5")


def make_workload(name, max_branches=None):
    """-> list of standin Branch objects.  Generation is host-side python (seconds to a minute), so it is cached under
    workloads/cache/ (git-ignored; /tmp when the tree is read-only); workload_hash_status() pins what was loaded to the
    tracked generator."""
    from workloads import standin_host as H

    cdir = os.path.join(ROOT, "workloads", "cache")
    try:
        os.makedirs(cdir, exist_ok=True)
    except OSError:
        cdir = "/tmp"
    cache = os.path.join(cdir, f"tbcuda_workload_{name}_{max_branches}.pkl")
    if os.path.exists(cache):
        with open(cache, "rb") as f:
            return pickle.load(f)
    kind, p, sc_target, mb = WORKLOADS[name]
    mb = max_branches or mb
    if kind == "regular":
        nv, edges = H.random_regular_graph(p["n"], p["d"], p["seed"])
        brs = H.slice_bfs(H.make_root(nv, edges, seed=p["seed"], ntrials=6), sc_target, max_branches=mb)
    elif kind == "ksg":
        nv, edges = H.random_ksg(p["m"], p["n"], p["rho"], p["seed"])
        brs = H.slice_bfs(H.make_root(nv, edges, seed=p["seed"], ntrials=4), sc_target, max_branches=mb)
    else:
        brs = []
        for i in range(p["count"] if not mb else min(mb, p["count"])):
            nv, edges = H.random_regular_graph(p["n"], p["d"], p["seed0"] + i)
            root = H.kernelize(H.make_root(nv, edges, seed=p["seed0"] + i, ntrials=1))
            sub = H.slice_bfs(root, sc_target, max_branches=1)
            brs.append(sub[0])
    tmp = cache + f".{os.getpid()}"
    with open(tmp, "wb") as f:
        pickle.dump(brs, f)
    os.replace(tmp, cache)
    return brs


def golden_record(name):
    p = os.path.join(ROOT, "tests", "golden", f"baseline_{name}.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def workload_hash_status(name, branches, max_branches=None):
    """The benched branch list against the committed hash of what the tracked generator produces
    (tests/golden/baseline_<name>.json): "ok", or raises -- a stale or foreign cache file must not be benchmarked."""
    from workloads import standin_host as H

    rec = golden_record(name)
    if rec is None or max_branches is not None:
        return "unpinned (no committed hash for this workload / branch count)"
    h = H.branch_list_hash(branches)
    if h != rec["hash"]:
        raise RuntimeError(f"workload {name}: the cached branch list does not match the committed hash "
                           f"(delete workloads/cache/ to regenerate it from workloads/standin_host.py)")
    return "ok: sha256 of the branch list == tests/golden/baseline_%s.json" % name


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def lpt_shards(costs, n):
    """longest-processing-time-first assignment of units to n ranks."""
    import heapq

    cost = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-cost, kind="stable").tolist()
    cost = cost.tolist()
    heap = [(0.0, r) for r in range(n)]  # (load, rank): the least-loaded rank, the lowest rank among equals
    owner = [0] * len(cost)
    for i in order:
        load, r = heapq.heappop(heap)
        owner[i] = r
        heapq.heappush(heap, (load + cost[i], r))
    return np.asarray(owner, dtype=np.int64)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


_DPX = None


def dpx_peak():
    exe = os.path.join(ROOT, "tensorbranching.jl_b200", "dpx_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)}


def dpx_peak_once():
    global _DPX
    if _DPX is None:
        _DPX = dpx_peak()
    return _DPX


def python_slice_labels(branch, k):
    """k labels to index-slice, chosen on the host without libtbcuda (the reference arm must not touch the engine):
    the labels carried by most large intermediates (workloads.standin_host.big_label_histogram)."""
    from workloads import standin_host as H

    sc, _ = H.tree_complexity(branch.ixs, branch.tree)
    hist = H.big_label_histogram(branch.ixs, branch.tree, max(0, int(sc) - 6))
    return [l for l, _ in sorted(hist.items(), key=lambda kv: (-kv[1], kv[0]))[:k]]


def feasible_assignments(branch, labels):
    es = set(map(tuple, branch.edges))
    bad = [(i, j) for i in range(len(labels)) for j in range(i) if (min(labels[i], labels[j]), max(labels[i], labels[j])) in es]
    return [a for a in range(1 << len(labels)) if not any((a >> i) & 1 and (a >> j) & 1 for i, j in bad)]


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_value_type(branches):
    """the fastest value type of the CPU port that is exact for this workload: int16 for unit weights"""
    return "i16" if all(b.weights is None for b in branches) else "f32"


def cpu_reference_run_sliced(branches, budget_s, k, value_type="f32"):
    """CPU sample for workloads whose single branches are minutes of CPU work: index slices (k labels fixed) of the
    first sc-maximal branch, one slice per core at a time.  -> (Gop/s, cores, sample, seconds, (branch index, labels,
    assignments, values))"""
    from oracle import c_oracle as CO
    from workloads import standin_host as H

    scs = [H.tree_complexity(b.ixs, b.tree)[0] if b.nv else 0 for b in branches[:8]]
    bi = int(np.argmax(scs))
    br = branches[bi]
    labels = python_slice_labels(br, k)
    assign = feasible_assignments(br, labels)
    cores = host_cores()
    t = time.perf_counter()
    v0, o0, th = CO.contract_index_slices(br, labels, assign[:cores], value_type)
    dt0 = max(time.perf_counter() - t, 1e-6)
    n = int(min(len(assign), max(cores, cores * int(budget_s / dt0))))
    t = time.perf_counter()
    vals, ops, th = CO.contract_index_slices(br, labels, assign[:n], value_type)
    dt = time.perf_counter() - t
    sample = (f"{n} of the {len(assign)} feasible index slices (k={len(labels)} labels fixed) of branch {bi} of the workload, "
              f"{dt:.1f} s")
    return float(ops.sum()) / dt * 1e-9, th, sample, dt, (bi, labels, assign[:n], vals)


def cpu_reference_run(branches, budget_s, value_type="f32"):
    """Time the oracle's C/OpenMP port on a bounded sample (heaviest-first prefix would bias; take
    branches in order until the budget).  -> (Gop/s, cores, sample description, seconds, values)"""
    from oracle import c_oracle as CO

    flats = [None if b.nv == 0 else CO.flatten(b) for b in branches]
    # calibrate on a prefix that gives every thread several branches, then size the sample to the budget (and grow it
    # once more if it still finished far too early: a prefix of light branches underestimates the rate)
    cores = host_cores()
    n = min(len(flats), max(8, 4 * cores))
    vals = ops = th = None
    dt = 0.0
    for _ in range(3):
        t = time.perf_counter()
        vals, ops, th = CO.contract_batch(flats[:n], value_type)
        dt = max(time.perf_counter() - t, 1e-6)
        if n >= len(flats) or dt >= 0.4 * budget_s:
            break
        n = int(min(len(flats), max(n + 1, 0.8 * n * budget_s / dt)))
    return float(ops.sum()) / dt * 1e-9, th, f"first {n} of {len(flats)} branches of the workload, {dt:.1f} s", dt, vals


def cpu_leg(workload, branches, budget_s):
    """the CPU port on a bounded sample of the workload, in its best exact value type -> (dict, seconds, values to compare)"""
    from oracle import c_oracle as CO

    vt = cpu_value_type(branches)
    if workload in CPU_SLICE_K:
        g, cores, sample, dt, check = cpu_reference_run_sliced(branches, budget_s, CPU_SLICE_K[workload], vt)
    else:
        g, cores, sample, dt, check = cpu_reference_run(branches, budget_s, vt)
    return {"value": g, "unit": "Gop/s", "cores": cores, "kind": "port", "sample": sample,
            "value_type": {"i16": "int16 (exact for unit weights; the port's fastest type, 2x the SIMD lanes of f32)", "f32": "f32"}[vt],
            "isa": CO.simd()}, dt, check


def with_f32_weights(branches, seed=15):
    """the weighted variant the reference tests with (Float32.(1.0 .+ rand(n)), /root/reference/test/dynamic_ob.jl:15,36):
    every branch gets w = float32(1 + U[0,1)) on its own vertices (seeded); the engine then computes in Tropical{Float32}"""
    import copy
    rng = np.random.default_rng(seed)
    out = []
    for b in branches:
        c = copy.copy(b)
        if b.nv:
            c.weights = (1.0 + rng.random(b.nv)).astype(np.float32)
            c.r = float(np.float32(b.r))
        out.append(c)
    return out


def workload_config(name, max_branches, scaling, slice_k, weights="unit"):
    wl_kind, wl_p, sc_target, wl_mb = WORKLOADS[name]
    wl_mb = max_branches or wl_mb
    wdesc = "unit weights" if weights == "unit" else "Float32 weights 1 + U[0,1) (test/dynamic_ob.jl:15), Tropical{Float32} on the device"
    config = {"workload": f"{name}: {wl_kind} {wl_p}, sc_target={sc_target}, {wdesc}, stand-in host branching" +
                          (f", first {wl_mb} finished branches of the depth-first slicer" if wl_mb else ""),
              "l2": "per-step working set (arena + descriptors) exceeds the 126 MB L2; no explicit flush"}
    if slice_k > 0:
        config["index_slicing"] = f"every branch cut into 2^{slice_k} index slices (tb_suggest_slices); infeasible assignments are not units"
    config["sharding"] = ("weak scaling: every rank contracts its own copy of the unit list, no data-path collective, one "
                          "all-reduce(max) over the result vector" if scaling == "weak" else
                          "strong scaling: ONE unit list dealt to the ranks longest-first (LPT) by tropical ops (tb_estimate), every "
                          "rank compiles and contracts only its shard, no data-path collective, one all-reduce(max) over the "
                          "result vector")
    config["resident"] = ("`value`: compiled plans, their descriptors AND their work lists (instance arrays, tile offsets, launch "
                          "geometry) are resident on the device; a step replays the lists and runs every contraction again "
                          "(TB_NO_LIST_CACHE=1 rebuilds and uploads the lists per step).  `e2e` starts from host branch objects: "
                          "cost estimate, plan compilation, uploads, contraction, results back")
    return config


# ------------------------------------------------------------------------------------------------ tbcuda arm
class Ctx:
    """what a workload measurement needs from the process: rank, world, engine"""

    def __init__(self, rank, world, local_rank, eng, value_type, weights="unit"):
        self.rank, self.world, self.local_rank, self.eng, self.value_type = rank, world, local_rank, eng, value_type
        self.weights = weights


def run_workload(cx, name, scaling, steps, warmup, max_branches=None, slice_k=None, e2e_on=True, cpu_budget=15.0,
                 cpu_on=True, sample_clocks=True, min_timed_s=0.0):
    """measure one workload on the job's GPUs -> dict (the bench line without the process-level keys; None off rank 0)"""
    import torch
    import torch.distributed as dist

    import tbcuda
    from tbcuda.multi_gpu import slice_range

    rank, world, eng = cx.rank, cx.world, cx.eng
    slice_k = slice_k if slice_k is not None else DEFAULT_SLICE_K.get(name, 0)
    config = workload_config(name, max_branches, scaling, slice_k, cx.weights)

    def to_sliced(b):
        return tbcuda.SlicedBranch.from_parts(b.nv, b.edges, b.weights, b.ixs, b.tree, b.r)

    branches = None
    if rank == 0:
        branches = make_workload(name, max_branches)
    if world > 1:
        dist.barrier()
    if rank != 0:
        branches = make_workload(name, max_branches)
    config["workload_hash"] = workload_hash_status(name, branches, max_branches)
    if cx.weights == "f32":
        branches = with_f32_weights(branches)
    n_br = len(branches)
    sliced = [to_sliced(b) for b in branches]
    weak = scaling == "weak"
    copies = world if weak else 1  # weak: the job holds one copy of the unit list per rank

    # units of work: one per branch, or (index slicing) one per feasible assignment of the k sliced labels of a branch
    units = []  # (branch index, {label: value} or None)
    slice_labels = {}
    for i, s in enumerate(sliced):
        if s.code is None or slice_k <= 0:
            units.append((i, None))
            continue
        slice_labels[i] = tbcuda.suggest_slices(s, -1, slice_k)[0]
        for a in feasible_assignments(branches[i], slice_labels[i]):
            units.append((i, {l: (a >> q) & 1 for q, l in enumerate(slice_labels[i])}))
    n_units = len(units)
    ub = np.array([u[0] for u in units], dtype=np.int64)

    def all_sum(vec):
        if world > 1:
            t = torch.tensor(list(vec), dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            return t.cpu().numpy()
        return np.asarray(vec, dtype=np.float64)

    def estimate_costs():
        """tropical ops of every branch from tb_estimate (label-set pass only): each rank estimates a strided share,
        one all-reduce(sum) gives every rank the whole vector"""
        c = np.zeros(n_br)
        c[rank::world] = tbcuda.estimate_many(sliced[rank::world], max(1, host_cores() // world))
        return all_sum(c)

    t0 = time.perf_counter()
    if weak or world == 1:
        owner = np.full(n_units, rank, dtype=np.int64)
    else:
        ucost = estimate_costs()[ub] / float(1 << slice_k if slice_k > 0 else 1)  # the slices of one branch cost the same
        owner = lpt_shards(ucost, world)
    mine = np.nonzero(owner == rank)[0]
    # plans of THIS rank's units only
    my_plans = [tbcuda.Plan(sliced[units[i][0]], np.float32, engine=eng, fixed=units[i][1]) if sliced[units[i][0]].code is not None
                else None for i in mine]
    plan_s = time.perf_counter() - t0
    my_stats = [p.info() if p is not None else None for p in my_plans]
    my_sliced = [sliced[ub[i]] for i in mine]

    def my_sum(field):
        return float(sum(getattr(s, field) for s in my_stats if s))

    total_ops = float(all_sum([my_sum("ops")])[0])  # weak: every rank holds (and counted) a full copy
    r_vec = np.array([b.r for b in branches], dtype=np.float64)
    r_units = r_vec[ub]

    res_dev = torch.full((copies * n_units,), -float("inf"), dtype=torch.float64, device="cuda")
    mine_dev = torch.from_numpy(mine + (rank * n_units if weak else 0)).cuda()

    def per_branch(unit_vals, idx):
        out = np.full(n_br, -np.inf)
        if slice_k <= 0:
            out[ub[idx]] = unit_vals  # one unit per branch
        else:
            np.maximum.at(out, ub[idx], unit_vals)
        return out

    def gather_units(vals, idx_dev=None, idx=None):
        """every rank's unit values -> the per-branch vector on every rank: ONE all-reduce(max)"""
        idx_dev = mine_dev if idx_dev is None else idx_dev
        idx = mine if idx is None else idx
        if world > 1:
            res_dev.fill_(-float("inf"))
            res_dev[idx_dev] = torch.from_numpy(np.ascontiguousarray(vals, dtype=np.float64)).cuda()
            dist.all_reduce(res_dev, op=dist.ReduceOp.MAX)
            full = res_dev.cpu().numpy().reshape(copies, n_units)
            # weak: the copies are the same instance, so every GPU must have produced the same values (bit-exact)
            assert (full == full[0]).all(), "ranks disagree on the value of a unit"
            return per_branch(full[0], np.arange(n_units))
        return per_branch(vals, idx)

    my_batch = tbcuda.PlanBatch(my_plans, None)  # handle array marshalled once, not once per step
    r_mine32 = r_units[mine].astype(np.float32)

    def step_resident():
        vals, status, _ = eng.contract_plans(my_batch)
        # + r in element_type (Float32) arithmetic, as contract_slices does (src/dynamic_ob.jl:43)
        return gather_units((vals.astype(np.float32) + r_mine32).astype(np.float64))

    def step_e2e():
        """the public call from host objects: (strong) estimate costs + LPT, compile, upload, contract, read back"""
        if slice_k <= 0:
            if weak or world == 1:
                idx, idx_dev, shard = mine, mine_dev, my_sliced
            else:
                idx = np.nonzero(lpt_shards(estimate_costs(), world) == rank)[0]
                idx_dev = torch.from_numpy(idx).cuda()
                shard = [sliced[i] for i in idx]
            vals = tbcuda.contract_slices(shard, np.float32, True, engine=eng).astype(np.float64)
            return gather_units(vals, idx_dev, idx)
        # index slicing through the public call: every rank contracts its contiguous range of each branch's 2^k
        # assignments (tb_contract_sliced), then one all-reduce(max) over the per-branch vector
        out = np.full(n_br, -np.inf)
        first, count = (0, 1 << slice_k) if weak else slice_range(1 << slice_k, world, rank)
        for i, s in enumerate(sliced):
            if s.code is None:
                out[i] = r_vec[i]
            elif count > 0:
                out[i] = eng.contract_index_sliced(s, slice_labels[i], first, count)[2] + r_vec[i]
        if world > 1:
            t = torch.from_numpy(out).cuda()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out = t.cpu().numpy()
        return out

    def timed(fn, n_steps, n_warm, sampler=None):
        out = None
        t_w = time.perf_counter()
        for _ in range(n_warm):
            out = fn()
        torch.cuda.synchronize()
        if min_timed_s > 0:
            # short workloads (a step of a millisecond): keep warming up until the clocks have ramped, and time enough
            # steps that the timed region lasts min_timed_s (the step count is reported)
            t_w = time.perf_counter()
            n_w = 0
            while n_w < 2 or time.perf_counter() - t_w < 0.4 * min_timed_s:
                out = fn()
                n_w += 1
            torch.cuda.synchronize()
            est = max((time.perf_counter() - t_w) / n_w, 1e-5)
            n_steps = max(n_steps, int(np.ceil(min_timed_s / est)))
        timed.steps = n_steps
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps)]
        e0.record()
        for q in range(n_steps):
            out = fn()
            marks[q].record()
        e1.record()
        torch.cuda.synchronize()
        per_step = [(e0 if q == 0 else marks[q - 1]).elapsed_time(marks[q]) for q in range(n_steps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n_steps, out, sorted(per_step)

    # ---- the timed region (no per-launch events inside it: timestamps between the launches of a lane cost ~10 % on the
    #      launch-rich workloads, measured on cfg2)
    sampler = ClockSampler(cx.local_rank) if (rank == 0 and sample_clocks) else None
    ms_step, result, per_step_ms = timed(step_resident, steps, max(warmup, 3), sampler)
    steps_timed = timed.steps
    clocks = sampler.stop() if sampler else None
    launches_step = eng.last_timing()[1]
    dev_ms_last = eng.last_timing()[0]
    # ---- kernel times for the roofline: ONE more step right behind the timed ones, same configuration (all stream lanes
    #      concurrent), with CUDA events around every launch on its own lane ...
    eng.profile(1)
    step_resident()
    prof = eng.last_profile()              # per kind (sum of launch durations, launches)
    prof_union = eng.last_profile_union()  # per kind: time on the device (union of the lanes' launch intervals)
    # ... and a single-lane pass (launches serialised: per-launch durations free of overlap) beside it
    eng.profile(2)
    step_resident()
    prof1 = eng.last_profile()
    eng.profile(0)

    e2e = None
    if e2e_on:
        ms_e2e, result_e2e, e2e_steps_ms = timed(step_e2e, steps, 3)
        h2d, d2h = eng.last_transfers()
        hb = all_sum([h2d, d2h])
        if slice_k > 0:  # tb_last_transfers covers one call; a sliced step makes one call per branch
            hb = hb * sum(1 for s in sliced if s.code is not None)
        e2e = {"value": total_ops / (ms_e2e * 1e-3) * 1e-9, "unit": "Gop/s", "h2d_bytes_per_step": int(hb[0]),
               "d2h_bytes_per_step": int(hb[1]), "ms_per_step": ms_e2e, "steps": timed.steps, "warmup": 3,
               "ms_per_step_min_median_max": [round(e2e_steps_ms[0], 4), round(e2e_steps_ms[len(e2e_steps_ms) // 2], 4),
                                              round(e2e_steps_ms[-1], 4)],
               "host_breakdown_rank0": eng.last_host_breakdown(),
               "slices_per_s": copies * n_units / (ms_e2e * 1e-3),
               "vs_resident": ms_step / ms_e2e}
        assert np.array_equal(result_e2e, result)

    out = None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        dpx = dpx_peak_once()
        i16 = cx.value_type == "i16"
        f32 = cx.value_type == "f32"
        peak_gops = dpx.get("fadd_fmnmx_f32_Gops" if f32 else ("viaddmax_s16x2_Gops" if i16 else "viaddmax_s32_Gops"))
        # the dominant kernel: the persistent max-plus GEMM kernel; in the dataflow executor its ONE launch per wave also
        # runs the wave's generic steps on its consumer warps, so its ops are the GEMM + generic steps' ops
        n_gen_launches = prof["generic"][1]
        k_ops = my_sum("gemm_ops") + (0.0 if n_gen_launches else my_sum("generic_ops"))
        k_ms_union, k_ms_sum, k_launches = prof_union["gemm"], prof["gemm"][0], prof["gemm"][1]
        k_ms_1lane = prof1["gemm"][0]
        ach = k_ops / (k_ms_union * 1e-3) * 1e-9 if k_ms_union > 0 else None
        ach1 = k_ops / (k_ms_1lane * 1e-3) * 1e-9 if k_ms_1lane > 0 else None
        traffic = traffic_note = traffic_algo = None
        tp = os.path.join(ROOT, "profiles", f"latest_ncu_traffic_{name}.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_launch_mean")
            traffic_algo = tj.get("algorithmic_bytes_per_launch_mean")
            traffic_note = tj.get("note")
        kname = ("k_gemm2<float> (FADD + FMNMX)" if f32 else ("k_gemm2h (packed int16x2)" if i16 else "k_gemm2<int32>")) + \
                (", persistent dataflow kernel: the tiled max-plus GEMM tiles and the generic tiles of a whole wave"
                 if not n_gen_launches else ", one launch per dependency level")
        # per-node hybrid roofline: every non-fused step needs at least max(ops / op rate, algorithmic bytes / HBM
        # bandwidth), the fused subtrees their ops at the op rate; the sum over this rank's plans is a lower bound of its step
        hybrid = None
        if peak_gops:
            p_ops, bw = peak_gops * 1e9, peaks["hbm_gbs"] * 1e9
            t_ops = t_mem = 0.0
            n_mem = n_nodes = 0
            for pl in my_plans:
                if pl is None:
                    continue
                ops_n = np.exp2(np.frombuffer(pl.raw(6), dtype=np.float32).astype(np.float64))
                byt_n = np.frombuffer(pl.raw(7), dtype=np.float64)
                tc, tm = ops_n / p_ops, byt_n / bw
                mem = tm > tc
                t_ops += float(tc[~mem].sum())
                t_mem += float(tm[mem].sum())
                n_mem += int(mem.sum())
                n_nodes += int(mem.size)
            t_fused = my_sum("fused_ops") / p_ops
            bound_ms = (t_ops + t_mem + t_fused) * 1e3
            hybrid = {"bound_ms": bound_ms, "frac_whole_step": bound_ms / ms_step if ms_step > 0 else None,
                      "compute_bound_ms": t_ops * 1e3, "memory_bound_ms": t_mem * 1e3, "fused_ms": t_fused * 1e3,
                      "memory_bound_nodes": n_mem, "nodes": n_nodes,
                      "note": "sum over the non-fused steps of this rank's plans of max(ops / peak op rate, algorithmic bytes / measured "
                              "HBM copy bandwidth) + fused-subtree ops / peak op rate, divided by the timed step"}
        roofline = {"bound": ("fp32 FADD+FMNMX (two issues per tropical op" if f32 else
                              ("dpx-int16x2 (VIADDMNMX.S16x2" if i16 else "dpx-int32 (VIADDMNMX")) +
                             " issue rate; the semiring is (max,+), tensor cores do not apply)",
                    "kernel": kname, "achieved": ach, "peak": peak_gops, "unit": "Gop/s",
                    "frac": (ach / peak_gops) if (ach and peak_gops) else None,
                    "frac_note": "kernel ops / time the kernel was on the device in one step run right behind the timed steps in the "
                                 "timed configuration (all stream lanes concurrent; CUDA events around every launch on its own lane, "
                                 "union of the lanes' intervals).  frac_whole_step uses the timed steps themselves",
                    "frac_single_lane": (ach1 / peak_gops) if (ach1 and peak_gops) else None,
                    "frac_whole_step": (total_ops / world / (ms_step * 1e-3) * 1e-9 / peak_gops) if peak_gops else None,
                    "traffic": traffic, "traffic_algorithmic_bytes_same_launches": traffic_algo, "traffic_note": traffic_note,
                    "algorithmic_bytes_per_launch": my_sum("gemm_bytes") / max(1, k_launches),
                    "peak_source": "tensorbranching.jl_b200/dpx_peak microbenchmark run inside bench.py (register-resident VIADDMNMX, all SMs)",
                    "avg_launch_ms": k_ms_sum / max(1, k_launches), "launches": k_launches,
                    "kernel_ms_on_device": k_ms_union, "kernel_ms_sum_of_launches": k_ms_sum, "kernel_ms_single_lane": k_ms_1lane,
                    "share_of_step": {k: v[0] for k, v in prof1.items()},
                    "hybrid": hybrid,
                    "hbm": {"peak": peaks["hbm_gbs"], "peak_source": peak_src, "unit": "GB/s",
                            "achieved_whole_step": my_sum("algo_bytes") / (ms_step * 1e-3) * 1e-9}}
        out = {"value": total_ops / (ms_step * 1e-3) * 1e-9, "unit": "Gop/s", "ms_per_step": ms_step,
               "ms_per_step_median_rank0": per_step_ms[len(per_step_ms) // 2], "ms_per_step_max_rank0": per_step_ms[-1],
               "dtype": "f32" if f32 else ("int16x2" if i16 else "int32"), "config": config,
               "slices_per_s": copies * n_units / (ms_step * 1e-3), "branches": copies * n_br, "units": copies * n_units,
               "total_ops": total_ops, "mis": float(np.max(result)),
               "gpu_launches": int(launches_step * steps_timed * world),  # rank 0's launches per step x ranks (LPT shards are alike)
               "steps_timed": steps_timed,
               "launches_per_step": int(launches_step), "device_ms_last_step": dev_ms_last,
               "plan_compile_s_this_rank": plan_s, "clocks": clocks, "roofline": roofline}
        gold = golden_record(name)
        if gold is not None and max_branches is None and cx.weights == "unit":
            out["agrees_with_golden"] = bool(np.array_equal(result, np.asarray(gold["values"])))
        if e2e:
            out["e2e"] = e2e
        if cpu_on:
            import __graft_entry__ as G
            G.build()
            cb, dt, check = cpu_leg(name, branches, cpu_budget)
            if name in CPU_SLICE_K:
                bi, labels, assign, vals = check
                gv = eng.contract_index_sliced(sliced[bi], labels, 0, max(assign) + 1)[0]
                cb["agrees_with_gpu"] = bool(np.array_equal(gv[np.asarray(assign)], vals))
            else:
                n = len(check)
                # unit weights: integers, exact in any type; f32 weights: the same Float32 arithmetic as contract_slices
                want = (check + r_vec[:n]) if cx.weights == "unit" else (check.astype(np.float32) + r_vec[:n].astype(np.float32)).astype(np.float64)
                cb["agrees_with_gpu"] = bool(np.array_equal(want, result[:n]))
            out["cpu_baseline"] = cb
    for p in my_plans:
        if p is not None:
            p.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="tbcuda", choices=["tbcuda", "reference"])
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--max-branches", type=int, default=None)
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): ONE unit list sharded over the ranks longest-first by tropical ops; weak: every rank "
                         "contracts its own copy of the workload's unit list (per-GPU work fixed, N x units in the job)")
    ap.add_argument("--slice-k", type=int, default=None,
                    help="index-slice every branch into 2^k units (default: per workload, 0 except cfg3)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short cfg2 / cfg3 / cfg5 runs appended at N=1")
    ap.add_argument("--value-type", default="i16", choices=["i32", "i16"],
                    help="i16 (default, what value_type AUTO picks for this workload) = packed int16x2; i32 = plan flag NO_I16")
    ap.add_argument("--weights", default="unit", choices=["unit", "f32"],
                    help="f32: Float32 vertex weights 1 + U[0,1) (the reference's weighted tests): Tropical{Float32} kernels (K3)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    slice_k = args.slice_k if args.slice_k is not None else DEFAULT_SLICE_K.get(args.workload, 0)
    config = workload_config(args.workload, args.max_branches, args.scaling, slice_k, args.weights)

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        import __graft_entry__ as G
        G.build()
        branches = make_workload(args.workload, args.max_branches)
        config["workload_hash"] = workload_hash_status(args.workload, branches, args.max_branches)
        if args.weights == "f32":
            branches = with_f32_weights(branches)
        per_step = max(3.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
        gops = []
        cb = None
        for i in range(args.warmup + args.steps):
            cb, dt, _ = cpu_leg(args.workload, branches, per_step)
            if i >= args.warmup:
                gops.append((cb["value"], dt))
        value = float(np.mean([g for g, _ in gops]))
        ms = float(np.mean([d for _, d in gops])) * 1e3
        cb["value"] = value
        line = {"impl": "reference", "metric": "tropical contraction throughput", "value": value, "unit": "Gop/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "int16" if cpu_value_type(branches) == "i16" else "f32", "data": "synthetic",
                "config": config, "cpu_baseline": cb,
                "e2e": {"value": value, "unit": "Gop/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- tbcuda arm
    import torch
    import torch.distributed as dist

    import __graft_entry__ as G
    if rank == 0:
        G.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import tbcuda

    eng = tbcuda.Engine(local_rank, plan_flags=(tbcuda.TB_PLAN_NO_I16 if args.value_type == "i32" else 0),
                        host_threads=max(1, host_cores() // world))  # ranks share the host's cores for plan compilation
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    cx = Ctx(rank, world, local_rank, eng, "f32" if args.weights == "f32" else args.value_type, args.weights)

    main_res = run_workload(cx, args.workload, args.scaling, args.steps, args.warmup, args.max_branches, args.slice_k,
                            e2e_on=not args.no_e2e, cpu_budget=args.cpu_budget,
                            cpu_on=(not args.no_cpu_baseline) and world == 1)
    others = {}
    if world == 1 and not args.no_other_configs:
        for wl in OTHER_CONFIGS:
            if wl == args.workload:
                continue
            n_steps = min(args.steps, 10)
            r = run_workload(cx, wl, args.scaling, n_steps, 3, None, None, e2e_on=not args.no_e2e,
                             cpu_budget=min(args.cpu_budget, 4.0), cpu_on=not args.no_cpu_baseline, sample_clocks=False,
                             min_timed_s=0.5)
            others[wl] = {"value": r["value"], "unit": "Gop/s", "ms_per_step": r["ms_per_step"], "steps": r["steps_timed"],
                          "ms_per_step_median": r["ms_per_step_median_rank0"], "ms_per_step_max": r["ms_per_step_max_rank0"],
                          "units": r["units"], "slices_per_s": r["slices_per_s"], "launches_per_step": r["launches_per_step"],
                          "mis": r["mis"], "agrees_with_golden": r.get("agrees_with_golden"),
                          "kernel_frac_of_dpx_peak": r["roofline"]["frac"],
                          "kernel_frac_single_lane": r["roofline"]["frac_single_lane"],
                          "whole_step_frac_of_dpx_peak": r["roofline"]["frac_whole_step"],
                          "whole_step_frac_of_hybrid_roofline": (r["roofline"].get("hybrid") or {}).get("frac_whole_step"),
                          "hybrid_roofline_memory_bound_ms": (r["roofline"].get("hybrid") or {}).get("memory_bound_ms"),
                          "e2e": ({"value": r["e2e"]["value"], "ms_per_step": r["e2e"]["ms_per_step"],
                                   "vs_resident": r["e2e"]["vs_resident"], "steps": r["e2e"]["steps"],
                                   "host_breakdown": r["e2e"]["host_breakdown_rank0"]} if "e2e" in r else None),
                          "cpu_port_gops": r.get("cpu_baseline", {}).get("value"),
                          "agrees_with_cpu": r.get("cpu_baseline", {}).get("agrees_with_gpu"),
                          "workload": r["config"]["workload"]}
    if rank == 0:
        line = {"metric": "tropical contraction throughput", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "data": "synthetic"}
        line.update(main_res)
        line["dpx_peak"] = dpx_peak_once()
        if others:
            line["other_configs"] = others
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
