"""CPU dry run of bench.py's tbcuda arm: the engine, torch.cuda and the timing events are replaced by stand-ins (values
come from the oracle) so that every line of main() executes without a GPU -- guards the JSON contract and catches
NameErrors / wrong keys before the driver's GPU run does.  The numbers it prints mean nothing."""
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _FakeEvent:
    def __init__(self, enable_timing=True):
        self.t = 0.0

    def record(self):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


class _FakeEngine:
    """implements what bench.py calls, with oracle values"""

    def __init__(self, device=0, **kw):
        self.handle = None

    def set_stream(self, s):
        pass

    def _value(self, plan):
        from oracle import tropical_oracle as O
        from workloads import standin_host as H
        br = plan._keep[0]
        fixed = None
        if len(plan._keep) > 2:
            fixed = {int(l): int(v) for l, v in zip(plan._keep[2], plan._keep[3])}
        hb = H.Branch(br.p.nv, br.p.edges, None, br.code.ixs, None, 0)
        left, right = br.code.node_left, br.code.node_right
        root, _ = O.contract_tree(br.code.ixs, list(left), list(right), None, np.float64, fixed=fixed)
        return float(np.asarray(root).reshape(-1)[0])

    def contract_plans(self, batch, r=None):
        vals = np.array([(self._value(p) if p is not None else 0.0) for p in batch.plans]) + (batch.r if batch.r is not None else 0.0)
        vals = np.asarray(vals, dtype=np.float64).reshape(len(batch.plans))
        return vals, np.zeros(len(vals), dtype=np.int32), float(vals.max()) if len(vals) else -np.inf

    def contract_index_sliced(self, branch, labels, first=0, count=None, element_type=np.float32, flags=0):
        import tbcuda
        count = (1 << len(labels)) - first if count is None else count
        vals = []
        for a in range(first, first + count):
            p = tbcuda.Plan(branch, fixed={l: (a >> i) & 1 for i, l in enumerate(labels)})
            vals.append(self._value(p))
        vals = np.array(vals)
        return vals, np.zeros(len(vals), dtype=np.int32), float(vals.max())

    def profile(self, mode=1):
        pass

    def last_profile(self):
        return {"fused": (0.1, 1), "generic": (0.0, 0), "gemm": (0.2, 1), "finalize": (0.01, 1)}

    def last_profile_union(self):
        return {"fused": 0.1, "generic": 0.0, "gemm": 0.2, "finalize": 0.01}

    def last_timing(self):
        return 0.2, 3

    def last_transfers(self):
        return 1000, 100

    def last_host_breakdown(self):
        return {"compile_ms": 0.0, "total_ms": 0.0}

    def close(self):
        pass


@pytest.mark.parametrize("argv", [["--workload", "cfg1", "--steps", "2", "--warmup", "1", "--cpu-budget", "0.2", "--scaling", "weak"],
                                  ["--workload", "cfg1", "--steps", "2", "--slice-k", "1", "--no-cpu-baseline", "--scaling", "weak",
                                   "--no-other-configs"],
                                  ["--workload", "cfg1", "--steps", "1", "--value-type", "i32", "--no-e2e", "--no-other-configs"]])
def test_bench_tbcuda_arm_dry_run(monkeypatch, argv):
    import torch

    sys.path.insert(0, ROOT)
    import bench
    import tbcuda
    from oracle import tropical_oracle as O

    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: type("S", (), {"cuda_stream": 0})())
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    real_full = torch.full
    monkeypatch.setattr(torch, "full", lambda *a, **k: real_full(*a, **{q: v for q, v in k.items() if q != "device"}))
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{q: v for q, v in k.items() if q != "device"}))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(tbcuda, "Engine", _FakeEngine)

    def fake_contract_slices(branches, element_type=np.float32, usecuda=True, engine=None):
        eng = _FakeEngine()
        return np.array([(eng._value(tbcuda.Plan(b)) if b.code is not None else 0.0) + b.r for b in branches]).astype(element_type)

    monkeypatch.setattr(tbcuda, "contract_slices", fake_contract_slices)
    monkeypatch.setattr(bench, "dpx_peak", lambda: {"viaddmax_s16x2_Gops": 35000.0, "viaddmax_s32_Gops": 18000.0})
    monkeypatch.setattr(bench, "_DPX", None)
    # the runs appended at N=1 (`other_configs`) on a tiny stand-in instead of cfg2 / cfg3 / cfg5
    monkeypatch.setitem(bench.WORKLOADS, "tiny", ("regular", dict(n=40, d=3, seed=5), 8, None))
    monkeypatch.setattr(bench, "OTHER_CONFIGS", ("tiny",))
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.main()
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "gpu_launches", "clocks", "roofline", "units", "branches", "mis"):
        assert key in line, key
    assert line["scaling"] == ("weak" if "weak" in argv else "strong") and line["n_gpus"] == 1  # strong is the default
    if "--no-other-configs" not in argv:
        oc = line["other_configs"]["tiny"]
        assert set(oc) >= {"value", "ms_per_step", "kernel_frac_of_dpx_peak", "e2e", "agrees_with_cpu"} and oc["agrees_with_cpu"]
    assert line["config"]["workload_hash"].startswith("ok")
    assert line["agrees_with_golden"] is True
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert "workload" in line["config"] and "sharding" in line["config"]
    if "--no-e2e" not in argv:
        assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    if "--no-cpu-baseline" not in argv:
        assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and line["cpu_baseline"]["agrees_with_gpu"]
    assert line["mis"] == 45.0  # cfg1: exact MIS of the seeded n=100 instance, through the oracle stand-in


def _world2_worker(rank, world, port, out_dir, scaling):
    """one rank of a 2-rank dry run: the same stand-ins, the process group on gloo instead of nccl"""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    import tbcuda

    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.current_stream = lambda *a: type("S", (), {"cuda_stream": 0})()
    torch.cuda.Event = _FakeEvent
    real_full, real_tensor = torch.full, torch.tensor
    torch.full = lambda *a, **k: real_full(*a, **{q: v for q, v in k.items() if q != "device"})
    torch.tensor = lambda *a, **k: real_tensor(*a, **{q: v for q, v in k.items() if q != "device"})
    torch.Tensor.cuda = lambda self, *a, **k: self
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, **k: real_init("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    tbcuda.Engine = _FakeEngine
    bench.dpx_peak = lambda: {"viaddmax_s16x2_Gops": 35000.0, "viaddmax_s32_Gops": 18000.0}
    bench._DPX = None

    def fake_contract_slices(branches, element_type=np.float32, usecuda=True, engine=None):
        eng = _FakeEngine()
        return np.array([(eng._value(tbcuda.Plan(b)) if b.code is not None else 0.0) + b.r for b in branches]).astype(element_type)

    tbcuda.contract_slices = fake_contract_slices
    sys.argv = ["bench.py", "--gpus", str(world), "--workload", "cfg1", "--steps", "2", "--warmup", "1", "--scaling", scaling,
                "--no-cpu-baseline"]
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.main()
    if rank == 0:
        with open(os.path.join(out_dir, "line.json"), "w") as f:
            f.write(buf.getvalue().strip().splitlines()[-1])


@pytest.mark.parametrize("scaling", ["weak", "strong"])
def test_bench_two_rank_dry_run(tmp_path, scaling):
    import socket

    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_world2_worker, args=(2, port, str(tmp_path), scaling), nprocs=2, join=True)
    line = json.loads(open(tmp_path / "line.json").read())
    assert line["n_gpus"] == 2 and line["scaling"] == scaling and line["mis"] == 45.0
    assert line["units"] == (92 if scaling == "weak" else 46) and line["branches"] == line["units"]
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
