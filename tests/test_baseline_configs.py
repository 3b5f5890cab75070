"""Every BASELINE.json config at FULL size, pinned by committed values (tests/golden/baseline_cfg*.json, written ONCE by
`python tests/golden/make_golden.py baseline` from the C oracle).  The property being pinned is the one of
/root/reference/test/slice.jl:32-33: contract_slices returns, per branch, exactly what the CPU contraction returns.

CPU part (-m "not gpu"): the branch lists bench.py contracts are the ones the goldens were made from (sha256 over their
canonical content: the tracked generator, not whatever cache file happens to ship), and the oracle reproduces a sample.
GPU part (-m gpu): bit-equality of the engine on all of cfg1 / cfg2 / cfg3 (+ its 2^3 index slices) / cfg4 / cfg5,
through both executors (dataflow, level-synchronous) and the engine's own per-call choice."""
import os

import numpy as np
import pytest

import bench
from helpers import load_golden, to_sliced
from workloads import standin_host as H

CONFIGS = ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]


def _workload(name):
    rec = load_golden(f"baseline_{name}.json")
    brs = bench.make_workload(name)
    return rec, brs


@pytest.mark.parametrize("name", CONFIGS)
def test_benched_branch_list_is_the_golden_one(name):
    rec, brs = _workload(name)
    assert len(brs) == rec["n"] == len(rec["values"])
    assert H.branch_list_hash(brs) == rec["hash"]


@pytest.mark.parametrize("name,count", [("cfg1", 46), ("cfg2", 96), ("cfg3", 1), ("cfg5", 64)])
def test_c_oracle_reproduces_a_sample_of_the_golden_values(name, count):
    from oracle import c_oracle as CO
    rec, brs = _workload(name)
    sample = brs[:count]
    for vt in ("f32", "i16"):
        got = CO.contract_slices(sample, np.float64, vt)
        assert np.array_equal(got, np.asarray(rec["values"][:count])), vt


@pytest.mark.gpu
@pytest.mark.parametrize("name", CONFIGS)
def test_gpu_equals_golden_at_full_size(tb, engine, engine_dataflow, engine_levelsync, name):
    rec, brs = _workload(name)
    sliced = [to_sliced(b) for b in brs]
    want = np.asarray(rec["values"])
    for eng in (engine, engine_dataflow, engine_levelsync):  # the engine's own choice, and both executors forced
        if name == "cfg4" and eng is engine:
            continue  # 1.5 s of device time per pass: the default engine picks the level-synchronous executor here
        got = tb.contract_slices(sliced, np.float32, True, engine=eng)
        assert np.array_equal(got.astype(np.float64), want)
        assert float(got.max()) == rec["mis"]


@pytest.mark.gpu
def test_gpu_cfg3_index_slices_equal_golden(tb, engine):
    rec, brs = _workload("cfg3")
    s = to_sliced(brs[0])
    vals, status, mx = engine.contract_index_sliced(s, rec["sliced_labels"])
    want = np.array([-np.inf if v is None else v - brs[0].r for v in rec["slice_values"]])
    assert not status.any()
    assert np.array_equal(vals, want)
    assert mx + brs[0].r == rec["values"][0]


@pytest.mark.gpu
def test_gpu_cfg4_resident_plans_equal_golden(tb, engine):
    """the resident-plan path bench.py times (tb_contract_batch over a PlanBatch), per branch"""
    rec, brs = _workload("cfg4")
    sliced = [to_sliced(b) for b in brs[:8]]
    plans = [tb.Plan(s, np.float32, engine=engine) for s in sliced]
    r = np.array([b.r for b in brs[:8]], dtype=np.float64)
    vals, status, _ = engine.contract_plans(tb.PlanBatch(plans, r))
    assert not status.any() and np.array_equal(vals, np.asarray(rec["values"][:8]))
    for p in plans:
        p.close()
