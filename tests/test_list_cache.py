"""Cached work lists of resident plans: contracting the same group of resident plans again (tb_contract_batch / tb_contract:
the loop body of contract_slices over pre-compiled plans, /root/reference/src/dynamic_ob.jl:38-46, as slice_dfs_lp-style
callers run it) replays the instance arrays and launch geometry kept on the device instead of rebuilding and uploading
them.  Every contraction is still executed; these tests check that a replay is never stale: other plan lists in between,
subsets, reordered lists, re-created plans (possibly at recycled addresses), arena growth, both executors, index-slice
assignments of one plan, and that a replay really uploads nothing."""
import gc

import numpy as np
import pytest

from helpers import golden_branches, load_golden, regular_root, to_sliced
from oracle import c_oracle as CO
from oracle import tropical_oracle as O
from workloads import standin_host as H

pytestmark = pytest.mark.gpu


def _golden(tb, engine, name="rr100_sc10_unit", flags=0):
    rec = load_golden(name + ".json")
    brs = golden_branches(rec)
    sl = [to_sliced(b) for b in brs]
    plans = [tb.Plan(s, flags=flags, engine=engine) if s.code is not None else None for s in sl]
    r = np.array([b.r for b in brs], dtype=np.float64)
    return plans, r, np.asarray(rec["values"], dtype=np.float64)


@pytest.mark.parametrize("which", ["default", "dataflow", "levelsync"])
def test_replays_equal_first_contraction_and_upload_nothing(tb, engine, engine_dataflow, engine_levelsync, which):
    eng = {"default": engine, "dataflow": engine_dataflow, "levelsync": engine_levelsync}[which]
    plans, r, want = _golden(tb, eng)
    v0, s0, _ = eng.contract_plans(plans, r)
    assert not s0.any() and np.array_equal(v0, want)
    h2d_first = eng.last_transfers()[0]
    # (the lists are kept when a group comes back: the second contraction still builds and uploads them)
    v1, s1, _ = eng.contract_plans(plans, r)
    assert not s1.any() and np.array_equal(v1, want) and 0 < eng.last_transfers()[0] < h2d_first
    for _ in range(3):
        v, s, mx = eng.contract_plans(plans, r)
        assert not s.any() and np.array_equal(v, want) and mx == want.max()
        assert eng.last_transfers()[0] == 0 < h2d_first  # descriptors resident, lists cached: nothing goes up
    # another r on the same plans (r is applied on the host)
    v, _, _ = eng.contract_plans(plans, r + 2.0)
    assert np.array_equal(v, want + 2.0)


def test_other_lists_in_between_subsets_and_orders(tb, engine):
    plans, r, want = _golden(tb, engine)
    n = len(plans)
    rng = np.random.default_rng(3)
    seqs = [list(range(n)), list(range(n // 2)), list(range(n // 2, n)), list(range(n)), list(range(n))[::-1],
            [int(i) for i in rng.permutation(n)], list(range(n // 2)), [0] * 5 + [n - 1] * 3, list(range(n))]
    for rep in range(2):
        for seq in seqs:
            v, s, _ = engine.contract_plans([plans[i] for i in seq], r[seq])
            assert not s.any() and np.array_equal(v, want[seq]), (rep, seq[:4])


def test_recreated_plans_never_hit_a_stale_entry(tb, engine):
    """plans destroyed and re-created (other branches, likely at recycled host / device addresses) between contractions"""
    rec_a, rec_b = load_golden("rr100_sc10_unit.json"), load_golden("ksg8x8_sc6.json")
    for rep in range(4):
        for rec in (rec_a, rec_b):
            brs = golden_branches(rec)
            sl = [to_sliced(b) for b in brs]
            plans = [tb.Plan(s, engine=engine) if s.code is not None else None for s in sl]
            r = np.array([b.r for b in brs], dtype=np.float64)
            for _ in range(2):
                v, s, _ = engine.contract_plans(plans, r)
                assert not s.any() and np.array_equal(v, np.asarray(rec["values"], dtype=np.float64))
            for p in plans:
                if p is not None:
                    p.close()
            del plans
            gc.collect()


def test_arena_growth_between_replays(tb):
    """a fresh engine with a small arena: light plans first, then a heavy plan that makes the arena grow (new base address
    or size => the cached lists of the light plans must not be replayed as they are), then the light plans again"""
    eng = tb.Engine(0, arena_bytes=0)
    try:
        plans, r, want = _golden(tb, eng)
        for _ in range(2):
            v, _, _ = eng.contract_plans(plans, r)
            assert np.array_equal(v, want)
        root = regular_root(150, 9)
        big = tb.Plan(to_sliced(root), engine=eng)
        exact = O.exact_mis_milp(root.nv, root.edges)
        assert eng.contract(big) == exact
        for _ in range(2):
            v, _, _ = eng.contract_plans(plans, r)
            assert np.array_equal(v, want)
            assert eng.contract(big) == exact
    finally:
        eng.close()


def test_single_plan_replays_with_intermediates_and_assignments(tb, engine):
    root = regular_root(60, 4)
    want = O.solve_slice(root, np.float64)
    p = tb.Plan(to_sliced(root), flags=tb.TB_PLAN_KEEP_INTERMEDIATES, engine=engine)
    for _ in range(3):
        assert engine.contract(p) == want
    node = [s.node for s in p.steps()][-1]
    _, data = engine.read_tensor(p, node)
    assert data.reshape(-1)[0] == want
    # index slicing: the assignments of one plan share every descriptor and differ in the leaf pool only
    labels = [3, 17, 40]
    base = tb.Plan(to_sliced(root), engine=engine, fixed={l: 0 for l in labels})
    asg = [{l: (a >> i) & 1 for i, l in enumerate(labels)} for a in range(8)]
    clones = [base.reassign(f) for f in asg]
    want_slices = np.array([O.solve_slice(root, np.float64, fixed=f) for f in asg])
    for _ in range(3):
        v, s, mx = engine.contract_plans(clones)
        assert not s.any() and np.array_equal(v, want_slices) and mx == want
