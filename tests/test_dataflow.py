"""The dataflow executor: ONE persistent launch per wave runs every non-fused step of every level; a tile waits on
the completion counters of the instances producing its operands instead of on a kernel boundary.  Replaces the recursive
executor behind solve (/root/reference/src/dynamic_ob.jl:32).  The engine picks it per call (few or light plans); here
it is forced (TB_DATAFLOW=1) and checked against the oracle and against the level-synchronous executor (TB_LEVEL_SYNC=1),
which shares no scheduling code with it."""
import numpy as np
import pytest

from helpers import device_tensor_as_ndarray, golden_branches, load_golden, reduce_to, regular_root, to_sliced
from oracle import c_oracle as CO
from oracle import tropical_oracle as O
from workloads import standin_host as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["rr100_sc10_unit", "rr100_sc10_f32", "rr30_disconnected", "ksg8x8_sc6", "ksg7x7_sc8_nokernel"])
def test_both_executors_on_golden(tb, engine_dataflow, engine_levelsync, name):
    engine = engine_dataflow
    rec = load_golden(name + ".json")
    et = np.dtype(rec["element_type"]).type
    brs = [to_sliced(b) for b in golden_branches(rec)]
    a = tb.contract_slices(brs, et, True, engine=engine)
    b = tb.contract_slices(brs, et, True, engine=engine_levelsync)
    assert np.array_equal(a, b) and np.array_equal(a.astype(np.float64), np.asarray(rec["values"]))


@pytest.mark.parametrize("flags", [0, 8, 16, 64, 64 | 8, 4, 2])
def test_many_level_branches_and_launch_count(tb, engine_dataflow, engine_levelsync, flags):
    engine = engine_dataflow
    """3-regular n=140 cut to sc 14: branches with 4-8 dependency levels of generic and GEMM steps.  Same vector from both
    executors and the oracle; the dataflow executor needs at most 3 launches per wave (fused, persistent, finalize)."""
    nv, edges = H.random_regular_graph(140, 3, 7)
    brs = H.slice_bfs(H.make_root(nv, edges, seed=7), 14)[:200]
    want = CO.contract_slices(brs, np.float32)
    sliced = [to_sliced(b) for b in brs]
    plans_a = [tb.Plan(s, flags=flags, engine=engine) if s.code is not None else None for s in sliced]
    plans_b = [tb.Plan(s, flags=flags, engine=engine_levelsync) if s.code is not None else None for s in sliced]
    r = np.array([b.r for b in brs], dtype=np.float64)
    va, sa, _ = engine.contract_plans(plans_a, r)
    la = engine.last_timing()[1]
    vb, sb, _ = engine_levelsync.contract_plans(plans_b, r)
    lb = engine_levelsync.last_timing()[1]
    assert not sa.any() and not sb.any()
    assert np.array_equal(va, want.astype(np.float64)) and np.array_equal(vb, va)
    levels = max(p.info().n_levels for p in plans_a if p is not None)
    if levels >= 4:
        assert la * 2 <= lb, (la, lb, levels)
    for p in plans_a + plans_b:
        if p is not None:
            p.close()


def test_dataflow_every_node_weighted_f32_with_gemm_steps(tb, engine_dataflow):
    engine = engine_dataflow
    """K3 (FADD + FMNMX tiled GEMM) on a Float32-weighted instance (w = 1 + rand, /root/reference/test/dynamic_ob.jl:15)
    large enough (sc >= 18) to have tiled GEMM steps: every node bit-exact against the oracle's intermediates."""
    rng = np.random.default_rng(7)
    nv, edges = H.random_regular_graph(130, 3, 5)
    w = (1 + rng.random(nv)).astype(np.float32)
    root = H.make_root(nv, edges, weights=w, seed=5)
    left, right = O.nested_to_postorder(root.tree, len(root.ixs))
    _, _, inter = O.contract_tree(root.ixs, left, right, w, np.float32, keep_intermediates=True)
    p = tb.Plan(to_sliced(root), flags=1, engine=engine)
    st = p.info()
    assert st.value_type == 2 and st.sc >= 18 and st.n_gemm_steps > 0
    engine.contract(p)
    kinds = set()
    for s in p.steps():
        if s.node not in inter:
            continue
        labels, data = engine.read_tensor(p, s.node)
        dl, darr = device_tensor_as_ndarray(labels, data)
        ol, oarr = inter[s.node]
        assert np.array_equal(reduce_to(dl, darr, ol).astype(np.float32), oarr), f"node {s.node} kind {s.kind}"
        kinds.add(s.kind)
    assert 2 in kinds
    # and without KEEP_INTERMEDIATES (fused subtrees + arena reuse): the same root value
    q = tb.Plan(to_sliced(root), engine=engine)
    assert q.info().n_gemm_steps > 0
    assert np.float32(engine.contract(q)) == CO.contract_slices([root], np.float32)[0]


def test_solo_waves_do_not_race_with_the_lanes(tb):
    """ADVICE r1 (engine.cu: cross-stream arena race): a plan larger than a lane's arena partition runs alone on the whole
    arena; the side lanes of the NEXT group must wait for it.  Small arena + alternating big / small plans + tiny waves
    force solo waves between many lane waves; repeated to give a race a chance to show."""
    big = regular_root(150, 1000)          # sc ~ 20: a few MB of arena
    want_big = CO.contract_slices([big], np.float32)[0]
    rec = load_golden("rr100_sc10_unit.json")
    small = [b for b in golden_branches(rec) if b.nv][:24]
    want_small = np.asarray(rec["values"])[[i for i, b in enumerate(golden_branches(rec)) if b.nv][:24]]
    probe = tb.Plan(to_sliced(big))
    need = probe.info().arena_elems * 2 + 4096
    probe.close()
    import os
    os.environ["TB_DATAFLOW"] = os.environ.get("TB_SOLO_TEST_EXECUTOR", "1")
    try:
        eng = tb.Engine(0, arena_bytes=int(need * 1.5), max_wave=4)  # 4 lanes x cap < need: the big plan is always solo
    finally:
        del os.environ["TB_DATAFLOW"]
    brs, want = [], []
    for rep in range(6):
        brs += small[rep * 4:(rep + 1) * 4] + [big]
        want += list(want_small[rep * 4:(rep + 1) * 4]) + [want_big + big.r]
    sliced = [to_sliced(b) for b in brs]
    for _ in range(5):
        got = tb.contract_slices(sliced, np.float32, True, engine=eng)
        assert np.array_equal(got.astype(np.float64), np.asarray(want, dtype=np.float64))
    eng.close()


def test_profile_modes_agree(tb, engine_dataflow):
    engine = engine_dataflow
    rec = load_golden("rr100_sc10_unit.json")
    brs = [to_sliced(b) for b in golden_branches(rec)]
    for mode in (1, 2):
        engine.profile(mode)
        got = tb.contract_slices(brs, np.float32, True, engine=engine)
        prof = engine.last_profile()
        uni = engine.last_profile_union()
        assert np.array_equal(got.astype(np.float64), np.asarray(rec["values"]))
        assert prof["fused"][1] >= 1 and prof["finalize"][1] >= 1
        for k in prof:
            assert uni[k] <= prof[k][0] + 1e-3  # a union is never longer than the sum
    engine.profile(0)


def test_default_engine_picks_the_executor_per_call(tb, engine):
    """few plans or light plans -> dataflow (3 launches per wave); many heavy plans -> one launch per level and kind"""
    root = regular_root(150, 1000)  # one heavy plan, several levels
    p = tb.Plan(to_sliced(root), engine=engine)
    want = CO.contract_slices([root], np.float32)[0]
    assert engine.contract(p) == want
    assert engine.last_timing()[1] <= 3 and p.info().n_levels >= 3
    p.close()
