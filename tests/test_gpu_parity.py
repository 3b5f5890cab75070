"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libtbcuda.so and is compared bit-exactly with the oracle / golden fixtures."""
import numpy as np
import pytest

from helpers import align_to, device_tensor_as_ndarray, golden_branches, load_golden, regular_root, to_sliced
from oracle import tropical_oracle as O
from workloads import standin_host as H

pytestmark = pytest.mark.gpu

FLAGS = [0, 2, 4, 8, 16, 2 | 8, 2 | 4, 4 | 16, 64, 64 | 8, 64 | 2, 64 | 4, 64 | 16]


@pytest.mark.parametrize("name", ["rr100_sc10_unit", "rr100_sc10_f32", "rr30_disconnected", "ksg8x8_sc6", "ksg7x7_sc8_nokernel"])
def test_contract_slices_on_golden(tb, engine, name):
    rec = load_golden(name + ".json")
    et = np.dtype(rec["element_type"]).type
    brs = golden_branches(rec)
    got = tb.contract_slices([to_sliced(b) for b in brs], et, True, engine=engine)
    assert got.dtype == np.dtype(et)
    assert np.array_equal(got.astype(np.float64), np.asarray(rec["values"]))  # bit-exact, per branch
    assert float(got.max()) == pytest.approx(rec["exact"], rel=1e-6)


@pytest.mark.parametrize("n,seed", [(12, 1), (30, 3), (60, 5), (100, 7), (120, 9)])
@pytest.mark.parametrize("flags", FLAGS)
def test_single_plan_all_kernel_paths(tb, engine, n, seed, flags):
    root = regular_root(n, seed)
    p = tb.Plan(to_sliced(root), flags=flags, engine=engine)
    assert engine.contract(p) == O.solve_slice(root, np.float64)
    p.close()


@pytest.mark.parametrize("flags", [0, 2, 8])
def test_weighted_f32_bit_exact(tb, engine, flags):
    rng = np.random.default_rng(3)
    nv, edges = H.random_regular_graph(80, 3, 21)
    w = (1 + rng.random(nv)).astype(np.float32)
    root = H.make_root(nv, edges, weights=w, seed=3)
    p = tb.Plan(to_sliced(root), flags=flags, engine=engine)
    assert np.float32(engine.contract(p)) == O.solve_slice(root, np.float32)


def test_float64_weights_with_float32_element_type(tb, engine):
    # /root/reference/test/slice.jl:41-47: ws = ones(n) (Float64), element_type Float32
    nv, edges = H.random_regular_graph(40, 3, 2)
    root = H.make_root(nv, edges, weights=np.ones(nv), seed=2)
    got = tb.solve_slice(to_sliced(root), np.float32, True, engine=engine)
    assert got == O.solve_slice(root, np.float32) == O.exact_mis_milp(nv, edges)


@pytest.mark.parametrize("flags", [1, 1 | 8, 1 | 4, 1 | 64, 1 | 64 | 8])
def test_every_node_bit_exact(tb, engine, flags):
    """node-by-node: every intermediate tensor equals the oracle's (SURVEY 8c oracle plan (1))."""
    root = regular_root(70, 8)
    left, right = O.nested_to_postorder(root.tree, len(root.ixs))
    _, _, inter = O.contract_tree(root.ixs, left, right, None, np.float64, keep_intermediates=True)
    p = tb.Plan(to_sliced(root), flags=flags, engine=engine)
    engine.contract(p)
    kinds = set()
    for s in p.steps():
        if s.node not in inter:
            continue  # synthetic split-K partial tensor: no counterpart in the reference's tree
        labels, data = engine.read_tensor(p, s.node)
        dl, darr = device_tensor_as_ndarray(labels, data)
        ol, oarr = inter[s.node]
        assert np.array_equal(align_to(dl, darr, ol), oarr), f"node {s.node} kind {s.kind}"
        kinds.add(s.kind)
    if (flags & 64) and not (flags & 4):
        assert 2 in kinds  # the tiled GEMM kernel was exercised (the packed-int16 tile needs larger nodes)


def test_every_node_bit_exact_f32(tb, engine):
    rng = np.random.default_rng(5)
    nv, edges = H.random_regular_graph(60, 3, 31)
    w = (1 + rng.random(nv)).astype(np.float32)
    root = H.make_root(nv, edges, weights=w, seed=4)
    left, right = O.nested_to_postorder(root.tree, len(root.ixs))
    _, _, inter = O.contract_tree(root.ixs, left, right, w, np.float32, keep_intermediates=True)
    p = tb.Plan(to_sliced(root), flags=1, engine=engine)
    engine.contract(p)
    for s in p.steps():
        if s.node not in inter:
            continue
        labels, data = engine.read_tensor(p, s.node)
        dl, darr = device_tensor_as_ndarray(labels, data)
        ol, oarr = inter[s.node]
        assert np.array_equal(align_to(dl, darr, ol).astype(np.float32), oarr), f"node {s.node}"


def test_branching_property_config2_shape(tb, engine):
    """3-regular n=120 sliced to sc 12: every branch equals the oracle, the max equals the exact MIS
    (the property of /root/reference/test/slice.jl:32-33 and test/dynamic_ob.jl:20)."""
    nv, edges = H.random_regular_graph(120, 3, 2)
    root = H.make_root(nv, edges, seed=2)
    brs = H.slice_bfs(root, 12)
    got = tb.contract_slices([to_sliced(b) for b in brs], np.float32, True, engine=engine)
    want = O.contract_slices(brs, np.float32)
    assert np.array_equal(got, want)
    assert float(got.max()) == O.exact_mis_milp(nv, edges)


def test_large_tensors_sc20(tb, engine):
    """one n=150 branch at sc ~20: big GEMM / generic nodes, checked against the oracle root value."""
    root = regular_root(150, 1000)
    want = O.solve_slice(root, np.float64)
    for base in (0, 64):  # packed int16 (default) and int32
        p = tb.Plan(to_sliced(root), flags=base, engine=engine)
        st = p.info()
        assert st.sc >= 16 and st.n_gemm_steps > 0 and st.value_type == (1 if base else 3)
        assert engine.contract(p) == want
        # and with the GEMM kernel disabled: identical
        q = tb.Plan(to_sliced(root), flags=base | 4, engine=engine)
        assert engine.contract(q) == want


@pytest.mark.parametrize("n,seed", [(130, 5), (150, 1000)])
def test_split_k_and_every_node_large(tb, engine, n, seed):
    """sc ~22 branches: long reductions are split into partial + reduce steps; every ORIGINAL node of the
    tree still holds exactly the oracle's tensor."""
    from oracle import c_oracle as CO
    root = regular_root(n, seed)
    want = CO.contract_slices([root], np.float32)[0]
    vals = {}
    for flags in (0, 16, 4, 8, 64, 64 | 16, 64 | 8):
        p = tb.Plan(to_sliced(root), flags=flags, engine=engine)
        vals[flags] = engine.contract(p)
        p.close()
    assert all(v == want for v in vals.values()), (vals, want)
    p = tb.Plan(to_sliced(root), flags=64, engine=engine)
    assert any(s.node >= 2 * len(root.ixs) - 1 for s in p.steps())  # split-K really happened


@pytest.mark.parametrize("name", ["rr100_sc10_unit", "rr30_disconnected", "ksg8x8_sc6", "ksg7x7_sc8_nokernel"])
def test_contract_slices_packed_int16(tb, name):
    """K2: the packed int16x2 value type (plan flag PREFER_I16) returns the same per-branch vector."""
    rec = load_golden(name + ".json")
    brs = golden_branches(rec)
    eng = tb.Engine(0, plan_flags=tb.TB_PLAN_PREFER_I16)
    got = tb.contract_slices([to_sliced(b) for b in brs], np.float32, True, engine=eng)
    assert np.array_equal(got.astype(np.float64), np.asarray(rec["values"]))
    for b in brs[:3]:
        if b.nv:
            assert tb.Plan(to_sliced(b), engine=eng).info().value_type == 3
    # and the int32 value type on the same branches
    eng32 = tb.Engine(0, plan_flags=tb.TB_PLAN_NO_I16)
    got32 = tb.contract_slices([to_sliced(b) for b in brs], np.float32, True, engine=eng32)
    assert np.array_equal(got32, got)
    eng.close()
    eng32.close()


@pytest.mark.parametrize("n,seed", [(130, 5), (150, 1000)])
def test_packed_int16_large(tb, n, seed):
    from oracle import c_oracle as CO
    root = regular_root(n, seed)
    want = CO.contract_slices([root], np.float32)[0]
    eng = tb.Engine(0, plan_flags=tb.TB_PLAN_PREFER_I16)
    for flags in (0, 8, 16):
        p = tb.Plan(to_sliced(root), flags=flags, engine=eng)
        st = p.info()
        assert st.value_type == 3 and st.n_gemm_steps > 0
        assert eng.contract(p) == want
        p.close()
    eng.close()


def test_int16_falls_back_when_weights_do_not_fit(tb, engine):
    nv, edges = H.random_regular_graph(40, 3, 2)
    w = np.full(nv, 1000, dtype=np.int64)  # sum |w| = 40000 >= 8192 -> int32
    root = H.make_root(nv, edges, weights=w, seed=2)
    p = tb.Plan(to_sliced(root), engine=engine)
    assert p.info().value_type == 1  # AUTO falls back to int32
    assert engine.contract(p) == O.exact_mis_milp(nv, edges, w)


def test_gemm_v1_kernel_agrees(tb):
    """the non-persistent cp.async GEMM (TB_GEMM_V1=1) and the persistent TMA GEMM give identical tensors."""
    import os
    root = regular_root(120, 9)
    os.environ["TB_GEMM_V1"] = "1"
    try:
        e1 = tb.Engine(0)
    finally:
        del os.environ["TB_GEMM_V1"]
    e2 = tb.Engine(0)
    for flags in (1 | 64, 1 | 8 | 64):
        p1 = tb.Plan(to_sliced(root), flags=flags, engine=e1)
        p2 = tb.Plan(to_sliced(root), flags=flags, engine=e2)
        assert e1.contract(p1) == e2.contract(p2) == O.solve_slice(root, np.float64)
        for s in p1.steps():
            if s.kind == 2:
                l1, d1 = e1.read_tensor(p1, s.node)
                l2, d2 = e2.read_tensor(p2, s.node)
                assert l1 == l2 and np.array_equal(d1, d2)
    e1.close()
    e2.close()


def test_many_lanes_and_small_waves(tb):
    """multi-stream lanes with tiny waves: same per-branch vector as the oracle."""
    rec = load_golden("rr100_sc10_unit.json")
    brs = golden_branches(rec)
    eng = tb.Engine(0, max_wave=4)
    got = tb.contract_slices([to_sliced(b) for b in brs], np.float32, True, engine=eng)
    assert np.array_equal(got.astype(np.float64), np.asarray(rec["values"]))
    eng.close()


def test_large_tensors_sc24_vs_c_oracle(tb, engine):
    """sc = 24 (16 Mi-element tensors): root value equals the C oracle's, both value types."""
    from oracle import c_oracle as CO
    root = regular_root(160, 3)
    want = CO.contract_slices([root], np.float32)[0]
    for flags in (0, 64):
        p = tb.Plan(to_sliced(root), flags=flags, engine=engine)
        assert p.info().sc >= 24
        assert engine.contract(p) == want
        p.close()


def test_full_size_sc28_property(tb, engine):
    """BASELINE config-4 tensor size (sc = 28: 2^28-element intermediates, ~1.5 GB arena).  No CPU restatement
    finishes in seconds at this size, so the check is the size-independent property the reference's tests use:
    the contracted value equals the exact MIS (independent HiGHS MILP)."""
    nv, edges = H.random_regular_graph(180, 3, 21)
    root = H.make_root(nv, edges, seed=21)
    p = tb.Plan(to_sliced(root), engine=engine)
    st = p.info()
    assert st.sc >= 28 and st.arena_elems * 2 > 1e9
    assert engine.contract(p) == O.exact_mis_milp(nv, edges)


def test_mixed_value_types_in_one_batch(tb, engine):
    """one contract_slices call with packed-int16, int32 (weights too large for int16), f32 and empty branches."""
    rng = np.random.default_rng(11)
    brs, want = [], []
    for i, (n, kind) in enumerate([(40, "unit"), (50, "big"), (40, "f32"), (0, "empty"), (60, "unit"), (30, "f32")]):
        if kind == "empty":
            brs.append(tb.SlicedBranch(tb.MISProblem(0, [], None), None, 5))
            want.append(np.float32(5))
            continue
        nv, edges = H.random_regular_graph(n, 3, 100 + i)
        w = None if kind == "unit" else (np.full(nv, 1000, dtype=np.int64) if kind == "big" else (1 + rng.random(nv)).astype(np.float32))
        root = H.make_root(nv, edges, weights=w, seed=i)
        root.r = i
        brs.append(to_sliced(root))
        want.append(np.float32(O.solve_slice(root, np.float32) + np.float32(i)))
    got = tb.contract_slices(brs, np.float32, True, engine=engine)
    assert np.array_equal(got, np.asarray(want, dtype=np.float32))


def test_corner_cases(tb, engine):
    b1 = tb.SlicedBranch(tb.MISProblem(1, [], None), tb.CompressedEinsum([(0,)], (), None), 3)
    b2 = tb.SlicedBranch(tb.MISProblem(2, [], None), tb.CompressedEinsum([(0,), (1,)], (), (0, 1)), 0)
    w = np.array([2.5, 1.0, 4.0], dtype=np.float32)
    b3 = tb.SlicedBranch(tb.MISProblem(3, [(0, 1)], w), tb.CompressedEinsum([(0,), (1,), (2,), (0, 1)], (), ((0, (1, 3)), 2)), 0.5)
    b4 = tb.SlicedBranch(tb.MISProblem(0, [], None), None, 9)  # empty graph => r (src/dynamic_ob.jl:39-40)
    got = tb.contract_slices([b1, b2, b3, b4], np.float32, True, engine=engine)
    assert list(got) == [4.0, 2.0, 7.0, 9.0]
    assert tb.contract_slices([], np.float32, True, engine=engine).shape == (0,)


def test_batch_api_and_max(tb, engine):
    rec = load_golden("rr100_sc10_unit.json")
    brs = golden_branches(rec)
    plans = [tb.Plan(to_sliced(b), engine=engine) if b.nv else None for b in brs]
    r = np.array([b.r for b in brs], dtype=np.float64)
    vals, status, mx = engine.contract_plans(plans, r)
    assert np.array_equal(vals, np.asarray(rec["values"])) and not status.any()
    assert mx == rec["exact"]
    # idempotence: a second run over resident plans gives the same vector
    vals2, _, _ = engine.contract_plans(plans, r)
    assert np.array_equal(vals, vals2)
    ms, launches = engine.last_timing()
    assert launches > 0 and ms > 0
    # the same list marshalled once (PlanBatch): identical results on every call
    batch = tb.PlanBatch(plans, r)
    for _ in range(3):
        v3, st3, mx3 = engine.contract_plans(batch)
        assert np.array_equal(v3, vals) and not st3.any() and mx3 == mx


@pytest.mark.parametrize("rank", [0, 1, 3, 7, 11, 12, 13, 20, 22])
def test_permute_bits(tb, engine, rank):
    rng = np.random.default_rng(rank)
    x = rng.integers(-1000, 1000, size=1 << rank, dtype=np.int32)
    # reversal, identity (no bit moves: the tile is filled up with the low bits), a rotation, random permutations
    perms = [list(range(rank))[::-1], list(range(rank)), [(i + 5) % max(rank, 1) for i in range(rank)],
             list(rng.permutation(rank)), list(rng.permutation(rank))]
    for perm in perms:
        perm = [int(v) for v in perm]
        got = engine.permute_bits(x, perm)
        dst = np.arange(1 << rank, dtype=np.int64)
        src = np.zeros_like(dst)
        for i, pbit in enumerate(perm):
            src |= ((dst >> i) & 1) << pbit
        assert np.array_equal(got, x[src])


def test_interleaved_calls_reuse_pooled_plans(tb):
    """tb_contract_networks keeps the plan objects of a call and compiles the next call's branches into them: calls with
    different branch lists (other sizes, value types, empty graphs) must not see anything of their predecessors"""
    eng = tb.Engine(0)
    sets = []
    for name in ["rr100_sc10_unit", "ksg8x8_sc6", "rr100_sc10_f32", "rr30_disconnected"]:
        rec = load_golden(name + ".json")
        sets.append((np.dtype(rec["element_type"]).type, [to_sliced(b) for b in golden_branches(rec)], np.asarray(rec["values"])))
    for rep in range(3):
        for et, brs, want in sets + sets[::-1]:
            half = brs[: max(1, len(brs) // (rep + 1))]  # shrinking lists: some pooled objects stay unused
            got = tb.contract_slices(half, et, True, engine=eng)
            assert np.array_equal(got.astype(np.float64), want[: len(half)])
    # the streaming entry point shares the pool
    et, brs, want = sets[0]
    with tb.BranchStream(eng, capacity=len(brs), element_type=et) as st:
        st.push(brs[: len(brs) // 2])
        st.push(brs[len(brs) // 2:])
        vals = st.finish()
    assert np.array_equal(np.asarray(vals, dtype=np.float64), want)
    eng.close()
