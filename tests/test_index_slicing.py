"""Index slicing of one heavy branch: fixing k labels gives 2^k independent contractions of the
same tree whose max is the unsliced value.  CPU side: the compiled plans are run by the numpy descriptor
interpreter and compared with the oracle slice by slice; GPU side (-m gpu): through tb_contract_sliced."""
import itertools

import numpy as np
import pytest

import desc_interp as DI
from helpers import golden_branches, load_golden, regular_root, to_sliced
from oracle import tropical_oracle as O
from workloads import standin_host as H


def _assignments(labels):
    for a in range(1 << len(labels)):
        yield a, {l: (a >> i) & 1 for i, l in enumerate(labels)}


@pytest.mark.parametrize("n,seed,k", [(12, 1, 2), (30, 3, 3), (60, 5, 3)])
@pytest.mark.parametrize("flags", [0, 2, 8, 64, 64 | 4])
def test_sliced_plans_equal_oracle_slices(tb, n, seed, k, flags):
    root = regular_root(n, seed)
    br = to_sliced(root)
    labels, sc_after, tc_after = tb.suggest_slices(br, -1, k)
    assert len(labels) == k and len(set(labels)) == k
    full = O.solve_slice(root, np.float64)
    best = -np.inf
    for a, fixed in _assignments(labels):
        p = tb.Plan(br, flags=flags, fixed=fixed)
        st = p.info()
        assert st.sc == sc_after and st.tc == pytest.approx(tc_after)
        val, _ = DI.run_plan(p)
        assert val == O.solve_slice(root, np.float64, fixed=fixed), (a, fixed)
        best = max(best, val)
    assert best == full


@pytest.mark.parametrize("n,seed,k,flags", [(30, 3, 3, 0), (60, 5, 4, 0), (60, 5, 3, 64), (40, 9, 3, 8)])
def test_reassigned_plan_equals_compiled_plan(tb, n, seed, k, flags):
    """tb_plan_reassign: one compilation, the other assignments are copies with patched leaf-pool words -- the same
    descriptors byte for byte, the same pool as a fresh compilation, the oracle's slice value."""
    root = regular_root(n, seed)
    br = to_sliced(root)
    labels, _, _ = tb.suggest_slices(br, -1, k)
    base = tb.Plan(br, flags=flags, fixed={l: 0 for l in labels})
    for a, fixed in _assignments(labels):
        q = base.reassign(fixed)
        fresh = tb.Plan(br, flags=flags, fixed=fixed)
        for which in range(6):
            assert q.raw(which) == fresh.raw(which), (a, which)
        val, _ = DI.run_plan(q)
        assert val == O.solve_slice(root, np.float64, fixed=fixed)
    with pytest.raises(tb.TBError):
        tb.Plan(br).reassign({})  # not a sliced plan


def test_sliced_weighted_f32(tb):
    rng = np.random.default_rng(5)
    nv, edges = H.random_regular_graph(40, 3, 9)
    w = (1 + rng.random(nv)).astype(np.float32)
    root = H.make_root(nv, edges, weights=w, seed=9)
    br = to_sliced(root)
    labels, _, _ = tb.suggest_slices(br, -1, 3)
    vals = []
    for a, fixed in _assignments(labels):
        p = tb.Plan(br, fixed=fixed)
        assert p.info().value_type == 2
        v, _ = DI.run_plan(p)
        assert np.float32(v) == O.solve_slice(root, np.float32, fixed=fixed)
        vals.append(np.float32(v))
    assert max(vals) == O.solve_slice(root, np.float32)


def test_adjacent_fixed_vertices_and_whole_edge_fixed(tb):
    """both ends of an edge fixed: (1,1) is -inf, everything else contracts; a fixed isolated vertex is a scalar."""
    ixs = [(0,), (1,), (2,), (0, 1), (1, 2)]
    tree = (((0, 3), 1), ((4, 2)))
    b = tb.SlicedBranch(tb.MISProblem(3, [(0, 1), (1, 2)], None), tb.CompressedEinsum(ixs, (), tree), 0)
    sb = H.Branch(3, [(0, 1), (1, 2)], None, ixs, tree, 0)
    for x, y in itertools.product((0, 1), (0, 1)):
        fixed = {0: x, 1: y}
        v, _ = DI.run_plan(tb.Plan(b, fixed=fixed))
        assert v == O.solve_slice(sb, np.float64, fixed=fixed)
        assert (v == -np.inf) == (x == 1 and y == 1)
    # all labels fixed: a network of scalars
    v, _ = DI.run_plan(tb.Plan(b, fixed={0: 1, 1: 0, 2: 1}))
    assert v == 2.0


def test_suggest_slices_reaches_sc_target(tb):
    root = regular_root(100, 7)
    br = to_sliced(root)
    sc0 = tb.sc(br)
    labels, sc_after, tc_after = tb.suggest_slices(br, int(sc0) - 3, 16)
    assert sc_after <= sc0 - 3 and 3 <= len(labels) <= 16
    p = tb.Plan(br, fixed={l: 0 for l in labels})
    assert p.info().sc == sc_after and p.info().tc == pytest.approx(tc_after)
    # nothing to do when the target is already met
    assert tb.suggest_slices(br, int(sc0), 16)[0] == []


def test_sliced_argument_errors(tb):
    br = to_sliced(regular_root(12, 1))
    for bad in ({99: 0}, {-1: 1}, {0: 2}):
        with pytest.raises(tb.TBError) as e:
            tb.Plan(br, fixed=bad)
        assert e.value.code == -1


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rr100_sc10_unit", "rr100_sc10_f32", "ksg7x7_sc8_nokernel"])
def test_gpu_contract_sliced_matches_oracle_per_slice(tb, engine, name):
    rec = load_golden(name + ".json")
    et = np.dtype(rec["element_type"]).type
    brs = [b for b in golden_branches(rec) if b.nv >= 12][:3]
    assert brs
    for b in brs:
        br = to_sliced(b)
        labels, _, _ = tb.suggest_slices(br, -1, 4)
        vals, status, mx = engine.contract_index_sliced(br, labels, element_type=et)
        assert (status == 0).all()
        for a, fixed in _assignments(labels):
            assert et(vals[a]) == O.solve_slice(b, et, fixed=fixed), (a, fixed)
        assert et(mx) == O.solve_slice(b, et)
        # a sub-range (what one rank of a multi-GPU job computes)
        sub, _, smx = engine.contract_index_sliced(br, labels, first=5, count=6, element_type=et)
        assert np.array_equal(sub, vals[5:11]) and smx == vals[5:11].max()
        assert tb.solve_slice_index_sliced(br, labels, et, engine=engine) == tb.solve_slice(br, et, engine=engine)


@pytest.mark.gpu
def test_gpu_contract_sliced_large(tb, engine):
    """sc 24 root sliced 2^5 ways: max over slices == unsliced value == C oracle."""
    from oracle import c_oracle as CO
    root = regular_root(160, 3)
    br = to_sliced(root)
    want = CO.contract_slices([root], np.float32)[0]
    labels, sc_after, _ = tb.suggest_slices(br, -1, 5)
    vals, status, mx = engine.contract_index_sliced(br, labels)
    assert (status == 0).all() and np.float32(mx) == want
    assert sc_after < tb.sc(br)


@pytest.mark.gpu
def test_gpu_full_size_branch_over_40_labels(tb, engine):
    """BASELINE config-4 size: a branch of the n=500 / sc 28 workload whose heaviest node involves 42 labels
    (2^42 tropical ops in one contraction).  No CPU restatement reaches the whole branch; the size-independent property
    is that the max over its 2^10 index slices (nodes of <= 32 labels, a sample of them checked against the C oracle)
    equals the unsliced contraction."""
    from oracle import c_oracle as CO
    nv, edges = H.random_regular_graph(500, 3, 4)
    brs = H.slice_bfs(H.make_root(nv, edges, seed=4, ntrials=6), 28, max_branches=4)
    b = brs[3]
    br = to_sliced(b)
    p = tb.Plan(br, engine=engine)
    st = p.info()
    steps = p.steps()
    assert st.sc == 28 and max(s.n_m + s.n_n + s.n_b + s.n_k + s.n_ka + s.n_kb for s in steps) > 40
    whole = engine.contract(p)
    p.close()
    labels, sc_after, _ = tb.suggest_slices(br, -1, 10)
    vals, status, mx = engine.contract_index_sliced(br, labels)
    assert (status == 0).all() and mx == whole
    live = [a for a in range(1 << 10) if vals[a] > -np.inf][:16]
    cpu, _, _ = CO.contract_index_slices(b, labels, live)
    assert np.array_equal(cpu, vals[live])


def test_suggest_slices_edge_cases(tb):
    import ctypes as C
    from tbcuda import _lib as L
    from tbcuda.contract import _network_of
    lib = tb.load()
    root = regular_root(30, 3)
    br = to_sliced(root)
    # zero labels asked for: nothing picked, complexity of the unsliced tree reported
    labels, sc, tc = tb.suggest_slices(br, -1, 0)
    assert labels == [] and sc == tb.sc(br) and tc == pytest.approx(tb.tc(br))
    # open labels are never sliced
    open_labels = [0, 1, 2, 3]
    bo = tb.SlicedBranch(tb.MISProblem(root.nv, root.edges, None), tb.CompressedEinsum(root.ixs, open_labels, root.tree), 0)
    labels, _, _ = tb.suggest_slices(bo, -1, 6)
    assert not set(labels) & set(open_labels)
    # a label cannot be both open and fixed
    with pytest.raises(tb.TBError) as e:
        tb.Plan(bo, fixed={0: 1})
    assert e.value.code == -1
    # a network that already carries fixed labels is refused; a broken tree gets the tree error code
    net, _ = _network_of(br, np.float32, 0)
    bad = L.tb_network.from_buffer_copy(bytes(net))
    fl = np.asarray([0], dtype=np.int32)
    fv = np.asarray([0], dtype=np.uint8)
    bad.n_fixed, bad.fixed_labels, bad.fixed_values = 1, fl.ctypes.data_as(C.POINTER(C.c_int32)), fv.ctypes.data_as(C.POINTER(C.c_uint8))
    out = (C.c_int32 * 4)()
    assert lib.tb_suggest_slices(None, C.byref(bad), -1, 2, out, None, None) == -1
    left = br.code.node_left.copy()
    left[-1] = left[-2]  # a tensor used twice
    broken = L.tb_network.from_buffer_copy(bytes(net))
    broken.node_left = left.ctypes.data_as(C.POINTER(C.c_int32))
    assert lib.tb_suggest_slices(None, C.byref(broken), -1, 2, out, None, None) == -2
    assert lib.tb_suggest_slices(None, None, -1, 2, out, None, None) == -1


def test_sliced_leaves_get_private_pool_slots(tb):
    """the descriptors of a sliced plan do not depend on the assignment: only the leaf-pool section does"""
    root = regular_root(40, 9)
    br = to_sliced(root)
    labels, _, _ = tb.suggest_slices(br, -1, 3)
    plans = [tb.Plan(br, fixed={l: (a >> i) & 1 for i, l in enumerate(labels)}) for a in range(8)]
    for which in (1, 2, 3, 4, 5):
        assert len({p.raw(which) for p in plans}) == 1
    assert len({p.raw(0) for p in plans}) > 1
