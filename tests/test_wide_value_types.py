"""The two 8-byte value types (generic + fused kernels only):

* TB_VALUE_F64 -- Tropical{Float64}: `element_type = Float64` with real weights, which the reference's boundary accepts
  (solve_slice(branch, element_type, usecuda), /root/reference/src/dynamic_ob.jl:30) and its tests use off the path
  (/root/reference/test/decompose.jl:13-32).
* TB_VALUE_SIZE_CONFIG -- size + one optimal configuration per element, the configuration-enumerating half of the
  branching tables (branching_table(p, TensorNetworkSolver(), region), /root/reference/src/branch.jl:79; regions of at most
  n_max = 20 vertices, src/types.jl:10): for every boundary configuration of a region the best size AND a vertex set that
  attains it.

CPU tests interpret the compiled descriptors (tests/desc_interp.py); GPU tests go through the C ABI."""
import itertools
import struct

import numpy as np
import pytest

import desc_interp as DI
from helpers import align_to, device_tensor_as_ndarray, regular_root, to_sliced
from oracle import tropical_oracle as O
from workloads import standin_host as H


def _f64_root(n, seed):
    rng = np.random.default_rng(seed)
    nv, edges = H.random_regular_graph(n, 3, seed)
    w = 1.0 + rng.random(nv)  # Float64 weights with all 52 mantissa bits in use
    return H.make_root(nv, edges, weights=w, seed=seed)


@pytest.mark.parametrize("n,seed", [(12, 1), (40, 3), (70, 5)])
@pytest.mark.parametrize("flags", [0, 2, 8])
def test_float64_plan_interpreted(tb, n, seed, flags):
    root = _f64_root(n, seed)
    p = tb.Plan(to_sliced(root), np.float64, flags=flags)
    st = p.info()
    assert st.value_type == tb.TB_VALUE_F64 and st.n_gemm_steps == 0
    got, _ = DI.run_plan(p)
    assert got == O.solve_slice(root, np.float64)  # bit-exact: one rounding per a+b, max is exact
    assert float(np.float32(got)) != float(got)  # the answer really needs Float64


def _region(tb, n, seed, n_open):
    root = regular_root(n, seed)
    rng = np.random.default_rng(seed + 100)
    open_labels = sorted(int(v) for v in rng.choice(root.nv, size=n_open, replace=False))
    br = tb.SlicedBranch(tb.MISProblem(root.nv, root.edges, root.weights), tb.CompressedEinsum(root.ixs, open_labels, root.tree), 0)
    return root, br, open_labels


def _check_table(root, open_labels, labels, sizes, cfgs):
    """sizes == the oracle's open-boundary tensor; every configuration is an independent set of that size that agrees with
    its boundary configuration on the open vertices"""
    left, right = O.nested_to_postorder(root.tree, len(root.ixs))
    t, labs = O.contract_tree(root.ixs, left, right, None, np.float64, open_labels=tuple(open_labels))
    dl, darr = device_tensor_as_ndarray(labels, sizes)
    assert np.array_equal(align_to(dl, darr, list(labs)), np.asarray(t))
    adj = set(map(tuple, root.edges))
    for idx in range(len(sizes)):
        if not np.isfinite(sizes[idx]):
            assert cfgs[idx] == 0
            continue
        chosen = [v for v in range(root.nv) if (int(cfgs[idx]) >> v) & 1]
        assert len(chosen) == sizes[idx]
        assert not any((min(u, v), max(u, v)) in adj for u, v in itertools.combinations(chosen, 2))
        for bit, lab in enumerate(labels):  # boundary vertex lab is in the set iff bit `bit` of the configuration index
            assert ((idx >> bit) & 1) == ((int(cfgs[idx]) >> lab) & 1)


@pytest.mark.parametrize("n,seed,n_open", [(10, 1, 3), (16, 2, 4), (20, 3, 5), (24, 4, 6), (32, 6, 5)])
def test_size_config_table_interpreted(tb, n, seed, n_open):
    root, br, open_labels = _region(tb, n, seed, n_open)
    p = tb.Plan(br, value_type=tb.TB_VALUE_SIZE_CONFIG)
    st = p.info()
    assert st.value_type == tb.TB_VALUE_SIZE_CONFIG and st.root_rank == n_open and st.n_gemm_steps == 0
    _, arena = DI.run_plan(p)
    root_off = struct.unpack("4q", p.raw(5))[1]
    s = [x for x in p.steps() if x.rank_c == n_open and x.c_offset == root_off][-1]
    labels = [s.labels_c[i] for i in range(s.rank_c)]
    raw = arena[root_off:root_off + (1 << n_open)]
    _check_table(root, open_labels, labels, DI.to_float(raw, 5), DI.to_config(raw))


def test_size_config_needs_at_most_32_labels(tb):
    root = regular_root(40, 3)
    with pytest.raises(tb.TBError) as e:
        tb.Plan(to_sliced(root), value_type=tb.TB_VALUE_SIZE_CONFIG)
    assert e.value.code == -3


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", [(12, 1), (70, 5), (110, 7)])
def test_gpu_float64(tb, engine, n, seed):
    root = _f64_root(n, seed)
    got = tb.solve_slice(to_sliced(root), np.float64, True, engine=engine)
    assert got.dtype == np.float64 and got == O.solve_slice(root, np.float64)
    # every node, against the oracle's Float64 intermediates
    left, right = O.nested_to_postorder(root.tree, len(root.ixs))
    _, _, inter = O.contract_tree(root.ixs, left, right, np.asarray(root.weights), np.float64, keep_intermediates=True)
    p = tb.Plan(to_sliced(root), np.float64, flags=1, engine=engine)
    engine.contract(p)
    for s in p.steps():
        if s.node in inter:
            labels, data = engine.read_tensor(p, s.node)
            dl, darr = device_tensor_as_ndarray(labels, data)
            ol, oarr = inter[s.node]
            assert np.array_equal(align_to(dl, darr, ol), oarr), f"node {s.node}"
    # a Float64 branch inside a mixed contract_slices call
    brs = [to_sliced(root), to_sliced(_f64_root(30, 9))]
    vals = tb.contract_slices(brs, np.float64, True, engine=engine)
    assert vals[0] == got and vals[1] == O.solve_slice(_f64_root(30, 9), np.float64)


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,n_open", [(10, 1, 3), (20, 3, 5), (24, 4, 6), (32, 6, 8)])
def test_gpu_size_config_table(tb, engine, n, seed, n_open):
    root, br, open_labels = _region(tb, n, seed, n_open)
    p = tb.Plan(br, value_type=tb.TB_VALUE_SIZE_CONFIG, engine=engine)
    labels, sizes, cfgs = engine.contract_table(p)
    assert sorted(labels) == open_labels
    _check_table(root, open_labels, labels, sizes, cfgs)
    # the closed network: the scalar is the MIS size, through tb_contract as for any plan
    q = tb.Plan(to_sliced(root), value_type=tb.TB_VALUE_SIZE_CONFIG, engine=engine)
    assert engine.contract(q) == O.solve_slice(root, np.float64)


def _keep_by_definition(sizes):
    return O.mis_compactify_keep(sizes)


def test_oracle_compactify_small_cases():
    inf = np.inf
    # one boundary vertex: choosing it survives only if it gains something
    assert list(O.mis_compactify_keep([2.0, 2.0])) == [True, False]
    assert list(O.mis_compactify_keep([2.0, 3.0])) == [True, True]
    # two boundary vertices (index = b1 b0): 3 is dominated by 1, 2 is infeasible
    assert list(O.mis_compactify_keep([1.0, 2.0, -inf, 2.0])) == [True, True, False, False]
    # the empty configuration is infeasible: nothing dominates the singletons
    assert list(O.mis_compactify_keep([-inf, 1.0, 1.0, 1.0])) == [False, True, True, False]


@pytest.mark.gpu
@pytest.mark.parametrize("rank,seed", [(0, 0), (1, 1), (3, 2), (6, 3), (9, 4), (11, 5)])
def test_gpu_compactify_random_tables(tb, engine, rank, seed):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(0, 6, size=1 << rank).astype(np.float64)
    sizes[rng.random(sizes.size) < 0.3] = -np.inf  # infeasible boundary configurations
    assert np.array_equal(engine.compactify_table(sizes), _keep_by_definition(sizes))
    # ties count as dominated; an all-equal table keeps only the empty configuration
    flat = np.full(1 << rank, 2.0)
    want = np.zeros(1 << rank, dtype=bool)
    want[0] = True
    assert np.array_equal(engine.compactify_table(flat), want)


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,n_open", [(10, 1, 3), (20, 3, 5), (24, 4, 6), (32, 6, 8)])
def test_gpu_branching_table(tb, engine, n, seed, n_open):
    """the table the set-cover solver receives: surviving boundary configurations, one optimal set each"""
    root, br, open_labels = _region(tb, n, seed, n_open)
    p = tb.Plan(br, value_type=tb.TB_VALUE_SIZE_CONFIG, engine=engine)
    labels, rows = engine.branching_table(p)
    _, sizes, cfgs = engine.contract_table(p)
    keep = _keep_by_definition(sizes)
    assert [a for a, _, _ in rows] == list(np.nonzero(keep)[0])
    adj = {(min(u, v), max(u, v)) for u, v in root.edges}
    for a, size, mask in rows:
        chosen = [v for v in range(root.nv) if (mask >> v) & 1]
        assert len(chosen) == size == sizes[a]
        assert all((min(u, v), max(u, v)) not in adj for i, u in enumerate(chosen) for v in chosen[i + 1:])
        for q, l in enumerate(labels):  # the set agrees with its boundary configuration
            assert ((mask >> l) & 1) == ((a >> q) & 1)
    # every boundary configuration is covered by a surviving row that chooses a subset of its boundary vertices and is
    # at least as large (that is what makes dropping the dominated rows safe)
    for a in range(sizes.size):
        if sizes[a] > -np.inf:
            assert any((b & a) == b and sb >= sizes[a] for b, sb, _ in rows)
