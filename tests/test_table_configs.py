"""ALL optimal configurations of the branching tables: `branching_table(p, TensorNetworkSolver(), region)`
(/root/reference/src/branch.jl:79, default solver /root/reference/src/types.jl:46, regions of at most n_max = 20 vertices
src/types.jl:10) contracts the region's network with the ConfigsMax element type [upstream GenericTensorNetworks], so every
row of the table lists every optimal vertex set of its boundary configuration.

CPU: the oracle's restatement of that set algebra along the tree (`contract_tree_configs`) against plain enumeration
(`table_configs_bruteforce`).  GPU: tb_table_configs through the C ABI against both, and Engine.branching_table(all_configs)."""
import numpy as np
import pytest

from helpers import regular_root
from oracle import tropical_oracle as O


def _region(tb, n, seed, n_open, weights=None):
    root = regular_root(n, seed)
    rng = np.random.default_rng(seed + 100)
    open_labels = sorted(int(v) for v in rng.choice(root.nv, size=n_open, replace=False))
    w = root.weights if weights is None else weights(root.nv, rng)
    br = tb.SlicedBranch(tb.MISProblem(root.nv, root.edges, w), tb.CompressedEinsum(root.ixs, open_labels, root.tree), 0)
    return root, br, open_labels, w


def _int_weights(nv, rng):
    return rng.integers(1, 4, size=nv).astype(np.int32)


def _real_weights(nv, rng):
    return (1.0 + rng.random(nv)).astype(np.float64)


def _tied_real_weights(nv, rng):
    # few distinct real values: plenty of exact ties between different vertex sets
    return rng.choice(np.array([1.0, 1.5, 2.25]), size=nv)


@pytest.mark.parametrize("n,seed,n_open", [(8, 1, 2), (10, 1, 3), (12, 2, 4), (14, 3, 5), (14, 4, 0)])
@pytest.mark.parametrize("weights", [None, _int_weights])
def test_oracle_set_algebra_equals_enumeration(tb, n, seed, n_open, weights):
    root, _, open_labels, w = _region(tb, n, seed, n_open, weights)
    left, right = O.nested_to_postorder(root.tree, len(root.ixs))
    labs, res = O.contract_tree_configs(root.ixs, left, right, w, open_labels)
    sizes, rows = O.table_configs_bruteforce(root.nv, root.edges, w, open_labels)
    # the sizes are the open-boundary tensor of the tropical contraction
    t, tl = O.contract_tree(root.ixs, left, right, None if w is None else np.asarray(w, dtype=np.float64), np.float64,
                            open_labels=tuple(open_labels))
    assert tuple(tl) == tuple(open_labels)
    for a in range(1 << n_open):
        key = tuple((a >> i) & 1 for i in range(n_open))
        assert res[key][0] == sizes[a] == np.asarray(t)[key]
        assert sorted(res[key][1]) == rows[a]


def test_oracle_known_tables():
    # a path u - x - v with the end vertices open: the interior vertex is chosen only when both ends are out
    sizes, rows = O.table_configs_bruteforce(3, [(0, 1), (1, 2)], None, [0, 2])
    assert list(sizes) == [1.0, 1.0, 1.0, 2.0]
    assert rows == [[0b010], [0b001], [0b100], [0b101]]
    # a triangle plus a pendant pair, nothing open: all optimal sets of the whole region in one row
    sizes, rows = O.table_configs_bruteforce(5, [(0, 1), (1, 2), (0, 2), (3, 4)], None, [])
    assert list(sizes) == [2.0] and len(rows[0]) == 6


def _check_against_oracle(root, w, labels, sizes, row_off, cfgs, keep=None):
    want_sizes, want_rows = O.table_configs_bruteforce(root.nv, root.edges, w, labels)
    assert np.array_equal(sizes, want_sizes)
    assert row_off[0] == 0 and row_off[-1] == len(cfgs)
    for a in range(len(want_rows)):
        got = [int(c) for c in cfgs[row_off[a]:row_off[a + 1]]]
        if keep is not None and not keep[a]:
            assert got == []
        else:
            assert got == want_rows[a], f"row {a}"  # same sets AND the documented order (ascending)


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,n_open", [(6, 1, 0), (8, 1, 2), (10, 1, 3), (14, 3, 5), (16, 2, 4), (18, 5, 6), (20, 3, 5)])
@pytest.mark.parametrize("weights", [None, _int_weights, _real_weights, _tied_real_weights])
def test_gpu_table_configs(tb, engine, n, seed, n_open, weights):
    root, br, open_labels, w = _region(tb, n, seed, n_open, weights)
    labels = list(reversed(open_labels))  # any bit order of the rows
    sizes, row_off, cfgs = engine.table_configs(br, labels)
    _check_against_oracle(root, w, labels, sizes, row_off, cfgs)
    ms, launches = engine.last_timing()
    assert launches == 7 and ms > 0  # init, optimum, sizes, count, scan, row offsets, write
    if n <= 14 and weights is not _real_weights and weights is not _tied_real_weights:
        # the reference's algebra along the tree gives the same rows
        left, right = O.nested_to_postorder(root.tree, len(root.ixs))
        _, res = O.contract_tree_configs(root.ixs, left, right, w, labels)
        for a in range(1 << n_open):
            key = tuple((a >> i) & 1 for i in range(n_open))
            assert sorted(res[key][1]) == [int(c) for c in cfgs[row_off[a]:row_off[a + 1]]]


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,n_open", [(10, 1, 3), (20, 3, 5), (24, 4, 6), (26, 6, 8)])
def test_gpu_branching_table_all_configs(tb, engine, n, seed, n_open):
    """contraction (sizes) -> mis_compactify -> all optimal sets of the surviving rows; the enumerated row optima must equal
    the contracted sizes (branching_table raises otherwise); the one-configuration table is a member of every row"""
    root, br, open_labels, w = _region(tb, n, seed, n_open)
    p = tb.Plan(br, value_type=tb.TB_VALUE_SIZE_CONFIG, engine=engine)
    labels, rows = engine.branching_table(p, all_configs=True)
    _, one_rows = engine.branching_table(p)
    _, sizes, _ = engine.contract_table(p)
    keep = O.mis_compactify_keep(sizes)
    assert [a for a, _, _ in rows] == [a for a, _, _ in one_rows] == list(np.nonzero(keep)[0])
    adj = [0] * root.nv
    for u, v in root.edges:
        adj[u] |= 1 << v
        adj[v] |= 1 << u
    for (a, size, masks), (_, _, one) in zip(rows, one_rows):
        assert masks == sorted(set(masks)) and one in masks
        for m in masks:
            assert bin(m).count("1") == size == sizes[a]
            assert not any((m >> v) & 1 and adj[v] & m for v in range(root.nv))
            assert all(((m >> l) & 1) == ((a >> q) & 1) for q, l in enumerate(labels))
    if n <= 20:
        _, want_rows = O.table_configs_bruteforce(root.nv, root.edges, None, labels)
        assert [masks for _, _, masks in rows] == [want_rows[a] for a, _, _ in rows]
    # the same table from a sizes-only plan (any tropical value type)
    q = tb.Plan(br, engine=engine)
    l2, rows2 = engine.branching_table(q, all_configs=True)
    order = {l: i for i, l in enumerate(l2)}
    remap = lambda a: sum(((a >> i) & 1) << order[l] for i, l in enumerate(labels))  # noqa: E731
    assert sorted((remap(a), s, tuple(m)) for a, s, m in rows) == sorted((a, s, tuple(m)) for a, s, m in rows2)


@pytest.mark.gpu
def test_gpu_table_configs_large_region_and_errors(tb, engine):
    # 30 vertices, 6 open: 2^24 interior configurations per row, 2^30 vertex sets filtered
    root, br, open_labels, w = _region(tb, 30, 7, 6)
    sizes, row_off, cfgs = engine.table_configs(br, open_labels)
    left, right = O.nested_to_postorder(root.tree, len(root.ixs))
    t, tl = O.contract_tree(root.ixs, left, right, None, np.float64, open_labels=tuple(open_labels))
    want = np.asarray(t).transpose(list(range(len(tl)))[::-1]).reshape(-1)  # index bit i = open_labels[i]
    assert np.array_equal(sizes, want)
    adj = [0] * root.nv
    for u, v in root.edges:
        adj[u] |= 1 << v
        adj[v] |= 1 << u
    assert len(cfgs) == row_off[-1] > 0
    for a in range(1 << 6):
        row = cfgs[row_off[a]:row_off[a + 1]]
        assert (len(row) > 0) == np.isfinite(sizes[a])
        assert np.all(np.diff(row.astype(np.int64)) > 0)
        for m in map(int, row[:50]):
            assert bin(m).count("1") == sizes[a] and not any((m >> v) & 1 and adj[v] & m for v in range(root.nv))
    # keep flags drop rows
    keep = np.zeros(1 << 6, dtype=np.uint8)
    keep[5] = 1
    s2, off2, c2 = engine.table_configs(br, open_labels, keep)
    assert np.array_equal(s2, sizes) and np.array_equal(c2, cfgs[row_off[5]:row_off[6]])
    assert off2[5] == 0 and off2[6] == off2[-1] == len(c2)
    # errors: a region of more than 32 vertices, a repeated boundary label
    big = regular_root(40, 3)
    bb = tb.SlicedBranch(tb.MISProblem(big.nv, big.edges, big.weights), tb.CompressedEinsum(big.ixs, [0, 1], big.tree), 0)
    with pytest.raises(tb.TBError) as e:
        engine.table_configs(bb, [0, 1])
    assert e.value.code == -3
    with pytest.raises(tb.TBError) as e:
        engine.table_configs(br, [1, 1])
    assert e.value.code == -1


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,n_open", [(8, 1, 0), (10, 1, 3), (16, 2, 4), (20, 3, 5), (20, 8, 8)])
@pytest.mark.parametrize("weights", [None, _int_weights])
def test_gpu_region_table_single_call(tb, engine, n, seed, n_open, weights):
    """tb_branching_table = optimum pass + mis_compactify + all optimal configurations in one call: equal to the oracle's
    table (enumeration + the all-pairs restatement of mis_compactify) and to the chained calls through the contraction"""
    root, br, open_labels, w = _region(tb, n, seed, n_open, weights)
    sizes, keep, rows = engine.region_table(br, open_labels)
    want_sizes, want_rows = O.table_configs_bruteforce(root.nv, root.edges, w, open_labels)
    want_keep = O.mis_compactify_keep(want_sizes)
    assert np.array_equal(sizes, want_sizes) and np.array_equal(keep, want_keep)
    assert rows == [(a, want_sizes[a], want_rows[a]) for a in np.nonzero(want_keep)[0]]
    assert engine.last_timing()[1] == 6 + n_open + 1 + 1  # + mis_compactify: one stage per boundary vertex, keep flags
    if weights is None:
        p = tb.Plan(br, engine=engine)
        labels, rows2 = engine.branching_table(p, all_configs=True)
        order = {l: i for i, l in enumerate(open_labels)}
        remap = lambda a: sum(((a >> i) & 1) << order[l] for i, l in enumerate(labels))  # noqa: E731
        assert sorted((remap(a), s, tuple(m)) for a, s, m in rows2) == sorted((a, s, tuple(m)) for a, s, m in rows)


@pytest.mark.gpu
def test_gpu_region_table_larger_than_first_guess(tb, engine):
    """13 disjoint edges, nothing open: 2^13 optimal sets in one row (more than the mirror's first buffer: one retry)"""
    from workloads import standin_host as H
    edges = [(2 * i, 2 * i + 1) for i in range(13)]
    root = H.make_root(26, edges, seed=1)
    br = tb.SlicedBranch(tb.MISProblem(26, edges, None), tb.CompressedEinsum(root.ixs, [], root.tree), 0)
    sizes, keep, rows = engine.region_table(br, [])
    assert list(sizes) == [13.0] and list(keep) == [True]
    (a, size, masks), = rows
    assert a == 0 and size == 13 and len(masks) == 1 << 13 and masks == sorted(set(masks))
    assert all(bin(m).count("1") == 13 and not (m & (m >> 1) & 0x1555555) for m in masks)
    # the C call itself: a short buffer fails with the total written
    import ctypes as C
    from tensorbranching_lib import L
    from tbcuda.contract import _network_of
    net, _ = _network_of(br, None, 0)
    lab = np.zeros(1, dtype=np.int32)
    off = np.zeros(2, dtype=np.int64)
    buf = np.zeros(16, dtype=np.uint32)
    total = C.c_int64()
    rc = L.load().tb_branching_table(engine.handle, C.byref(net), lab.ctypes.data_as(C.POINTER(C.c_int32)), 0, None, None,
                                     off.ctypes.data_as(C.POINTER(C.c_int64)), buf.ctypes.data_as(C.POINTER(C.c_uint32)), 16,
                                     C.byref(total))
    assert rc == L.TB_ERR_BAD_ARGUMENT and total.value == 1 << 13 and list(off) == [0, 1 << 13]


@pytest.mark.gpu
def test_gpu_region_tables_batched(tb, engine):
    """many regions in the same launches (different sizes, ranks and weight kinds, a region without boundary, one with
    nothing but boundary): every region's table equals its own single call and the oracle"""
    specs = [(10, 1, 3, None), (16, 2, 4, _int_weights), (8, 1, 0, None), (20, 3, 5, None), (6, 2, 6, None), (14, 3, 5, _real_weights),
             (18, 5, 6, _tied_real_weights), (12, 2, 4, None), (20, 8, 8, _int_weights)] + [(10 + 2 * (i % 3), 20 + i, i % 5, None) for i in range(40)]
    regions = [_region(tb, n, seed, n_open, w) for n, seed, n_open, w in specs]
    got = engine.region_tables([r[1] for r in regions], [r[2] for r in regions])
    ms, launches = engine.last_timing()
    assert launches == 6 + 8 + 1 + 1 and len(got) == len(specs)
    for (root, br, open_labels, w), (sizes, keep, rows) in zip(regions, got):
        want_sizes, want_rows = O.table_configs_bruteforce(root.nv, root.edges, w, open_labels)
        want_keep = O.mis_compactify_keep(want_sizes)
        assert np.array_equal(sizes, want_sizes) and np.array_equal(keep, want_keep)
        assert rows == [(a, want_sizes[a], want_rows[a]) for a in np.nonzero(want_keep)[0]]
    one = engine.region_table(regions[3][1], regions[3][2])
    assert np.array_equal(one[0], got[3][0]) and np.array_equal(one[1], got[3][1]) and one[2] == got[3][2]
    assert engine.region_tables([], []) == []
    # a bad region fails the call with its error code
    with pytest.raises(tb.TBError) as e:
        engine.region_tables([regions[0][1], regions[1][1]], [regions[0][2], [1, 1]])
    assert e.value.code == -1


def test_oracle_set_algebra_random_graphs_and_trees():
    """the restatement of the reference's ConfigsMax algebra follows ANY tree of ANY graph: random sparse graphs (isolated
    vertices, several components), random boundary sets, trees from the stand-in optimiser with different seeds"""
    from workloads import standin_host as H
    rng = np.random.default_rng(11)
    for trial in range(25):
        nv = int(rng.integers(3, 12))
        m = int(rng.integers(0, 2 * nv))
        edges = sorted({(int(min(u, v)), int(max(u, v))) for u, v in rng.integers(0, nv, size=(m, 2)) if u != v})
        w = None if trial % 3 == 0 else rng.integers(1, 4, size=nv).astype(np.int32)
        root = H.make_root(nv, edges, weights=w, seed=trial)
        n_open = int(rng.integers(0, min(nv, 5) + 1))
        open_labels = [int(v) for v in rng.choice(nv, size=n_open, replace=False)]
        left, right = O.nested_to_postorder(root.tree, len(root.ixs))
        _, res = O.contract_tree_configs(root.ixs, left, right, w, open_labels)
        sizes, rows = O.table_configs_bruteforce(nv, edges, w, open_labels)
        for a in range(1 << n_open):
            key = tuple((a >> i) & 1 for i in range(n_open))
            assert res[key][0] == sizes[a], (trial, a)
            assert sorted(res[key][1]) == rows[a], (trial, a)
        # every vertex set of the region sits in at most one row, and a row is never empty when its size is finite
        allsets = [s for r in rows for s in r]
        assert len(allsets) == len(set(allsets))
        assert all((len(r) > 0) == bool(np.isfinite(sizes[a])) for a, r in enumerate(rows))
