"""Property tests (hypothesis): random graphs, RANDOM contraction trees (not the greedy order: outer products, labels
private to one operand, very unbalanced shapes), unit / integer / float32 weights, optional index slicing and open
labels.  The compiled plan, run by the numpy descriptor interpreter, must equal the oracle, and the oracle must equal
the brute-force MIS (the invariant the reference's tests pin: test/utils.jl:34,37,61, test/decompose.jl:57-85)."""
import os

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import desc_interp as DI
from helpers import to_sliced
from oracle import tropical_oracle as O
from workloads import standin_host as H


@st.composite
def networks(draw):
    big = bool(os.environ.get("TB_HYP_BIG"))  # stress runs: larger, denser graphs
    nv = draw(st.integers(2, 14 if big else 11))
    pairs = [(u, v) for u in range(nv) for v in range(u + 1, nv)]
    edges = sorted(draw(st.sets(st.sampled_from(pairs), max_size=min(len(pairs), (3 if big else 2) * nv))))
    kind = draw(st.sampled_from(["unit", "int", "f32"]))
    if kind == "unit":
        w = None
    elif kind == "int":
        w = np.asarray(draw(st.lists(st.integers(0, 9), min_size=nv, max_size=nv)), dtype=np.int64)
    else:
        w = np.asarray(draw(st.lists(st.floats(0.5, 4.0, width=32), min_size=nv, max_size=nv)), dtype=np.float32)
    ixs = H.mis_ixs(nv, edges)
    # random binary tree: merge two random entries until one is left
    items = list(range(len(ixs)))
    order = draw(st.permutations(items))
    pool = list(order)
    picks = draw(st.lists(st.integers(0, 10 ** 6), min_size=len(pool), max_size=len(pool)))
    q = 0
    while len(pool) > 1:
        i = picks[q % len(picks)] % len(pool)
        a = pool.pop(i)
        j = picks[(q + 1) % len(picks)] % len(pool)
        b = pool.pop(j)
        pool.append((a, b))
        q += 2
    tree = pool[0]
    flags = draw(st.sampled_from([0, 2, 4, 8, 16, 64, 2 | 8]))
    return nv, edges, w, ixs, tree, flags


SETTINGS = dict(max_examples=int(os.environ.get("TB_HYP_EXAMPLES", "150")), derandomize=not os.environ.get("TB_HYP_RANDOM"), deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])


@settings(**SETTINGS)
@given(networks())
def test_random_tree_plan_equals_oracle_and_bruteforce(tb, net):
    nv, edges, w, ixs, tree, flags = net
    b = H.Branch(nv, edges, w, ixs, tree, 0)
    et = np.float32 if (w is not None and w.dtype == np.float32) else np.float64
    want = O.solve_slice(b, et)
    val, _ = DI.run_plan(tb.Plan(to_sliced(b), flags=flags))
    assert et(val) == want
    if et is np.float64:
        assert want == O.exact_mis_bruteforce(nv, edges, w)


@settings(**SETTINGS)
@given(networks(), st.data())
def test_random_tree_index_slices(tb, net, data):
    nv, edges, w, ixs, tree, flags = net
    b = H.Branch(nv, edges, w, ixs, tree, 0)
    et = np.float32 if (w is not None and w.dtype == np.float32) else np.float64
    k = data.draw(st.integers(1, min(3, nv)))
    labels = data.draw(st.permutations(list(range(nv))))[:k]
    best = -np.inf
    for a in range(1 << k):
        fixed = {l: (a >> i) & 1 for i, l in enumerate(labels)}
        want = O.solve_slice(b, et, fixed=fixed)
        val, _ = DI.run_plan(tb.Plan(to_sliced(b), flags=flags, fixed=fixed))
        assert et(val) == want, fixed
        best = max(best, want)
    assert best == O.solve_slice(b, et)


@settings(**SETTINGS)
@given(networks(), st.data())
def test_random_tree_open_labels(tb, net, data):
    import struct
    from helpers import align_to, device_tensor_as_ndarray
    nv, edges, w, ixs, tree, flags = net
    et = np.float32 if (w is not None and w.dtype == np.float32) else np.float64
    k = data.draw(st.integers(1, min(3, nv)))
    open_labels = sorted(data.draw(st.permutations(list(range(nv))))[:k])
    br = tb.SlicedBranch(tb.MISProblem(nv, edges, w), tb.CompressedEinsum(ixs, open_labels, tree), 0)
    p = tb.Plan(br, flags=flags)
    st_ = p.info()
    _, arena = DI.run_plan(p)
    root_off = struct.unpack("4q", p.raw(5))[1]
    s = [x for x in p.steps() if x.rank_c == k and x.c_offset == root_off][-1]
    labels = [s.labels_c[i] for i in range(s.rank_c)]
    data_ = DI.to_float(arena[root_off:root_off + (1 << k)], st_.value_type)
    dl, darr = device_tensor_as_ndarray(labels, data_)
    left, right = O.nested_to_postorder(tree, len(ixs))
    wv = None if w is None else np.asarray(w).astype(et)
    t, labs = O.contract_tree(ixs, left, right, wv, et, open_labels=tuple(open_labels))
    assert sorted(dl) == open_labels
    assert np.array_equal(align_to(dl, darr, list(labs)).astype(et), np.asarray(t))


def _random_network(rng, nv_max=12):
    """numpy-seeded twin of the hypothesis strategy above (for the GPU batch test)."""
    nv = int(rng.integers(2, nv_max + 1))
    pairs = [(u, v) for u in range(nv) for v in range(u + 1, nv)]
    m = int(rng.integers(0, min(len(pairs), 2 * nv) + 1))
    edges = sorted(pairs[i] for i in rng.choice(len(pairs), size=m, replace=False)) if m else []
    kind = ("unit", "int", "f32")[int(rng.integers(0, 3))]
    w = None if kind == "unit" else (rng.integers(0, 10, size=nv).astype(np.int64) if kind == "int"
                                     else (0.5 + 3.5 * rng.random(nv)).astype(np.float32))
    ixs = H.mis_ixs(nv, edges)
    pool = [int(i) for i in rng.permutation(len(ixs))]
    while len(pool) > 1:
        a = pool.pop(int(rng.integers(0, len(pool))))
        b = pool.pop(int(rng.integers(0, len(pool))))
        pool.append((a, b))
    return H.Branch(nv, edges, w, ixs, pool[0], float(rng.integers(0, 5)))


import pytest  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_gpu_random_trees_batch(tb, engine, seed):
    """300 random networks with random (not greedy) trees and mixed weight kinds in ONE contract_slices call per
    element type: every value equals the oracle's, which equals the brute-force optimum."""
    rng = np.random.default_rng(seed)
    nets = [_random_network(rng) for _ in range(300)]
    for et, sel in ((np.float64, [b for b in nets if b.weights is None or b.weights.dtype != np.float32]),
                    (np.float32, [b for b in nets if b.weights is not None and b.weights.dtype == np.float32])):
        got = tb.contract_slices([to_sliced(b) for b in sel], np.float32 if et is np.float32 else np.float64, True, engine=engine)
        want = np.array([et(O.solve_slice(b, et) + et(b.r)) for b in sel])
        assert np.array_equal(got.astype(et), want)
        if et is np.float64:
            exact = np.array([O.exact_mis_bruteforce(b.nv, b.edges, b.weights) + b.r for b in sel])
            assert np.array_equal(want, exact)
    # index slices and open boundaries of a few of them
    for b in nets[:25]:
        if b.nv < 4:
            continue
        br = to_sliced(b)
        labels = [int(x) for x in rng.permutation(b.nv)[:2]]
        et = np.float32 if (b.weights is not None and b.weights.dtype == np.float32) else np.float64
        vals, status, mx = engine.contract_index_sliced(br, labels, element_type=np.float32)
        want = np.array([O.solve_slice(b, et, fixed={l: (a >> i) & 1 for i, l in enumerate(labels)}) for a in range(4)])
        assert np.array_equal(vals.astype(et), want.astype(et))
