"""CPU tests of the oracle itself: pinned to independent exact solvers and the golden fixtures."""
import numpy as np
import pytest

from helpers import golden_branches, load_golden, regular_root
from oracle import tropical_oracle as O
from workloads import standin_host as H


@pytest.mark.parametrize("n,seed", [(10, 1), (14, 2), (16, 3)])
def test_oracle_vs_bruteforce(n, seed):
    root = regular_root(n, seed)
    assert O.solve_slice(root, np.float64) == O.exact_mis_bruteforce(root.nv, root.edges)


@pytest.mark.parametrize("n,seed", [(30, 4), (40, 5), (60, 6)])
def test_oracle_vs_clique_and_milp(n, seed):
    root = regular_root(n, seed)
    v = O.solve_slice(root, np.float32)
    assert v == O.exact_mis_milp(root.nv, root.edges)
    if n <= 40:
        assert v == O.exact_mis_clique(root.nv, root.edges)


def test_oracle_weighted_f32():
    rng = np.random.default_rng(0)
    nv, edges = H.random_regular_graph(30, 3, 7)
    w = (1 + rng.random(nv)).astype(np.float32)
    root = H.make_root(nv, edges, weights=w, seed=1)
    assert O.solve_slice(root, np.float32) == pytest.approx(O.exact_mis_milp(nv, edges, w), rel=1e-6)


def test_oracle_disconnected_and_isolated():
    # mirrors /root/reference/test/decompose.jl:62-85: 1- and 2-vertex components
    nv, edges = H.random_regular_graph(20, 3, 9)
    edges = [(u, v) for u, v in edges if 3 not in (u, v) and 7 not in (u, v) and 11 not in (u, v)] + [(3, 7)]
    root = H.make_root(nv, sorted(edges), seed=2)
    assert O.solve_slice(root, np.float32) == O.exact_mis_milp(nv, edges)


@pytest.mark.parametrize("name", ["rr100_sc10_unit", "rr100_sc10_f32", "rr30_disconnected", "ksg8x8_sc6", "ksg7x7_sc8_nokernel"])
def test_oracle_matches_golden(name):
    rec = load_golden(name + ".json")
    brs = golden_branches(rec)
    vals = O.contract_slices(brs, np.dtype(rec["element_type"]).type)
    assert np.array_equal(vals.astype(np.float64), np.asarray(rec["values"]))
    assert float(vals.max()) == pytest.approx(rec["exact"], rel=1e-6)


def test_branching_property_matches_reference_test():
    # /root/reference/test/slice.jl:32-33: max over contract_slices == unsliced contraction
    root = regular_root(60, 12)
    brs = H.slice_bfs(root, 6)
    assert len(brs) > 1
    assert float(O.contract_slices(brs, np.float32).max()) == float(O.solve_slice(root, np.float32))


def test_empty_graph_branch_is_r():
    # /root/reference/src/dynamic_ob.jl:39-40
    b = H.Branch(0, [], None, [], None, 7)
    assert O.contract_slices([b], np.float32)[0] == np.float32(7)
