"""The C/OpenMP restatement (oracle/c/tropical_ref.c: the checker at sizes numpy cannot reach and the CPU baseline of
bench.py) pinned against the numpy restatement, the golden fixtures and the exact solvers."""
import numpy as np
import pytest

from helpers import golden_branches, load_golden, regular_root
from oracle import c_oracle as CO
from oracle import tropical_oracle as O
from workloads import standin_host as H


@pytest.mark.parametrize("name", ["rr100_sc10_unit", "rr100_sc10_f32", "rr30_disconnected", "ksg8x8_sc6", "ksg7x7_sc8_nokernel"])
def test_c_oracle_on_golden(name):
    rec = load_golden(name + ".json")
    et = np.dtype(rec["element_type"]).type
    brs = golden_branches(rec)
    got = CO.contract_slices(brs, et)
    assert np.array_equal(got.astype(np.float64), np.asarray(rec["values"]))
    assert float(got.max()) == pytest.approx(rec["exact"], rel=1e-6)


@pytest.mark.parametrize("n,seed", [(30, 3), (60, 5), (100, 7), (140, 11)])
def test_c_oracle_equals_numpy_oracle_and_exact(n, seed):
    root = regular_root(n, seed)
    got = CO.contract_slices([root], np.float32)[0]
    if n <= 100:
        assert got == O.solve_slice(root, np.float32)
    assert float(got) == O.exact_mis_milp(root.nv, root.edges)


def test_c_oracle_weighted_f32_bit_exact():
    rng = np.random.default_rng(4)
    nv, edges = H.random_regular_graph(70, 3, 17)
    w = (1 + rng.random(nv)).astype(np.float32)
    root = H.make_root(nv, edges, weights=w, seed=17)
    assert CO.contract_slices([root], np.float32)[0] == O.solve_slice(root, np.float32)


def test_c_oracle_many_vs_few_branches_paths():
    """many branches: one per thread; fewer branches than threads: the threads share each GEMM -- same values."""
    rec = load_golden("rr100_sc10_unit.json")
    brs = [b for b in golden_branches(rec) if b.nv > 0]
    flats = [CO.flatten(b) for b in brs]
    many, ops_many, th = CO.contract_batch(flats * 3)
    one_by_one = np.array([CO.contract_batch([f])[0][0] for f in flats])
    assert np.array_equal(many[:len(flats)], one_by_one) and np.array_equal(many[:len(flats)], many[len(flats):2 * len(flats)])
    assert th >= 1 and ops_many[0] > 0


@pytest.mark.parametrize("n,seed,k", [(40, 2, 3), (80, 6, 4)])
def test_c_oracle_index_slices(n, seed, k):
    import tbcuda
    from helpers import to_sliced
    root = regular_root(n, seed)
    labels, _, _ = tbcuda.suggest_slices(to_sliced(root), -1, k)
    vals, ops, _ = CO.contract_index_slices(root, labels, range(1 << k))
    want = [O.solve_slice(root, np.float32, fixed={l: (a >> i) & 1 for i, l in enumerate(labels)}) for a in range(1 << k)]
    assert np.array_equal(vals, np.asarray(want, dtype=np.float64))
    assert vals.max() == O.solve_slice(root, np.float32)
    p = tbcuda.Plan(to_sliced(root), fixed={l: 0 for l in labels})
    assert ops[0] == p.info().ops  # the port counts the same algorithmic ops as the plan compiler


def test_c_oracle_simd_kernel_is_compiled_in():
    lib = CO.load()
    import ctypes as C
    lib.tref_simd.restype = C.c_char_p
    assert lib.tref_simd() in (b"avx2", b"avx512")
