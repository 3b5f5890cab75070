"""Open-boundary contractions (iy non-empty; SURVEY 8f #3, first half): the root tensor over the boundary labels of a
region -- max independent-set size inside the region for every boundary configuration -- equals the oracle's, and its
maximum equals the closed contraction."""
import numpy as np
import pytest

import desc_interp as DI
from helpers import align_to, device_tensor_as_ndarray, regular_root
from oracle import tropical_oracle as O


def _open_branch(tb, n, seed, n_open, weights=None):
    root = regular_root(n, seed, weights=weights)
    rng = np.random.default_rng(seed)
    open_labels = sorted(int(v) for v in rng.choice(root.nv, size=n_open, replace=False))
    br = tb.SlicedBranch(tb.MISProblem(root.nv, root.edges, root.weights),
                         tb.CompressedEinsum(root.ixs, open_labels, root.tree), 0)
    left, right = O.nested_to_postorder(root.tree, len(root.ixs))
    return root, br, open_labels, left, right


def _oracle_root(root, open_labels, left, right, dtype):
    w = None if root.weights is None else np.asarray(root.weights).astype(dtype)
    t, labs = O.contract_tree(root.ixs, left, right, w, dtype, open_labels=tuple(open_labels))
    return list(labs), np.asarray(t)


@pytest.mark.parametrize("n,seed,n_open", [(12, 1, 2), (30, 3, 4), (60, 5, 6)])
@pytest.mark.parametrize("flags", [0, 2, 8, 64])
def test_open_root_tensor_interpreted(tb, n, seed, n_open, flags):
    import struct
    root, br, open_labels, left, right = _open_branch(tb, n, seed, n_open)
    p = tb.Plan(br, flags=flags)
    st = p.info()
    assert st.root_rank == n_open
    _, arena = DI.run_plan(p)
    root_off = struct.unpack("4q", p.raw(5))[1]
    # layout of the root = labels_c of the step that produces it
    s = [x for x in p.steps() if x.rank_c == n_open and x.c_offset == root_off][-1]
    labels = [s.labels_c[i] for i in range(s.rank_c)]
    data = DI.to_float(arena[root_off:root_off + (1 << n_open)], st.value_type)
    dl, darr = device_tensor_as_ndarray(labels, data)
    ol, oarr = _oracle_root(root, open_labels, left, right, np.float64)
    assert sorted(dl) == sorted(ol) == open_labels
    assert np.array_equal(align_to(dl, darr, ol), oarr)
    assert oarr.max() == O.solve_slice(root, np.float64)


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,n_open,flags", [(30, 3, 4, 0), (60, 5, 6, 0), (100, 7, 8, 0), (100, 7, 8, 64), (60, 5, 5, 2)])
def test_gpu_contract_tensor(tb, engine, n, seed, n_open, flags):
    root, br, open_labels, left, right = _open_branch(tb, n, seed, n_open)
    p = tb.Plan(br, flags=flags, engine=engine)
    labels, data = engine.contract_tensor(p)
    dl, darr = device_tensor_as_ndarray(labels, data)
    ol, oarr = _oracle_root(root, open_labels, left, right, np.float64)
    assert sorted(dl) == open_labels
    assert np.array_equal(align_to(dl, darr, ol), oarr)
    assert data.max() == O.solve_slice(root, np.float64)
    p.close()


@pytest.mark.gpu
def test_gpu_contract_tensor_weighted_f32(tb, engine):
    rng = np.random.default_rng(2)
    w = (1 + rng.random(40)).astype(np.float32)
    root, br, open_labels, left, right = _open_branch(tb, 40, 9, 5, weights=w)
    p = tb.Plan(br, engine=engine)
    labels, data = engine.contract_tensor(p)
    dl, darr = device_tensor_as_ndarray(labels, data)
    ol, oarr = _oracle_root(root, open_labels, left, right, np.float32)
    assert np.array_equal(align_to(dl, darr, ol).astype(np.float32), oarr)
    p.close()
