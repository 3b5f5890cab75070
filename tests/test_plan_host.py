"""CPU tests of the host side: C-ABI surface, plan compiler (via the numpy descriptor interpreter),
error behaviour.  No compute call touches a GPU here."""
import ctypes as C
import re
import os

import numpy as np
import pytest

import desc_interp as DI
from helpers import align_to, device_tensor_as_ndarray, golden_branches, load_golden, regular_root, to_sliced
from oracle import tropical_oracle as O
from workloads import standin_host as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(tb):
    lib = tb.load()
    hdr = open(os.path.join(ROOT, "include", "tbcuda.h")).read()
    declared = set(re.findall(r"\b(tb_[a-z_]+)\s*\(", hdr))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/tbcuda.h but not exported"
    assert b"sm_100a" in lib.tb_version()


def test_init_without_gpu_fails_loudly(tb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(tb.TBError) as e:
        tb.Engine(0)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_usecuda_false_is_refused(tb):
    root = regular_root(12, 1)
    with pytest.raises(tb.TBError):
        tb.contract_slices([to_sliced(root)], np.float32, False)


FLAGS = [0, 2, 4, 8, 16, 2 | 8, 4 | 8, 64, 64 | 8, 64 | 2, 64 | 4, 64 | 16]


@pytest.mark.parametrize("n,seed", [(12, 1), (30, 3), (60, 5), (100, 7)])
@pytest.mark.parametrize("flags", FLAGS)
def test_plan_interpreted_equals_oracle(tb, n, seed, flags):
    root = regular_root(n, seed)
    p = tb.Plan(to_sliced(root), flags=flags)
    val, _ = DI.run_plan(p)
    assert val == O.solve_slice(root, np.float64)
    st = p.info()
    sc, tc = H.tree_complexity(root.ixs, root.tree)
    assert st.sc == sc and st.tc == pytest.approx(tc)
    if flags & 2:
        assert st.n_fused_steps == 0
    if flags & 4:
        assert st.n_gemm_steps == 0


def test_plan_weighted_f32_bit_exact(tb):
    rng = np.random.default_rng(3)
    nv, edges = H.random_regular_graph(60, 3, 21)
    w = (1 + rng.random(nv)).astype(np.float32)
    root = H.make_root(nv, edges, weights=w, seed=3)
    for flags in (0, 2, 8):
        p = tb.Plan(to_sliced(root), flags=flags)
        assert p.info().value_type == 2
        val, _ = DI.run_plan(p)
        assert np.float32(val) == O.solve_slice(root, np.float32)  # bit-exact: same tree, one rounding per add


def test_plan_every_node_matches_oracle(tb):
    """KEEP_INTERMEDIATES: every node's tensor, in the plan's layout, equals the oracle's."""
    root = regular_root(40, 8)
    left, right = O.nested_to_postorder(root.tree, len(root.ixs))
    _, _, inter = O.contract_tree(root.ixs, left, right, None, np.float64, keep_intermediates=True)
    for flags in (1, 1 | 8, 1 | 64):
        p = tb.Plan(to_sliced(root), flags=flags)
        vt = p.info().value_type
        assert vt == (1 if flags & 64 else 3)
        _, arena = DI.run_plan(p)
        for s in p.steps():
            if s.node not in inter:
                continue
            labels = [s.labels_c[i] for i in range(s.rank_c)]
            data = DI.to_float(arena[s.c_offset:s.c_offset + (1 << s.rank_c)], vt)
            dl, darr = device_tensor_as_ndarray(labels, data)
            ol, oarr = inter[s.node]
            assert sorted(dl) == sorted(ol)
            assert np.array_equal(align_to(dl, darr, ol), oarr), f"node {s.node}"


@pytest.mark.parametrize("name", ["rr100_sc10_unit", "rr100_sc10_f32", "rr30_disconnected", "ksg8x8_sc6", "ksg7x7_sc8_nokernel"])
def test_plan_on_golden(tb, name):
    rec = load_golden(name + ".json")
    et = np.dtype(rec["element_type"]).type
    for b, want in list(zip(golden_branches(rec), rec["values"]))[:12]:
        if b.nv == 0:
            continue
        p = tb.Plan(to_sliced(b))
        val, _ = DI.run_plan(p)
        assert et(et(val) + et(b.r)) == et(want)


def test_corner_cases(tb):
    # single isolated vertex: one leaf, no tree
    b = tb.SlicedBranch(tb.MISProblem(1, [], None), tb.CompressedEinsum([(0,)], (), None), 0)
    assert DI.run_plan(tb.Plan(b))[0] == 1.0
    # two isolated vertices -> outer product of two scalars
    b = tb.SlicedBranch(tb.MISProblem(2, [], None), tb.CompressedEinsum([(0,), (1,)], (), (0, 1)), 0)
    assert DI.run_plan(tb.Plan(b))[0] == 2.0
    # one edge (2-vertex component) + isolated vertex with weights
    w = np.array([2.5, 1.0, 4.0], dtype=np.float32)
    b = tb.SlicedBranch(tb.MISProblem(3, [(0, 1)], w), tb.CompressedEinsum([(0,), (1,), (2,), (0, 1)], (), ((0, (1, 3)), 2)), 0)
    assert DI.run_plan(tb.Plan(b))[0] == 6.5


def test_error_codes(tb):
    lib = tb.load()
    # non-binary / malformed trees
    with pytest.raises(ValueError):
        tb.CompressedEinsum([(0,), (1,), (0, 1)], (), (0, 1, 2))
    ce = tb.CompressedEinsum([(0,), (1,), (0, 1)], (), None, flat=(np.array([0, 0]), np.array([1, 3])))
    with pytest.raises(tb.TBError) as e:
        tb.Plan(tb.SlicedBranch(tb.MISProblem(2, [(0, 1)], None), ce, 0))
    assert e.value.code == -2
    # leaf of rank 3 is not an IndependentSet tensor
    ce = tb.CompressedEinsum([(0, 1, 2), (1,)], (), (0, 1))
    with pytest.raises(tb.TBError) as e:
        tb.Plan(tb.SlicedBranch(tb.MISProblem(3, [], None), ce, 0))
    assert e.value.code == -3
    # label out of range
    ce = tb.CompressedEinsum([(0,), (5,)], (), (0, 1))
    with pytest.raises(tb.TBError) as e:
        tb.Plan(tb.SlicedBranch(tb.MISProblem(2, [], None), ce, 0))
    assert e.value.code == -1
    assert lib.tb_last_error(None)


def test_arena_reuse_is_smaller_than_keep(tb):
    root = regular_root(100, 7)
    a = tb.Plan(to_sliced(root)).info().arena_elems
    b = tb.Plan(to_sliced(root), flags=1).info().arena_elems
    assert a < b


@pytest.mark.parametrize("n,seed", [(12, 1), (30, 3), (60, 5), (100, 7), (160, 3)])
def test_memory_estimators_match_reference_formulas(tb, n, seed):
    """tb_plan_info.peak_memory_log2 / all_memory_log2 == the reference's contraction_peak_memory /
    contraction_all_memory (src/utils.jl:197-229) restated in the oracle; independent of plan flags."""
    root = regular_root(n, seed)
    want_peak = O.contraction_peak_memory(root.ixs, root.tree)
    want_all = O.contraction_all_memory(root.ixs, root.tree)
    for flags in (0, 2, 16, 64):
        st = tb.Plan(to_sliced(root), flags=flags).info()
        assert st.peak_memory_log2 == pytest.approx(want_peak, abs=1e-12)
        assert st.all_memory_log2 == pytest.approx(want_all, abs=1e-12)
    br = to_sliced(root)
    assert tb.contraction_peak_memory(br) == pytest.approx(want_peak) and tb.contraction_all_memory(br) == pytest.approx(want_all)
    assert want_peak <= want_all + 1 and st.sc <= want_peak


def _arena_blocks(plan):
    """{node: (offset, size, left, right)} of the tensors that live in the HBM arena (non-fused steps + fused-subtree roots)"""
    steps = plan.steps()
    parent = {}
    kind = {}
    for s in steps:
        parent[s.left] = s.node
        parent[s.right] = s.node
        kind[s.node] = s.kind
    out = {}
    for s in steps:
        in_arena = s.kind != 0 or s.node not in parent or kind[parent[s.node]] != 0
        if in_arena:
            out[s.node] = (s.c_offset, (((1 << s.rank_c) + 63) // 64) * 64, s.left, s.right)
    return out, parent


@pytest.mark.parametrize("n,seed,flags", [(100, 7, 0), (130, 5, 0), (150, 1000, 0), (150, 1000, 16), (150, 1000, 8), (120, 9, 64),
                                          (160, 3, 0), (140, 11, 2)])
def test_arena_reuse_is_safe_for_the_dataflow_executor(n, seed, flags):
    """The dataflow executor orders steps ONLY operand -> consumer.  Two arena tensors may therefore share bytes only
    if one lies in the interior of the other's subtree (strictly below its operands): everything there is dead -- and
    complete -- once the operands are complete.  Siblings / cousins run concurrently and must be disjoint."""
    import tbcuda
    from helpers import regular_root, to_sliced
    p = tbcuda.Plan(to_sliced(regular_root(n, seed)), flags=flags)
    blocks, parent = _arena_blocks(p)

    def ancestors(t):
        out = []
        while t in parent:
            t = parent[t]
            out.append(t)
        return out

    nodes = sorted(blocks)
    n_shared = 0
    for i, x in enumerate(nodes):
        ox, sx, _, _ = blocks[x]
        anc = ancestors(x)
        for y in nodes[i + 1:]:
            oy, sy, ly, ry = blocks[y]
            if ox + sx <= oy or oy + sy <= ox:
                continue
            n_shared += 1
            # y was allocated later (children precede parents): x must be strictly below y's operands
            assert y in anc and x not in (ly, ry), (x, y)
    assert p.info().arena_elems >= max(o + s for o, s, _, _ in blocks.values())
    if n >= 130 and not flags:
        assert n_shared > 0  # reuse does happen
    p.close()


def test_estimate_equals_plan_ops():
    """tb_estimate (label-set pass only) returns exactly the ops / sc of the compiled plan: what the sharders balance"""
    import tbcuda
    from helpers import golden_branches, load_golden, regular_root, to_sliced
    for n, seed in [(12, 1), (60, 5), (100, 7), (150, 1000)]:
        s = to_sliced(regular_root(n, seed))
        p = tbcuda.Plan(s)
        st = p.info()
        assert tbcuda.estimate(s) == (st.ops, st.sc)
        p.close()
    brs = golden_branches(load_golden("rr100_sc10_unit.json"))
    assert tbcuda.estimate(to_sliced([b for b in brs if b.nv == 0][0] if any(b.nv == 0 for b in brs) else brs[0]))[0] >= 0
    from tbcuda.multi_gpu import branch_cost
    assert branch_cost(to_sliced(regular_root(60, 5))) == tbcuda.Plan(to_sliced(regular_root(60, 5))).info().ops


def test_network_cache_follows_code_and_weights_objects():
    """the cached tb_network holds raw pointers into branch.code / weights; replacing either must rebuild it"""
    import numpy as np
    import tbcuda
    from tbcuda import contract as Cn
    from helpers import regular_root, to_sliced
    a, b = to_sliced(regular_root(30, 3)), to_sliced(regular_root(40, 4))
    n1 = Cn._network_bytes(a, np.float32)
    assert Cn._network_bytes(a, np.float32) is n1  # cached
    a.code = b.code  # the host swaps the tree of a branch
    a.p = b.p
    n2 = Cn._network_bytes(a, np.float32)
    assert n2 != n1 and n2 == Cn._network_bytes(b, np.float32)
    p = tbcuda.Plan(a)
    assert p.info().ops == tbcuda.Plan(b).info().ops


def test_per_step_ops_and_bytes_export(tb):
    """tb_plan_export_raw sections 6 / 7 (what bench.py's per-node hybrid roofline reads): one entry per non-fused step,
    consistent with the plan's totals"""
    for n, seed in ((60, 5), (100, 7), (130, 5)):
        p = tb.Plan(to_sliced(regular_root(n, seed)))
        st = p.info()
        log2_ops = np.frombuffer(p.raw(6), dtype=np.float32)
        nbytes = np.frombuffer(p.raw(7), dtype=np.float64)
        assert log2_ops.size == nbytes.size == st.n_gemm_steps + st.n_generic_steps
        if log2_ops.size:
            assert np.all(log2_ops == np.round(log2_ops)) and np.all(nbytes > 0)
            # second halves of split reductions are engine overhead: their ops are not in the plan's algorithmic totals
            assert np.exp2(log2_ops.astype(np.float64)).sum() >= st.gemm_ops + st.generic_ops
            assert nbytes.sum() >= st.gemm_bytes
        p.close()
