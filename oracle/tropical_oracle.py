"""CPU ORACLE (test infrastructure, NOT product code) for the tropical-contraction hot path.

PARITY STATUS: **vector-level parity unpinned**.  The reference has no golden vectors for this path
and cannot run here (no Julia; the arithmetic lives in un-vendored GenericTensorNetworks 4.1 /
OMEinsum 0.9 / TropicalNumbers 0.6 / TropicalGEMM, /root/reference/Project.toml:30-51).  What pins
this oracle instead is the *property* every reference test asserts for the path -- the contracted
value equals the exact maximum (weighted) independent set of the branch graph
(/root/reference/test/slice.jl:32-33,47; test/dynamic_ob.jl:20,38,51; test/utils.jl:34,37,61;
test/decompose.jl:57,71,84) -- checked here against independent exact solvers (``exact_mis_*``).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (tensorbranching.jl_b200, libtbcuda.so) never does.

What is restated (file:line = the reference call site whose behaviour is followed):

* ``solve_slice``       /root/reference/src/dynamic_ob.jl:30-34   contract the branch network, return scalar
* ``contract_slices``   /root/reference/src/dynamic_ob.jl:36-48   per-branch value + r, empty graph => r
* ``leaf_tensor``       generate_tensors(Tropical{T}(1), IndependentSet(g, w)) [upstream GenericTensorNetworks,
                        recalled]: vertex v -> [0, w_v]; edge (u,v) -> [[0, 0], [0, -inf]]
* ``contract_pair``     OMEinsum binary rule [upstream, recalled]: labels of (A, B, out) split into
                        m (A only, kept) / n (B only, kept) / b (shared, kept) / k (dropped);
                        C[m,n,b] = max_k A[m,k,b] + B[k,n,b]   (tropical GEMM: (+) = max, (x) = +)
* output labels of a node = labels of its operands that still occur outside its subtree or in iy
  (parse_eincode on the ContractionTree, /root/reference/src/types.jl:75-79)

Tensors are numpy arrays of shape (2,)*rank with axis i <-> labels[i]; all label sizes are 2
(uniformsize(code, 2), /root/reference/src/types.jl:118).
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

NEG_INF = -np.inf


# --------------------------------------------------------------------------------------------
# leaves
# --------------------------------------------------------------------------------------------
def leaf_tensor(labels: Sequence[int], weights, dtype) -> np.ndarray:
    """Tropical leaf of the IndependentSet network.  1 label: [one, one^w] = [0, w];
    2 labels: [[one, one], [one, zero]] = [[0, 0], [0, -inf]]."""
    if len(labels) == 1:
        w = 1 if weights is None else weights[labels[0]]
        return np.array([0, w], dtype=dtype)
    if len(labels) == 2:
        if labels[0] == labels[1]:
            raise ValueError("self-loop edge tensor")
        return np.array([[0, 0], [0, NEG_INF]], dtype=dtype)
    raise ValueError(f"leaf with {len(labels)} labels is not an IndependentSet tensor")


# --------------------------------------------------------------------------------------------
# tree bookkeeping
# --------------------------------------------------------------------------------------------
def nested_to_postorder(tree, n_leaves: int) -> Tuple[List[int], List[int]]:
    """Nested (left, right) tuples with int leaves -> child arrays; node j has id n_leaves + j."""
    left: List[int] = []
    right: List[int] = []
    stack = [(tree, False)]
    ids: List[int] = []
    while stack:
        t, done = stack.pop()
        if not isinstance(t, tuple):
            ids.append(int(t))
        elif not done:
            stack.append((t, True))
            stack.append((t[1], False))
            stack.append((t[0], False))
        else:
            r_id = ids.pop()
            l_id = ids.pop()
            left.append(l_id)
            right.append(r_id)
            ids.append(n_leaves + len(left) - 1)
    return left, right


def node_output_labels(ixs, left, right, open_labels=()) -> List[Tuple[int, ...]]:
    """For every tensor id (leaves then internal nodes) the labels it carries.  A label survives a
    contraction iff it still occurs in a leaf outside the node's subtree, or is open."""
    n_leaves = len(ixs)
    total: Dict[int, int] = {}
    for ix in ixs:
        for l in set(ix):
            total[l] = total.get(l, 0) + 1
    for l in open_labels:
        total[l] = total.get(l, 0) + 1
    labs: List[Tuple[int, ...]] = [tuple(ix) for ix in ixs]
    counts: List[Optional[Dict[int, int]]] = [{l: 1 for l in set(ix)} for ix in ixs]
    for j in range(len(left)):
        a, b = left[j], right[j]
        ca, cb = counts[a], counts[b]
        merged = dict(ca)
        for l, c in cb.items():
            merged[l] = merged.get(l, 0) + c
        union = list(dict.fromkeys(list(labs[a]) + list(labs[b])))
        out = tuple(sorted(l for l in union if merged[l] < total[l]))
        counts.append({l: merged[l] for l in out})
        counts[a] = counts[b] = None
        labs.append(out)
    return labs


# --------------------------------------------------------------------------------------------
# the binary rule
# --------------------------------------------------------------------------------------------
def contract_pair(A: np.ndarray, la: Sequence[int], B: np.ndarray, lb: Sequence[int],
                  lo: Sequence[int]) -> np.ndarray:
    """C[lo] = max over dropped labels of A[la] + B[lb].  Permute to (m,k,b)/(k,n,b), tropical GEMM
    per batch element, permute back to `lo` -- the OMEinsum binary rule [upstream, recalled]."""
    la, lb, lo = list(la), list(lb), list(lo)
    sa, sb, so = set(la), set(lb), set(lo)
    # labels private to one operand and dropped are reduced first (unary max)
    for l in [l for l in la if l not in sb and l not in so]:
        A = A.max(axis=la.index(l))
        la.remove(l)
    for l in [l for l in lb if l not in sa and l not in so]:
        B = B.max(axis=lb.index(l))
        lb.remove(l)
    sa, sb = set(la), set(lb)
    m = [l for l in la if l not in sb]
    n = [l for l in lb if l not in sa]
    bt = [l for l in la if l in sb and l in so]
    k = [l for l in la if l in sb and l not in so]
    Am = np.transpose(A, [la.index(l) for l in m + k + bt]).reshape(2 ** len(m), 2 ** len(k), 2 ** len(bt))
    Bm = np.transpose(B, [lb.index(l) for l in k + n + bt]).reshape(2 ** len(k), 2 ** len(n), 2 ** len(bt))
    C = np.full((2 ** len(m), 2 ** len(n), 2 ** len(bt)), NEG_INF, dtype=A.dtype)
    for kk in range(Am.shape[1]):
        # one rounding per a+b, max is exact: order of the k loop does not change the result
        np.maximum(C, Am[:, kk, None, :] + Bm[None, kk, :, :], out=C)
    cur = m + n + bt
    C = C.reshape((2,) * len(cur))
    return np.transpose(C, [cur.index(l) for l in lo]) if cur else C


def contract_tree(ixs, left, right, weights=None, dtype=np.float64, open_labels=(),
                  keep_intermediates: bool = False, fixed: Optional[Dict[int, int]] = None):
    """Evaluate the whole tree.  Returns (root_tensor, root_labels[, {tensor id: (labels, array)}]).
    fixed = {label: value}: index slicing -- every leaf carrying a fixed label is restricted to that index
    and the label leaves the network (the max over all assignments of the fixed labels is the unsliced value)."""
    n_leaves = len(ixs)
    vals: List[Optional[np.ndarray]] = [leaf_tensor(ix, weights, dtype) for ix in ixs]
    if fixed:
        vals = [np.asarray(v[tuple(fixed[l] if l in fixed else slice(None) for l in ix)]) for ix, v in zip(ixs, vals)]
        ixs = [tuple(l for l in ix if l not in fixed) for ix in ixs]
    labs = node_output_labels(ixs, left, right, open_labels)
    if open_labels:
        labs[-1] = tuple(open_labels) if len(left) else labs[-1]
    inter = {}
    with np.errstate(invalid="ignore"):
        for j in range(len(left)):
            a, b = left[j], right[j]
            C = contract_pair(vals[a], labs[a], vals[b], labs[b], labs[n_leaves + j])
            vals.append(C)
            if keep_intermediates:
                inter[n_leaves + j] = (labs[n_leaves + j], C)
            else:
                vals[a] = vals[b] = None
        if len(left) == 0:
            # single-leaf network: drop every non-open label
            root = vals[0]
            keep = [l for l in labs[0] if l in set(open_labels)]
            for l in [l for l in labs[0] if l not in set(open_labels)]:
                root = root.max(axis=list(labs[0]).index(l)) if root.ndim else root
            root = np.asarray(root)
            res = (root, tuple(keep))
        else:
            res = (vals[-1], labs[-1])
    if keep_intermediates:
        return res[0], res[1], inter
    return res


# --------------------------------------------------------------------------------------------
# the boundary functions
# --------------------------------------------------------------------------------------------
def solve_slice(branch, element_type=np.float32, fixed: Optional[Dict[int, int]] = None):
    """/root/reference/src/dynamic_ob.jl:30-34 (fixed: one index slice of it, see contract_tree)."""
    left, right = nested_to_postorder(branch.tree, len(branch.ixs))
    w = None if branch.weights is None else np.asarray(branch.weights).astype(element_type)
    root, _ = contract_tree(branch.ixs, left, right, w, element_type, fixed=fixed)
    return element_type(np.asarray(root).reshape(-1)[0])


def contract_slices(branches, element_type=np.float32) -> np.ndarray:
    """/root/reference/src/dynamic_ob.jl:36-48: empty graph => r, else solve_slice + r."""
    res = []
    for br in branches:
        if br.nv == 0:
            res.append(element_type(br.r))
        else:
            res.append(element_type(solve_slice(br, element_type) + element_type(br.r)))
    return np.asarray(res, dtype=element_type)


# --------------------------------------------------------------------------------------------
# independent exact solvers for the invariant (what the reference's tests compare against)
# --------------------------------------------------------------------------------------------
def exact_mis_bruteforce(nv: int, edges, weights=None) -> float:
    assert nv <= 22
    best = 0.0
    w = np.ones(nv) if weights is None else np.asarray(weights, dtype=np.float64)
    emask = [(1 << u) | (1 << v) for u, v in edges]
    for s in range(1 << nv):
        if any((s & e) == e for e in emask):
            continue
        tot = sum(w[i] for i in range(nv) if (s >> i) & 1)
        best = max(best, tot)
    return best


def exact_mis_milp(nv: int, edges, weights=None) -> float:
    """max sum w_v x_v  s.t. x_u + x_v <= 1, x binary (HiGHS via scipy) -- plays the role of mis2 /
    the unsliced contraction in the reference's tests."""
    from scipy.optimize import Bounds, LinearConstraint, milp
    from scipy.sparse import lil_matrix

    if nv == 0:
        return 0.0
    w = np.ones(nv) if weights is None else np.asarray(weights, dtype=np.float64)
    if len(edges) == 0:
        return float(np.clip(w, 0, None).sum())
    A = lil_matrix((len(edges), nv))
    for i, (u, v) in enumerate(edges):
        A[i, u] = 1
        A[i, v] = 1
    res = milp(c=-w, constraints=LinearConstraint(A.tocsr(), -np.inf, 1), integrality=np.ones(nv),
               bounds=Bounds(0, 1))
    assert res.success, res.message
    x = np.round(res.x)
    return float(w @ x)


def exact_mis_clique(nv: int, edges, weights=None) -> float:
    """Max-weight clique of the complement graph (networkx), integer weights only, nv <= ~60."""
    import networkx as nx

    g = nx.Graph()
    g.add_nodes_from(range(nv))
    g.add_edges_from(edges)
    gc = nx.complement(g)
    for v in gc.nodes:
        gc.nodes[v]["weight"] = 1 if weights is None else int(weights[v])
    _, wt = nx.max_weight_clique(gc, weight="weight")
    return float(wt)


# --------------------------------------------------------------------------------------------
# the reference's memory estimators (restated; used to check tb_plan_info)
# --------------------------------------------------------------------------------------------
def contraction_all_memory(ixs, tree) -> float:
    """/root/reference/src/utils.jl:222-229: log2 of the summed sizes of every intermediate (all label sizes 2)."""
    left, right = nested_to_postorder(tree, len(ixs))
    labs = node_output_labels(ixs, left, right)
    return float(np.log2(sum(2.0 ** len(labs[len(ixs) + j]) for j in range(len(left)))))


def contraction_peak_memory(ixs, tree) -> float:
    """/root/reference/src/utils.jl:197-219: depth-first walk (first operand first); the running total starts at the
    summed leaf sizes, every node adds its result and releases the operands of its child nodes; log2 of the maximum."""
    n_leaves = len(ixs)
    left, right = nested_to_postorder(tree, n_leaves)
    labs = node_output_labels(ixs, left, right)
    tscs = [sum(2.0 ** len(set(ix)) for ix in ixs)]

    def walk(t):  # t = tensor id; returns the summed size of t's operands (released by t's parent)
        if t < n_leaves:
            return 0.0
        j = t - n_leaves
        freed = 0.0
        for c in (left[j], right[j]):
            freed += walk(c)
        future = sum(2.0 ** len(labs[c]) for c in (left[j], right[j]))
        tscs.append(tscs[-1] + 2.0 ** len(labs[t]) - freed)
        return future

    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 4 * n_leaves + 100))
    walk(n_leaves + len(left) - 1)
    return float(np.log2(max(tscs)))


def mis_compactify_keep(sizes):
    """[upstream, recalled] GenericTensorNetworks `mis_compactify!`, as OptimalBranchingMIS `reduced_alpha_configs` applies
    it to the open-boundary size tensor before `branching_table` builds its rows (reached from
    /root/reference/src/branch.jl:79): entry a is set to tropical zero iff some entry b != a chooses a subset of a's
    boundary vertices ((b & a) == b) and sizes[b] >= sizes[a].  Plain all-pairs restatement over the flat table (index bit
    i = boundary vertex i); returns the boolean mask of the entries that survive (infeasible entries do not)."""
    sizes = np.asarray(sizes, dtype=np.float64).ravel()
    keep = np.zeros(sizes.size, dtype=bool)
    for a in range(sizes.size):
        if sizes[a] == -np.inf:
            continue
        keep[a] = not any(b != a and (b & a) == b and sizes[b] >= sizes[a] for b in range(a))
    return keep


# --------------------------------------------------------------------------------------------
# all optimal configurations (the rows of the reference's branching tables)
# --------------------------------------------------------------------------------------------
def contract_tree_configs(ixs, left, right, weights=None, open_labels=()):
    """[upstream, recalled] the contraction `solve(problem, ConfigsMax(; bounded=false))` that OptimalBranchingMIS'
    `reduced_alpha_configs` runs for `branching_table(p, TensorNetworkSolver(), region)` (/root/reference/src/branch.jl:79):
    the same tree, the same binary rule, but every element is a pair (size, SET of configurations) -- GenericTensorNetworks'
    CountingTropical{T, ConfigEnumerator}: (x) adds the sizes and joins every pair of configurations (bitwise or),
    (+) keeps the larger size and unites the sets on a tie; tropical zero = (-inf, {}).  Leaves: vertex v -> [(0, {0}),
    (w_v, {1 << v})], edge -> [[one, one], [one, zero]] with one = (0, {0}).
    Pure Python over dictionaries (small regions only).  Returns (labels, {boundary assignment tuple: (size, frozenset of
    vertex masks)}) for the open labels in the order given."""
    zero = (NEG_INF, frozenset())
    one = (0.0, frozenset([0]))

    def add(x, y):
        if x[0] > y[0]:
            return x
        if y[0] > x[0]:
            return y
        return (x[0], x[1] | y[1])

    def mul(x, y):
        if x[0] == NEG_INF or y[0] == NEG_INF:
            return zero
        return (x[0] + y[0], frozenset(a | b for a in x[1] for b in y[1]))

    def leaf(ix):
        if len(ix) == 1:
            w = 1.0 if weights is None else float(weights[ix[0]])
            return {(0,): one, (1,): (w, frozenset([1 << ix[0]]))}
        return {(0, 0): one, (0, 1): one, (1, 0): one, (1, 1): zero}

    n_leaves = len(ixs)
    labs = node_output_labels(ixs, left, right, open_labels)
    vals = [leaf(tuple(ix)) for ix in ixs]
    for j in range(len(left)):
        a, b = left[j], right[j]
        la, lb, lo = labs[a], labs[b], labs[n_leaves + j]
        allv = list(dict.fromkeys(list(la) + list(lb)))
        out = {}
        for bits in itertools.product((0, 1), repeat=len(allv)):
            asg = dict(zip(allv, bits))
            v = mul(vals[a][tuple(asg[l] for l in la)], vals[b][tuple(asg[l] for l in lb)])
            key = tuple(asg[l] for l in lo)
            out[key] = add(out.get(key, zero), v)
        vals.append(out)
        vals[a] = vals[b] = None
    root, rl = vals[-1], labs[-1]
    res = {}
    for key, v in root.items():  # single-leaf networks may carry non-open labels: reduce them; then order as asked
        asg = dict(zip(rl, key))
        k2 = tuple(asg[l] for l in open_labels)
        res[k2] = add(res.get(k2, zero), v)
    return tuple(open_labels), res


def table_configs_bruteforce(nv: int, edges, weights, boundary):
    """Independent check of the above: for every assignment of the boundary vertices, the best weight of an independent
    set that agrees with it and ALL vertex masks that attain it, by enumerating the 2^nv vertex sets.
    -> (sizes float64[2^rank] (index bit i = boundary[i]), [sorted list of masks per entry])"""
    adj = [0] * nv
    for u, v in edges:
        adj[u] |= 1 << v
        adj[v] |= 1 << u
    w = [1.0] * nv if weights is None else [float(x) for x in weights]
    rank = len(boundary)
    sizes = np.full(1 << rank, NEG_INF)
    rows: List[List[int]] = [[] for _ in range(1 << rank)]
    for s in range(1 << nv):
        if any((s >> v) & 1 and adj[v] & s for v in range(nv)):
            continue
        tot = 0.0
        for v in range(nv):  # ascending vertex order (the device sums in the same order: exact equality for real weights)
            if (s >> v) & 1:
                tot += w[v]
        a = sum(((s >> boundary[i]) & 1) << i for i in range(rank))
        if tot > sizes[a]:
            sizes[a] = tot
            rows[a] = [s]
        elif tot == sizes[a]:
            rows[a].append(s)
    return sizes, rows
