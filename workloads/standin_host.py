"""Stand-in for the Julia host that sits ABOVE the hot-path boundary.

The reference's host side (kernelize.jl, branch.jl, slice.jl, refine.jl, the TreeSA optimiser
from OMEinsumContractionOrders and the OptimalBranching rules) cannot run in this image (no
Julia).  The engine only needs what that host hands across the boundary -- a list of
``SlicedBranch`` records (graph, weights, leaf label lists, binary contraction tree, offset r;
/root/reference/src/types.jl:51-103).  This module produces records of the same *shape* from
seeded synthetic graphs so that tests and bench.py have something to contract:

* ``random_regular_graph`` / ``random_ksg``   instance generators (random_ksg restates
  /root/reference/src/utils.jl:191-195: pick round(m*n*rho) lattice sites, connect sites at
  Chebyshev distance 1).
* ``greedy_tree``     pairwise greedy contraction order (stands in for TreeSA,
  /root/reference/src/dynamic_ob.jl:50-54).
* ``kernelize``       removal-only reduction rules iterated to a fix point (stands in for
  /root/reference/src/kernelize.jl:4-17; unit weights: degree-0/1, triangle degree-2, domination).
* ``slice_bfs``       vertex branching until every branch has sc <= sc_target (stands in for
  /root/reference/src/slice.jl:68-90 + src/branch.jl:212-232; the tree of a child branch is the
  parent's tree with the removed tensors deleted and re-binarised, the same operation as
  remove_tensors!/reform_tree! in /root/reference/src/utils.jl:46-120).

It is a workload GENERATOR, not a port: branching decisions differ from OptimalBranching's, but any
valid branching satisfies  max_i (MIS(branch_i) + r_i) == MIS(g), which is the property the
reference's own tests pin (test/slice.jl:32-33, test/dynamic_ob.jl:20).
"""
from __future__ import annotations

import heapq
import random
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

Tree = object  # int leaf id (0-based index into ixs) or (left, right) tuple


# --------------------------------------------------------------------------------------------
# graphs
# --------------------------------------------------------------------------------------------
def random_regular_graph(n: int, d: int, seed: int) -> Tuple[int, List[Tuple[int, int]]]:
    import networkx as nx

    g = nx.random_regular_graph(d, n, seed=seed)
    edges = sorted((min(u, v), max(u, v)) for u, v in g.edges())
    return n, edges


def random_ksg(m: int, n: int, rho: float, seed: int) -> Tuple[int, List[Tuple[int, int]]]:
    """King's subgraph: round(m*n*rho) random sites of an m x n lattice, 8-neighbour adjacency."""
    rng = np.random.default_rng(seed)
    nsites = int(round(m * n * rho))
    sites = np.sort(rng.choice(m * n, size=nsites, replace=False))
    pos = {int(s): i for i, s in enumerate(sites)}
    edges = []
    for s, i in pos.items():
        x, y = divmod(s, n)
        for dx, dy in ((0, 1), (1, -1), (1, 0), (1, 1)):
            xx, yy = x + dx, y + dy
            if 0 <= xx < m and 0 <= yy < n:
                j = pos.get(xx * n + yy)
                if j is not None:
                    edges.append((min(i, j), max(i, j)))
    return nsites, sorted(edges)


def mis_ixs(nv: int, edges: Sequence[Tuple[int, int]]) -> List[Tuple[int, ...]]:
    """Leaf label lists of the IndependentSet network: one 1-label tensor per vertex, then one
    2-label tensor per edge [upstream GenericTensorNetworks layout, recalled]."""
    return [(v,) for v in range(nv)] + [(u, v) for u, v in edges]


# --------------------------------------------------------------------------------------------
# data model handed across the boundary
# --------------------------------------------------------------------------------------------
@dataclass
class Branch:
    """One SlicedBranch (src/types.jl:85-103) in plain-Python form, labels 0-based."""
    nv: int
    edges: List[Tuple[int, int]]
    weights: Optional[np.ndarray]          # None == UnitWeight
    ixs: List[Tuple[int, ...]]             # CompressedEinsum.ixs
    tree: Tree                             # CompressedEinsum.ct (None when nv == 0)
    r: float = 0
    meta: dict = field(default_factory=dict)


# --------------------------------------------------------------------------------------------
# contraction trees
# --------------------------------------------------------------------------------------------
def greedy_tree(ixs: Sequence[Tuple[int, ...]], seed: int = 0, alpha: float = 0.0,
                temperature: float = 0.0) -> Tree:
    """Pairwise greedy order: repeatedly contract the pair minimising
    2^|out| - alpha*(2^|a| + 2^|b|); leftovers (disconnected parts) are joined by outer products."""
    rng = random.Random(seed)
    labels: Dict[int, frozenset] = {i: frozenset(ix) for i, ix in enumerate(ixs)}
    trees: Dict[int, Tree] = {i: i for i in range(len(ixs))}
    where: Dict[int, set] = {}
    for i, ls in labels.items():
        for l in ls:
            where.setdefault(l, set()).add(i)
    next_id = len(ixs)

    def out_labels(a: int, b: int) -> frozenset:
        la, lb = labels[a], labels[b]
        keep = []
        for l in la | lb:
            cnt = len(where[l]) - (l in la) - (l in lb)
            if cnt > 0:
                keep.append(l)
        return frozenset(keep)

    def cost(a: int, b: int) -> float:
        o = out_labels(a, b)
        c = 2.0 ** len(o) - alpha * (2.0 ** len(labels[a]) + 2.0 ** len(labels[b]))
        if temperature > 0:
            c += temperature * rng.random()
        return c

    heap: list = []

    def push_pairs(a: int) -> None:
        nbrs = set()
        for l in labels[a]:
            nbrs |= where[l]
        nbrs.discard(a)
        for b in nbrs:
            heapq.heappush(heap, (cost(a, b), rng.random(), a, b))

    for i in list(labels):
        nbrs = set()
        for l in labels[i]:
            nbrs |= where[l]
        for b in nbrs:
            if b > i:
                heapq.heappush(heap, (cost(i, b), rng.random(), i, b))

    while heap:
        _, _, a, b = heapq.heappop(heap)
        if a not in labels or b not in labels:
            continue
        o = out_labels(a, b)
        for l in labels[a]:
            where[l].discard(a)
        for l in labels[b]:
            where[l].discard(b)
        c = next_id
        next_id += 1
        ta, tb = trees.pop(a), trees.pop(b)
        del labels[a], labels[b]
        labels[c] = o
        trees[c] = (ta, tb)
        for l in o:
            where[l].add(c)
        push_pairs(c)

    rest = [trees[k] for k in sorted(trees)]
    t = rest[0]
    for x in rest[1:]:
        t = (t, x)
    return t


def tree_leaves(tree: Tree) -> List[int]:
    out, stack = [], [tree]
    while stack:
        t = stack.pop()
        if isinstance(t, tuple):
            stack.append(t[1])
            stack.append(t[0])
        else:
            out.append(t)
    return out


def tree_to_postorder(tree: Tree, n_leaves: int) -> Tuple[np.ndarray, np.ndarray]:
    """Flatten to the tb_network arrays: internal node j (id n_leaves+j) = (left[j], right[j]),
    children always have smaller ids; the last node is the root."""
    left, right = [], []

    # iterative post-order
    stack = [(tree, False)]
    ids: list = []
    while stack:
        t, done = stack.pop()
        if not isinstance(t, tuple):
            ids.append(int(t))
            continue
        if not done:
            stack.append((t, True))
            stack.append((t[1], False))
            stack.append((t[0], False))
        else:
            r_id = ids.pop()
            l_id = ids.pop()
            left.append(l_id)
            right.append(r_id)
            ids.append(n_leaves + len(left) - 1)
    return np.asarray(left, dtype=np.int32), np.asarray(right, dtype=np.int32)


def tree_complexity(ixs: Sequence[Tuple[int, ...]], tree: Tree, return_nodes: bool = False):
    """sc / tc in the reference's sense (src/types.jl:115-121, all label sizes 2):
    sc = max rank of any tensor, tc = log2(sum over nodes of 2^(#labels involved))."""
    if tree is None:
        return (0.0, 0.0, []) if return_nodes else (0.0, 0.0)
    left, right = tree_to_postorder(tree, len(ixs))
    n_leaves = len(ixs)
    # count of leaves containing each label, to decide what survives a contraction
    total: Dict[int, int] = {}
    for ix in ixs:
        for l in set(ix):
            total[l] = total.get(l, 0) + 1
    lab: List[Optional[frozenset]] = [frozenset(ix) for ix in ixs] + [None] * len(left)
    cnt: List[Optional[Dict[int, int]]] = [None] * (n_leaves + len(left))
    for i, ix in enumerate(ixs):
        cnt[i] = {l: 1 for l in set(ix)}
    sc = max((len(set(ix)) for ix in ixs), default=0)
    ops = 0.0
    nodes = []
    for j in range(len(left)):
        a, b = int(left[j]), int(right[j])
        ca, cb = cnt[a], cnt[b]
        if len(ca) < len(cb):
            ca, cb = cb, ca
        merged = ca  # reuse (small-to-large)
        for l, c in cb.items():
            merged[l] = merged.get(l, 0) + c
        la, lb = lab[a], lab[b]
        union = la | lb
        out = frozenset(l for l in union if merged[l] < total[l])
        for l in union - out:
            del merged[l]
        cnt[n_leaves + j] = merged
        cnt[a] = cnt[b] = None
        lab[n_leaves + j] = out
        sc = max(sc, len(out))
        ops += 2.0 ** len(union)
        if return_nodes:
            shared = la & lb
            nb = len(shared & out)
            nm = len((la - lb) & out)
            nn = len((lb - la) & out)
            nodes.append(dict(ra=len(la), rb=len(lb), rc=len(out), m=nm, n=nn, b=nb,
                              k=len(union) - len(out), tc=len(union)))
        lab[a] = lab[b] = None
    tc = float(np.log2(ops)) if ops > 0 else 0.0
    if return_nodes:
        return float(sc), tc, nodes
    return float(sc), tc


def big_label_histogram(ixs, tree, threshold: int) -> Dict[int, int]:
    """label -> number of intermediates of rank > threshold containing it (intent of sc_score,
    /root/reference/src/branch.jl:132-147)."""
    left, right = tree_to_postorder(tree, len(ixs))
    n_leaves = len(ixs)
    total: Dict[int, int] = {}
    for ix in ixs:
        for l in set(ix):
            total[l] = total.get(l, 0) + 1
    cnt: list = [None] * (n_leaves + len(left))
    for i, ix in enumerate(ixs):
        cnt[i] = {l: 1 for l in set(ix)}
    hist: Dict[int, int] = {}
    for j in range(len(left)):
        a, b = int(left[j]), int(right[j])
        ca, cb = cnt[a], cnt[b]
        if len(ca) < len(cb):
            ca, cb = cb, ca
        for l, c in cb.items():
            ca[l] = ca.get(l, 0) + c
        for l in [l for l, c in ca.items() if c >= total[l]]:
            del ca[l]
        cnt[n_leaves + j] = ca
        cnt[a] = cnt[b] = None
        if len(ca) > threshold:
            w = 1 << min(len(ca) - threshold, 20)
            for l in ca:
                hist[l] = hist.get(l, 0) + w
    return hist


def remove_vertices(br: Branch, removed: Sequence[int], dr: float = 0) -> Branch:
    """Child branch on the induced subgraph without `removed`: delete every leaf tensor touching a
    removed vertex, re-binarise the tree, renumber vertices 0..nv'-1 (vmap of generate_branch,
    /root/reference/src/branch.jl:212-232)."""
    removed = set(int(v) for v in removed)
    keep = [v for v in range(br.nv) if v not in removed]
    vmap = {v: i for i, v in enumerate(keep)}
    new_ixs: List[Tuple[int, ...]] = []
    leaf_map: Dict[int, int] = {}
    for i, ix in enumerate(br.ixs):
        if any(l in removed for l in ix):
            continue
        leaf_map[i] = len(new_ixs)
        new_ixs.append(tuple(vmap[l] for l in ix))

    def prune(t):
        # iterative prune to avoid recursion limits on path-like trees
        stack = [(t, False)]
        res: list = []
        while stack:
            node, done = stack.pop()
            if not isinstance(node, tuple):
                res.append(leaf_map.get(node))
                continue
            if not done:
                stack.append((node, True))
                stack.append((node[1], False))
                stack.append((node[0], False))
            else:
                r_ = res.pop()
                l_ = res.pop()
                if l_ is None:
                    res.append(r_)
                elif r_ is None:
                    res.append(l_)
                else:
                    res.append((l_, r_))
        return res[0]

    new_tree = prune(br.tree) if br.tree is not None else None
    edges = [(vmap[u], vmap[v]) for u, v in br.edges if u not in removed and v not in removed]
    w = None if br.weights is None else np.asarray(br.weights)[keep]
    return Branch(nv=len(keep), edges=edges, weights=w, ixs=new_ixs, tree=new_tree, r=br.r + dr,
                  meta=dict(br.meta))


# --------------------------------------------------------------------------------------------
# reductions + branching
# --------------------------------------------------------------------------------------------
def _adjacency(nv, edges):
    adj = [set() for _ in range(nv)]
    for u, v in edges:
        adj[u].add(v)
        adj[v].add(u)
    return adj


def kernelize(br: Branch) -> Branch:
    """Removal-only reductions to a fix point.  Unit weights: isolated vertex, pendant vertex,
    degree-2 vertex in a triangle, domination (N[u] subset N[v] => drop v).  Weighted: isolated
    vertex with w>0, pendant vertex with w_v >= w_u."""
    while True:
        adj = _adjacency(br.nv, br.edges)
        removed: set = set()
        gain = 0.0
        unit = br.weights is None
        w = (lambda v: 1) if unit else (lambda v: float(br.weights[v]))
        for v in range(br.nv):
            if v in removed:
                continue
            nb = [u for u in adj[v] if u not in removed]
            if len(nb) == 0:
                if w(v) > 0:
                    gain += w(v)
                removed.add(v)
            elif len(nb) == 1:
                u = nb[0]
                if w(v) >= w(u):
                    gain += w(v)
                    removed.update((v, u))
            elif unit and len(nb) == 2 and nb[1] in adj[nb[0]]:
                gain += 1
                removed.update((v, nb[0], nb[1]))
            if removed and len(removed) > 64:
                break
        if not removed and unit:
            # domination: N[u] subset N[v]  =>  v can be dropped
            for v in range(br.nv):
                nv_closed = adj[v] | {v}
                for u in adj[v]:
                    if len(adj[u]) <= len(adj[v]) and (adj[u] | {u}) <= nv_closed:
                        removed.add(v)
                        break
                if removed:
                    break
        if not removed:
            return br
        if unit:
            gain = int(gain)
        br = remove_vertices(br, sorted(removed), gain)


def slice_bfs(root: Branch, sc_target: int, max_branches: Optional[int] = None,
              reduce: bool = True) -> List[Branch]:
    """Vertex branching (v out | v in, N(v) out) until sc <= sc_target for every branch."""
    root = kernelize(root) if reduce else root
    unfinished = [root]
    finished: List[Branch] = []
    while unfinished:
        br = unfinished.pop()
        if br.nv == 0 or br.tree is None:
            finished.append(Branch(0, [], None, [], None, br.r, br.meta))
            continue
        sc, _ = tree_complexity(br.ixs, br.tree)
        if sc <= sc_target:
            finished.append(br)
            if max_branches is not None and len(finished) >= max_branches:
                break
            continue
        hist = big_label_histogram(br.ixs, br.tree, sc_target)
        adj = _adjacency(br.nv, br.edges)
        v = max(hist, key=lambda l: (hist[l], len(adj[l]), -l))
        wv = 1 if br.weights is None else float(br.weights[v])
        c_out = remove_vertices(br, [v], 0)
        c_in = remove_vertices(br, [v] + sorted(adj[v]), wv)
        for c in (c_out, c_in):
            unfinished.append(kernelize(c) if reduce else c)
    return finished


def make_root(nv, edges, weights=None, seed=0, ntrials=4) -> Branch:
    ixs = mis_ixs(nv, edges)
    best = None
    for t in range(ntrials):
        tree = greedy_tree(ixs, seed=seed * 1000 + t, temperature=0.0 if t == 0 else 0.5)
        sc, tc = tree_complexity(ixs, tree)
        if best is None or (sc, tc) < best[0]:
            best = ((sc, tc), tree)
    return Branch(nv=nv, edges=list(edges), weights=weights, ixs=ixs, tree=best[1], r=0)


def branch_list_hash(branches: Sequence[Branch]) -> str:
    """sha256 over the canonical content of a branch list (graph, weights, leaf labels, tree, r): pins the benched /
    golden workloads to what this (tracked) generator produces, whatever cache file they were loaded from."""
    import hashlib

    h = hashlib.sha256()
    for b in branches:
        h.update(repr((b.nv, [tuple(e) for e in b.edges], None if b.weights is None else [float(x) for x in b.weights],
                       [tuple(i) for i in b.ixs], b.tree, float(b.r))).encode())
    return h.hexdigest()
