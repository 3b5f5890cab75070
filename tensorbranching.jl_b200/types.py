"""Host-side mirror of the reference's data model for the hot path
(/root/reference/src/types.jl:51-121).  Labels and tensor ids are 0-based here (the reference is
1-based); the on-disk reader in io.py converts.

The mirror keeps the tree in the flat form the C ABI takes (CSR leaf labels + post-order child
arrays), built once when the branch is constructed -- the analogue of `compress`
(src/types.jl:64-69) -- so that contract_slices only passes pointers.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np


class UnitWeight:
    """UnitWeight(n) of ProblemReductions [upstream]: every vertex weighs 1."""

    def __init__(self, n: int):
        self.n = int(n)

    def __len__(self):
        return self.n

    def __eq__(self, other):
        return isinstance(other, UnitWeight) and other.n == self.n

    def __repr__(self):
        return f"UnitWeight({self.n})"


@dataclass
class MISProblem:
    """MISProblem(g, weights) [upstream OptimalBranchingMIS]; g = (nv, edge list)."""
    nv: int
    edges: List[Tuple[int, int]]
    weights: object  # UnitWeight or numpy vector

    def __post_init__(self):
        if self.weights is None:
            self.weights = UnitWeight(self.nv)


def _flatten_tree(tree, n_leaves: int):
    left: List[int] = []
    right: List[int] = []
    stack = [(tree, False)]
    ids: List[int] = []
    while stack:
        t, done = stack.pop()
        if not isinstance(t, (tuple, list)):
            ids.append(int(t))
        elif not done:
            if len(t) != 2:
                raise ValueError("contraction tree is not binary")
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


class CompressedEinsum:
    """CompressedEinsum{LT}(ixs, iy, ct) (src/types.jl:51-58).  `ct` is a nested (left, right) tuple
    with int leaves = index into ixs (src/decompose.jl:107-109), or already-flat child arrays."""

    def __init__(self, ixs: Sequence[Sequence[int]], iy: Sequence[int] = (), ct=None, *, flat=None):
        self.ixs = [tuple(int(l) for l in ix) for ix in ixs]
        self.iy = tuple(int(l) for l in iy)
        self.ct = ct
        n = len(self.ixs)
        self.leaf_off = np.zeros(n + 1, dtype=np.int32)
        np.cumsum([len(ix) for ix in self.ixs], out=self.leaf_off[1:])
        self.leaf_labels = np.asarray([l for ix in self.ixs for l in ix], dtype=np.int32)
        if flat is not None:
            self.node_left = np.ascontiguousarray(flat[0], dtype=np.int32)
            self.node_right = np.ascontiguousarray(flat[1], dtype=np.int32)
        elif n == 1:
            self.node_left = np.zeros(0, dtype=np.int32)
            self.node_right = np.zeros(0, dtype=np.int32)
        else:
            self.node_left, self.node_right = _flatten_tree(ct, n)
        self.open_labels = np.asarray(self.iy, dtype=np.int32)
        if len(self.node_left) != max(n - 1, 0):
            raise ValueError(f"tree has {len(self.node_left)} internal nodes for {n} leaves (not binary / not spanning)")


def compress(ixs, iy, ct) -> CompressedEinsum:
    """compress(code) (src/types.jl:64-69)."""
    return CompressedEinsum(ixs, iy, ct)


class SlicedBranch:
    """SlicedBranch{INT,VT,RT}(p, code, r) (src/types.jl:85-103).  code is None <=> empty graph
    (src/branch.jl:224)."""

    def __init__(self, p: MISProblem, code: Optional[CompressedEinsum], r=0):
        self.p = p
        self.code = code
        self.r = r

    @classmethod
    def from_parts(cls, nv, edges, weights, ixs, tree, r=0):
        p = MISProblem(nv, list(edges), weights)
        code = None if (tree is None and len(ixs) != 1) or nv == 0 else CompressedEinsum(ixs, (), tree)
        return cls(p, code, r)

    def __repr__(self):
        kind = "simple graph" if isinstance(self.p.weights, UnitWeight) else "weighted graph"
        return f"SlicedBranch: graph {{{self.p.nv}, {len(self.p.edges)}}} {kind}; fixed weight: {self.r}"


def add_r(branch: SlicedBranch, r) -> SlicedBranch:
    """add_r (src/types.jl:113)."""
    return SlicedBranch(branch.p, branch.code, type(branch.r)(branch.r + r))
