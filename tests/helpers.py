"""Shared helpers for the test-suite."""
import json
import os

import numpy as np

import tbcuda
from workloads import standin_host as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def to_sliced(b) -> "tbcuda.SlicedBranch":
    return tbcuda.SlicedBranch.from_parts(b.nv, b.edges, b.weights, b.ixs, b.tree, b.r)


def regular_root(n, seed, weights=None, d=3):
    nv, edges = H.random_regular_graph(n, d, seed)
    return H.make_root(nv, edges, weights=weights, seed=seed)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def golden_branches(rec):
    """golden record -> list of standin Branch objects."""
    out = []
    for b in rec["branches"]:
        def tup(t):
            return tuple(tup(x) for x in t) if isinstance(t, list) else t
        w = None if b["weights"] is None else np.asarray(b["weights"], dtype=np.dtype(b["weight_dtype"]))
        out.append(H.Branch(nv=b["nv"], edges=[tuple(e) for e in b["edges"]], weights=w,
                            ixs=[tuple(ix) for ix in b["ixs"]], tree=None if b["tree"] is None else tup(b["tree"]),
                            r=b["r"]))
    return out


def device_tensor_as_ndarray(labels, data):
    """device layout (labels in bit order, bit 0 fastest) -> (labels tuple, ndarray with axis i <-> labels[i])."""
    rank = len(labels)
    arr = np.asarray(data).reshape((2,) * rank)  # C order: first axis = highest bit
    return tuple(reversed(labels)), arr


def align_to(labels_src, arr, labels_dst):
    """transpose arr (axes = labels_src) to axis order labels_dst."""
    if len(labels_src) == 0:
        return arr
    return np.transpose(arr, [list(labels_src).index(l) for l in labels_dst])


def reduce_to(labels_src, arr, labels_dst):
    """device tensor (axes = labels_src) -> axes labels_dst.  A node whose long reduction was split keeps some of the
    reduced labels as extra output labels for its consumer to reduce (plan compiler, folded split-K): max over them
    first -- the reference's node tensor is the fully reduced one."""
    labels_src = list(labels_src)
    extra = [i for i, l in enumerate(labels_src) if l not in labels_dst]
    if extra:
        arr = arr.max(axis=tuple(extra))
        labels_src = [l for l in labels_src if l in labels_dst]
    return align_to(tuple(labels_src), arr, labels_dst)
