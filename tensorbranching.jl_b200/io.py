"""Reader / writer for the reference's on-disk slice directory -- the language-neutral
hand-off `slice_bfs_rw` / `slice_dfs_lp` already produce and `load_all_finished` consumes:

    slices.csv            id,sc,tc,r[,solution]                     /root/reference/src/slice.jl:16,153,212
    graph_<id>.dot        Graphs.savegraph default (LGFormat text)  /root/reference/src/io.jl:82,85   [format upstream, recalled]
    eincode_<id>.json     OMEinsum.writejson of the NestedEinsum    /root/reference/src/io.jl:58-61  [format upstream, recalled]
    weights_<id>.txt      "UnitWeight\\n<n>" or "<T>\\n<w1>\\n..."      /root/reference/src/io.jl:1-44

A Julia run elsewhere can therefore feed this engine its TRUE slicing decisions:
    branches = load_all_finished(dir);  contract_slices(branches, Float32, True)
mirrors /root/reference/test/slice.jl:31-32.  Labels and tensor indices are 1-based on disk (Julia) and
0-based in memory.
"""
from __future__ import annotations

import csv
import json
import os
from typing import List

import numpy as np

from .types import CompressedEinsum, MISProblem, SlicedBranch, UnitWeight

_JULIA_T = {"Float64": np.float64, "Float32": np.float32, "Int64": np.int64, "Int32": np.int32}
_NP_T = {np.dtype(v).name: k for k, v in _JULIA_T.items()}


# ---------------------------------------------------------------- weights (src/io.jl:1-44)
def save_weights(filename, weights):
    with open(filename, "w") as f:
        if weights is None or isinstance(weights, UnitWeight):
            n = weights.n if isinstance(weights, UnitWeight) else 0
            f.write(f"UnitWeight\n{n}\n")
        else:
            w = np.asarray(weights)
            f.write(_NP_T[w.dtype.name] + "\n")
            for x in w:
                f.write(repr(x.item()) + "\n")


def load_weights(filename):
    with open(filename) as f:
        line1 = f.readline().strip()
        if line1 == "UnitWeight":
            return UnitWeight(int(f.readline().strip()))
        if line1 not in _JULIA_T:
            raise ValueError(f"Unknown weight type: {line1}")  # same failure as src/io.jl:37
        return np.asarray([float(x) if "Float" in line1 else int(x) for x in f.read().split()], dtype=_JULIA_T[line1])


# ---------------------------------------------------------------- graph (LGFormat text)
def save_graph(filename, nv, edges):
    with open(filename, "w") as f:
        f.write(f"{nv},{len(edges)},u,graph,2,Int64,simplegraph\n")
        for u, v in edges:
            f.write(f"{u + 1},{v + 1}\n")


def load_graph(filename):
    with open(filename) as f:
        head = f.readline().strip().split(",")
        nv, ne = int(head[0]), int(head[1])
        edges = []
        for _ in range(ne):
            u, v = f.readline().strip().split(",")[:2]
            edges.append((int(u) - 1, int(v) - 1))
    return nv, edges


# ---------------------------------------------------------------- code (OMEinsum JSON)
def _tree_to_dict(tree, ixs, out_labels_of):
    if not isinstance(tree, tuple):
        return {"isleaf": True, "tensorindex": int(tree) + 1}, list(ixs[tree])
    (dl, ll), (dr, lr) = _tree_to_dict(tree[0], ixs, out_labels_of), _tree_to_dict(tree[1], ixs, out_labels_of)
    iy = out_labels_of(tree)
    return ({"isleaf": False, "args": [dl, dr],
             "eins": {"ixs": [[l + 1 for l in ll], [l + 1 for l in lr]], "iy": [l + 1 for l in iy]}}, iy)


def save_code(filename, code: CompressedEinsum):
    """writejson(filename, uncompress(code)) (src/io.jl:58-61, 86)."""
    if code is None:
        with open(filename, "w") as f:
            f.write("nothing")
        return
    # output labels of every subtree: labels occurring both inside and outside it
    total = {}
    for ix in code.ixs:
        for l in set(ix):
            total[l] = total.get(l, 0) + 1
    cache = {}

    def counts(t):
        if not isinstance(t, tuple):
            return {l: 1 for l in set(code.ixs[t])}
        k = id(t)
        if k not in cache:
            a, b = counts(t[0]), counts(t[1])
            m = dict(a)
            for l, c in b.items():
                m[l] = m.get(l, 0) + c
            cache[k] = m
        return cache[k]

    def out_labels_of(t):
        return sorted(l for l, c in counts(t).items() if c < total[l] or l in code.iy)

    tree = code.ct if code.ct is not None else 0
    d, _ = _tree_to_dict(tree, code.ixs, out_labels_of)
    doc = {"label-type": "Int64", "inputs": [[l + 1 for l in ix] for ix in code.ixs], "output": [l + 1 for l in code.iy], "tree": d}
    with open(filename, "w") as f:
        json.dump(doc, f)


def load_code(filename):
    """readjson (src/io.jl:63-79) -> CompressedEinsum (compress is applied on load, as the SlicedBranch
    constructor does, src/types.jl:94-101)."""
    with open(filename) as f:
        txt = f.read()
    if txt.strip() == "nothing":
        return None
    doc = json.loads(txt)
    ixs = [[int(l) - 1 for l in ix] for ix in doc["inputs"]]
    iy = [int(l) - 1 for l in doc["output"]]

    def conv(d):
        if d["isleaf"]:
            return int(d["tensorindex"]) - 1
        if len(d["args"]) != 2:
            raise ValueError("eincode is not binary")  # the reference asserts is_binary on compress
        return (conv(d["args"][0]), conv(d["args"][1]))

    tree = conv(doc["tree"])
    return CompressedEinsum(ixs, iy, tree if isinstance(tree, tuple) else None)


# ---------------------------------------------------------------- branches (src/io.jl:81-121)
def save_finished(dirname, branch: SlicedBranch, id_: int):
    save_graph(os.path.join(dirname, f"graph_{id_}.dot"), branch.p.nv, branch.p.edges)
    save_code(os.path.join(dirname, f"eincode_{id_}.json"), branch.code)
    save_weights(os.path.join(dirname, f"weights_{id_}.txt"), branch.p.weights)


def load_finished(dirname, id_: int):
    nv, edges = load_graph(os.path.join(dirname, f"graph_{id_}.dot"))
    code = load_code(os.path.join(dirname, f"eincode_{id_}.json"))
    weights = load_weights(os.path.join(dirname, f"weights_{id_}.txt"))
    return (nv, edges), code, weights


def save_slices(dirname, branches: List[SlicedBranch], complexities=None):
    """What slice_bfs_rw leaves behind (src/slice.jl:148-232): slices.csv + one file triple per branch."""
    os.makedirs(dirname, exist_ok=True)
    with open(os.path.join(dirname, "slices.csv"), "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["id", "sc", "tc", "r"])
        for i, br in enumerate(branches):
            sc, tc = complexities[i] if complexities else (0.0, 0.0)
            wr.writerow([i + 1, sc, tc, br.r])
            save_finished(dirname, br, i + 1)


def load_all_finished(dirname) -> List[SlicedBranch]:
    """load_all_finished (src/io.jl:113-121)."""
    out = []
    with open(os.path.join(dirname, "slices.csv")) as f:
        for row in csv.DictReader(f):
            id_ = int(row["id"])
            (nv, edges), code, weights = load_finished(dirname, id_)
            r = float(row["r"])
            r = int(r) if r == int(r) and "." not in row["r"] else r
            out.append(SlicedBranch(MISProblem(nv, edges, weights), code if nv > 0 else None, r))
    return out
