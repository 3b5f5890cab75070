"""Generates the committed golden fixtures.

The reference holds no golden vectors for this path and cannot run here (no Julia), so the fixtures
pin the PROPERTY its tests assert (value == exact MIS / MWIS, /root/reference/test/slice.jl:32-33,
test/dynamic_ob.jl:20, test/utils.jl:34): each record stores the branch list produced by the
stand-in host for a seeded instance, the per-branch values (numpy oracle, following the stored
trees), and the instance's optimum from an independent exact solver (HiGHS MILP; clique search for
the small ones).  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tropical_oracle as O  # noqa: E402
from workloads import standin_host as H  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def lst(t):
    return [lst(x) for x in t] if isinstance(t, tuple) else t


def record(name, nv, edges, weights, sc_target, seed, element_type, reduce=True):
    root = H.make_root(nv, edges, weights=weights, seed=seed)
    brs = H.slice_bfs(root, sc_target, reduce=reduce)
    vals = O.contract_slices(brs, element_type)
    exact = O.exact_mis_milp(nv, edges, weights)
    assert abs(float(vals.max()) - exact) <= 1e-4 * max(1.0, abs(exact)), (vals.max(), exact)
    rec = dict(name=name, nv=nv, edges=[list(e) for e in edges], sc_target=sc_target, seed=seed,
               weights=None if weights is None else [float(x) for x in weights],
               weight_dtype=None if weights is None else str(np.asarray(weights).dtype),
               element_type=np.dtype(element_type).name, exact=exact,
               values=[float(v) for v in vals],
               branches=[dict(nv=b.nv, edges=[list(e) for e in b.edges],
                              weights=None if b.weights is None else [float(x) for x in b.weights],
                              weight_dtype=None if b.weights is None else str(np.asarray(b.weights).dtype),
                              ixs=[list(ix) for ix in b.ixs], tree=lst(b.tree), r=float(b.r)) for b in brs])
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(rec, f, separators=(",", ":"))
    print(name, "branches", len(brs), "exact", exact)


if __name__ == "__main__":
    # config 1 of BASELINE.json: README example shape (3-regular n=100, sc_target=10)
    nv, edges = H.random_regular_graph(100, 3, 1)
    record("rr100_sc10_unit", nv, edges, None, 10, 1, np.float32)
    # weighted variant mirroring test/dynamic_ob.jl:15  (Float32 weights 1 + rand)
    rng = np.random.default_rng(15)
    w = (1.0 + rng.random(nv)).astype(np.float32)
    record("rr100_sc10_f32", nv, edges, w, 10, 1, np.float32)
    # disconnected / isolated-vertex / 2-vertex-component corner cases (test/decompose.jl:62-85)
    nv2, edges2 = H.random_regular_graph(30, 3, 5)
    edges2 = [(u, v) for u, v in edges2 if 3 not in (u, v) and 7 not in (u, v)] + [(3, 7)]
    root = None
    record("rr30_disconnected", nv2, sorted(edges2), None, 6, 5, np.float32)
    # small KSG
    nv3, edges3 = H.random_ksg(8, 8, 0.8, 3)
    record("ksg8x8_sc6", nv3, edges3, None, 6, 3, np.float32)
    # the same kind of instance WITHOUT the stand-in kernelisation (which solves small KSGs outright): dense degree-8 network
    nv4, edges4 = H.random_ksg(7, 7, 0.8, 4)
    record("ksg7x7_sc8_nokernel", nv4, edges4, None, 8, 4, np.float32, reduce=False)
