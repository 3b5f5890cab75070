"""Generates the committed golden fixtures.

The reference holds no golden vectors for this path and cannot run here (no Julia), so the fixtures
pin the PROPERTY its tests assert (value == exact MIS / MWIS, /root/reference/test/slice.jl:32-33,
test/dynamic_ob.jl:20, test/utils.jl:34): each record stores the branch list produced by the
stand-in host for a seeded instance, the per-branch values (numpy oracle, following the stored
trees), and the instance's optimum from an independent exact solver (HiGHS MILP; clique search for
the small ones).  Run from the repo root:  python tests/golden/make_golden.py
(`python tests/golden/make_golden.py sliced` regenerates only the index-slicing / open-boundary fixtures.)
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


def record_sliced_open(name, nv, edges, weights, seed, element_type, k, n_open):
    """One branch (the unsliced root of the instance) with (a) the values of the 2^k index slices of k labels and (b) the
    open-boundary root tensor over n_open labels -- both from the numpy oracle, the labels picked without libtbcuda."""
    root = H.make_root(nv, edges, weights=weights, seed=seed)
    sc, _ = H.tree_complexity(root.ixs, root.tree)
    hist = H.big_label_histogram(root.ixs, root.tree, max(0, int(sc) - 4))
    labels = [l for l, _ in sorted(hist.items(), key=lambda kv: (-kv[1], kv[0]))[:k]]
    slice_values = [float(O.solve_slice(root, element_type, fixed={l: (a >> i) & 1 for i, l in enumerate(labels)}))
                    for a in range(1 << k)]
    whole = float(O.solve_slice(root, element_type))
    assert max(slice_values) == whole
    rng = np.random.default_rng(seed)
    open_labels = sorted(int(v) for v in rng.choice(nv, size=n_open, replace=False))
    left, right = O.nested_to_postorder(root.tree, len(root.ixs))
    w = None if weights is None else np.asarray(weights).astype(element_type)
    t, labs = O.contract_tree(root.ixs, left, right, w, element_type, open_labels=tuple(open_labels))
    assert float(np.max(t)) == whole
    rec = dict(name=name, element_type=np.dtype(element_type).name, value=whole,
               sliced_labels=labels, slice_values=[v if np.isfinite(v) else None for v in slice_values],
               open_labels=open_labels, open_tensor_labels=[int(l) for l in labs],
               open_tensor=[float(x) for x in np.asarray(t).reshape(-1)],  # C order over open_tensor_labels
               branch=dict(nv=root.nv, edges=[list(e) for e in root.edges],
                           weights=None if root.weights is None else [float(x) for x in root.weights],
                           weight_dtype=None if root.weights is None else str(np.asarray(root.weights).dtype),
                           ixs=[list(ix) for ix in root.ixs], tree=lst(root.tree), r=float(root.r)))
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(rec, f, separators=(",", ":"))
    print(name, "value", whole, "labels", labels, "open", open_labels)


def record_baseline(names):
    """BASELINE.json configs 2-5 as bench.py runs them (bench.WORKLOADS): one value per branch (value + r, what
    contract_slices returns) from the C oracle, ONCE, so that `pytest -m gpu` can assert bit-equality at full size.
    Integer-valued workloads run in the oracle's int16 value type (exact, 2-4x faster than f32); a prefix of every
    workload is cross-checked against the f32 value type (Tropical{Float32}, the reference's element_type).  The record
    also pins the branch list itself (sha256 over its canonical content, workloads.standin_host.branch_list_hash)."""
    import time

    import bench
    from oracle import c_oracle as CO

    for name in names:
        brs = bench.make_workload(name)
        t0 = time.time()
        flats = [None if b.nv == 0 else CO.flatten(b) for b in brs]
        vals, ops, th = CO.contract_batch(flats, "auto")
        dt = time.time() - t0
        r = np.array([float(b.r) for b in brs])
        n_chk = {"cfg4": 2}.get(name, min(len(brs), 256))
        v32, _, _ = CO.contract_batch(flats[:n_chk], "f32")
        assert np.array_equal(v32, vals[:n_chk]), "int16 and f32 value types of the C oracle disagree"
        rec = dict(workload=name, spec=[bench.WORKLOADS[name][0], bench.WORKLOADS[name][1], bench.WORKLOADS[name][2],
                                        bench.WORKLOADS[name][3]],
                   hash=H.branch_list_hash(brs), n=len(brs), values=[float(v) for v in (vals + r)],
                   total_ops=float(ops.sum()), mis=float((vals + r).max()),
                   oracle=f"oracle/c/tropical_ref.c value type auto (int16), {CO.simd()}, {th} threads, {dt:.0f} s; first {n_chk} "
                          "branches also contracted in f32: identical")
        if name == "cfg3":  # index slices of the single branch: labels chosen without libtbcuda
            k = bench.DEFAULT_SLICE_K[name]
            labels = bench.python_slice_labels(brs[0], k)
            sv, _, _ = CO.contract_index_slices(brs[0], labels, range(1 << k), "auto")
            feas = set(bench.feasible_assignments(brs[0], labels))
            rec["sliced_labels"] = [int(l) for l in labels]
            rec["slice_values"] = [float(v) + float(brs[0].r) if (a in feas and np.isfinite(v)) else None for a, v in enumerate(sv)]
            assert max(v for v in rec["slice_values"] if v is not None) == rec["values"][0]
        with open(os.path.join(OUT, f"baseline_{name}.json"), "w") as f:
            json.dump(rec, f, separators=(",", ":"))
        print(name, "branches", len(brs), "mis", rec["mis"], f"{dt:.0f}s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "baseline":  # full-size values of the BASELINE configs (minutes of CPU)
        record_baseline(sys.argv[2:] or ["cfg1", "cfg2", "cfg3", "cfg5", "cfg4"])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sliced":  # only the index-slicing / open-boundary fixtures
        nv, edges = H.random_regular_graph(60, 3, 5)
        record_sliced_open("rr60_sliced_open_unit", nv, edges, None, 5, np.float32, 4, 5)
        rng = np.random.default_rng(9)
        nv, edges = H.random_regular_graph(40, 3, 9)
        record_sliced_open("rr40_sliced_open_f32", nv, edges, (1.0 + rng.random(nv)).astype(np.float32), 9, np.float32, 3, 4)
        sys.exit(0)
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
