"""Committed golden vectors for index slicing and open-boundary contraction (tests/golden/rr*_sliced_open_*.json, made
by `python tests/golden/make_golden.py sliced` from the numpy oracle): the compiled plans (descriptor interpreter, CPU),
the C oracle and the GPU engine (through the C ABI) must reproduce them exactly."""
import struct

import numpy as np
import pytest

import desc_interp as DI
from helpers import align_to, device_tensor_as_ndarray, load_golden
from oracle import c_oracle as CO
from workloads import standin_host as H

NAMES = ["rr60_sliced_open_unit", "rr40_sliced_open_f32"]


def _load(tb, name):
    rec = load_golden(name + ".json")
    b = rec["branch"]

    def tup(t):
        return tuple(tup(x) for x in t) if isinstance(t, list) else t
    w = None if b["weights"] is None else np.asarray(b["weights"], dtype=np.dtype(b["weight_dtype"]))
    hb = H.Branch(nv=b["nv"], edges=[tuple(e) for e in b["edges"]], weights=w, ixs=[tuple(ix) for ix in b["ixs"]],
                  tree=tup(b["tree"]), r=b["r"])
    et = np.dtype(rec["element_type"]).type
    want_slices = np.array([-np.inf if v is None else v for v in rec["slice_values"]])
    return rec, hb, et, want_slices


def _sliced(tb, hb, open_labels=()):
    return tb.SlicedBranch(tb.MISProblem(hb.nv, hb.edges, hb.weights), tb.CompressedEinsum(hb.ixs, open_labels, hb.tree), hb.r)


@pytest.mark.parametrize("name", NAMES)
def test_interpreted_plans_reproduce_golden(tb, name):
    rec, hb, et, want = _load(tb, name)
    br = _sliced(tb, hb)
    labels = rec["sliced_labels"]
    for a in range(1 << len(labels)):
        v, _ = DI.run_plan(tb.Plan(br, fixed={l: (a >> i) & 1 for i, l in enumerate(labels)}))
        assert et(v) == et(want[a]), a
    # open boundary
    p = tb.Plan(_sliced(tb, hb, rec["open_labels"]))
    k = len(rec["open_labels"])
    _, arena = DI.run_plan(p)
    root_off = struct.unpack("4q", p.raw(5))[1]
    s = [x for x in p.steps() if x.rank_c == k and x.c_offset == root_off][-1]
    dl, darr = device_tensor_as_ndarray([s.labels_c[i] for i in range(k)],
                                        DI.to_float(arena[root_off:root_off + (1 << k)], p.info().value_type))
    want_t = np.asarray(rec["open_tensor"]).reshape((2,) * k)
    assert np.array_equal(align_to(dl, darr, rec["open_tensor_labels"]).astype(et), want_t.astype(et))


@pytest.mark.parametrize("name", NAMES)
def test_c_oracle_reproduces_golden_slices(tb, name):
    rec, hb, et, want = _load(tb, name)
    vals, _, _ = CO.contract_index_slices(hb, rec["sliced_labels"], range(len(want)))
    assert np.array_equal(vals.astype(et), want.astype(et))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_golden(tb, engine, name):
    rec, hb, et, want = _load(tb, name)
    br = _sliced(tb, hb)
    vals, status, mx = engine.contract_index_sliced(br, rec["sliced_labels"], element_type=et)
    assert (status == 0).all()
    assert np.array_equal(vals.astype(et), want.astype(et)) and et(mx) == et(rec["value"])
    p = tb.Plan(_sliced(tb, hb, rec["open_labels"]), engine=engine)
    labels, data = engine.contract_tensor(p)
    dl, darr = device_tensor_as_ndarray(labels, data)
    k = len(rec["open_labels"])
    want_t = np.asarray(rec["open_tensor"]).reshape((2,) * k)
    assert np.array_equal(align_to(dl, darr, rec["open_tensor_labels"]).astype(et), want_t.astype(et))
    p.close()
