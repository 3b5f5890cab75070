"""Round trip of the reference's on-disk slice directory (mirrors /root/reference/test/io.jl:9-38 and
the load_all_finished -> contract_slices pipeline of /root/reference/test/slice.jl:30-33)."""
import numpy as np
import pytest

import desc_interp as DI
from helpers import golden_branches, load_golden, to_sliced


@pytest.mark.parametrize("name", ["rr100_sc10_unit", "rr100_sc10_f32"])
def test_slice_directory_round_trip(tb, tmp_path, name):
    from tbcuda import io as tio

    rec = load_golden(name + ".json")
    brs = [to_sliced(b) for b in golden_branches(rec)][:10]
    tio.save_slices(str(tmp_path), brs)
    back = tio.load_all_finished(str(tmp_path))
    assert len(back) == len(brs)
    for a, b, want in zip(brs, back, rec["values"]):
        assert a.p.nv == b.p.nv and sorted(a.p.edges) == sorted(b.p.edges) and a.r == b.r
        if a.code is None:
            assert b.code is None
            continue
        assert a.code.ixs == b.code.ixs
        assert np.array_equal(a.code.node_left, b.code.node_left) and np.array_equal(a.code.node_right, b.code.node_right)
        if isinstance(a.p.weights, np.ndarray):
            assert np.array_equal(a.p.weights, b.p.weights) and a.p.weights.dtype == b.p.weights.dtype
        else:
            assert a.p.weights == b.p.weights
        val, _ = DI.run_plan(tb.Plan(b))
        assert np.float32(np.float32(val) + np.float32(b.r)) == np.float32(want)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int32])
def test_weights_round_trip(tb, tmp_path, dtype):
    from tbcuda import io as tio

    rng = np.random.default_rng(1)
    w = (rng.random(100) * 100).astype(dtype)
    fn = str(tmp_path / "weights.txt")
    tio.save_weights(fn, w)
    back = tio.load_weights(fn)
    assert back.dtype == w.dtype and np.array_equal(back, w)
    tio.save_weights(fn, tb.UnitWeight(100))
    assert tio.load_weights(fn) == tb.UnitWeight(100)


@pytest.mark.gpu
def test_load_all_finished_then_contract_slices(tb, engine, tmp_path):
    from tbcuda import io as tio

    rec = load_golden("rr100_sc10_unit.json")
    brs = [to_sliced(b) for b in golden_branches(rec)]
    tio.save_slices(str(tmp_path), brs)
    res = tb.contract_slices(tio.load_all_finished(str(tmp_path)), np.float32, True, engine=engine)
    assert float(res.max()) == rec["exact"]
    assert np.array_equal(res.astype(np.float64), np.asarray(rec["values"]))
