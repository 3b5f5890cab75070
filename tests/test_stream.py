"""Streaming hand-off (tb_stream_*): branches pushed in rounds, as the reference's slicer finishes them
(src/slice.jl:79-86), give exactly what one contract_slices call over the concatenation gives."""
import ctypes as C

import numpy as np
import pytest

from helpers import golden_branches, load_golden, to_sliced
from oracle import tropical_oracle as O


def test_stream_null_arguments_are_refused(tb):
    lib = tb.load()
    h = C.c_void_p()
    assert lib.tb_stream_begin(None, 10, C.byref(h)) == -1
    assert lib.tb_stream_push(None, None, None, 0) == -1
    assert lib.tb_stream_finish(None, None, None, 0, None, None) == -1


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rr100_sc10_unit", "rr100_sc10_f32", "rr30_disconnected"])
def test_stream_equals_contract_slices(tb, engine, name):
    rec = load_golden(name + ".json")
    et = np.dtype(rec["element_type"]).type
    brs = [to_sliced(b) for b in golden_branches(rec)]
    want = np.asarray(rec["values"])
    with tb.BranchStream(engine, capacity=len(brs) + 5, element_type=et) as st:
        # uneven rounds, including an empty one
        cuts = sorted(min(c, len(brs)) for c in (0, 1, 1, 4, len(brs) // 2, len(brs)))
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            st.push(brs[lo:hi])
    assert np.array_equal(st.values.astype(np.float64), want)
    # the context is usable again after the stream closed
    assert np.array_equal(tb.contract_slices(brs, et, True, engine=engine).astype(np.float64), want)


@pytest.mark.gpu
def test_stream_blocks_other_calls_and_capacity(tb, engine):
    rec = load_golden("rr100_sc10_unit.json")
    brs = [to_sliced(b) for b in golden_branches(rec)]
    st = tb.BranchStream(engine, capacity=3)
    st.push(brs[:2])
    with pytest.raises(tb.TBError):          # the context belongs to the stream
        tb.contract_slices(brs[:1], np.float32, True, engine=engine)
    with pytest.raises(tb.TBError):          # over capacity
        st.push(brs[2:6])
    vals = st.finish()
    assert np.array_equal(vals.astype(np.float64), np.asarray(rec["values"][:2]))
    assert np.array_equal(tb.contract_slices(brs[:3], np.float32, True, engine=engine).astype(np.float64),
                          np.asarray(rec["values"][:3]))


@pytest.mark.gpu
def test_plans_may_outlive_their_context(tb):
    """tb_shutdown detaches resident plans: destroying them afterwards (a GC's order) is safe, and a detached plan
    can be contracted again on another context."""
    rec = load_golden("rr100_sc10_unit.json")
    brs = [to_sliced(b) for b in golden_branches(rec) if b.nv > 0][:40]
    eng = tb.Engine(0)
    plans = [tb.Plan(b, engine=eng) for b in brs]
    want, _, _ = eng.contract_plans(plans)
    eng.close()                       # plans still hold device-side descriptors of the dead context
    for p in plans[:20]:
        p.close()
    eng2 = tb.Engine(0)
    got, _, _ = eng2.contract_plans(plans[20:])
    assert np.array_equal(got, want[20:])
    eng2.close()
    for p in plans[20:]:
        p.close()
