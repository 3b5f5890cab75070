"""Multi-GPU inside the library (tb_init_multi): ONE process, one sub-context per device, LPT sharding in C,
one ncclAllReduce(ncclMax) over the result vector in place of maximum(res) (/root/reference/src/dynamic_ob.jl:27).

On a single-GPU box the sharding path is exercised on a repeated device ([0, 0, 0]: host combine, NCCL cannot span one GPU
twice); with >= 2 GPUs (gpurun --gpus 2) the same tests run over real devices and NCCL."""
import numpy as np
import pytest

from helpers import golden_branches, load_golden, regular_root, to_sliced
from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(n)))
    return lists


@pytest.mark.parametrize("which", [0, 1, 2])
def test_contract_slices_on_a_multi_device_engine(tb, which):
    lists = _device_lists()
    if which >= len(lists):
        pytest.skip("needs >= 2 GPUs")
    devices = lists[which]
    eng = tb.Engine(devices=devices)
    assert eng.n_devices == len(devices)
    for name in ("rr100_sc10_unit", "rr100_sc10_f32", "rr30_disconnected"):
        rec = load_golden(name + ".json")
        et = np.dtype(rec["element_type"]).type
        brs = [to_sliced(b) for b in golden_branches(rec)]
        got = tb.contract_slices(brs, et, True, engine=eng)
        assert np.array_equal(got.astype(np.float64), np.asarray(rec["values"])), (name, devices)
    assert tb.contract_slices([], np.float32, True, engine=eng).shape == (0,)
    # more devices than branches: some devices get nothing and still take part in the reduction
    rec = load_golden("rr100_sc10_unit.json")
    brs = [to_sliced(b) for b in golden_branches(rec)][:2]
    got = tb.contract_slices(brs, np.float32, True, engine=eng)
    assert np.array_equal(got.astype(np.float64), np.asarray(rec["values"][:2]))
    eng.close()


@pytest.mark.parametrize("which", [1, 2])
def test_resident_plans_stay_on_their_device(tb, which):
    lists = _device_lists()
    if which >= len(lists):
        pytest.skip("needs >= 2 GPUs")
    eng = tb.Engine(devices=lists[which])
    rec = load_golden("rr100_sc10_unit.json")
    brs = golden_branches(rec)
    plans = [tb.Plan(to_sliced(b), engine=eng) if b.nv else None for b in brs]
    r = np.array([b.r for b in brs], dtype=np.float64)
    batch = tb.PlanBatch(plans, r)
    for _ in range(3):  # the second and third call find every plan resident on the device the first call chose
        vals, status, mx = eng.contract_plans(batch)
        assert not status.any() and np.array_equal(vals, np.asarray(rec["values"])) and mx == rec["exact"]
    ms, launches = eng.last_timing()
    assert ms > 0 and launches > 0
    for p in plans:
        if p is not None:
            p.close()
    eng.close()


@pytest.mark.parametrize("which", [1, 2])
def test_index_slices_and_slice_budget(tb, which):
    lists = _device_lists()
    if which >= len(lists):
        pytest.skip("needs >= 2 GPUs")
    devices = lists[which]
    root = regular_root(130, 5)
    want = CO.contract_slices([root], np.float32)[0]
    s = to_sliced(root)
    eng = tb.Engine(devices=devices)
    labels, _, _ = tb.suggest_slices(s, -1, 4)
    vals, status, mx = eng.contract_index_sliced(s, labels)
    assert not status.any() and mx == want and vals.max() == want
    single = tb.Engine(0)
    v1, _, _ = single.contract_index_sliced(s, labels)
    assert np.array_equal(vals, v1)  # every one of the 2^4 slice values, whatever device contracted it
    single.close()
    eng.close()
    # slice_budget: ONE branch on several devices is cut into index slices by the library itself (BASELINE config 3)
    eng = tb.Engine(devices=devices, slice_budget=5)
    got = tb.contract_slices([s], np.float32, True, engine=eng)
    assert got[0] == want
    assert eng.last_timing()[1] >= 2 * len(devices)  # several slices were launched, not one contraction
    eng.close()


def test_multi_device_errors(tb):
    with pytest.raises(tb.TBError) as e:
        tb.Engine(devices=[0, 99])
    assert e.value.code in (-1, -5)
    eng = tb.Engine(devices=[0, 0])
    with pytest.raises(tb.TBError) as e:
        eng.set_stream(0)
    assert e.value.code == -3
    eng.close()
