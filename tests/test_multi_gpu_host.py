"""world_size-2 gloo test of the N>1 path on CPU: LPT sharding + the single all-reduce(max).
The local contraction is replaced by the oracle here (no GPU in this container); the GPU ranks use
the engine (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from helpers import golden_branches, load_golden, to_sliced

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import tbcuda  # noqa: F401
    from tbcuda.multi_gpu import contract_slices_distributed
    from oracle import tropical_oracle as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rec = load_golden("rr100_sc10_unit.json")
    brs = golden_branches(rec)
    sliced = [to_sliced(b) for b in brs]
    by_id = {id(s): b for s, b in zip(sliced, brs)}
    seen = []

    def local(shard):
        seen.append(len(shard))
        return O.contract_slices([by_id[id(s)] for s in shard], np.float32)

    full = contract_slices_distributed(sliced, np.float32, local_contract=local)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full)
    np.save(os.path.join(out_dir, f"n{rank}.npy"), np.asarray(seen))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_contract_slices(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    rec = load_golden("rr100_sc10_unit.json")
    r0 = np.load(tmp_path / "r0.npy")
    r1 = np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1)
    assert np.array_equal(r0.astype(np.float64), np.asarray(rec["values"]))
    n0, n1 = int(np.load(tmp_path / "n0.npy")[0]), int(np.load(tmp_path / "n1.npy")[0])
    assert n0 + n1 == len(rec["values"]) and n0 > 0 and n1 > 0  # no unit contracted twice, both ranks worked


def _worker_sliced(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import tbcuda
    from tbcuda.multi_gpu import solve_slice_index_sliced_distributed
    from oracle import tropical_oracle as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    b = [x for x in golden_branches(load_golden("rr100_sc10_unit.json")) if x.nv >= 12][0]
    br = to_sliced(b)
    labels, _, _ = tbcuda.suggest_slices(br, -1, 3)
    seen = []

    def local(first, count):
        seen.append((first, count))
        return [O.solve_slice(b, np.float64, fixed={l: (a >> i) & 1 for i, l in enumerate(labels)})
                for a in range(first, first + count)]

    val, full = solve_slice_index_sliced_distributed(br, labels, np.float32, local_contract=local)
    np.save(os.path.join(out_dir, f"s{rank}.npy"), np.concatenate([[val, O.solve_slice(b, np.float32)], full,
                                                                   np.asarray(seen[0], dtype=np.float64)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_index_sliced_branch(tmp_path):
    """one branch, 2^3 index slices dealt to 2 ranks, one all-reduce(max): both ranks get the unsliced value."""
    port = _free_port()
    mp.spawn(_worker_sliced, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    s0, s1 = np.load(tmp_path / "s0.npy"), np.load(tmp_path / "s1.npy")
    assert np.array_equal(s0[:-2], s1[:-2])
    assert s0[0] == s0[1] and s0[2:10].max() == s0[0]
    assert tuple(s0[-2:]) == (0, 4) and tuple(s1[-2:]) == (4, 4)


def test_lpt_balances():
    from tbcuda.multi_gpu import shard_lpt

    rng = np.random.default_rng(0)
    costs = 2.0 ** rng.uniform(10, 30, size=500)
    for world in (2, 4, 8):
        owner = shard_lpt(costs, world)
        loads = np.array([costs[owner == r].sum() for r in range(world)])
        assert loads.max() <= max(costs.max(), 1.05 * loads.mean())
