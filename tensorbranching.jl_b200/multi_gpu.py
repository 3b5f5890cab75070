"""Multi-GPU form of contract_slices with one PROCESS per GPU (torch.distributed): branches sharded by estimated cost,
ONE all-reduce(max) over the per-branch result vector.

The branches are independent (/root/reference/src/dynamic_ob.jl:38-46 is a plain loop), so there is no data-path
collective: every rank contracts its own shard on its own GPU; the only exchange is the result vector, which callers
consume per branch (src/slice.jl:39-48) and reduce with maximum (src/dynamic_ob.jl:27).  NCCL on GPU ranks; gloo works
for CPU-side tests of the plumbing.

A host that drives all GPUs from ONE process (the Julia host of the reference) needs none of this: `Engine(devices=[...])`
(tb_init_multi) does the same sharding and the all-reduce inside libtbcuda.so.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Sequence

import numpy as np


def shard_lpt(costs: Sequence[float], world: int) -> np.ndarray:
    """Longest-processing-time-first: sort units by cost, give each to the least-loaded rank.
    Returns owner[i] in [0, world).  Deterministic, identical on every rank."""
    import heapq

    cost = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-cost, kind="stable").tolist()
    cost = cost.tolist()
    heap = [(0.0, r) for r in range(world)]  # (load, rank): the least-loaded rank, the lowest rank among equals
    owner = [0] * len(cost)
    for i in order:
        load, r = heapq.heappop(heap)
        owner[i] = r
        heapq.heappush(heap, (load + cost[i], r))
    return np.asarray(owner, dtype=np.int64)


def branch_cost(branch) -> float:
    """Tropical ops of a branch (the reference's 2^tc, src/types.jl:120) from the label-set pass of the plan compiler
    (tb_estimate): what the LPT sharder balances.  0 for an empty graph."""
    from .contract import estimate

    return estimate(branch)[0]


def _default_device(device):
    import torch
    import torch.distributed as dist

    if device is not None:
        return device
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return "cpu"


def _default_engine(engine):
    """the rank's own GPU: LOCAL_RANK if a launcher set it, else torch's current device (never silently GPU 0 for all ranks)"""
    if engine is not None:
        return engine
    from . import contract as Cn

    dev = os.environ.get("LOCAL_RANK")
    if dev is None:
        try:
            import torch
            dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
        except Exception:  # noqa: BLE001
            dev = 0
    dev = int(dev)
    if Cn._default_engine is None or Cn._default_engine.device != dev:
        Cn._default_engine = Cn.Engine(dev)
    return Cn._default_engine


def allreduce_max_vector(local_values: np.ndarray, mine: np.ndarray, n: int, device=None, group=None) -> np.ndarray:
    """Every rank contributes values for its indices `mine`; returns the full length-n vector."""
    import torch
    import torch.distributed as dist

    full = torch.full((n,), -float("inf"), dtype=torch.float64, device=_default_device(device))
    if len(mine):
        full[torch.as_tensor(mine, device=full.device)] = torch.as_tensor(local_values, dtype=torch.float64, device=full.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(full, op=dist.ReduceOp.MAX, group=group)
    return full.cpu().numpy()


def distributed_costs(branches, cost_fn: Callable = branch_cost, device=None, group=None, threads: int = 0) -> np.ndarray:
    """cost of every branch, computed ONCE across the job: rank r estimates branches r, r + world, ... and one all-reduce
    (sum over a zero vector) gives every rank the whole vector -- so the sharding decision costs 1/world of a pass."""
    import torch
    import torch.distributed as dist

    n = len(branches)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    costs = np.zeros(n, dtype=np.float64)
    if cost_fn is branch_cost:  # one multi-threaded C call for this rank's share
        from .contract import estimate_many
        costs[rank::world] = estimate_many(branches[rank::world], threads)
    else:
        for i in range(rank, n, world):
            costs[i] = cost_fn(branches[i])
    if world > 1:
        t = torch.as_tensor(costs, device=_default_device(device))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        costs = t.cpu().numpy()
    return costs


def contract_slices_distributed(branches, element_type=np.float32, engine=None, costs: Optional[Sequence[float]] = None,
                                group=None, local_contract: Optional[Callable] = None, device=None) -> np.ndarray:
    """contract_slices over all ranks of the process group.  Every rank passes the SAME branch list and
    gets the SAME full result vector back.  `costs` default to the branches' tropical ops (tb_estimate, computed once
    across the job).  `local_contract(list_of_branches) -> values` defaults to this rank's engine (tests on CPU ranks
    inject a stand-in to exercise the sharding / collective plumbing)."""
    import torch.distributed as dist

    from .contract import contract_slices

    n = len(branches)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if costs is None:
        costs = distributed_costs(branches, device=device, group=group) if world > 1 else np.zeros(n)
    owner = shard_lpt(costs, world)
    mine = np.nonzero(owner == rank)[0]
    shard = [branches[i] for i in mine]
    if local_contract is None:
        vals = contract_slices(shard, element_type, True, engine=_default_engine(engine))
    else:
        vals = local_contract(shard)
    full = allreduce_max_vector(np.asarray(vals, dtype=np.float64), mine, n, device=device, group=group)
    return full.astype(element_type)


def slice_range(n_assign: int, world: int, rank: int):
    """Contiguous share [first, first + count) of the 2^k assignments for `rank` (slices of one branch cost the same)."""
    base, extra = divmod(n_assign, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def solve_slice_index_sliced_distributed(branch, sliced_labels, element_type=np.float32, engine=None, group=None,
                                         local_contract: Optional[Callable] = None, device=None):
    """ONE heavy branch over all ranks: the 2^k assignments of `sliced_labels` are dealt to the ranks in
    contiguous ranges, every rank contracts its range (tb_contract_sliced), and one all-reduce(max) over the length-2^k
    vector gives every rank every slice value.  Returns (value_of_the_branch, per-slice values), r not included.
    `local_contract(first, count) -> values` replaces the engine in CPU tests of the plumbing."""
    import torch.distributed as dist

    k = len(sliced_labels)
    n = 1 << k
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    first, count = slice_range(n, world, rank)
    if count == 0:
        vals = np.empty(0)
    elif local_contract is None:
        vals, status, _ = _default_engine(engine).contract_index_sliced(branch, sliced_labels, first, count, element_type)
    else:
        vals = local_contract(first, count)
    full = allreduce_max_vector(np.asarray(vals, dtype=np.float64), np.arange(first, first + count), n, device=device,
                                group=group)
    return np.dtype(element_type).type(full.max()), full.astype(element_type)
