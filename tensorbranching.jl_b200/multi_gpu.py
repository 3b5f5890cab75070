"""Multi-GPU form of contract_slices: one process per GPU (torch.distributed), branches sharded by
estimated cost, ONE all-reduce(max) over the per-branch result vector.

The branches are independent (/root/reference/src/dynamic_ob.jl:38-46 is a plain loop), so there is no
data-path collective: every rank contracts its own shard on its own GPU; the only exchange is the
result vector, which callers consume per branch (src/slice.jl:39-48) and reduce with maximum
(src/dynamic_ob.jl:27).  NCCL on GPU ranks; gloo works for CPU-side tests of the plumbing.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np


def shard_lpt(costs: Sequence[float], world: int) -> np.ndarray:
    """Longest-processing-time-first: sort units by cost, give each to the least-loaded rank.
    Returns owner[i] in [0, world).  Deterministic, identical on every rank."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world)
    owner = np.zeros(len(costs), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += costs[i]
    return owner


def branch_cost(branch) -> float:
    """2^tc proxy without compiling: sum over leaves is not enough, so use nv^2-ish fallbacks only when
    no plan statistics are available.  Prefer plan stats (tb_plan_info.ops)."""
    return float(max(1, branch.p.nv)) ** 2


def allreduce_max_vector(local_values: np.ndarray, mine: np.ndarray, n: int, device=None, group=None) -> np.ndarray:
    """Every rank contributes values for its indices `mine`; returns the full length-n vector."""
    import torch
    import torch.distributed as dist

    full = torch.full((n,), -float("inf"), dtype=torch.float64, device=device or "cpu")
    if len(mine):
        full[torch.as_tensor(mine, device=full.device)] = torch.as_tensor(local_values, dtype=torch.float64, device=full.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(full, op=dist.ReduceOp.MAX, group=group)
    return full.cpu().numpy()


def contract_slices_distributed(branches, element_type=np.float32, engine=None, costs: Optional[Sequence[float]] = None,
                                group=None, local_contract: Optional[Callable] = None, device=None) -> np.ndarray:
    """contract_slices over all ranks of the process group.  Every rank passes the SAME branch list and
    gets the SAME full result vector back.  `local_contract(list_of_branches) -> values` defaults to this
    rank's engine (tests on CPU ranks inject a stand-in to exercise the sharding / collective plumbing)."""
    import torch.distributed as dist

    from .contract import contract_slices

    n = len(branches)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if costs is None:
        costs = [branch_cost(b) for b in branches]
    owner = shard_lpt(costs, world)
    mine = np.nonzero(owner == rank)[0]
    shard = [branches[i] for i in mine]
    if local_contract is None:
        vals = contract_slices(shard, element_type, True, engine=engine)
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
        from .contract import default_engine
        eng = engine or default_engine()
        vals, status, _ = eng.contract_index_sliced(branch, sliced_labels, first, count, element_type)
    else:
        vals = local_contract(first, count)
    full = allreduce_max_vector(np.asarray(vals, dtype=np.float64), np.arange(first, first + count), n, device=device,
                                group=group)
    return np.dtype(element_type).type(full.max()), full.astype(element_type)
