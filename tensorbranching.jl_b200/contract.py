"""Host-side mirror of the reference's hot-path entry points, backed by libtbcuda.so.

    solve_slice(branch, element_type, usecuda)        /root/reference/src/dynamic_ob.jl:30-34
    contract_slices(branches, element_type, usecuda)  /root/reference/src/dynamic_ob.jl:36-48
    complexity / sc / tc                              /root/reference/src/types.jl:115-121

Same names, argument meaning and error behaviour.  `usecuda` is the switch position this library
occupies: usecuda=False raises (the CPU path is the reference's own TropicalGEMM route, which this
package deliberately does not contain -- there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _lib as L
from .types import CompressedEinsum, MISProblem, SlicedBranch, UnitWeight


def _weight_desc(weights, n_labels):
    """-> (keepalive array or None, dtype code)."""
    if weights is None or isinstance(weights, UnitWeight):
        return None, L.TB_WEIGHT_UNIT
    w = np.asarray(weights)
    if w.shape != (n_labels,):
        raise ValueError(f"weights has shape {w.shape}, expected ({n_labels},)")
    if w.dtype == np.float32:
        return np.ascontiguousarray(w), L.TB_WEIGHT_F32
    if w.dtype == np.float64:
        return np.ascontiguousarray(w), L.TB_WEIGHT_F64
    if w.dtype == np.int32:
        return np.ascontiguousarray(w), L.TB_WEIGHT_I32
    if np.issubdtype(w.dtype, np.integer):
        return np.ascontiguousarray(w, dtype=np.int64), L.TB_WEIGHT_I64
    return np.ascontiguousarray(w, dtype=np.float64), L.TB_WEIGHT_F64


def _value_type_for(element_type, weight_code):
    """element_type of the reference -> tb_value_type.  Integer-valued problems (UnitWeight / integer
    weights) are computed exactly in integers whatever float container is asked for; real weights are computed in
    Tropical{Float32}, or in Tropical{Float64} when element_type is Float64 (generic + fused kernels only: there is no
    tiled GEMM kernel for 8-byte values)."""
    et = np.dtype(element_type) if element_type is not None else None
    if weight_code in (L.TB_WEIGHT_UNIT, L.TB_WEIGHT_I32, L.TB_WEIGHT_I64):
        return L.TB_VALUE_AUTO  # packed int16 when the weights fit, else int32
    if et is not None and et == np.float64:
        return L.TB_VALUE_F64
    return L.TB_VALUE_F32


def _network_of(branch: SlicedBranch, element_type, flags=0, keep: Optional[list] = None):
    """tb_network view of a branch.  The struct only holds pointers into the branch's flat arrays
    (built once by CompressedEinsum, the analogue of `compress`), so it is cached on the branch."""
    key = (np.dtype(element_type).name if element_type is not None else None, flags)
    cache = branch.__dict__.setdefault("_net_cache", {})
    # the cached struct holds raw pointers into branch.code's arrays and the weight vector: replacing either on the branch
    # must not reuse them
    ident = (id(branch.code), id(branch.p.weights))
    if cache.get("ident") != ident:
        cache.clear()
        cache["ident"] = ident
    hit = cache.get(key)
    if hit is not None:
        return hit
    code = branch.code
    net = L.tb_network()
    net.n_labels = branch.p.nv
    net.n_leaves = len(code.ixs)
    net.leaf_off = code.leaf_off.ctypes.data_as(C.POINTER(C.c_int32))
    net.leaf_labels = code.leaf_labels.ctypes.data_as(C.POINTER(C.c_int32))
    net.n_open = len(code.open_labels)
    net.open_labels = code.open_labels.ctypes.data_as(C.POINTER(C.c_int32))
    net.node_left = code.node_left.ctypes.data_as(C.POINTER(C.c_int32))
    net.node_right = code.node_right.ctypes.data_as(C.POINTER(C.c_int32))
    w, wcode = _weight_desc(branch.p.weights, branch.p.nv)
    if keep is not None and w is not None:
        keep.append(w)
    net.weights = w.ctypes.data if w is not None else None
    net.weight_dtype = wcode
    net.value_type = _value_type_for(element_type, wcode)
    net.flags = flags
    cache[key] = (net, w)  # w is kept alive by the cache entry
    return net, w


_EMPTY_NET_BYTES = bytes(C.sizeof(L.tb_network))


def _network_bytes(branch: SlicedBranch, element_type, flags=0) -> bytes:
    key = ("bytes", np.dtype(element_type).name if element_type is not None else None, flags)
    cache = branch.__dict__.setdefault("_net_cache", {})
    if cache.get("ident") != (id(branch.code), id(branch.p.weights)):
        cache.clear()  # _network_of below re-keys the cache on the current code / weights objects
    hit = cache.get(key)
    if hit is None:
        net, _ = _network_of(branch, element_type, flags)
        hit = bytes(net)
        cache[key] = hit
    return hit


class Plan:
    """Compiled, device-resident form of one branch's contraction (tb_plan)."""

    def __init__(self, branch: SlicedBranch, element_type=np.float32, flags=0, engine: "Engine" = None,
                 fixed: Optional[dict] = None, value_type: Optional[int] = None):
        """fixed: {label: 0 | 1} -- index slicing, the labels this contraction holds at one value.
        value_type: a tb_value_type overriding what element_type implies (TB_VALUE_SIZE_CONFIG for the branching tables)."""
        lib = L.load()
        self._lib = lib
        self.handle = C.c_void_p()
        net, w = _network_of(branch, element_type, flags)
        self._keep = (branch, w)
        if value_type is not None:
            net = L.tb_network.from_buffer_copy(bytes(net))
            net.value_type = int(value_type)
        if fixed:
            net = L.tb_network.from_buffer_copy(bytes(net))
            fl = np.asarray(list(fixed.keys()), dtype=np.int32)
            fv = np.asarray([fixed[k] for k in fixed], dtype=np.uint8)
            net.n_fixed = len(fl)
            net.fixed_labels = fl.ctypes.data_as(C.POINTER(C.c_int32))
            net.fixed_values = fv.ctypes.data_as(C.POINTER(C.c_uint8))
            self._keep = (branch, w, fl, fv)
        L.check(lib.tb_plan_create(engine.handle if engine else None, C.byref(net), C.byref(self.handle)),
                engine.handle if engine else None)

    def reassign(self, fixed: dict) -> "Plan":
        """tb_plan_reassign: this plan (created with fixed labels) for another assignment of the same labels."""
        fl = self._keep[2] if len(self._keep) > 2 else []
        fv = np.asarray([fixed[int(l)] for l in fl] + [0], dtype=np.uint8)  # (+1 pad: never an empty buffer)
        q = Plan.__new__(Plan)
        q._lib = self._lib
        q.handle = C.c_void_p()
        q._keep = self._keep
        L.check(self._lib.tb_plan_reassign(self.handle, fv.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(q.handle)))
        return q

    def info(self) -> L.tb_plan_stats:
        st = L.tb_plan_stats()
        L.check(self._lib.tb_plan_info(self.handle, C.byref(st)))
        return st

    def steps(self) -> List[L.tb_step_info]:
        n = L.check(self._lib.tb_plan_export(self.handle, None, 0))
        arr = (L.tb_step_info * max(n, 1))()
        L.check(self._lib.tb_plan_export(self.handle, arr, n))
        return list(arr[:n])

    def raw(self, which: int) -> bytes:
        n = self._lib.tb_plan_export_raw(self.handle, which, None, 0)
        L.check(int(n))
        buf = C.create_string_buffer(max(int(n), 1))
        self._lib.tb_plan_export_raw(self.handle, which, buf, n)
        return buf.raw[:n]

    def close(self):
        if self.handle:
            self._lib.tb_plan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """tb_ctx: one engine per CUDA device."""

    def __init__(self, device: int = 0, arena_bytes: int = 0, max_wave: int = 0, host_threads: int = 0,
                 plan_flags: int = 0, devices: Optional[Sequence[int]] = None, streams_per_device: int = 0,
                 slice_budget: int = 0, timing: int = 0):
        """devices=[...]: a multi-GPU engine in this process (tb_init_multi): contract_slices / contract_plans shard the
        branches over the devices and combine with one ncclAllReduce(max) inside the library."""
        lib = L.load()
        self._lib = lib
        opts = L.tb_options(device=device, arena_bytes=arena_bytes, max_wave=max_wave,
                            host_threads=host_threads, plan_flags=plan_flags, streams_per_device=streams_per_device,
                            slice_budget=slice_budget, timing=timing)
        self.handle = C.c_void_p()
        if devices is not None:
            devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
            L.check(lib.tb_init_multi(devs, len(devices), C.byref(opts), C.byref(self.handle)))
            device = int(devices[0])
        else:
            L.check(lib.tb_init(C.byref(opts), C.byref(self.handle)))
        self.device = device
        self.n_devices = lib.tb_device_count(self.handle)

    def close(self):
        if self.handle:
            self._lib.tb_shutdown(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- single plan -------------------------------------------------------------------------
    def contract(self, plan: Plan) -> float:
        out = C.c_double()
        L.check(self._lib.tb_contract(self.handle, plan.handle, C.byref(out)), self.handle)
        return out.value

    def read_tensor(self, plan: Plan, node: int):
        """-> (labels in bit order, flat float64 array of 2^rank values)."""
        rank = C.c_int32()
        labels = (C.c_int32 * 32)()
        L.check(self._lib.tb_plan_read_tensor(self.handle, plan.handle, node, None, 0, labels, C.byref(rank)), self.handle)
        n = 1 << rank.value
        data = np.empty(n, dtype=np.float64)
        L.check(self._lib.tb_plan_read_tensor(self.handle, plan.handle, node, data.ctypes.data_as(C.POINTER(C.c_double)), n,
                                              labels, C.byref(rank)), self.handle)
        return list(labels[:rank.value]), data

    def contract_tensor(self, plan: Plan):
        """tb_contract_tensor: the root tensor of a plan with open labels -> (labels bit-0-first, 2^rank float64)."""
        labels = (C.c_int32 * 32)()
        rank = C.c_int32()
        L.check(self._lib.tb_contract_tensor(self.handle, plan.handle, None, 0, labels, C.byref(rank)), self.handle)
        data = np.empty(1 << rank.value, dtype=np.float64)
        L.check(self._lib.tb_contract_tensor(self.handle, plan.handle, data.ctypes.data_as(C.POINTER(C.c_double)),
                                             data.size, labels, C.byref(rank)), self.handle)
        return list(labels[:rank.value]), data

    def contract_table(self, plan: Plan):
        """tb_contract_table (plan created with value_type=TB_VALUE_SIZE_CONFIG and open labels): per boundary configuration
        the best size and one optimal vertex set.  -> (labels bit-0-first, sizes float64[2^rank], configs uint32[2^rank])"""
        labels = (C.c_int32 * 32)()
        rank = C.c_int32()
        L.check(self._lib.tb_contract_table(self.handle, plan.handle, None, None, 0, labels, C.byref(rank)), self.handle)
        n = 1 << rank.value
        sizes = np.empty(n, dtype=np.float64)
        cfgs = np.zeros(n, dtype=np.uint32)
        L.check(self._lib.tb_contract_table(self.handle, plan.handle, sizes.ctypes.data_as(C.POINTER(C.c_double)),
                                            cfgs.ctypes.data_as(C.POINTER(C.c_uint32)), n, labels, C.byref(rank)), self.handle)
        return list(labels[:rank.value]), sizes, cfgs

    def compactify_table(self, sizes: np.ndarray) -> np.ndarray:
        """tb_compactify_table (mis_compactify of the reference's table solver): which of the 2^rank boundary configurations
        survive -- feasible and not dominated by a configuration that chooses a subset of their boundary vertices and is at
        least as large.  -> bool[2^rank]"""
        sizes = np.ascontiguousarray(sizes, dtype=np.float64)
        rank = int(sizes.size).bit_length() - 1
        if sizes.size != 1 << rank:
            raise ValueError("a boundary table has 2^rank entries")
        keep = np.zeros(sizes.size, dtype=np.uint8)
        L.check(self._lib.tb_compactify_table(self.handle, rank, sizes.ctypes.data_as(C.POINTER(C.c_double)),
                                              keep.ctypes.data_as(C.POINTER(C.c_uint8))), self.handle)
        return keep.astype(bool)

    def table_configs(self, branch, labels, keep: Optional[np.ndarray] = None):
        """tb_table_configs: ALL optimal vertex sets of every boundary configuration of a region (the ConfigsMax rows of the
        reference's table solver).  branch = the region's SlicedBranch (or a Plan made from it), labels = the boundary
        vertices in the bit order of the rows (contract_table's labels), keep = compactify_table's flags.
        -> (sizes float64[2^rank], row_off int64[2^rank + 1], configs uint32[total])"""
        if isinstance(branch, Plan):
            branch = branch._keep[0]
        net, w = _network_of(branch, None, 0)
        rank = len(labels)
        n = 1 << rank
        lab = np.asarray(list(labels) + [0], dtype=np.int32)  # (+1 pad: never an empty buffer)
        kp = None
        if keep is not None:
            keep = np.ascontiguousarray(keep, dtype=np.uint8)
            if keep.size != n:
                raise ValueError("keep has one flag per boundary configuration")
            kp = keep.ctypes.data_as(C.POINTER(C.c_uint8))
        sizes, row_off, cfgs = self._table_call(self._lib.tb_table_configs, net, lab, rank, kp)
        del w
        return sizes, row_off, cfgs

    def _table_call(self, fn, net, lab, rank, keep_ptr):
        """tb_table_configs / tb_branching_table with a first guess for the number of configurations; one retry with the exact
        size when the table is larger (the failed call has already written the total)."""
        n = 1 << rank
        sizes = np.empty(n, dtype=np.float64)
        row_off = np.zeros(n + 1, dtype=np.int64)
        total = C.c_int64()
        cfgs = np.zeros(4096, dtype=np.uint32)
        for attempt in (0, 1):
            rc = fn(self.handle, C.byref(net), lab.ctypes.data_as(C.POINTER(C.c_int32)), rank, keep_ptr,
                    sizes.ctypes.data_as(C.POINTER(C.c_double)), row_off.ctypes.data_as(C.POINTER(C.c_int64)),
                    cfgs.ctypes.data_as(C.POINTER(C.c_uint32)), cfgs.size, C.byref(total))
            if rc == L.TB_ERR_BAD_ARGUMENT and attempt == 0 and total.value > cfgs.size:
                cfgs = np.zeros(total.value, dtype=np.uint32)
                continue
            L.check(rc, self.handle)
            break
        return sizes, row_off, cfgs[:total.value].copy()

    def region_table(self, branch, boundary):
        """tb_branching_table: the whole `branching_table(p, TensorNetworkSolver(), region)` of the reference (src/branch.jl:79)
        in one call -- row optima, mis_compactify and all optimal configurations of the surviving rows, all on the device.
        branch = the region's SlicedBranch (only its graph and weights are read), boundary = its open vertices (bit order of
        the rows).  -> (sizes float64[2^rank], keep bool[2^rank], rows = [(boundary bits, size, [vertex masks, ascending])])"""
        if isinstance(branch, Plan):
            branch = branch._keep[0]
        net, w = _network_of(branch, None, 0)
        rank = len(boundary)
        lab = np.asarray(list(boundary) + [0], dtype=np.int32)
        keep = np.zeros(1 << rank, dtype=np.uint8)
        sizes, row_off, cfgs = self._table_call(self._lib.tb_branching_table, net, lab, rank, keep.ctypes.data_as(C.POINTER(C.c_uint8)))
        del w
        keep = keep.astype(bool)
        return sizes, keep, [(int(a), float(sizes[a]), [int(c) for c in cfgs[row_off[a]:row_off[a + 1]]]) for a in np.nonzero(keep)[0]]

    def region_tables(self, branches, boundaries):
        """tb_branching_tables: the tables of many regions in the same launches.
        -> one (sizes, keep, rows) per region, each as region_table returns it"""
        n = len(branches)
        if n == 0:
            return []
        branches = [b._keep[0] if isinstance(b, Plan) else b for b in branches]
        pairs = [_network_of(b, None, 0) for b in branches]
        nets = (L.tb_network * n)(*[p[0] for p in pairs])
        ranks = [len(b) for b in boundaries]
        boff = np.zeros(n + 1, dtype=np.int32)
        boff[1:] = np.cumsum(ranks)
        lab = np.asarray([v for b in boundaries for v in b] + [0], dtype=np.int32)
        base = np.zeros(n + 1, dtype=np.int64)
        base[1:] = np.cumsum([1 << r for r in ranks])
        n_rows = int(base[-1])
        keep = np.zeros(n_rows, dtype=np.uint8)
        sizes = np.empty(n_rows, dtype=np.float64)
        row_off = np.zeros(n_rows + 1, dtype=np.int64)
        total = C.c_int64()
        cfgs = np.zeros(max(4096, 64 * n), dtype=np.uint32)
        for attempt in (0, 1):
            rc = self._lib.tb_branching_tables(self.handle, nets, boff.ctypes.data_as(C.POINTER(C.c_int32)),
                                               lab.ctypes.data_as(C.POINTER(C.c_int32)), n, keep.ctypes.data_as(C.POINTER(C.c_uint8)),
                                               sizes.ctypes.data_as(C.POINTER(C.c_double)), row_off.ctypes.data_as(C.POINTER(C.c_int64)),
                                               cfgs.ctypes.data_as(C.POINTER(C.c_uint32)), cfgs.size, C.byref(total))
            if rc == L.TB_ERR_BAD_ARGUMENT and attempt == 0 and total.value > cfgs.size:
                cfgs = np.zeros(total.value, dtype=np.uint32)
                continue
            L.check(rc, self.handle)
            break
        del pairs
        out = []
        for i in range(n):
            lo, hi = int(base[i]), int(base[i + 1])
            kp = keep[lo:hi].astype(bool)
            rows = [(int(a), float(sizes[lo + a]), [int(c) for c in cfgs[row_off[lo + a]:row_off[lo + a + 1]]]) for a in np.nonzero(kp)[0]]
            out.append((sizes[lo:hi].copy(), kp, rows))
        return out

    def branching_table(self, plan: Plan, all_configs: bool = False):
        """The table `branching_table(p, TensorNetworkSolver(), region)` hands to the set-cover solver (src/branch.jl:79):
        contract the region's network (boundary vertices open), drop the dominated boundary configurations.
        all_configs=False: ONE optimal configuration per row (size + configuration elements carried through the contraction)
        -> (boundary labels bit-0-first, rows) with rows = [(boundary bits, size, vertex mask of one optimal set), ...]
        all_configs=True: every optimal configuration of every surviving row, as the reference's ConfigsMax tables hold them
        -> (labels, rows) with rows = [(boundary bits, size, [vertex masks, ascending]), ...]"""
        if plan.info().value_type == L.TB_VALUE_SIZE_CONFIG:
            labels, sizes, cfgs = self.contract_table(plan)
        else:
            labels, sizes = self.contract_tensor(plan)
            cfgs = None
        keep = self.compactify_table(sizes)
        if not all_configs:
            if cfgs is None:
                raise ValueError("one configuration per row needs a plan created with value_type=TB_VALUE_SIZE_CONFIG")
            return labels, [(int(a), float(sizes[a]), int(cfgs[a])) for a in np.nonzero(keep)[0]]
        own_sizes, row_off, allc = self.table_configs(plan, labels, keep)
        if not np.array_equal(own_sizes, sizes):
            raise L.TBError(L.TB_ERR_INTERNAL, "the enumerated row optima differ from the contracted sizes")
        return labels, [(int(a), float(sizes[a]), [int(c) for c in allc[row_off[a]:row_off[a + 1]]]) for a in np.nonzero(keep)[0]]

    # -- batches -----------------------------------------------------------------------------
    def contract_plans(self, plans, r: Optional[np.ndarray] = None):
        """tb_contract_batch over resident plans (None = empty graph).  `plans` may be a PlanBatch: the handle array and
        the output buffers are then marshalled once and reused by every call (the returned arrays are overwritten by
        the next call on the same batch)."""
        if isinstance(plans, PlanBatch):
            b = plans
            mx = C.c_double()
            L.check(self._lib.tb_contract_batch(self.handle, b.arr, b.rp, b.n, b.outp, b.statusp, C.byref(mx)), self.handle)
            return b.out, b.status, mx.value
        n = len(plans)
        arr = (C.c_void_p * max(n, 1))(*[p.handle if p is not None else None for p in plans])
        out = np.empty(n, dtype=np.float64)
        status = np.zeros(n, dtype=np.int32)
        mx = C.c_double()
        rp = None
        if r is not None:
            r = np.ascontiguousarray(r, dtype=np.float64)
            rp = r.ctypes.data_as(C.POINTER(C.c_double))
        L.check(self._lib.tb_contract_batch(self.handle, arr, rp, n, out.ctypes.data_as(C.POINTER(C.c_double)),
                                            status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(mx)), self.handle)
        return out, status, mx.value

    def contract_branches(self, branches: Sequence[SlicedBranch], element_type=np.float32, flags=0):
        """The whole of contract_slices through ONE C call (compile + upload + contract):
        returns the contracted values WITHOUT r (float64)."""
        n = len(branches)
        # one tb_network record per branch; the records are cached as bytes on the branch objects, so a call only
        # joins them (the pointers inside stay valid as long as the branches are alive)
        key = ("bytes", np.dtype(element_type).name if element_type is not None else None, flags)
        parts = []
        for br in branches:
            c = br.__dict__.get("_net_cache")
            ident = (id(br.code), id(br.p.weights))
            if c is not None and c.get("ident") == ident:  # hot path: the branch already carries its record
                hit = c.get(key)
                if hit is not None:
                    parts.append(hit)
                    continue
            if br.p.nv == 0 or br.code is None:
                c = br.__dict__.setdefault("_net_cache", {})
                c.clear()
                c["ident"] = ident
                c[key] = _EMPTY_NET_BYTES
                parts.append(_EMPTY_NET_BYTES)
            else:
                parts.append(_network_bytes(br, element_type, flags))
        nets = (L.tb_network * max(n, 1)).from_buffer_copy(b"".join(parts) if n else _EMPTY_NET_BYTES)
        out = np.empty(n, dtype=np.float64)
        status = np.zeros(n, dtype=np.int32)
        mx = C.c_double()
        L.check(self._lib.tb_contract_networks(self.handle, nets, None, n, out.ctypes.data_as(C.POINTER(C.c_double)),
                                               status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(mx)), self.handle)
        return out, status

    def contract_index_sliced(self, branch: SlicedBranch, sliced_labels: Sequence[int], first: int = 0,
                              count: Optional[int] = None, element_type=np.float32, flags=0):
        """tb_contract_sliced: the 2^k assignments [first, first+count) of `sliced_labels` of ONE branch, each a
        contraction of the same tree without those labels.  -> (values WITHOUT r, status, max)."""
        lab = np.ascontiguousarray(sliced_labels, dtype=np.int32)
        k = len(lab)
        if count is None:
            count = (1 << k) - first
        net, _ = _network_of(branch, element_type, flags)
        out = np.empty(max(count, 0), dtype=np.float64)
        status = np.zeros(max(count, 0), dtype=np.int32)
        mx = C.c_double()
        L.check(self._lib.tb_contract_sliced(self.handle, C.byref(net), lab.ctypes.data_as(C.POINTER(C.c_int32)), k,
                                             first, count, 0.0, out.ctypes.data_as(C.POINTER(C.c_double)),
                                             status.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(mx)), self.handle)
        return out, status, mx.value

    def last_timing(self):
        ms = C.c_double()
        nl = C.c_int64()
        L.check(self._lib.tb_last_timing(self.handle, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def set_stream(self, cuda_stream_handle: int):
        L.check(self._lib.tb_set_stream(self.handle, C.c_void_p(cuda_stream_handle)), self.handle)

    def profile(self, mode=1):
        """tb_profile: 0 off, 1 per-launch events with the lanes concurrent, 2 single lane (launches serialised)."""
        L.check(self._lib.tb_profile(self.handle, int(mode)), self.handle)

    def last_profile_union(self):
        """per kind: how long that kind of kernel was on the device in the last call (union of its launches' intervals)."""
        ms = (C.c_double * 4)()
        L.check(self._lib.tb_last_profile_union(self.handle, ms))
        return dict(zip(("fused", "generic", "gemm", "finalize"), ms))

    def last_profile(self):
        ms = (C.c_double * 4)()
        nl = (C.c_int64 * 4)()
        L.check(self._lib.tb_last_profile(self.handle, ms, nl))
        names = ("fused", "generic", "gemm", "finalize")
        return {k: (ms[i], nl[i]) for i, k in enumerate(names)}

    def last_host_breakdown(self):
        ms = (C.c_double * 6)()
        L.check(self._lib.tb_last_host_breakdown(self.handle, ms))
        names = ("compile_ms", "upload_ms", "worklist_ms", "launch_wait_ms", "teardown_ms", "total_ms")
        return {k: ms[i] for i, k in enumerate(names)}

    def last_transfers(self):
        a = C.c_int64()
        b = C.c_int64()
        L.check(self._lib.tb_last_transfers(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def permute_bits(self, x: np.ndarray, perm: Sequence[int]) -> np.ndarray:
        x = np.ascontiguousarray(x)
        assert x.dtype.itemsize == 4
        rank = len(perm)
        assert x.size == 1 << rank
        out = np.empty_like(x)
        p = (C.c_int32 * max(rank, 1))(*perm)
        L.check(self._lib.tb_permute_bits(self.handle, x.ctypes.data, out.ctypes.data, rank, p), self.handle)
        return out


class PlanBatch:
    """A fixed list of resident plans (+ r) marshalled once for repeated Engine.contract_plans calls."""

    def __init__(self, plans: Sequence[Optional[Plan]], r: Optional[np.ndarray] = None):
        self.plans = list(plans)  # keeps the plans alive
        self.n = len(self.plans)
        self.arr = (C.c_void_p * max(self.n, 1))(*[p.handle if p is not None else None for p in self.plans])
        self.r = None if r is None else np.ascontiguousarray(r, dtype=np.float64)
        self.rp = None if self.r is None else self.r.ctypes.data_as(C.POINTER(C.c_double))
        self.out = np.empty(self.n, dtype=np.float64)
        self.status = np.zeros(self.n, dtype=np.int32)
        self.outp = self.out.ctypes.data_as(C.POINTER(C.c_double))
        self.statusp = self.status.ctypes.data_as(C.POINTER(C.c_int32))


class BranchStream:
    """Streaming hand-off (tb_stream_*): push branches as the slicer finishes them
    (src/slice.jl:79-86), the GPU contracts them while the host keeps slicing; finish() returns what
    contract_slices would have returned for the concatenation of everything pushed.

        with tbcuda.BranchStream(engine, capacity=100000) as st:
            for finished in slicer_rounds():
                st.push(finished)
        values = st.values            # element_type vector, push order
    """

    def __init__(self, engine: "Engine" = None, capacity: int = 1 << 20, element_type=np.float32, flags=0):
        self._eng = engine or default_engine()
        self._lib = self._eng._lib
        self._et = element_type
        self._flags = flags
        self._branches: List[SlicedBranch] = []
        self.handle = C.c_void_p()
        self.values = None
        L.check(self._lib.tb_stream_begin(self._eng.handle, capacity, C.byref(self.handle)), self._eng.handle)

    def push(self, branches: Sequence[SlicedBranch]):
        n = len(branches)
        if n == 0:
            return
        parts = [_EMPTY_NET_BYTES if (br.p.nv == 0 or br.code is None) else _network_bytes(br, self._et, self._flags)
                 for br in branches]
        nets = (L.tb_network * n).from_buffer_copy(b"".join(parts))
        L.check(self._lib.tb_stream_push(self.handle, nets, None, n), self._eng.handle)  # inputs are copied by the compiler
        self._branches.extend(branches)

    def finish(self) -> np.ndarray:
        if not self.handle:
            return self.values
        n = len(self._branches)
        out = np.empty(max(n, 1), dtype=np.float64)
        status = np.zeros(max(n, 1), dtype=np.int32)
        got = C.c_int64()
        mx = C.c_double()
        h, self.handle = self.handle, C.c_void_p()
        L.check(self._lib.tb_stream_finish(h, out.ctypes.data_as(C.POINTER(C.c_double)),
                                           status.ctypes.data_as(C.POINTER(C.c_int32)), max(n, 1), C.byref(got),
                                           C.byref(mx)), self._eng.handle)
        et = self._et
        r = np.array([br.r for br in self._branches], dtype=np.float64).astype(et)
        empty = np.array([br.code is None or br.p.nv == 0 for br in self._branches], dtype=bool)
        res = out[:n].astype(et) + r  # same arithmetic as contract_slices (src/dynamic_ob.jl:39-44)
        res[empty] = r[empty]
        self.values = res
        return res

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.finish()
        elif self.handle:  # close the stream without masking the original exception
            h, self.handle = self.handle, C.c_void_p()
            self._lib.tb_stream_finish(h, None, None, 0, None, None)
        return False


_default_engine: Optional[Engine] = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


def _require_cuda(usecuda):
    if not usecuda:
        raise L.TBError(L.TB_ERR_UNSUPPORTED,
                        "usecuda=False selects the reference's own CPU route (TropicalGEMM.jl); "
                        "tensorbranching.jl_b200 implements only the usecuda=True position and has no CPU fallback")


def solve_slice(branch: SlicedBranch, element_type=np.float32, usecuda: bool = True, engine: Engine = None):
    """solve_slice (src/dynamic_ob.jl:30-34): the contracted value of one branch, as element_type."""
    _require_cuda(usecuda)
    eng = engine or default_engine()
    vals, status = eng.contract_branches([branch], element_type)
    return np.dtype(element_type).type(vals[0])


def contract_slices(branches: Sequence[SlicedBranch], element_type=np.float32, usecuda: bool = True,
                    engine: Engine = None) -> np.ndarray:
    """contract_slices (src/dynamic_ob.jl:36-48): one value per branch, in input order:
    element_type(r) for an empty graph, else solve_slice + element_type(r) in element_type arithmetic."""
    _require_cuda(usecuda)
    et = np.dtype(element_type).type
    eng = engine or default_engine()
    vals, status = eng.contract_branches(branches, element_type)
    n = len(branches)
    r = np.fromiter((br.r for br in branches), dtype=np.float64, count=n).astype(element_type)
    res = vals.astype(element_type) + r  # element_type arithmetic, as t + element_type(branch.r) in the reference
    # empty graph => element_type(r) (src/dynamic_ob.jl:39-40): the engine contracts nothing and returns 0 for those
    # entries, so 0 + r is already the answer
    return res


def estimate(branch: SlicedBranch):
    """tb_estimate: (tropical ops, sc) of a branch from the label-set pass alone -- the cost a sharder needs."""
    if branch.code is None or branch.p.nv == 0:
        return 0.0, 0.0
    lib = L.load()
    net, _ = _network_of(branch, np.float32, 0)
    ops = C.c_double()
    sc_ = C.c_double()
    L.check(lib.tb_estimate(C.byref(net), C.byref(ops), C.byref(sc_)))
    return ops.value, sc_.value


def estimate_many(branches: Sequence[SlicedBranch], threads: int = 0) -> np.ndarray:
    """tb_estimate_many: tropical ops of every branch (0 for an empty graph) in ONE multi-threaded C call."""
    n = len(branches)
    if n == 0:
        return np.zeros(0)
    lib = L.load()
    parts = [_EMPTY_NET_BYTES if (br.p.nv == 0 or br.code is None) else _network_bytes(br, np.float32, 0) for br in branches]
    nets = (L.tb_network * n).from_buffer_copy(b"".join(parts))
    ops = np.zeros(n, dtype=np.float64)
    L.check(lib.tb_estimate_many(nets, n, threads, ops.ctypes.data_as(C.POINTER(C.c_double)), None))
    return ops


def suggest_slices(branch: SlicedBranch, sc_target: int = -1, max_sliced: int = 8):
    """tb_suggest_slices: greedy choice of the labels to index-slice.  -> (labels, sc, tc) per slice afterwards."""
    lib = L.load()
    net, _ = _network_of(branch, np.float32, 0)
    out = (C.c_int32 * max(max_sliced, 1))()
    sc_ = C.c_double()
    tc_ = C.c_double()
    n = L.check(lib.tb_suggest_slices(None, C.byref(net), sc_target, max_sliced, out, C.byref(sc_), C.byref(tc_)))
    return list(out[:n]), sc_.value, tc_.value


def solve_slice_index_sliced(branch: SlicedBranch, sliced_labels: Sequence[int], element_type=np.float32,
                             usecuda: bool = True, engine: Engine = None):
    """solve_slice (src/dynamic_ob.jl:30-34) of ONE heavy branch computed as the max over the 2^k index slices of
    `sliced_labels`: the same value, from 2^k independent contractions that can be sharded."""
    _require_cuda(usecuda)
    eng = engine or default_engine()
    _, status, mx = eng.contract_index_sliced(branch, sliced_labels, element_type=element_type)
    return np.dtype(element_type).type(mx)


def complexity(branch: SlicedBranch):
    """complexity(branch) (src/types.jl:115-119): (tc, sc) of the branch's tree; zeros for code=None."""
    if branch.code is None:
        return dict(tc=0.0, sc=0.0, rwc=0.0)
    p = Plan(branch)
    st = p.info()
    p.close()
    return dict(tc=st.tc, sc=st.sc)


def contraction_peak_memory(branch: SlicedBranch) -> float:
    """contraction_peak_memory(code, uniformsize(code, 2)) (src/utils.jl:197-219), log2 elements."""
    p = Plan(branch)
    v = p.info().peak_memory_log2
    p.close()
    return v


def contraction_all_memory(branch: SlicedBranch) -> float:
    """contraction_all_memory (src/utils.jl:222-229), log2 elements."""
    p = Plan(branch)
    v = p.info().all_memory_log2
    p.close()
    return v


def tc(branch):  # src/types.jl:120
    return complexity(branch)["tc"]


def sc(branch):  # src/types.jl:121
    return complexity(branch)["sc"]
