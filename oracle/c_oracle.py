"""ctypes wrapper of oracle/c/libtropical_ref.so (plain-C + OpenMP restatement; test infrastructure and
the CPU baseline of bench.py).  Built by __graft_entry__.build()."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "c", "libtropical_ref.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run __graft_entry__.build()")
        _lib = C.CDLL(LIB_PATH)
        _lib.tref_contract_batch.restype = C.c_int
        try:  # all host cores, even when a launcher (torchrun) exported OMP_NUM_THREADS=1
            _lib.tref_set_threads(len(os.sched_getaffinity(0)))
        except AttributeError:
            _lib.tref_set_threads(os.cpu_count() or 1)
    return _lib


def flatten(branch):
    """Branch (workloads.standin_host) or tbcuda.SlicedBranch -> (n_labels, leaf_off, leaf_labels, left, right, weights)."""
    if hasattr(branch, "code"):  # SlicedBranch
        code = branch.code
        w = branch.p.weights
        w = None if not isinstance(w, np.ndarray) else np.ascontiguousarray(w, dtype=np.float64)
        return (branch.p.nv, code.leaf_off, code.leaf_labels, code.node_left, code.node_right, w)
    from .tropical_oracle import nested_to_postorder
    off = np.zeros(len(branch.ixs) + 1, dtype=np.int32)
    np.cumsum([len(ix) for ix in branch.ixs], out=off[1:])
    labs = np.asarray([l for ix in branch.ixs for l in ix], dtype=np.int32)
    left, right = nested_to_postorder(branch.tree, len(branch.ixs)) if len(branch.ixs) > 1 else ([], [])
    w = None if branch.weights is None else np.ascontiguousarray(branch.weights, dtype=np.float64)
    return (branch.nv, off, labs, np.asarray(left, dtype=np.int32), np.asarray(right, dtype=np.int32), w)


VALUE_TYPES = {"f32": 0, "i16": 1, "auto": 2}


def simd() -> str:
    """ISA the micro-kernels dispatch to on this host ("avx512" / "avx2" / "scalar")."""
    lib = load()
    lib.tref_simd.restype = C.c_char_p
    return lib.tref_simd().decode()


def contract_batch(flat_list, value_type="f32"):
    """flat_list: list of flatten() tuples (None for an empty graph).  -> (values float64, ops float64, threads).
    value_type: "f32" = Tropical{Float32} (the reference's default element_type); "i16" = int16 with a -2^14 sentinel
    (caller guarantees integer weights with sum |w| < 8192); "auto" = i16 where that holds, else f32."""
    lib = load()
    n = len(flat_list)
    ip = C.POINTER(C.c_int32)
    dp = C.POINTER(C.c_double)
    nl = (C.c_int32 * max(n, 1))()
    nv = (C.c_int32 * max(n, 1))()
    a_off = (ip * max(n, 1))()
    a_lab = (ip * max(n, 1))()
    a_l = (ip * max(n, 1))()
    a_r = (ip * max(n, 1))()
    a_w = (dp * max(n, 1))()
    any_w = False
    keep = []
    for i, f in enumerate(flat_list):
        if f is None:
            nl[i] = 0
            continue
        nlab, off, labs, left, right, w = f
        off = np.ascontiguousarray(off, dtype=np.int32)
        labs = np.ascontiguousarray(labs, dtype=np.int32)
        left = np.ascontiguousarray(left, dtype=np.int32)
        right = np.ascontiguousarray(right, dtype=np.int32)
        keep += [off, labs, left, right, w]
        nv[i] = nlab
        nl[i] = len(off) - 1
        a_off[i] = off.ctypes.data_as(ip)
        a_lab[i] = labs.ctypes.data_as(ip)
        a_l[i] = left.ctypes.data_as(ip)
        a_r[i] = right.ctypes.data_as(ip)
        if w is not None:
            any_w = True
            a_w[i] = w.ctypes.data_as(dp)
    vals = np.zeros(n, dtype=np.float64)
    ops = np.zeros(n, dtype=np.float64)
    lib.tref_contract_batch_vt.restype = C.c_int
    th = lib.tref_contract_batch_vt(n, nv, nl, a_off, a_lab, a_l, a_r, a_w if any_w else None, VALUE_TYPES[value_type],
                                    vals.ctypes.data_as(dp), ops.ctypes.data_as(dp))
    return vals, ops, th


def contract_slices(branches, element_type=np.float32, value_type="f32"):
    """contract_slices (/root/reference/src/dynamic_ob.jl:36-48) on the C oracle."""
    flats = [None if b.nv == 0 else flatten(b) for b in branches]
    vals, _, _ = contract_batch(flats, value_type)
    et = np.dtype(element_type).type
    return np.asarray([et(b.r) if b.nv == 0 else et(et(v) + et(b.r)) for b, v in zip(branches, vals)], dtype=element_type)


def contract_index_slices(branch, sliced_labels, assignments, value_type="f32"):
    """Index slices of ONE branch on the C oracle, one slice per OpenMP thread: assignment a holds
    sliced_labels[i] at bit i of a.  -> (values float64 WITHOUT r, ops float64, threads)."""
    lib = load()
    nlab, off, labs, left, right, w = flatten(branch)
    off = np.ascontiguousarray(off, dtype=np.int32)
    labs = np.ascontiguousarray(labs, dtype=np.int32)
    left = np.ascontiguousarray(left, dtype=np.int32)
    right = np.ascontiguousarray(right, dtype=np.int32)
    assignments = list(assignments)
    n = len(assignments)
    fixed = np.full((max(n, 1), max(nlab, 1)), -1, dtype=np.int8)
    for q, a in enumerate(assignments):
        for i, l in enumerate(sliced_labels):
            fixed[q, l] = (a >> i) & 1
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    vals = np.zeros(n, dtype=np.float64)
    ops = np.zeros(n, dtype=np.float64)
    lib.tref_contract_slices_of_vt.restype = C.c_int
    th = lib.tref_contract_slices_of_vt(n, nlab, len(off) - 1, off.ctypes.data_as(ip), labs.ctypes.data_as(ip),
                                        left.ctypes.data_as(ip), right.ctypes.data_as(ip),
                                        w.ctypes.data_as(dp) if w is not None else None,
                                        fixed.ctypes.data_as(C.POINTER(C.c_int8)), VALUE_TYPES[value_type],
                                        vals.ctypes.data_as(dp), ops.ctypes.data_as(dp))
    return vals, ops, th
