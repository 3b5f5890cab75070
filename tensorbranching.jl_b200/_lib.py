"""ctypes binding of libtbcuda.so (include/tbcuda.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TBCUDA_LIB") or os.path.join(_HERE, "libtbcuda.so")  # TBCUDA_LIB: diagnostics builds (e.g. -DTB_KPROF)

TB_OK = 0
TB_ERR_BAD_ARGUMENT = -1
TB_ERR_NOT_BINARY_TREE = -2
TB_ERR_UNSUPPORTED = -3
TB_ERR_OUT_OF_MEMORY = -4
TB_ERR_CUDA = -5
TB_ERR_NCCL = -6
TB_ERR_INTERNAL = -7

TB_VALUE_AUTO, TB_VALUE_I32, TB_VALUE_F32, TB_VALUE_I16X2, TB_VALUE_F64, TB_VALUE_SIZE_CONFIG = 0, 1, 2, 3, 4, 5
TB_WEIGHT_UNIT, TB_WEIGHT_I32, TB_WEIGHT_I64, TB_WEIGHT_F32, TB_WEIGHT_F64 = 0, 1, 2, 3, 4
TB_PLAN_KEEP_INTERMEDIATES = 1
TB_PLAN_NO_FUSED_SUBTREES = 2
TB_PLAN_NO_GEMM = 4
TB_PLAN_SCRAMBLE_LAYOUT = 8
TB_PLAN_NO_SPLIT_K = 16
TB_PLAN_PREFER_I16 = 32
TB_PLAN_NO_I16 = 64


class tb_options(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_devices", C.c_int32), ("arena_bytes", C.c_int64),
                ("max_wave", C.c_int32), ("host_threads", C.c_int32), ("plan_flags", C.c_uint32),
                ("streams_per_device", C.c_int32), ("devices", C.POINTER(C.c_int32)), ("slice_budget", C.c_int32),
                ("timing", C.c_int32)]


class tb_network(C.Structure):
    _fields_ = [("n_labels", C.c_int32), ("n_leaves", C.c_int32),
                ("leaf_off", C.POINTER(C.c_int32)), ("leaf_labels", C.POINTER(C.c_int32)),
                ("n_open", C.c_int32), ("open_labels", C.POINTER(C.c_int32)),
                ("node_left", C.POINTER(C.c_int32)), ("node_right", C.POINTER(C.c_int32)),
                ("weights", C.c_void_p), ("weight_dtype", C.c_int32), ("value_type", C.c_int32),
                ("flags", C.c_uint32), ("n_fixed", C.c_int32),
                ("fixed_labels", C.POINTER(C.c_int32)), ("fixed_values", C.POINTER(C.c_uint8))]


class tb_plan_stats(C.Structure):
    _fields_ = [("sc", C.c_double), ("tc", C.c_double), ("ops", C.c_double), ("algo_bytes", C.c_double),
                ("arena_elems", C.c_int64), ("n_nodes", C.c_int32), ("n_levels", C.c_int32),
                ("n_fused_subtrees", C.c_int32), ("n_fused_steps", C.c_int32), ("n_gemm_steps", C.c_int32),
                ("n_generic_steps", C.c_int32), ("value_type", C.c_int32), ("root_rank", C.c_int32),
                ("gemm_ops", C.c_double), ("fused_ops", C.c_double), ("generic_ops", C.c_double),
                ("gemm_bytes", C.c_double), ("peak_memory_log2", C.c_double), ("all_memory_log2", C.c_double)]


class tb_step_info(C.Structure):
    _fields_ = [("node", C.c_int32), ("left", C.c_int32), ("right", C.c_int32), ("kind", C.c_int32),
                ("level", C.c_int32), ("rank_a", C.c_int32), ("rank_b", C.c_int32), ("rank_c", C.c_int32),
                ("n_m", C.c_int32), ("n_n", C.c_int32), ("n_b", C.c_int32), ("n_k", C.c_int32),
                ("n_ka", C.c_int32), ("n_kb", C.c_int32), ("tile_m", C.c_int32), ("tile_n", C.c_int32),
                ("c_offset", C.c_int64), ("labels_a", C.c_int32 * 32), ("labels_b", C.c_int32 * 32),
                ("labels_c", C.c_int32 * 32)]


# every symbol include/tbcuda.h declares
EXPORTS = ["tb_version", "tb_init", "tb_init_multi", "tb_device_count", "tb_estimate", "tb_estimate_many", "tb_shutdown", "tb_last_error", "tb_plan_create", "tb_plan_destroy",
           "tb_plan_info", "tb_plan_export", "tb_plan_export_raw", "tb_contract", "tb_contract_batch",
           "tb_contract_networks", "tb_contract_sliced", "tb_suggest_slices", "tb_stream_begin", "tb_stream_push", "tb_stream_finish", "tb_contract_tensor", "tb_contract_table", "tb_compactify_table", "tb_table_configs", "tb_branching_table", "tb_branching_tables", "tb_plan_reassign", "tb_plan_read_tensor", "tb_last_timing", "tb_permute_bits", "tb_set_stream", "tb_profile",
           "tb_last_profile", "tb_last_profile_union", "tb_last_transfers", "tb_last_host_breakdown"]

_lib = None


class TBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libtbcuda error {code}: {msg}")
        self.code = code


def load():
    """Load libtbcuda.so (built by __graft_entry__.build()).  No fallback of any kind."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA engine is the only implementation of this path)")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.tb_version.restype = C.c_char_p
    lib.tb_last_error.restype = C.c_char_p
    lib.tb_last_error.argtypes = [vp]
    lib.tb_init.argtypes = [C.POINTER(tb_options), C.POINTER(vp)]
    lib.tb_init_multi.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.POINTER(tb_options), C.POINTER(vp)]
    lib.tb_device_count.argtypes = [vp]
    lib.tb_estimate.argtypes = [C.POINTER(tb_network), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.tb_estimate_many.argtypes = [C.POINTER(tb_network), C.c_int64, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.tb_shutdown.argtypes = [vp]
    lib.tb_plan_create.argtypes = [vp, C.POINTER(tb_network), C.POINTER(vp)]
    lib.tb_plan_destroy.argtypes = [vp]
    lib.tb_plan_info.argtypes = [vp, C.POINTER(tb_plan_stats)]
    lib.tb_plan_export.argtypes = [vp, C.POINTER(tb_step_info), C.c_int32]
    lib.tb_plan_export_raw.argtypes = [vp, C.c_int32, vp, C.c_int64]
    lib.tb_plan_export_raw.restype = C.c_int64
    lib.tb_contract.argtypes = [vp, vp, C.POINTER(C.c_double)]
    lib.tb_contract_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_double), C.c_int64,
                                      C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    lib.tb_contract_networks.argtypes = [vp, C.POINTER(tb_network), C.POINTER(C.c_double), C.c_int64,
                                         C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    lib.tb_contract_sliced.argtypes = [vp, C.POINTER(tb_network), C.POINTER(C.c_int32), C.c_int32, C.c_int64, C.c_int64,
                                       C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    lib.tb_contract_tensor.argtypes = [vp, vp, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.tb_contract_table.argtypes = [vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_int64, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32)]
    lib.tb_compactify_table.argtypes = [vp, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint8)]
    lib.tb_table_configs.argtypes = [vp, C.POINTER(tb_network), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_uint8),
                                     C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.c_int64,
                                     C.POINTER(C.c_int64)]
    lib.tb_branching_table.argtypes = lib.tb_table_configs.argtypes
    lib.tb_branching_tables.argtypes = [vp, C.POINTER(tb_network), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int64,
                                        C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint32),
                                        C.c_int64, C.POINTER(C.c_int64)]
    lib.tb_plan_reassign.argtypes = [vp, C.POINTER(C.c_uint8), C.POINTER(vp)]
    lib.tb_stream_begin.argtypes = [vp, C.c_int64, C.POINTER(vp)]
    lib.tb_stream_push.argtypes = [vp, C.POINTER(tb_network), C.POINTER(C.c_double), C.c_int64]
    lib.tb_stream_finish.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_int64),
                                     C.POINTER(C.c_double)]
    lib.tb_suggest_slices.argtypes = [vp, C.POINTER(tb_network), C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.tb_plan_read_tensor.argtypes = [vp, vp, C.c_int32, C.POINTER(C.c_double), C.c_int64,
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.tb_last_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.tb_set_stream.argtypes = [vp, vp]
    lib.tb_profile.argtypes = [vp, C.c_int]
    lib.tb_last_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.tb_last_profile_union.argtypes = [vp, C.POINTER(C.c_double)]
    lib.tb_last_host_breakdown.argtypes = [vp, C.POINTER(C.c_double)]
    lib.tb_last_transfers.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.tb_permute_bits.argtypes = [vp, vp, vp, C.c_int32, C.POINTER(C.c_int32)]
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc < 0:
        msg = load().tb_last_error(ctx).decode(errors="replace")
        raise TBError(rc, msg)
    return rc
