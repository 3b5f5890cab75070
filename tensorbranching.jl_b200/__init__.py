"""tensorbranching.jl_b200 -- B200-native engine for the tropical-contraction hot path of
TensorBranching.jl (solve_slice / contract_slices, /root/reference/src/dynamic_ob.jl:30-48).

The directory name is not a valid Python identifier; import it through the `tbcuda` shim at the
repository root (`import tbcuda`).
"""
from ._lib import (TB_VALUE_F64, TB_VALUE_SIZE_CONFIG, TB_PLAN_KEEP_INTERMEDIATES, TB_PLAN_NO_FUSED_SUBTREES, TB_PLAN_NO_GEMM,
                   TB_PLAN_NO_I16, TB_PLAN_NO_SPLIT_K, TB_PLAN_PREFER_I16, TB_PLAN_SCRAMBLE_LAYOUT, TBError, load)
from .contract import (BranchStream, Engine, Plan, PlanBatch, complexity, contraction_all_memory, contraction_peak_memory, contract_slices, default_engine, estimate, estimate_many, sc, solve_slice,
                       solve_slice_index_sliced, suggest_slices, tc)
from .types import CompressedEinsum, MISProblem, SlicedBranch, UnitWeight, add_r, compress

__all__ = ["estimate", "estimate_many", "BranchStream", "PlanBatch", "Engine", "Plan", "complexity", "contract_slices", "default_engine", "sc", "solve_slice", "tc",
           "solve_slice_index_sliced", "suggest_slices", "contraction_peak_memory", "contraction_all_memory",
           "CompressedEinsum", "MISProblem", "SlicedBranch", "UnitWeight", "add_r", "compress", "TBError", "load",
           "TB_VALUE_F64", "TB_VALUE_SIZE_CONFIG", "TB_PLAN_KEEP_INTERMEDIATES", "TB_PLAN_NO_FUSED_SUBTREES", "TB_PLAN_NO_GEMM", "TB_PLAN_SCRAMBLE_LAYOUT", "TB_PLAN_NO_SPLIT_K", "TB_PLAN_PREFER_I16", "TB_PLAN_NO_I16"]
