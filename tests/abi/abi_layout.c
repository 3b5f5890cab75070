/* Prints sizeof / offsetof of every struct that crosses the C ABI of libtbcuda.so, as JSON.  tests/test_abi_layout.py
 * compares the output with the ctypes mirror (tensorbranching.jl_b200/_lib.py) and with the struct definitions parsed out
 * of julia/TBCuda.jl -- the only executable evidence for the Julia binding this image allows (no Julia toolchain). */
#include <stddef.h>
#include <stdio.h>

#include "tbcuda.h"

#define F(S, f) printf("    \"%s\": [%zu, %zu],\n", #f, offsetof(S, f), sizeof(((S*)0)->f))
#define BEGIN(S) printf("  \"%s\": {\n", #S)
#define END(S, last) printf("    \"__sizeof__\": [%zu, 0]\n  }%s\n", sizeof(S), last ? "" : ",")

int main(void) {
    printf("{\n");
    BEGIN(tb_options);
    F(tb_options, device); F(tb_options, n_devices); F(tb_options, arena_bytes); F(tb_options, max_wave);
    F(tb_options, host_threads); F(tb_options, plan_flags); F(tb_options, streams_per_device); F(tb_options, devices);
    F(tb_options, slice_budget); F(tb_options, timing);
    END(tb_options, 0);
    BEGIN(tb_network);
    F(tb_network, n_labels); F(tb_network, n_leaves); F(tb_network, leaf_off); F(tb_network, leaf_labels);
    F(tb_network, n_open); F(tb_network, open_labels); F(tb_network, node_left); F(tb_network, node_right);
    F(tb_network, weights); F(tb_network, weight_dtype); F(tb_network, value_type); F(tb_network, flags);
    F(tb_network, n_fixed); F(tb_network, fixed_labels); F(tb_network, fixed_values);
    END(tb_network, 0);
    BEGIN(tb_plan_stats);
    F(tb_plan_stats, sc); F(tb_plan_stats, tc); F(tb_plan_stats, ops); F(tb_plan_stats, algo_bytes);
    F(tb_plan_stats, arena_elems); F(tb_plan_stats, n_nodes); F(tb_plan_stats, n_levels);
    F(tb_plan_stats, n_fused_subtrees); F(tb_plan_stats, n_fused_steps); F(tb_plan_stats, n_gemm_steps);
    F(tb_plan_stats, n_generic_steps); F(tb_plan_stats, value_type); F(tb_plan_stats, root_rank);
    F(tb_plan_stats, gemm_ops); F(tb_plan_stats, fused_ops); F(tb_plan_stats, generic_ops); F(tb_plan_stats, gemm_bytes);
    F(tb_plan_stats, peak_memory_log2); F(tb_plan_stats, all_memory_log2);
    END(tb_plan_stats, 0);
    BEGIN(tb_step_info);
    F(tb_step_info, node); F(tb_step_info, left); F(tb_step_info, right); F(tb_step_info, kind); F(tb_step_info, level);
    F(tb_step_info, rank_a); F(tb_step_info, rank_b); F(tb_step_info, rank_c); F(tb_step_info, n_m); F(tb_step_info, n_n);
    F(tb_step_info, n_b); F(tb_step_info, n_k); F(tb_step_info, n_ka); F(tb_step_info, n_kb); F(tb_step_info, tile_m);
    F(tb_step_info, tile_n); F(tb_step_info, c_offset); F(tb_step_info, labels_a); F(tb_step_info, labels_b);
    F(tb_step_info, labels_c);
    END(tb_step_info, 1);
    printf("}\n");
    return 0;
}
