// plan.hpp -- host-side plan compiler: contraction tree -> flat, layout-resolved step list.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/tbcuda.h"
#include "desc.h"

struct tb_ctx;

namespace tb {

struct Plan {
    // resolved inputs
    int n_labels = 0;
    int n_leaves = 0;  // including the synthetic unit leaf of a single-leaf network
    int n_nodes = 0;
    uint32_t flags = 0;
    bool temporary = false;  // compiled with TB_PLAN_TEMPORARY (a temporary of the *_networks / stream / sliced calls: contracted once)
    int value_type = TB_VALUE_I32;

    // per tensor id (leaves, then nodes, then synthetic split-K tensors)
    std::vector<int32_t> lay_off;              // layout of tensor t = lay_data[lay_off[t] .. +lay_n[t])
    std::vector<uint8_t> lay_n;                //   labels in address-bit order (bit 0 first)
    std::vector<int32_t> lay_data;
    int rank(int t) const { return lay_n[t]; }
    const int32_t* layout(int t) const { return lay_data.data() + lay_off[t]; }
    std::vector<int8_t> loc;                   // LOC_*
    std::vector<int64_t> off;                  // element offset within loc
    std::vector<int32_t> level;                // -1 for leaves / interior of fused subtrees

    // descriptors
    std::vector<uint64_t> pool;  // one slot per pool element (int32 / float bits in the low half, double / int64 bits in full);
                                 // narrowed to the value type's width on upload
    // index slicing: the pool words that depend on the values of the fixed labels.  A plan compiled for one assignment
    // becomes the plan of another by assign() alone (same descriptors, same layouts).
    struct PoolPatch {
        int32_t off;      // pool slot of the sliced leaf
        uint8_t kind;     // 0 vertex leaf, 1 edge leaf with one end fixed, 2 edge leaf with both ends fixed
        int32_t fa, fb;   // positions of the leaf's fixed labels in the fixed-label list (fb = -1: none)
        double w;         // vertex weight (kind 0)
        int32_t vtx;      // the vertex (kind 0): its bit of the configuration mask (TB_VALUE_SIZE_CONFIG)
    };
    std::vector<PoolPatch> patches;
    int n_fixed = 0;
    void assign(const uint8_t* values);  // values[i] = 0 / 1 for fixed label i
    static uint64_t encode_value(int value_type, double x, bool neg_inf, int config_bit = -1);
    static int elem_size_of(int vt) { return vt == TB_VALUE_I16X2 ? 2 : (vt == TB_VALUE_F64 || vt == TB_VALUE_SIZE_CONFIG) ? 8 : 4; }
    int elem_size() const { return elem_size_of(value_type); }
    size_t pool_bytes() const { return pool.size() * (size_t)elem_size(); }
    void write_pool(uint8_t* dst) const;  // device representation of the pool
    std::vector<SubStep> sub_steps;
    std::vector<SubTree> subtrees;
    std::vector<BigStep> big_steps;        // sorted by level (levels start at 1)
    std::vector<int32_t> big_level_begin;  // big_steps index where level L starts, size n_levels + 2
    std::vector<float> big_log2_ops;  // per big step: log2 of its tropical ops (per-launch roofline records)
    std::vector<double> big_bytes;    // per big step: bytes it moves (operands read once + result written once)
    std::vector<int32_t> big_dep_a, big_dep_b;  // per big step: index of the big step producing operand A / B, -1 = leaf pool
                                                // or fused subtree (complete before any big step starts); always < own index
    int n_levels = 0;                      // number of levels with big steps (levels 1..n_levels)
    int64_t arena_elems = 0;
    int64_t root_off = 0;
    int32_t root_id = 0;

    tb_plan_stats stats{};
    // execution order (fused steps first, then big steps by level); tb_step_info is built on demand
    struct StepRec {
        int32_t node, left, right;
        int8_t kind;
        int32_t level;
        uint8_t nm, nn, nb, nk, nka, nkb, tm, tn;
    };
    std::vector<StepRec> recs;
    int n_tensors = 0;  // leaves + nodes + synthetic (split-K) tensors
    tb_step_info step_info(size_t i) const;

    // back to the state of a new plan, keeping the capacity of every array: compiling into a recycled plan allocates nothing
    void recycle();
    // the part of a compiled plan the executor needs -- descriptors, launch geometry, statistics -- copied into `dst` with
    // exact-size arrays; layouts, the tensor table and the step records stay behind (temporaries of contract_slices)
    void copy_descriptors_to(Plan& dst) const;
    // host memory held by the arrays copy_descriptors_to fills (capacities, so a recycled plan reports what it pins)
    size_t descriptor_capacity_bytes() const {
        return pool.capacity() * sizeof(uint64_t) + patches.capacity() * sizeof(PoolPatch) + sub_steps.capacity() * sizeof(SubStep) +
               subtrees.capacity() * sizeof(SubTree) + big_steps.capacity() * sizeof(BigStep) +
               big_level_begin.capacity() * sizeof(int32_t) + big_log2_ops.capacity() * sizeof(float) +
               big_bytes.capacity() * sizeof(double) + (big_dep_a.capacity() + big_dep_b.capacity()) * sizeof(int32_t) + sizeof(Plan);
    }

    // device residency (managed by the engine)
    tb_ctx* owner = nullptr;
    Plan* res_prev = nullptr;  // intrusive list of the plans resident on `owner` (tb_shutdown detaches them, so a plan
    Plan* res_next = nullptr;  // may safely outlive its context)
    void* d_blob = nullptr;
    uint64_t upload_uid = 0;  // unique per upload of the descriptors (0 = not resident): what cached work lists are keyed on
    size_t blob_bytes = 0;
    size_t sub_blob_off = 0, big_blob_off = 0;
};

// internal flag (never part of the ABI): the plan is a temporary of tb_contract_networks / tb_stream_push -- nobody will
// ask for its step records or the reference's memory estimators, so they are not computed
constexpr uint32_t TB_PLAN_TEMPORARY = 1u << 31;
// internal flag: stop after the label-set pass; only stats.ops / stats.tc / stats.sc are filled (tb_estimate: the cost a
// sharder needs, at ~1/4 of the price of a full compilation)
constexpr uint32_t TB_PLAN_ESTIMATE_ONLY = 1u << 30;

// returns a tb_status; on failure `err` holds the message
int compile_plan(const tb_network& net, uint32_t extra_flags, Plan& plan, std::string& err);

// greedy label choice for index slicing (tb_suggest_slices); returns the number of labels picked or a tb_status
int suggest_slices(const tb_network& net, int sc_target, int max_sliced, int32_t* out_labels, double* out_sc,
                   double* out_tc, std::string& err);

// serialise pool | SubStep[] | BigStep[] (16-byte aligned sections) for upload
void build_blob(Plan& plan, std::vector<uint8_t>& blob);

}  // namespace tb

// the opaque C handle
struct tb_plan {
    tb::Plan p;
};
