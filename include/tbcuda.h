/*
 * tbcuda.h -- C ABI of libtbcuda.so, the B200 (sm_100a) engine behind the tropical-contraction hot
 * path of TensorBranching.jl.
 *
 * The library stands where the reference's
 *     solve_slice(branch, element_type, usecuda)        /root/reference/src/dynamic_ob.jl:30-34
 *     contract_slices(branches, element_type, usecuda)  /root/reference/src/dynamic_ob.jl:36-48
 * stand today (callers: dynamic_ob_mis src/dynamic_ob.jl:24, slice_dfs_lp src/slice.jl:39, user code
 * after load_all_finished src/io.jl:113-121).  The reference has no FFI of its own for this path (it
 * calls GenericTensorNetworks.solve in-process); INTEGRATION.md shows the Julia `ccall` stubs a
 * maintainer would add.  Every entry point below names the reference behaviour it replaces.
 *
 * Conventions
 *   - plain C, no exceptions cross the ABI; every function returning int returns a tb_status
 *     (0 = ok, negative = error); the message is available from tb_last_error().
 *   - all ids are 0-based.  Inputs are caller-owned and only read during the call (they are copied
 *     by tb_plan_create).  Handles are library-owned and freed only by the matching destroy call.
 *     Outputs go to caller-allocated buffers.  Device memory never crosses the ABI.
 *   - one tb_ctx is used from one host thread at a time; calls block until results are on the host.
 *     Distinct contexts are independent.  The library installs no signal handlers; its worker threads end with
 *     the call that spawned them, except one short-lived helper that frees the host side of a call's temporary
 *     plans, which is joined by the next call or by tb_shutdown (no thread outlives tb_shutdown).
 *   - a tensor "layout" is a list of labels in address-bit order, bit 0 (fastest) first -- i.e.
 *     Julia's column-major dimension order.  Every label has size 2 (uniformsize(code, 2),
 *     /root/reference/src/types.jl:118).
 */
#ifndef TBCUDA_H
#define TBCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tb_ctx tb_ctx;
typedef struct tb_plan tb_plan;

typedef enum tb_status {
    TB_OK = 0,
    TB_ERR_BAD_ARGUMENT = -1,
    TB_ERR_NOT_BINARY_TREE = -2,   /* tree arrays do not describe a binary tree over the leaves */
    TB_ERR_UNSUPPORTED = -3,       /* leaf rank > 2, tensor rank > 31, value range too large ... */
    TB_ERR_OUT_OF_MEMORY = -4,
    TB_ERR_CUDA = -5,              /* no usable device / a CUDA call failed; there is NO CPU fallback */
    TB_ERR_NCCL = -6,
    TB_ERR_INTERNAL = -7
} tb_status;

/* value types of the tropical numbers (element_type of the reference, src/dynamic_ob.jl:6) */
typedef enum tb_value_type {
    TB_VALUE_AUTO = 0,   /* integer weights: packed int16 if sum |w| < 8192 (unless TB_PLAN_NO_I16), else int32; real weights: f32 */
    TB_VALUE_I32 = 1,    /* exact; -inf is the sentinel -2^30 */
    TB_VALUE_F32 = 2,    /* Tropical{Float32}; -inf is IEEE -inf */
    TB_VALUE_I16X2 = 3,  /* int16 values, packed pairs in the GEMM (VIADDMNMX.S16x2); -inf is -2^14; needs sum |w| < 2^13 */
    TB_VALUE_F64 = 4,    /* Tropical{Float64}: element_type = Float64 with real weights.  Never picked by AUTO; generic +
                            fused kernels only (no tiled GEMM kernel for 8-byte values) */
    TB_VALUE_SIZE_CONFIG = 5 /* size + ONE optimal configuration per element (the SingleConfigMax element type of the
                            branching tables, src/branch.jl:79): int64 = size << 32 | vertex mask.  Needs n_labels <= 32 and
                            integer weights; read back with tb_contract_table */
} tb_value_type;

typedef enum tb_weight_dtype {
    TB_WEIGHT_UNIT = 0,  /* UnitWeight: weights pointer ignored, every vertex weighs 1 */
    TB_WEIGHT_I32 = 1,
    TB_WEIGHT_I64 = 2,
    TB_WEIGHT_F32 = 3,
    TB_WEIGHT_F64 = 4
} tb_weight_dtype;

/* plan flags */
#define TB_PLAN_KEEP_INTERMEDIATES 1u /* every node gets its own arena region (no reuse) so that
                                         tb_plan_read_tensor can return it after tb_contract */
#define TB_PLAN_NO_FUSED_SUBTREES 2u  /* testing: run every node as its own step */
#define TB_PLAN_NO_GEMM 4u            /* testing: never choose the tiled max-plus GEMM kernel */
#define TB_PLAN_SCRAMBLE_LAYOUT 8u    /* testing: pseudo-random (valid) operand layouts */
#define TB_PLAN_NO_SPLIT_K 16u        /* testing: never split a long reduction into partial + reduce steps */
#define TB_PLAN_PREFER_I16 32u        /* (default behaviour, kept for explicitness) value_type AUTO picks packed int16
                                         (TB_VALUE_I16X2) when the weights are integers with sum |w| < 8192: identical
                                         results, 2x DPX rate, half the bytes */
#define TB_PLAN_NO_I16 64u            /* value_type AUTO never picks packed int16 (int32 for integer weights) */

/* The hot path's only knobs in the reference are element_type and usecuda (src/dynamic_ob.jl:6,30,36); everything
 * else an engine needs is here ("Config / flags"): no environment-variable magic is required. */
typedef struct tb_options {
    int32_t device;          /* CUDA device ordinal (single-device context) */
    int32_t n_devices;       /* 0 / 1: one device (`device`); > 1: tb_init returns a multi-GPU context over devices[0 .. n_devices)
                                (exactly tb_init_multi) */
    int64_t arena_bytes;     /* HBM arena for intermediates, per device; 0 = 60 % of free memory at first use */
    int32_t max_wave;        /* max branches contracted concurrently in one wave (0 = by plan weight: 128 / 256) */
    int32_t host_threads;    /* plan-compiler threads for the *_networks calls, all devices together (0 = all cores) */
    uint32_t plan_flags;     /* default flags OR-ed into every plan */
    int32_t streams_per_device; /* stream lanes = waves in flight per GPU, 1 .. 8 (0 = default 4) */
    const int32_t* devices;  /* n_devices > 1: the device ordinals (NULL = 0 .. n_devices-1); read during tb_init only */
    int32_t slice_budget;    /* index-slicing budget of a multi-GPU context: when a call has fewer branches than 2 x devices,
                                every branch is cut into up to 2^slice_budget index slices (tb_suggest_slices picks the labels)
                                so that all devices get work; 0 = never slice on the engine's own initiative */
    int32_t timing;          /* tb_profile mode from the start: 0 off, 1 per-launch events (lanes concurrent), 2 single lane */
} tb_options;

/*
 * One branch network = what a SlicedBranch carries (src/types.jl:85-103): the leaf label lists
 * `ixs` and the binary ContractionTree of its CompressedEinsum (src/types.jl:51-58), plus weights.
 *   leaf i has labels leaf_labels[leaf_off[i] .. leaf_off[i+1]); 1 label = vertex tensor [0, w_v],
 *   2 labels = edge tensor [[0,0],[0,-inf]] (generate_tensors of IndependentSet [upstream]).
 *   internal node j (tensor id n_leaves + j) contracts tensor ids node_left[j], node_right[j];
 *   children always have smaller ids (post-order); the last node is the root.  n_leaves - 1 nodes.
 *   open_labels (iy) is empty on the solve_slice path; if given, the root keeps those labels.
 *   weights[v] is the weight of label/vertex v (n_labels entries) unless weight_dtype == UNIT.
 *   Index slicing: fixing label v to x restricts every leaf that carries v to its x-th slice (the vertex
 *   tensor [0, w_v] becomes the scalar 0 or w_v, the edge tensor a row of it) and v disappears from
 *   the network, so every tensor that carried it halves.  The 2^k assignments of k fixed labels are
 *   independent contractions of the same tree; the max of their values is the unsliced value.
 *   tb_contract_sliced enumerates them, tb_suggest_slices picks the labels.
 */
typedef struct tb_network {
    int32_t n_labels;
    int32_t n_leaves;
    const int32_t* leaf_off;
    const int32_t* leaf_labels;
    int32_t n_open;
    const int32_t* open_labels;
    const int32_t* node_left;
    const int32_t* node_right;
    const void* weights;
    int32_t weight_dtype;    /* tb_weight_dtype */
    int32_t value_type;      /* tb_value_type */
    uint32_t flags;          /* TB_PLAN_* */
    int32_t n_fixed;         /* index slicing: labels fixed to one value in this contraction; 0 = none */
    const int32_t* fixed_labels; /* n_fixed distinct labels ... */
    const uint8_t* fixed_values; /* ... and the value (0 or 1) each one is fixed to */
} tb_network;

/* what complexity(branch) reports in the reference (src/types.jl:115-121) + engine facts */
typedef struct tb_plan_stats {
    double sc;               /* max tensor rank */
    double tc;               /* log2(sum over nodes of 2^(labels involved)) */
    double ops;              /* tropical ops = sum over nodes 2^(m+n+k+b) */
    double algo_bytes;       /* sum over nodes elem_size*(2^rank A + 2^rank B + 2^rank C) */
    int64_t arena_elems;     /* peak arena footprint in elements */
    int32_t n_nodes;
    int32_t n_levels;        /* dependency levels of the non-fused nodes */
    int32_t n_fused_subtrees;
    int32_t n_fused_steps;
    int32_t n_gemm_steps;
    int32_t n_generic_steps;
    int32_t value_type;      /* resolved tb_value_type */
    int32_t root_rank;
    double gemm_ops;         /* ops carried by the tiled GEMM kernel */
    double fused_ops;
    double generic_ops;
    double gemm_bytes;       /* algorithmic bytes (operands read once + result written once) of the GEMM steps */
    double peak_memory_log2; /* contraction_peak_memory(code, size 2) of the reference (src/utils.jl:197-219), in elements */
    double all_memory_log2;  /* contraction_all_memory (src/utils.jl:222-229): log2 of the sum of all intermediates */
} tb_plan_stats;

/* one exported step (for inspection / the oracle-side plan checker / estimators) */
typedef struct tb_step_info {
    int32_t node;            /* tensor id produced */
    int32_t left, right;     /* operand tensor ids */
    int32_t kind;            /* 0 fused-subtree step, 1 generic, 2 gemm */
    int32_t level;
    int32_t rank_a, rank_b, rank_c;
    int32_t n_m, n_n, n_b, n_k, n_ka, n_kb;
    int32_t tile_m, tile_n;
    int64_t c_offset;        /* arena / smem element offset */
    int32_t labels_a[32];    /* layouts, bit 0 first */
    int32_t labels_b[32];
    int32_t labels_c[32];
} tb_step_info;

const char* tb_version(void);

/* lifetime of the engine on one device.  Fails with TB_ERR_CUDA when no device is usable. */
int tb_init(const tb_options* opts, tb_ctx** out_ctx);
int tb_shutdown(tb_ctx* ctx);

/* Multi-GPU engine in ONE process (no Distributed.jl / MPI needed on the host side): one sub-context per
 * device, each with its own arena, stream lanes and share of the plan-compiler threads.  On such a context
 *   tb_contract_networks / tb_contract_batch  deal the branches to the devices longest-first (LPT) by their tropical
 *       ops (for resident plans: a plan stays on the device it first ran on), every device contracts its share, and ONE
 *       ncclAllReduce(ncclMax) over the per-branch result vector -- pre-filled with -inf on every device -- replaces
 *       maximum(res) (src/dynamic_ob.jl:27); no tensor crosses NVLink;
 *   tb_contract_sliced  deals the 2^k assignments of one heavy branch in contiguous ranges;
 *   tb_contract / tb_contract_tensor / tb_plan_read_tensor / tb_stream_*  run on the first device.
 * devices == NULL means 0 .. n_devices-1.  NCCL (libnccl.so.2) is loaded with dlopen here: TB_ERR_NCCL if it is missing
 * or a NCCL call fails; a single-device context never needs it.  A device may be listed more than once (exercising
 * the sharding on a single-GPU box): NCCL cannot form such a communicator, so the per-device vectors are then combined
 * on the host instead -- a testing configuration, not a product path.  opts->device / n_devices / devices are ignored. */
int tb_init_multi(const int32_t* devices, int32_t n_devices, const tb_options* opts, tb_ctx** out_ctx);
/* number of devices of a context (1 for tb_init contexts) */
int tb_device_count(const tb_ctx* ctx);

/* the cost a sharder needs, without compiling: tropical ops (the reference's 2^tc, src/types.jl:120) and sc (:121) from
 * the label-set pass alone, ~4x cheaper than tb_plan_create.  Host-only.  out_sc may be NULL. */
int tb_estimate(const tb_network* net, double* out_ops, double* out_sc);
/* the same for a whole branch list on `threads` host threads (0 = all cores); out_sc may be NULL; n_leaves == 0 gives 0 */
int tb_estimate_many(const tb_network* nets, int64_t n, int32_t threads, double* out_ops, double* out_sc);
const char* tb_last_error(const tb_ctx* ctx); /* ctx may be NULL: last error of the calling thread */

/* replaces uncompress(branch.code) + GenericTensorNetwork(...) (src/dynamic_ob.jl:31, src/types.jl:75-79):
 * compiles the tree into a device-resident step list.  ctx may be NULL (host-only compile; the plan is
 * uploaded on first use). */
int tb_plan_create(tb_ctx* ctx, const tb_network* net, tb_plan** out_plan);
int tb_plan_destroy(tb_plan* plan);
int tb_plan_info(const tb_plan* plan, tb_plan_stats* out);
/* exports up to `cap` steps in execution order; returns the number of steps (or a negative status) */
int tb_plan_export(const tb_plan* plan, tb_step_info* out, int32_t cap);

/* raw descriptor sections of a compiled plan, for the test-side descriptor interpreter and for
 * debugging: which = 0 leaf pool (u32 words), 1 fused steps (48-byte records), 2 fused subtrees
 * (24-byte records), 3 level steps (144-byte records), 4 level index (i32), 5 header (i64 words:
 * arena_elems, root_off, n_levels, value_type), 6 log2 of the tropical ops of every level step (f32), 7 algorithmic bytes
 * of every level step (f64: operands read once + result written once) -- 6 and 7 are what a per-node roofline
 * max(ops / op rate, bytes / bandwidth) needs.  Copies min(cap, size) bytes, returns the size. */
int64_t tb_plan_export_raw(const tb_plan* plan, int32_t which, void* out, int64_t cap);

/* replaces solve(net, SizeMax(), T, usecuda)[].n (src/dynamic_ob.jl:32): the root scalar of one plan */
int tb_contract(tb_ctx* ctx, tb_plan* plan, double* out_value);

/* replaces the loop of contract_slices (src/dynamic_ob.jl:38-46) and maximum(res) (:27).
 * plans[i] == NULL means "empty graph": the value is r[i] (src/dynamic_ob.jl:39-40).
 * r may be NULL (all zero).  out_values[i] = value_i + r[i] in double; out_status may be NULL;
 * out_max may be NULL.  A failed branch gets a negative status and NaN, never a silent wrong max. */
int tb_contract_batch(tb_ctx* ctx, tb_plan* const* plans, const double* r, int64_t n,
                      double* out_values, int32_t* out_status, double* out_max);

/* the whole of contract_slices in one call: compile (multi-threaded), upload, contract, discard.
 * nets[i].n_leaves == 0 means "empty graph". */
int tb_contract_networks(tb_ctx* ctx, const tb_network* nets, const double* r, int64_t n,
                         double* out_values, int32_t* out_status, double* out_max);

/* Streaming hand-off: the host's slicer decides finished vs unfinished per round
 * (src/slice.jl:79-86) and hands finished branches over while it keeps slicing the rest.  tb_stream_push compiles,
 * uploads and ENQUEUES the branches and returns while the GPU contracts them; tb_stream_finish waits and returns one
 * value per pushed branch, in push order (value + r, as tb_contract_networks), and their max.  `capacity` bounds the
 * total number of branches of the stream.  The context cannot run other contract calls while a stream is open.
 * tb_stream_finish always closes the stream, also after a failed push. */
typedef struct tb_stream tb_stream;
int tb_stream_begin(tb_ctx* ctx, int64_t capacity, tb_stream** out_stream);
int tb_stream_push(tb_stream* stream, const tb_network* nets, const double* r, int64_t n);
int tb_stream_finish(tb_stream* stream, double* out_values, int32_t* out_status, int64_t cap, int64_t* out_n,
                     double* out_max);

/* Index slicing of ONE heavy branch (the 2^k slice assignments of BASELINE.json north_star (4)).
 * The network's own n_fixed must be 0.  Assignment a in [first, first + count) fixes sliced_labels[i] to
 * bit i of a; every assignment is a contraction of the same tree with n_sliced labels removed, all of
 * them run as one batch (tb_contract_networks).  out_values[a - first] = value of assignment a + r
 * (-inf when the assignment selects two adjacent vertices: such slices are not contracted at all);
 * out_max = max over the range.  Ranks of a multi-GPU job call this with disjoint ranges and combine with
 * one all-reduce(max).  out_values and out_status may be NULL. */
int tb_contract_sliced(tb_ctx* ctx, const tb_network* net, const int32_t* sliced_labels, int32_t n_sliced,
                       int64_t first, int64_t count, double r, double* out_values, int32_t* out_status,
                       double* out_max);

/* A plan created with fixed labels, re-targeted to another assignment of the SAME labels (fixed_values[i] for
 * fixed_labels[i] of the network it was created from): the 2^k assignments share every descriptor and layout and
 * differ only in a few leaf-pool words, so this is a copy plus a patch, ~30x cheaper than tb_plan_create.
 * tb_contract_sliced uses it internally (one compilation per call). */
int tb_plan_reassign(const tb_plan* base, const uint8_t* fixed_values, tb_plan** out_plan);

/* Pick up to max_sliced labels to slice, greedily.  sc_target >= 0 (memory-driven): until the largest tensor has
 * rank <= sc_target, each pick is the label held by most tensors of the current top rank (ties: most ops removed).
 * sc_target < 0 (parallelism-driven): always max_sliced labels, each pick the label whose removal takes away most
 * ops, i.e. the least total overhead of the 2^k slices.  Host-only (no device work; ctx may be NULL).  Returns the number
 * of labels written to out_labels (<= max_sliced) or a negative tb_status; out_sc / out_tc (may be NULL)
 * receive the per-slice complexity after slicing. */
int tb_suggest_slices(tb_ctx* ctx, const tb_network* net, int32_t sc_target, int32_t max_sliced,
                      int32_t* out_labels, double* out_sc, double* out_tc);

/* Open-boundary contraction (first half): a network created with open_labels (iy non-empty) keeps
 * those labels at the root -- for a region of the graph with its boundary vertices open this root is the tensor of
 * best sizes per boundary configuration that the branching tables of TensorNetworkSolver are read from
 * (src/branch.jl:79, src/types.jl:46; the configurations themselves: tb_contract_table, tb_table_configs,
 * tb_branching_table below).  Contracts the plan and returns the whole root tensor: 2^rank doubles (-inf = tropical zero),
 * layout in out_labels, bit 0 first.  out_data == NULL only queries rank / labels. */
int tb_contract_tensor(tb_ctx* ctx, tb_plan* plan, double* out_data, int64_t cap, int32_t* out_labels,
                       int32_t* out_rank);

/* The configuration-enumerating half of the branching table (branching_table(p, TensorNetworkSolver(),
 * region), src/branch.jl:79 [upstream OptimalBranchingMIS]): for a region's network created with value_type
 * TB_VALUE_SIZE_CONFIG and its boundary vertices as open_labels, contracts the plan and returns, for every one of the
 * 2^rank boundary configurations (layout in out_labels, bit 0 first): out_sizes = the best size of an independent set of
 * the region compatible with it (-inf if none) and out_configs = ONE such optimal set as a vertex bit mask (bit v =
 * vertex / label v, boundary vertices included).  Ties are broken towards the larger mask, so the answer is deterministic.
 * out_sizes == NULL only queries rank / labels. */
int tb_contract_table(tb_ctx* ctx, tb_plan* plan, double* out_sizes, uint32_t* out_configs, int64_t cap, int32_t* out_labels,
                      int32_t* out_rank);

/* The reduction step of the same table solver ([upstream, recalled] GenericTensorNetworks mis_compactify!, applied by
 * OptimalBranchingMIS reduced_alpha_configs before the table is built; reached from src/branch.jl:79): boundary
 * configuration a is dominated -- and its row dropped from the branching table -- when some other configuration b chooses
 * a subset of a's boundary vertices (b & a == b, b != a) and sizes[b] >= sizes[a].  sizes = the 2^rank sizes returned by
 * tb_contract_table / tb_contract_tensor (index bit i = boundary vertex out_labels[i]; -inf = infeasible);
 * out_keep[a] = 1 for the rows that survive, 0 for dominated or infeasible ones.  Runs on the device as a subset-max
 * transform (rank + 1 launches, O(rank 2^rank)) instead of the reference's all-pairs loop (O(4^rank)). */
int tb_compactify_table(tb_ctx* ctx, int32_t rank, const double* sizes, uint8_t* out_keep);

/* ALL optimal configurations of the table ([upstream, recalled] the ConfigsMax element type of GenericTensorNetworks that
 * OptimalBranchingMIS' table solver contracts with; every row of the BranchingTable built at src/branch.jl:79 lists all
 * optimal vertex sets of its boundary configuration, and the set-cover solver may satisfy a row with any one of them).
 * net = the region's network (its leaves give the graph and the weights; the tree is not used, n_labels <= 32),
 * boundary_labels[0 .. rank) = the open vertices in the bit order of the rows (pass tb_contract_table's out_labels, so that
 * row a here is entry a of its sizes), keep = tb_compactify_table's flags (NULL = every feasible row).
 *   out_sizes[a]   (optional) the optimum of row a, computed independently of the contraction -- equal to
 *                  tb_contract_tensor's sizes (the tests assert it); -inf = infeasible
 *   out_row_off    2^rank + 1 offsets: row a owns out_configs[out_row_off[a] .. out_row_off[a+1]) (empty if dropped)
 *   out_configs    the vertex masks (bit v = vertex v, boundary vertices included), rows in order, each row ascending in
 *                  its interior configuration; NULL (or cap == 0) only counts; cap < total is TB_ERR_BAD_ARGUMENT
 *   out_total      number of configurations in the table
 * A region has <= n_max = 20 vertices by default (src/types.jl:10), so instead of carrying configuration SETS through the
 * contraction the device filters the 2^n vertex sets of the region against the row optima (one CTA per boundary
 * configuration and chunk of interior configurations; three passes: optimum, count, ordered write). */
int tb_table_configs(tb_ctx* ctx, const tb_network* net, const int32_t* boundary_labels, int32_t rank, const uint8_t* keep,
                     double* out_sizes, int64_t* out_row_off, uint32_t* out_configs, int64_t cap, int64_t* out_total);

/* The whole of branching_table(p, TensorNetworkSolver(), region) (src/branch.jl:79; [upstream, recalled] OptimalBranchingMIS
 * reduced_alpha_configs: SizeMax tensor -> mis_compactify! -> ConfigsMax rows of the surviving entries) in ONE call and one
 * round trip: row optima of the region, dominated rows dropped on the device (out_keep, optional, 2^rank flags as
 * tb_compactify_table returns them), all optimal configurations of the surviving rows.  Arguments as tb_table_configs;
 * boundary_labels = the region's open vertices in any order (they define the bit order of the rows).  When cap < total the
 * call fails with TB_ERR_BAD_ARGUMENT after writing out_total / out_row_off / out_sizes / out_keep: retry with a larger buffer
 * (tables have tens of rows, so a first guess of a few thousand entries practically always holds). */
int tb_branching_table(tb_ctx* ctx, const tb_network* net, const int32_t* boundary_labels, int32_t rank, uint8_t* out_keep,
                       double* out_sizes, int64_t* out_row_off, uint32_t* out_configs, int64_t cap, int64_t* out_total);

/* The tables of MANY regions in the same launches (the reducer and the slicer of the reference ask for the table of one
 * small region per vertex / per branching step: TensorNetworkReducer src/dynamic_ob.jl:6, table solver src/types.jl:46 --
 * independent requests, so a host that has several at hand sends them together and pays the launch latency once).
 * Region i = nets[i] with the open vertices boundary_labels[boundary_off[i] .. boundary_off[i+1]); its rows are
 * [R_i, R_i + 2^rank_i) with R_i = the sum of 2^rank_j over j < i; out_keep / out_sizes have R_n entries, out_row_off R_n + 1
 * (one CSR over the rows of all regions); everything else as tb_branching_table. */
int tb_branching_tables(tb_ctx* ctx, const tb_network* nets, const int32_t* boundary_off, const int32_t* boundary_labels, int64_t n,
                        uint8_t* out_keep, double* out_sizes, int64_t* out_row_off, uint32_t* out_configs, int64_t cap,
                        int64_t* out_total);

/* after tb_contract on a TB_PLAN_KEEP_INTERMEDIATES plan: copy tensor `node` (any internal node id,
 * or the root) to the host as doubles (-inf for tropical zero), 2^rank elements, and its layout. */
int tb_plan_read_tensor(tb_ctx* ctx, tb_plan* plan, int32_t node, double* out_data, int64_t cap,
                        int32_t* out_labels, int32_t* out_rank);

/* timing of the last tb_contract / tb_contract_batch on this ctx, measured with CUDA events on the
 * engine's stream: total device ms and number of kernel launches. */
int tb_last_timing(const tb_ctx* ctx, double* out_device_ms, int64_t* out_launches);

/* run the engine on a caller-owned CUDA stream (cudaStream_t), e.g. the harness's current stream, so
 * that the caller's events bracket the engine's work.  The stream is not destroyed by tb_shutdown. */
int tb_set_stream(tb_ctx* ctx, void* cuda_stream);

/* opt-in per-launch CUDA-event timing, accumulated by kernel kind over the last contract call:
 * 0 fused subtrees, 1 generic (level-synchronous launches only), 2 the persistent max-plus GEMM kernel (in the default
 * dataflow executor this ONE launch per wave also runs the wave's generic steps on its consumer warps), 3 finalize.
 * mode 0 = off; 1 = events around every launch on its own stream lane, lanes stay concurrent (the timed configuration);
 * 2 = the same on a single lane (launches serialised: per-launch durations free of overlap).
 * tb_last_profile: per kind the SUM of the launch durations; tb_last_profile_union: per kind the length of the union
 * of the launches' [start, end] intervals, i.e. how long that kind was on the device (equal to the sum in mode 2). */
int tb_profile(tb_ctx* ctx, int mode);
int tb_last_profile(const tb_ctx* ctx, double* ms_by_kind /*[4]*/, int64_t* launches_by_kind /*[4]*/);
int tb_last_profile_union(const tb_ctx* ctx, double* ms_by_kind /*[4]*/);

/* host wall-clock breakdown (ms) of the last tb_contract_networks call:
 * [0] plan compilation, [1] descriptor upload, [2] work-list build, [3] launch + wait, [4] plan teardown, [5] total */
int tb_last_host_breakdown(const tb_ctx* ctx, double* ms6);

/* host<->device bytes moved by the last contract call (descriptors + work lists up, results down) */
int tb_last_transfers(const tb_ctx* ctx, int64_t* h2d_bytes, int64_t* d2h_bytes);

/* K5: out[dst] = in[src] where bit i of dst is bit perm[i] of src (2^rank elements of 4 bytes),
 * host buffers; exercises the bit-permutation transpose kernel (OMEinsum permutedims [upstream]). */
int tb_permute_bits(tb_ctx* ctx, const void* in, void* out, int32_t rank, const int32_t* perm);

#ifdef __cplusplus
}
#endif
#endif /* TBCUDA_H */
