// engine.cu -- tb_ctx, plan residency, the wave executor and the C ABI of libtbcuda.so.
//
// Execution model (replaces the serial `for branch in branches` loop of contract_slices,
// /root/reference/src/dynamic_ob.jl:38-46, and OMEinsum's recursive, allocating executor):
//   * branches are grouped into WAVES that fit the HBM arena together;
//   * inside a wave, steps of ALL branches that sit on the same dependency level run in ONE launch
//     (work lists of (branch, step) instances; a CTA finds its instance by binary search);
//     level 0 = all fused small subtrees, levels >= 1 = generic + tiled-GEMM steps;
//   * root scalars are gathered by k_finalize into one result vector, copied back once.
// There is no CPU fallback: without a CUDA device tb_init fails with TB_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_set>
#include <vector>

#include "kernels.cuh"
#include "plan.hpp"

using namespace tb;

namespace {
thread_local std::string g_tls_error;
inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct BlobChunk {
    void* d = nullptr;
    size_t cap = 0, used = 0;
    int live = 0;
};
}  // namespace

// A loop the launching thread shares with the call's compile workers: while a call is being compiled, the workers look
// here between two plans and take chunks of the loop (serialising descriptor blobs into pinned memory is the launching
// thread's largest serial cost, profiles/x3_e2e_cfg2_*.json).  Everything is under one mutex: a few dozen chunk
// hand-outs per batch.
struct HelperJob {
    std::mutex mu;
    std::function<void(int64_t, int64_t)> fn;  // [begin, end)
    int64_t n = 0, next = 0, grain = 1, pending = 0;  // pending = chunks handed out and not finished yet
    bool active = false;
    bool has_helpers = false;  // set by the call that runs worker threads
    bool help_once() {
        int64_t b, e;
        {
            std::lock_guard<std::mutex> lk(mu);
            if (!active || next >= n) return false;
            b = next;
            e = std::min(n, b + grain);
            next = e;
            ++pending;
        }
        fn(b, e);
        {
            std::lock_guard<std::mutex> lk(mu);
            --pending;
        }
        return true;
    }
    // owner: runs fn over [0, count) with whoever helps, returns when every chunk is finished
    void run(int64_t count, int64_t chunk, std::function<void(int64_t, int64_t)> f) {
        if (!has_helpers || count <= chunk) {
            f(0, count);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            fn = std::move(f);
            n = count;
            next = 0;
            grain = std::max<int64_t>(1, chunk);
            pending = 0;
            active = true;
        }
        while (help_once()) {
        }
        for (;;) {
            {
                std::lock_guard<std::mutex> lk(mu);
                if (pending == 0) {
                    active = false;
                    break;
                }
            }
            std::this_thread::yield();
        }
    }
};

constexpr size_t kNoDone = (size_t)-1;
constexpr uint32_t kFusedSmemMax = 96 * 1024;  // dynamic shared memory limit of k_fused_subtrees (descriptors + pool + data)
struct Launch {
    int lane;        // stream lane the launch goes to
    int kind;        // 0 fused, 1 generic, 2 gemm, 3 finalize
    int vt;
    size_t inst_off; // byte offset of the instance array in the staging buffer
    size_t starts_off;
    size_t counter_off;  // gemm v2: a zeroed u32 tile counter inside the staging buffer
    size_t done_off = kNoDone;  // dataflow launch: per-instance completion counters inside the staging buffer
    int n_insts;
    uint32_t grid;
    uint32_t smem;
    double ops = 0, bytes = 0;  // big steps only: tropical ops / bytes moved by the launch (per-launch roofline records)
};

struct tb_ctx {
    // ---- multi-GPU context (tb_init_multi): one sub-context per device, this object only coordinates
    std::vector<tb_ctx*> subs;
    std::vector<void*> comms;          // ncclComm_t per sub-context (empty: single device, or host combine)
    bool host_combine = false;         // the device list repeats a device (testing on a single-GPU box): no NCCL possible
    tb_ctx* parent = nullptr;          // sub-context: the multi context that owns it
    const int64_t* index_map = nullptr;  // sub-context inside a multi call: local branch index -> index in the call's result vector
    bool prefill_results = false;        // sub-context inside a multi call: the result vector starts as -inf (input of the max-reduce)
    Plan* resident = nullptr;  // head of the intrusive list of plans whose descriptors live on this context
    bool stream_open = false;  // a tb_stream owns the context between tb_stream_begin and tb_stream_finish
    int call_wave = 0;  // wave size of the current call (a small call is cut into more, smaller waves: all lanes busy)
    std::thread reaper;  // frees the host side of the previous call's temporary plans that do not go back to the pool
    // temporary plans of tb_contract_networks / tb_stream_push are recycled: the next call compiles into the same objects
    // (arrays keep their capacity), so a steady stream of calls neither allocates nor frees host memory per branch
    void* permute_buf = nullptr;  // tb_permute_bits: source | destination, kept between calls
    size_t permute_cap = 0;
    // Work lists of resident plans (tb_contract_batch / tb_contract): a group of plans contracted again with the same arena
    // placement needs the same instance arrays, tile offsets and launch geometry, so they are kept on the device -- a pristine
    // copy and a working copy (the kernels count tiles and completions down inside it) -- and replayed: no host pass over the
    // plans, no upload.  Keyed on the exact plan sequence (pointer + upload id), value type, arena and lane state.
    struct ListCacheEntry {
        std::vector<uint64_t> key;
        std::vector<Launch> launches;
        size_t first_solo_launch = (size_t)-1;
        void* d_pristine = nullptr;
        void* d_work = nullptr;
        size_t bytes = 0;
        int lane_after = 0;
        uint64_t used = 0;
    };
    std::vector<ListCacheEntry> list_cache;
    std::vector<uint64_t> list_seen;  // hashes of the groups contracted once: an entry is made when a group comes back, so a
    size_t list_seen_pos = 0;         // caller that never repeats a group (new plans every call) pays for no device copies
    size_t list_cache_bytes = 0;
    uint64_t list_cache_clock = 0;
    int64_t list_cache_hits = 0, list_cache_misses = 0;
    static constexpr size_t kListCacheMaxEntries = 64;
    static constexpr size_t kListCacheMaxBytes = 512u << 20;
    void* table_buf = nullptr;    // tb_table_configs / tb_branching_table: device scratch and output rows, kept between calls
    size_t table_cap = 0;         // (a host calls them once per branching step: no allocation in steady state)
    void* table_out = nullptr;
    size_t table_out_cap = 0;
    HelperJob helper;
    std::mutex pool_mu;
    std::vector<tb_plan*> plan_pool;
    size_t plan_pool_bytes = 0;                             // host memory the pooled plans hold (array capacities)
    static constexpr size_t kPlanPoolMax = 1u << 15;        // ~30 KB of descriptors each for sc 20 branches ...
    static constexpr size_t kPlanPoolMaxBytes = 512u << 20;  // ... and never more than this in total
    tb_options opts{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<BlobChunk> chunks;
    // staging ring: pinned host + device buffers for descriptor blobs and work lists.  A slot is reused only
    // after the event recorded behind its last consumer (copy or kernels) has completed.
    struct Slot {
        void* h = nullptr;
        size_t hcap = 0;
        void* d = nullptr;
        size_t dcap = 0;
        cudaEvent_t ev = nullptr;
        bool busy = false;
        uint64_t seq = 0;  // when the slot was last handed out (the busy slot with the smallest seq frees first)
    };
    static constexpr int kSlots = 6;
    uint64_t slot_seq = 0;
    Slot slots[kSlots];
    int slot_cursor = 0;
    cudaStream_t copy_stream = nullptr;    // H2D of blobs / work lists, never blocked behind compute
    cudaEvent_t ev_copy = nullptr;
    int lane_cursor = 0;                   // round-robin position of the next wave's lane (kept across batches)
    double* d_results = nullptr;
    double* h_results = nullptr;
    size_t results_cap = 0;
    std::string last_error;
    double last_ms = 0;
    int64_t last_launches = 0;
    int64_t last_plan_arena_base_elems = 0;  // where tb_contract placed the single plan
    int sm_count = 148;
    bool own_stream = true;
    static constexpr int kMaxLanes = 8;
    double host_ms[6] = {0, 0, 0, 0, 0, 0};  // last call: compile, upload, build lists, launch+wait, destroy, total
    // TB_TRACE_CALL=1: host-side stalls (allocations, waits for a staging slot) longer than 0.2 ms, with the call's clock
    bool trace_on = false;
    double trace_t0 = 0;
    std::vector<std::string> trace_notes;
    int gemm2_ctas_per_sm = 2;
    bool staged_epilogue = true;           // TB_EPI_DIRECT=1: scatter stores straight from registers (A/B testing)
    bool gemm_v1 = false;                  // TB_GEMM_V1=1: the non-persistent cp.async GEMM kernel (A/B testing)
    int n_lanes = 4;                       // waves in flight: lane 0 = main stream, others = side streams
    int waves_per_lane = 0;                // waves per lane a small call is cut into; 0 = by plan weight (TB_WAVES_PER_LANE)
    cudaStream_t side[kMaxLanes] = {};     // side[1..n_lanes-1]
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxLanes] = {};
    // Executor of the non-fused steps of a wave.  dataflow: ONE persistent launch runs every level (tiles wait on
    // completion counters); level-synchronous: one launch per dependency level and kernel kind.  Measured on B200
    // (profiles/r2e_*): dataflow wins where launches / level tails dominate (few or light plans: BASELINE configs 3 and 5),
    // the level-synchronous launches win on DPX-bound calls (configs 2 and 4): their GEMM instance carries no generic-tile
    // code (uniform-datapath main loop) and the HBM-bound generic steps get a kernel of their own with twice the warps.
    int dataflow_mode = 0;         // 0 = per call by plan count / weight, 1 = always dataflow (TB_DATAFLOW=1), 2 = never (TB_LEVEL_SYNC=1)
    bool dataflow = true;          // the choice for the current call
    int call_lanes = 1;            // stream lanes the current call can fill (its waves): the persistent dataflow kernels of
                                   // different lanes share the SMs, each takes 1 / call_lanes of the CTA slots
    bool solo_fence = false;       // a solo wave ran on the whole arena: the side lanes must wait for it before their next wave
    int profile_mode = 0;          // 1: events around every launch on its own lane (lanes stay concurrent); 2: single lane
    struct ProfRec {
        int kind;
        cudaEvent_t e0, e1;
        double ops, bytes;
        int n_insts;
        uint32_t grid;
    };
    std::vector<ProfRec> prof_recs;  // launches of the current call (profile_mode != 0)
    size_t prof_used = 0;            // events of prof_events handed out in the current call
    double prof_union_ms[4] = {0, 0, 0, 0};  // per kind: length of the union of the launches' [start, end] intervals
    double prof_ms[4] = {0, 0, 0, 0};
    int64_t prof_launches[4] = {0, 0, 0, 0};
    int64_t h2d_bytes = 0, d2h_bytes = 0;  // of the last contract call
    std::vector<cudaEvent_t> prof_events;
};

namespace {

int set_err(tb_ctx* ctx, int code, const std::string& msg) {
    g_tls_error = msg;
    if (ctx) ctx->last_error = msg;
    return code;
}

#define TB_CUDA(ctx, call)                                                                                   \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return set_err(ctx, e__ == cudaErrorMemoryAllocation ? TB_ERR_OUT_OF_MEMORY : TB_ERR_CUDA,       \
                           std::string(#call) + ": " + cudaGetErrorString(e__));                             \
    } while (0)

int sync_all_lanes(tb_ctx* ctx);

// worker threads of one call: always joined, also when the launching thread unwinds (an exception then still reaches the
// function-try-block of the ABI entry point instead of std::terminate on a joinable std::thread)
struct ThreadGroup {
    std::vector<std::thread> th;
    template <class F> void spawn(F&& f) { th.emplace_back(std::forward<F>(f)); }
    void join() {
        for (auto& t : th)
            if (t.joinable()) t.join();
        th.clear();
    }
    ~ThreadGroup() { join(); }
};

// compile_plan for worker threads: never throws (an exception escaping a std::thread is std::terminate)
int compile_guarded(const tb_network& net, uint32_t flags, Plan& P, std::string& err) noexcept {
    try {
        return compile_plan(net, flags, P, err);
    } catch (const std::bad_alloc&) {
        err = "host memory allocation failed";
        return TB_ERR_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        err = std::string("C++ exception: ") + e.what();
        return TB_ERR_INTERNAL;
    } catch (...) {
        err = "unknown C++ exception";
        return TB_ERR_INTERNAL;
    }
}
tb_plan* compile_new(const tb_network& net, uint32_t flags, int& code, std::string& err) noexcept {
    tb_plan* p = nullptr;
    try {
        p = new tb_plan();
    } catch (...) {
        code = TB_ERR_OUT_OF_MEMORY;
        err = "host memory allocation failed";
        return nullptr;
    }
    code = compile_guarded(net, flags, p->p, err);
    if (code) {
        delete p;
        return nullptr;
    }
    return p;
}

// A temporary of tb_contract_networks / tb_stream_push: compiled in the worker thread's own scratch plan (memory that
// stays hot in its cache), then only the descriptors are copied into a plan object from the context's pool.
tb_plan* compile_temporary(tb_ctx* ctx, const tb_network& net, uint32_t flags, int& code, std::string& err) noexcept {
    tb_plan* p = nullptr;
    try {
        static thread_local Plan scratch;
        scratch.recycle();
        code = compile_guarded(net, flags | TB_PLAN_TEMPORARY, scratch, err);
        if (code) return nullptr;
        {
            std::lock_guard<std::mutex> lk(ctx->pool_mu);
            if (!ctx->plan_pool.empty()) {
                p = ctx->plan_pool.back();
                ctx->plan_pool.pop_back();
                ctx->plan_pool_bytes -= std::min(ctx->plan_pool_bytes, p->p.descriptor_capacity_bytes());
            }
        }
        if (!p) p = new tb_plan();
        scratch.copy_descriptors_to(p->p);
        return p;
    } catch (const std::bad_alloc&) {
        code = TB_ERR_OUT_OF_MEMORY;
        err = "host memory allocation failed";
    } catch (...) {
        code = TB_ERR_INTERNAL;
        err = "unknown C++ exception";
    }
    delete p;
    return nullptr;
}

// NCCL, resolved with dlopen on first use (tb_init_multi): single-GPU users need no NCCL at all
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
    int (*CommDestroy)(void* comm) = nullptr;
    int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};
constexpr int kNcclFloat64 = 8, kNcclMax = 2;  // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            api.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "unknown");
            return;
        }
        auto sym = [&](const char* n) {
            void* p = dlsym(api.lib, n);
            if (!p && api.err.empty()) api.err = std::string("libnccl has no symbol ") + n;
            return p;
        };
        api.CommInitAll = (int (*)(void**, int, const int*))sym("ncclCommInitAll");
        api.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
        api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
        api.GroupStart = (int (*)())sym("ncclGroupStart");
        api.GroupEnd = (int (*)())sym("ncclGroupEnd");
        api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    });
    return &api;
}

#define TB_NCCL(ctx, call)                                                                                         \
    do {                                                                                                           \
        int e__ = (call);                                                                                          \
        if (e__ != 0) return set_err(ctx, TB_ERR_NCCL, std::string(#call) + ": " + nccl_api()->GetErrorString(e__)); \
    } while (0)


void link_resident(tb_ctx* ctx, Plan& P) {
    P.res_prev = nullptr;
    P.res_next = ctx->resident;
    if (ctx->resident) ctx->resident->res_prev = &P;
    ctx->resident = &P;
}
void unlink_resident(tb_ctx* ctx, Plan& P) {
    if (P.res_prev) P.res_prev->res_next = P.res_next;
    else if (ctx->resident == &P) ctx->resident = P.res_next;
    if (P.res_next) P.res_next->res_prev = P.res_prev;
    P.res_prev = P.res_next = nullptr;
}

// times a host-side operation that may stall the launching thread (TB_TRACE_CALL diagnostics)
struct StallTimer {
    tb_ctx* ctx;
    const char* what;
    double t0;
    StallTimer(tb_ctx* c, const char* w) : ctx(c), what(w), t0(c->trace_on ? now_ms() : 0) {}
    ~StallTimer() {
        if (!ctx->trace_on) return;
        const double t1 = now_ms();
        if (t1 - t0 > 0.2) {
            char buf[160];
            snprintf(buf, sizeof buf, "%s: %.3f ms at %.3f", what, t1 - t0, t0 - ctx->trace_t0);
            ctx->trace_notes.push_back(buf);
        }
    }
};

int ensure_arena(tb_ctx* ctx, size_t need_bytes) {
    if (ctx->arena && ctx->arena_bytes >= need_bytes) return TB_OK;
    size_t want = (size_t)ctx->opts.arena_bytes;
    if (want == 0) {
        size_t fr = 0, tot = 0;
        TB_CUDA(ctx, cudaMemGetInfo(&fr, &tot));
        want = (size_t)((double)fr * 0.6);
        want = std::min(want, (size_t)96 << 30);
    }
    if (ctx->arena && ctx->arena_bytes >= want && need_bytes > want)
        return set_err(ctx, TB_ERR_OUT_OF_MEMORY, "a single branch needs " + std::to_string(need_bytes) + " bytes of arena, more than configured");
    if (need_bytes > want) return set_err(ctx, TB_ERR_OUT_OF_MEMORY, "a single branch needs " + std::to_string(need_bytes) + " bytes of arena; arena limit is " + std::to_string(want));
    StallTimer stall(ctx, "arena (re)allocation");
    if (ctx->arena) {
        int rcs = sync_all_lanes(ctx);
        if (rcs) return rcs;
        cudaFree(ctx->arena);
        ctx->arena = nullptr;
        ctx->arena_bytes = 0;
    }
    // grow lazily: start small, double up to `want`
    size_t sz = std::max<size_t>(need_bytes, std::min<size_t>(want, (size_t)1 << 30));
    sz = std::min(std::max(sz, need_bytes), want);
    TB_CUDA(ctx, cudaMalloc(&ctx->arena, sz));
    ctx->arena_bytes = sz;
    return TB_OK;
}

int acquire_slot(tb_ctx* ctx, size_t hbytes, size_t dbytes, tb_ctx::Slot** out) {
    // prefer a free slot that is already big enough (pinned allocations cost milliseconds), then any free slot,
    // and only then wait for the oldest busy one
    int pick = -1;
    for (int q = 0; q < tb_ctx::kSlots; ++q) {
        tb_ctx::Slot& c = ctx->slots[q];
        if (c.busy && c.ev && cudaEventQuery(c.ev) == cudaSuccess) c.busy = false;
        if (c.busy) continue;
        const bool fits = c.hcap >= hbytes && c.dcap >= dbytes;
        if (pick < 0) pick = q;
        else {
            tb_ctx::Slot& p = ctx->slots[pick];
            const bool pfits = p.hcap >= hbytes && p.dcap >= dbytes;
            if ((fits && !pfits) || (fits == pfits && (fits ? c.hcap + c.dcap < p.hcap + p.dcap : c.hcap + c.dcap > p.hcap + p.dcap))) pick = q;
        }
    }
    cudaGetLastError();  // cudaEventQuery returns cudaErrorNotReady for pending events
    // Growing a slot re-pins host memory (~2 ms per MB) and frees device memory (a device-wide synchronisation), so once half
    // of the ring is big enough for this request the others are left alone: rather than growing a small free slot (or
    // waiting for a small busy one and growing it), wait for the big-enough slot that was handed out first.  Without this
    // a small slot that the warm-up calls never happened to pick grows in the middle of some later call.
    {
        int n_fit = 0, oldest_fit = -1;
        for (int q = 0; q < tb_ctx::kSlots; ++q) {
            const tb_ctx::Slot& c = ctx->slots[q];
            if (c.hcap < hbytes || c.dcap < dbytes) continue;
            ++n_fit;
            if (c.busy && (oldest_fit < 0 || c.seq < ctx->slots[oldest_fit].seq)) oldest_fit = q;
        }
        const bool pick_fits = pick >= 0 && ctx->slots[pick].hcap >= hbytes && ctx->slots[pick].dcap >= dbytes;
        if (!pick_fits && n_fit >= tb_ctx::kSlots / 2 && oldest_fit >= 0) pick = oldest_fit;
    }
    if (pick < 0) {
        pick = ctx->slot_cursor;
        ctx->slot_cursor = (ctx->slot_cursor + 1) % tb_ctx::kSlots;
    }
    tb_ctx::Slot& sl = ctx->slots[pick];
    sl.seq = ++ctx->slot_seq;
    if (!sl.ev) TB_CUDA(ctx, cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
    if (sl.busy) {
        StallTimer stall(ctx, "wait for a staging slot");
        TB_CUDA(ctx, cudaEventSynchronize(sl.ev));
        sl.busy = false;
    }
    // (Re)pinning host memory costs ~2 ms per MB (profiles/x3_trace_cfg2_before.txt: 6-15 ms stalls in the middle of a call),
    // so a slot that has to grow grows generously -- twice the request, at least 16 MB, at least the largest size any
    // slot has reached: the ring stops growing after the first call instead of creeping up for dozens of calls
    size_t max_h = 0, max_d = 0;
    for (const tb_ctx::Slot& c : ctx->slots) {
        max_h = std::max(max_h, c.hcap);
        max_d = std::max(max_d, c.dcap);
    }
    if (sl.hcap < hbytes) {
        StallTimer stall(ctx, "staging slot grows (pinned host)");
        if (sl.h) cudaFreeHost(sl.h);
        sl.h = nullptr;
        size_t cap = std::max(std::max(hbytes * 2, (size_t)16 << 20), max_h);
        TB_CUDA(ctx, cudaMallocHost(&sl.h, cap));
        sl.hcap = cap;
    }
    if (sl.dcap < dbytes) {
        StallTimer stall(ctx, "staging slot grows (device)");
        if (sl.d) cudaFree(sl.d);
        sl.d = nullptr;
        size_t cap = std::max(std::max(dbytes * 2, (size_t)8 << 20), max_d);
        TB_CUDA(ctx, cudaMalloc(&sl.d, cap));
        sl.dcap = cap;
    }
    sl.busy = true;  // reserved; the caller records sl.ev behind the last consumer
    *out = &sl;
    return TB_OK;
}

int sync_all_lanes(tb_ctx* ctx) {
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    for (int l = 1; l < ctx->n_lanes; ++l) TB_CUDA(ctx, cudaStreamSynchronize(ctx->side[l]));
    TB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& sl : ctx->slots) sl.busy = false;  // everything that used a slot has completed
    return TB_OK;
}

int ensure_results(tb_ctx* ctx, size_t n) {
    if (ctx->results_cap >= n) return TB_OK;
    if (ctx->d_results) cudaFree(ctx->d_results);
    if (ctx->h_results) cudaFreeHost(ctx->h_results);
    ctx->d_results = nullptr;
    ctx->h_results = nullptr;
    size_t cap = std::max<size_t>(n * 2, 1024);
    TB_CUDA(ctx, cudaMalloc(&ctx->d_results, cap * sizeof(double)));
    TB_CUDA(ctx, cudaMallocHost(&ctx->h_results, cap * sizeof(double)));
    ctx->results_cap = cap;
    return TB_OK;
}

// upload the descriptor blobs of all plans that are not resident yet (one staging copy per chunk)
std::atomic<uint64_t> g_upload_uid{0};

int ensure_uploaded(tb_ctx* ctx, tb_plan* const* plans, int64_t n) {
    std::vector<tb_plan*> todo;
    std::unordered_set<tb_plan*> seen;
    seen.reserve((size_t)n * 2);
    size_t total = 0;
    for (int64_t i = 0; i < n; ++i) {
        tb_plan* p = plans[i];
        if (!p) continue;
        if (p->p.d_blob) {
            if (p->p.owner != ctx) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "plan is resident on a different context");
            continue;
        }
        if (!seen.insert(p).second) continue;  // the same plan may appear more than once in a batch
        todo.push_back(p);
    }
    if (todo.empty()) return TB_OK;
    // section sizes are known from the plan; blobs are serialised straight into the pinned staging buffer
    auto a16 = [](size_t x) { return (x + 15) / 16 * 16; };
    auto blob_size = [&](const Plan& P) {
        return a16(P.pool_bytes()) + a16(P.sub_steps.size() * sizeof(SubStep)) + a16(P.big_steps.size() * sizeof(BigStep));
    };
    auto write_blob = [&](Plan& P, uint8_t* dst) {
        size_t pool_b = a16(P.pool_bytes()), sub_b = a16(P.sub_steps.size() * sizeof(SubStep));
        P.sub_blob_off = pool_b;
        P.big_blob_off = pool_b + sub_b;
        P.blob_bytes = blob_size(P);
        P.write_pool(dst);
        if (!P.sub_steps.empty()) std::memcpy(dst + P.sub_blob_off, P.sub_steps.data(), P.sub_steps.size() * sizeof(SubStep));
        if (!P.big_steps.empty()) std::memcpy(dst + P.big_blob_off, P.big_steps.data(), P.big_steps.size() * sizeof(BigStep));
    };
    std::vector<size_t> bsz(todo.size());
    for (size_t i = 0; i < todo.size(); ++i) {
        bsz[i] = (blob_size(todo[i]->p) + 255) / 256 * 256;
        total += bsz[i];
    }
    size_t pos = 0;
    while (pos < todo.size()) {
        BlobChunk* ck = ctx->chunks.empty() ? nullptr : &ctx->chunks.back();
        if (!ck || ck->used + bsz[pos] > ck->cap) {
            size_t rest = 0;
            for (size_t j = pos; j < todo.size(); ++j) rest += bsz[j];
            BlobChunk nc;
            size_t prev = ctx->chunks.empty() ? 0 : ctx->chunks.back().cap;
            nc.cap = std::max<size_t>(std::max(rest, 2 * prev), (size_t)32 << 20);
            {
                StallTimer stall(ctx, "descriptor chunk allocation");
                TB_CUDA(ctx, cudaMalloc(&nc.d, nc.cap));
            }
            ctx->chunks.push_back(nc);
            ck = &ctx->chunks.back();
        }
        size_t first = pos, bytes = 0;
        while (pos < todo.size() && ck->used + bytes + bsz[pos] <= ck->cap) bytes += bsz[pos++];
        tb_ctx::Slot* sl = nullptr;
        int rc = acquire_slot(ctx, bytes, 0, &sl);
        if (rc) return rc;
        size_t o = 0;
        std::vector<size_t> at(pos - first);
        for (size_t j = first; j < pos; ++j) {
            at[j - first] = o;
            todo[j]->p.d_blob = (uint8_t*)ck->d + ck->used + o;
            todo[j]->p.upload_uid = g_upload_uid.fetch_add(1, std::memory_order_relaxed) + 1;
            todo[j]->p.owner = ctx;
            link_resident(ctx, todo[j]->p);
            ck->live++;
            o += bsz[j];
        }
        // serialise into the pinned staging buffer; the call's compile workers take chunks of this loop
        ctx->helper.run((int64_t)(pos - first), 8, [&](int64_t b, int64_t e) {
            for (int64_t q = b; q < e; ++q) write_blob(todo[first + (size_t)q]->p, (uint8_t*)sl->h + at[(size_t)q]);
        });
        TB_CUDA(ctx, cudaMemcpyAsync((uint8_t*)ck->d + ck->used, sl->h, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        TB_CUDA(ctx, cudaEventRecord(sl->ev, ctx->copy_stream));
        sl->busy = true;
        ctx->h2d_bytes += (int64_t)bytes;
        ck->used += bytes;
    }
    (void)total;
    return TB_OK;
}

// CTAs of one persistent dataflow kernel: the kernels of the call's concurrent lanes co-reside, each on its share of the
// CTA slots, so that the dependency stalls and the thin last levels of one wave are covered by the other waves' tiles
uint32_t dataflow_grid_cap(const tb_ctx* ctx) {
    static const int forced = [] {
        const char* e = getenv("TB_DF_GRID");  // experiments: CTAs per dataflow kernel
        return e ? atoi(e) : 0;
    }();
    if (forced > 0) return (uint32_t)forced;
    // half of the CTA slots when several waves are in flight: two kernels co-reside, the others queue behind them.  An even
    // share per lane (slots / 4) leaves each kernel too few CTAs to cover its own dependency stalls
    // (profiles/x3_df_wave_lane_sweep.txt: cfg5 1.73 ms with 74 CTAs per kernel, 1.59 with 148, 1.58 with 296)
    const int slots = std::max(1, ctx->gemm2_ctas_per_sm) * ctx->sm_count;
    return (uint32_t)std::max(1, slots / std::min(2, std::max(1, ctx->call_lanes)));
}

template <typename T>
void launch_one(tb_ctx* ctx, const Launch& L, uint8_t* dbase) {
    cudaStream_t st = L.lane == 0 ? ctx->stream : ctx->side[L.lane];
    switch (L.kind) {
        case 0:
            k_fused_subtrees<T><<<L.grid, FUSED_THREADS, L.smem, st>>>((const SubInst*)(dbase + L.inst_off), L.n_insts);
            break;
        case 1:
            k_generic<T><<<L.grid, BIG_THREADS, 0, st>>>((const BigInst*)(dbase + L.inst_off), (const uint32_t*)(dbase + L.starts_off), L.n_insts);
            break;
        case 2:
            if constexpr (sizeof(T) == 8) {
                // 8-byte value types are compiled with TB_PLAN_NO_GEMM: no such launch exists
            } else if constexpr (std::is_same<T, int16_t>::value) {
                uint32_t grid = std::min<uint32_t>(L.grid, (uint32_t)(std::max(1, ctx->gemm2_ctas_per_sm) * ctx->sm_count));
                if (L.done_off != kNoDone) grid = std::min<uint32_t>(grid, dataflow_grid_cap(ctx));
                if (L.done_off == kNoDone)
                    k_gemm2h<false><<<grid, G2_THREADS, G2_SMEM_BYTES, st>>>((const BigInst*)(dbase + L.inst_off), (const uint32_t*)(dbase + L.starts_off),
                                                                            L.n_insts, L.grid, (unsigned int*)(dbase + L.counter_off), nullptr);
                else
                    k_gemm2h<true><<<grid, G2_THREADS, G2_SMEM_BYTES, st>>>((const BigInst*)(dbase + L.inst_off), (const uint32_t*)(dbase + L.starts_off),
                                                                           L.n_insts, L.grid, (unsigned int*)(dbase + L.counter_off),
                                                                           (unsigned int*)(dbase + L.done_off));
            } else if (ctx->gemm_v1) {
                k_gemm<T><<<L.grid, BIG_THREADS, GEMM_SMEM_BYTES, st>>>((const BigInst*)(dbase + L.inst_off), (const uint32_t*)(dbase + L.starts_off), L.n_insts);
            } else {
                uint32_t grid = std::min<uint32_t>(L.grid, (uint32_t)(std::max(1, ctx->gemm2_ctas_per_sm) * ctx->sm_count));
                if (L.done_off != kNoDone) grid = std::min<uint32_t>(grid, dataflow_grid_cap(ctx));
                if (L.done_off == kNoDone)
                    k_gemm2<T, false><<<grid, G2_THREADS, G2_SMEM_BYTES, st>>>((const BigInst*)(dbase + L.inst_off), (const uint32_t*)(dbase + L.starts_off),
                                                                              L.n_insts, L.grid, (unsigned int*)(dbase + L.counter_off),
                                                                              ctx->staged_epilogue ? 1 : 0, nullptr);
                else
                    k_gemm2<T, true><<<grid, G2_THREADS, G2_SMEM_BYTES, st>>>((const BigInst*)(dbase + L.inst_off), (const uint32_t*)(dbase + L.starts_off),
                                                                             L.n_insts, L.grid, (unsigned int*)(dbase + L.counter_off),
                                                                             ctx->staged_epilogue ? 1 : 0, (unsigned int*)(dbase + L.done_off));
            }
            break;
        case 3:
            k_finalize<T><<<(L.n_insts + 127) / 128, 128, 0, st>>>((const FinalInst*)(dbase + L.inst_off), L.n_insts, ctx->d_results);
            break;
    }
}

// contract plans[idx[0..m)] (all of one value type); results land in ctx->d_results[idx[i]]
int run_group(tb_ctx* ctx, tb_plan* const* plans, const std::vector<int64_t>& idx, int vt, std::vector<int32_t>& status,
              bool single_plan_mode) {
    if (idx.empty()) return TB_OK;
    const size_t elem = (size_t)Plan::elem_size_of(vt);
    const bool wide = elem == 8;  // Tropical{Float64} / size+configuration: no persistent GEMM kernel to host a dataflow launch
    const int max_wave = ctx->call_wave > 0 ? ctx->call_wave : (ctx->opts.max_wave > 0 ? ctx->opts.max_wave : 128);
    const int NL = (ctx->profile_mode == 2 || single_plan_mode) ? 1 : ctx->n_lanes;
    // ---- launches of a group: shared by the first contraction (lists just built, in a staging slot) and by the replays of
    // a cached group (lists resident on the device)
    auto issue_launches = [&](const std::vector<Launch>& launches, size_t first_solo_launch, uint8_t* dbase, tb_ctx::Slot* sl) -> int {
        int rc = TB_OK;
        TB_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
        TB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
        for (int l = 1; l < NL; ++l) TB_CUDA(ctx, cudaStreamWaitEvent(ctx->side[l], ctx->ev_copy, 0));
        // a solo wave of the previous group used the whole arena on the main stream: no side lane may touch its partition
        // before that wave has finished (the main stream itself is ordered behind it)
        if (ctx->solo_fence) {
            for (int l = 1; l < ctx->n_lanes; ++l) TB_CUDA(ctx, cudaStreamWaitEvent(ctx->side[l], ctx->ev_fork, 0));
            ctx->solo_fence = false;
        }
        auto prof_event = [&](cudaEvent_t* out) -> int {
            if (ctx->prof_used == ctx->prof_events.size()) {
                cudaEvent_t e;
                TB_CUDA(ctx, cudaEventCreate(&e));
                ctx->prof_events.push_back(e);
            }
            *out = ctx->prof_events[ctx->prof_used++];
            return TB_OK;
        };
        auto join_lanes = [&]() -> int {
            for (int l = 1; l < NL; ++l) {
                TB_CUDA(ctx, cudaEventRecord(ctx->ev_join[l], ctx->side[l]));
                TB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[l], 0));
            }
            return TB_OK;
        };
        bool joined = (NL == 1);
        for (size_t li = 0; li < launches.size(); ++li) {
            if (li == first_solo_launch && !joined) {
                rc = join_lanes();
                if (rc) return rc;
                joined = true;
            }
            const Launch& L = launches[li];
            cudaStream_t st = L.lane == 0 ? ctx->stream : ctx->side[L.lane];
            tb_ctx::ProfRec pr{L.kind, nullptr, nullptr, L.ops, L.bytes, L.n_insts, L.grid};
            if (ctx->profile_mode) {
                if ((rc = prof_event(&pr.e0)) || (rc = prof_event(&pr.e1))) return rc;
                TB_CUDA(ctx, cudaEventRecord(pr.e0, st));
            }
            if (vt == TB_VALUE_I32) launch_one<int32_t>(ctx, L, dbase);
            else if (vt == TB_VALUE_I16X2) launch_one<int16_t>(ctx, L, dbase);
            else if (vt == TB_VALUE_F64) launch_one<double>(ctx, L, dbase);
            else if (vt == TB_VALUE_SIZE_CONFIG) launch_one<long long>(ctx, L, dbase);
            else launch_one<float>(ctx, L, dbase);
            if (ctx->profile_mode) {
                TB_CUDA(ctx, cudaEventRecord(pr.e1, st));
                ctx->prof_recs.push_back(pr);
            }
        }
        // the main stream joins the lanes at the end of every group: the slot event (and the caller's final sync)
        // then cover every kernel that reads this group's work lists
        if (!joined) {
            rc = join_lanes();
            if (rc) return rc;
        }
        TB_CUDA(ctx, cudaGetLastError());
        if (sl) {
            TB_CUDA(ctx, cudaEventRecord(sl->ev, ctx->stream));
            sl->busy = true;
        }
        if (first_solo_launch != (size_t)-1) {
            TB_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
            ctx->solo_fence = true;
        }
        ctx->last_launches += (int64_t)launches.size();
        return TB_OK;
    };
    // ---- cached work lists (resident plans only: the temporaries of the *_networks calls are contracted once)
    static const bool list_cache_on = [] {
        const char* e = getenv("TB_NO_LIST_CACHE");  // A/B: rebuild and upload the lists on every call
        return !(e && atoi(e) != 0);
    }();
    bool cacheable = list_cache_on && !ctx->trace_on;
    for (size_t q = 0; q < idx.size() && cacheable; ++q) cacheable = !plans[idx[q]]->p.temporary;
    std::vector<uint64_t> key;
    const int lane_at_entry = ctx->lane_cursor;
    auto make_key = [&]() {
        key.clear();
        key.reserve(8 + idx.size() * 3);
        key.push_back((uint64_t)vt | ((uint64_t)NL << 8) | ((uint64_t)(ctx->dataflow ? 1 : 0) << 16) | ((uint64_t)(single_plan_mode ? 1 : 0) << 17) |
                      ((uint64_t)(uint32_t)max_wave << 32));
        key.push_back((uint64_t)(uintptr_t)ctx->arena);
        key.push_back((uint64_t)ctx->arena_bytes);
        key.push_back((uint64_t)lane_at_entry);
        for (int64_t i : idx) {
            key.push_back((uint64_t)(ctx->index_map ? ctx->index_map[i] : i));
            key.push_back((uint64_t)(uintptr_t)plans[i]);
            key.push_back(plans[i]->p.upload_uid);
        }
    };
    bool remember = false;
    auto key_hash = [&]() {
        uint64_t h = 1469598103934665603ull;
        for (uint64_t w : key) h = (h ^ w) * 1099511628211ull;
        return h;
    };
    if (cacheable) {
        make_key();
        for (auto& e : ctx->list_cache) {
            if (e.key != key) continue;
            const double t_l0 = now_ms();
            e.used = ++ctx->list_cache_clock;
            ++ctx->list_cache_hits;
            TB_CUDA(ctx, cudaMemcpyAsync(e.d_work, e.d_pristine, e.bytes, cudaMemcpyDeviceToDevice, ctx->copy_stream));
            int rc = issue_launches(e.launches, e.first_solo_launch, (uint8_t*)e.d_work, nullptr);
            if (rc) return rc;
            ctx->lane_cursor = e.lane_after;
            if (single_plan_mode) ctx->last_plan_arena_base_elems = 0;
            ctx->host_ms[3] += now_ms() - t_l0;
            return TB_OK;
        }
        ++ctx->list_cache_misses;
        bool seen = false;
        const uint64_t h = key_hash();
        for (uint64_t x : ctx->list_seen) seen = seen || x == h;
        if (!seen) {
            remember = true;    // first time: remember the group only (its hash is taken once the arena has its final
            cacheable = false;  // place for this call, below)
        }
    }
    // ---- waves
    size_t max_need = 0;
    for (int64_t i : idx) max_need = std::max(max_need, (size_t)plans[i]->p.arena_elems * elem + 256);
    int rc = ensure_arena(ctx, std::max<size_t>(max_need, 1 << 20));
    if (rc == TB_ERR_OUT_OF_MEMORY) {
        // keep going with what we have: oversize branches get a per-branch status below
        if (!ctx->arena) {
            rc = ensure_arena(ctx, 1 << 20);
            if (rc) return rc;
        }
    } else if (rc) {
        return rc;
    }
    // try to grow the arena so that NL full waves fit (bounded by the configured limit)
    {
        std::vector<size_t> needs;
        size_t total_all = 0;
        for (int64_t i : idx) {
            needs.push_back(((size_t)plans[i]->p.arena_elems * elem + 255) / 256 * 256);
            total_all += needs.back();
        }
        std::sort(needs.begin(), needs.end(), std::greater<size_t>());
        size_t top = 0;
        for (size_t q = 0; q < needs.size() && q < (size_t)max_wave; ++q) top += needs[q];
        size_t target = std::max((size_t)NL * needs[0], std::min(total_all, (size_t)NL * top)) + 4096;
        if (target > ctx->arena_bytes) {
            size_t want = (size_t)ctx->opts.arena_bytes;
            if (want == 0) {
                size_t fr = 0, tot = 0;
                cudaMemGetInfo(&fr, &tot);
                want = std::min((size_t)((double)(fr + ctx->arena_bytes) * 0.6), (size_t)96 << 30);
            }
            target = std::min(target, want);
            if (target > ctx->arena_bytes) {
                StallTimer stall(ctx, "arena grows for the waves of a batch");
                sync_all_lanes(ctx);
                void* na = nullptr;
                cudaFree(ctx->arena);
                ctx->arena = nullptr;
                if (cudaMalloc(&na, target) == cudaSuccess) {
                    ctx->arena = na;
                    ctx->arena_bytes = target;
                } else {
                    cudaGetLastError();
                    size_t back = std::max<size_t>(max_need, 1 << 20);
                    TB_CUDA(ctx, cudaMalloc(&ctx->arena, back));
                    ctx->arena_bytes = back;
                }
            }
        }
    }

    struct Wave {
        std::vector<int64_t> members;
        std::vector<size_t> base;  // arena byte offsets
        int levels = 0;
        int lane = 0;
    };
    // lanes own equal arena partitions; a plan too big for a partition runs alone on the whole arena
    // after everything else ("solo" waves, lane 0)
    std::vector<Wave> waves, solo;
    {
        const size_t cap = (ctx->arena_bytes / (size_t)NL) / 256 * 256;
        Wave cur;
        size_t used = 0;
        int lane = ctx->lane_cursor % NL;
        auto flush = [&]() {
            if (cur.members.empty()) return;
            cur.lane = lane;
            for (auto& b : cur.base) b += (size_t)lane * cap;
            waves.push_back(std::move(cur));
            cur = Wave();
            used = 0;
            lane = (lane + 1) % NL;
        };
        for (int64_t i : idx) {
            const Plan& P = plans[i]->p;
            size_t need = ((size_t)P.arena_elems * elem + 255) / 256 * 256;
            if (need > ctx->arena_bytes) {
                status[i] = TB_ERR_OUT_OF_MEMORY;
                cacheable = false;  // (a per-branch status is not part of a cached group)
                continue;
            }
            if (need > cap) {
                Wave w;
                w.members.push_back(i);
                w.base.push_back(0);
                w.levels = P.n_levels;
                solo.push_back(std::move(w));
                continue;
            }
            if (used + need > cap || (int)cur.members.size() >= max_wave) flush();
            cur.members.push_back(i);
            cur.base.push_back(used);
            cur.levels = std::max(cur.levels, P.n_levels);
            used += need;
        }
        flush();
        ctx->lane_cursor = lane;
    }
    const size_t n_lane_waves = waves.size();
    for (auto& w : solo) waves.push_back(std::move(w));
    if (single_plan_mode && !waves.empty()) ctx->last_plan_arena_base_elems = 0;

    // ---- build work lists for every wave into one host buffer
    const double t_b0 = now_ms();
    std::vector<uint8_t> host;
    std::vector<Launch> launches;
    auto align16 = [&]() { host.resize((host.size() + 15) / 16 * 16); };
    size_t first_solo_launch = (size_t)-1;
    // scratch of the level-synchronous lists, reused by every level of every wave
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> ls_buckets;  // (level, kind) -> (member, step)
    std::vector<BigInst> ls_insts, ls_sorted;
    std::vector<uint32_t> ls_starts, ls_nk, ls_tiles, ls_ord;
    for (size_t wi = 0; wi < waves.size(); ++wi) {
        const Wave& w = waves[wi];
        if (wi == n_lane_waves) first_solo_launch = launches.size();
        // level 0: fused subtrees
        {
            align16();
            Launch L{};
            L.lane = w.lane;
            L.kind = 0;
            L.vt = vt;
            L.inst_off = host.size();
            uint32_t smem_elems = 0;  // largest shared-memory footprint of a subtree (bytes)
            std::vector<SubInst> insts;
            for (size_t m = 0; m < w.members.size(); ++m) {
                const Plan& P = plans[w.members[m]]->p;
                uint8_t* blob = (uint8_t*)P.d_blob;
                uint8_t* ab = (uint8_t*)ctx->arena + w.base[m];
                for (const SubTree& st : P.subtrees) {
                    SubInst si{};
                    si.steps = (const SubStep*)(blob + P.sub_blob_off) + st.first_step;
                    si.pool = blob;
                    si.out = ab + st.out_off * elem;
                    si.n_steps = st.n_steps;
                    si.n_pool = P.pool.size() <= 4096 ? (uint32_t)P.pool.size() : 0u;
                    uint32_t need = fused_desc_bytes(si.n_steps) + fused_pool_bytes(si.n_pool, (uint32_t)elem) + st.smem_elems * (uint32_t)elem;
                    if (need > kFusedSmemMax) {  // drop the pool copy first
                        si.n_pool = 0;
                        need = fused_desc_bytes(si.n_steps) + st.smem_elems * (uint32_t)elem;
                        if (need > kFusedSmemMax) return set_err(ctx, TB_ERR_UNSUPPORTED, "a fused subtree has too many steps for shared memory");
                    }
                    insts.push_back(si);
                    smem_elems = std::max(smem_elems, need);
                }
            }
            if (!insts.empty()) {
                L.n_insts = (int)insts.size();
                L.grid = (uint32_t)insts.size();
                L.smem = std::max<uint32_t>(smem_elems, 16);  // bytes: descriptors + pool copy + data
                size_t o = host.size();
                host.resize(o + insts.size() * sizeof(SubInst));
                std::memcpy(host.data() + o, insts.data(), insts.size() * sizeof(SubInst));
                launches.push_back(L);
            }
        }
        // one pass over the members sorts their big steps into (level, kind) buckets -- member-major inside a bucket, as the
        // launches list them -- instead of one scan of all members per level and kind
        ls_buckets.resize(std::max(ls_buckets.size(), (size_t)(w.levels + 1) * 2));
        for (auto& bk : ls_buckets) bk.clear();
        for (size_t m = 0; m < w.members.size(); ++m) {
            const Plan& P = plans[w.members[m]]->p;
            for (int lv = 1; lv <= P.n_levels; ++lv)
                for (int s = P.big_level_begin[lv]; s < P.big_level_begin[lv + 1]; ++s)
                    ls_buckets[(size_t)lv * 2 + (P.big_steps[s].kind == KIND_GEMM ? 1 : 0)].push_back({(uint32_t)m, (uint32_t)s});
        }
        if (ctx->dataflow && !wide) {
            // ---- dataflow: every big step of every member, level-major, in ONE persistent launch.  A tile waits on the
            // completion counters of the instances that produce its operands (BigInst::dep_a / dep_b) instead of on a
            // kernel boundary; tiles are handed out in this (topological) order, so the kernel cannot deadlock.
            std::vector<BigInst> insts;
            std::vector<uint32_t> starts, done_init;
            std::vector<std::vector<int32_t>> inst_of(w.members.size());  // member -> big step -> instance index
            for (size_t m = 0; m < w.members.size(); ++m) inst_of[m].assign(plans[w.members[m]]->p.big_steps.size(), -1);
            uint64_t tiles = 0;
            double df_ops = 0, df_bytes = 0;
            std::vector<std::pair<uint32_t, uint32_t>> lvl;  // (member, step) of one level and kind
            for (int lv = 1; lv <= w.levels; ++lv) {
                for (int kind : {(int)KIND_GENERIC, (int)KIND_GEMM}) {
                    lvl = ls_buckets[(size_t)lv * 2 + (kind == KIND_GEMM ? 1 : 0)];
                    if (kind == KIND_GEMM)  // longest reductions first inside a level
                        std::stable_sort(lvl.begin(), lvl.end(), [&](const std::pair<uint32_t, uint32_t>& x, const std::pair<uint32_t, uint32_t>& y) {
                            return plans[w.members[x.first]]->p.big_steps[x.second].nk > plans[w.members[y.first]]->p.big_steps[y.second].nk;
                        });
                    for (const auto& ms : lvl) {
                        const Plan& P = plans[w.members[ms.first]]->p;
                        const BigStep& st = P.big_steps[ms.second];
                        uint8_t* blob = (uint8_t*)P.d_blob;
                        BigInst bi{};
                        bi.step = (const BigStep*)(blob + P.big_blob_off) + ms.second;
                        bi.pool = blob;
                        bi.arena = (uint8_t*)ctx->arena + w.base[ms.first];
                        bi.tile_start = (uint32_t)tiles;
                        const int32_t da = P.big_dep_a[ms.second], db = P.big_dep_b[ms.second];
                        bi.dep_a = da >= 0 ? inst_of[ms.first][(size_t)da] : -1;
                        bi.dep_b = db >= 0 ? inst_of[ms.first][(size_t)db] : -1;
                        inst_of[ms.first][ms.second] = (int32_t)insts.size();
                        df_ops += (double)(1ull << (int)P.big_log2_ops[ms.second]);  // log2 ops <= 62 (plan compiler)
                        df_bytes += P.big_bytes[ms.second];
                        insts.push_back(bi);
                        starts.push_back((uint32_t)tiles);
                        done_init.push_back(st.n_tiles * (uint32_t)(G2_CONSUMERS / 32));
                        tiles += st.n_tiles;
                    }
                }
            }
            if (tiles > 0x7fffffffull) return set_err(ctx, TB_ERR_UNSUPPORTED, "a wave needs more than 2^31 tiles");
            if (!insts.empty()) {
                align16();
                Launch L{};
                L.lane = w.lane;
                L.kind = KIND_GEMM;
                L.vt = vt;
                L.inst_off = host.size();
                host.resize(host.size() + insts.size() * sizeof(BigInst));
                std::memcpy(host.data() + L.inst_off, insts.data(), insts.size() * sizeof(BigInst));
                align16();
                L.starts_off = host.size();
                host.resize(host.size() + starts.size() * 4);
                std::memcpy(host.data() + L.starts_off, starts.data(), starts.size() * 4);
                align16();
                L.done_off = host.size();
                host.resize(host.size() + done_init.size() * 4);
                std::memcpy(host.data() + L.done_off, done_init.data(), done_init.size() * 4);
                align16();
                L.counter_off = host.size();
                host.resize(host.size() + 16, 0);
                L.n_insts = (int)insts.size();
                L.grid = (uint32_t)tiles;
                L.ops = df_ops;
                L.bytes = df_bytes;
                launches.push_back(L);
            }
        } else {
        for (int lv = 1; lv <= w.levels; ++lv) {
            for (int kind : {(int)KIND_GENERIC, (int)KIND_GEMM}) {
                std::vector<BigInst>& insts = ls_insts;
                std::vector<uint32_t>&starts = ls_starts, &inst_nk = ls_nk, &inst_tiles = ls_tiles;
                insts.clear();
                starts.clear();
                inst_nk.clear();
                inst_tiles.clear();
                uint64_t tiles = 0;
                double ls_ops = 0, ls_bytes = 0;
                for (const auto& ms : ls_buckets[(size_t)lv * 2 + (kind == KIND_GEMM ? 1 : 0)]) {
                    const size_t m = ms.first;
                    const int s = (int)ms.second;
                    const Plan& P = plans[w.members[m]]->p;
                    uint8_t* blob = (uint8_t*)P.d_blob;
                    const BigStep* dsteps = (const BigStep*)(blob + P.big_blob_off);
                    BigInst bi{};
                    bi.step = dsteps + s;
                    bi.pool = blob;
                    bi.arena = (uint8_t*)ctx->arena + w.base[m];
                    bi.tile_start = (uint32_t)tiles;
                    insts.push_back(bi);
                    starts.push_back((uint32_t)tiles);
                    inst_nk.push_back(P.big_steps[s].nk);
                    inst_tiles.push_back(P.big_steps[s].n_tiles);
                    tiles += P.big_steps[s].n_tiles;
                    ls_ops += (double)(1ull << (int)P.big_log2_ops[(size_t)s]);
                    ls_bytes += P.big_bytes[(size_t)s];
                }
                if (insts.empty()) continue;
                if (tiles > 0x7fffffffull) return set_err(ctx, TB_ERR_UNSUPPORTED, "a level needs more than 2^31 CTAs");
                if (kind == KIND_GEMM) {
                    // longest reductions first: the dynamic tile scheduler then fills the tail with short tiles
                    std::vector<uint32_t>& ord = ls_ord;
                    ord.resize(insts.size());
                    for (size_t q = 0; q < ord.size(); ++q) ord[q] = (uint32_t)q;
                    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t x, uint32_t y) { return inst_nk[x] > inst_nk[y]; });
                    std::vector<BigInst>& si = ls_sorted;
                    si.resize(insts.size());
                    uint64_t acc_t = 0;
                    for (size_t q = 0; q < ord.size(); ++q) {
                        si[q] = insts[ord[q]];
                        si[q].tile_start = (uint32_t)acc_t;
                        starts[q] = (uint32_t)acc_t;
                        acc_t += inst_tiles[ord[q]];
                    }
                    insts.swap(si);
                }
                align16();
                Launch L{};
                L.lane = w.lane;
                L.kind = kind;
                L.vt = vt;
                L.inst_off = host.size();
                host.resize(host.size() + insts.size() * sizeof(BigInst));
                std::memcpy(host.data() + L.inst_off, insts.data(), insts.size() * sizeof(BigInst));
                align16();
                L.starts_off = host.size();
                host.resize(host.size() + starts.size() * 4);
                std::memcpy(host.data() + L.starts_off, starts.data(), starts.size() * 4);
                align16();
                L.counter_off = host.size();
                host.resize(host.size() + 16, 0);
                L.n_insts = (int)insts.size();
                L.grid = (uint32_t)tiles;
                L.ops = ls_ops;
                L.bytes = ls_bytes;
                launches.push_back(L);
            }
        }
        }
        // finalize
        {
            align16();
            Launch L{};
            L.lane = w.lane;
            L.kind = 3;
            L.vt = vt;
            L.inst_off = host.size();
            std::vector<FinalInst> fin;
            for (size_t m = 0; m < w.members.size(); ++m) {
                const Plan& P = plans[w.members[m]]->p;
                FinalInst f{};
                f.src = (uint8_t*)ctx->arena + w.base[m] + P.root_off * elem;
                f.out_index = ctx->index_map ? ctx->index_map[w.members[m]] : w.members[m];
                fin.push_back(f);
            }
            L.n_insts = (int)fin.size();
            host.resize(host.size() + fin.size() * sizeof(FinalInst));
            std::memcpy(host.data() + L.inst_off, fin.data(), fin.size() * sizeof(FinalInst));
            launches.push_back(L);
        }
    }
    if (remember) {
        make_key();  // (the arena may have grown since the lookup)
        const uint64_t h = key_hash();
        if (ctx->list_seen.size() < 256) ctx->list_seen.push_back(h);
        else ctx->list_seen[ctx->list_seen_pos++ % 256] = h;
    }
    if (launches.empty()) return TB_OK;
    const double t_l0 = now_ms();
    ctx->host_ms[2] += t_l0 - t_b0;
    tb_ctx::Slot* sl = nullptr;
    rc = acquire_slot(ctx, host.size(), host.size(), &sl);
    if (rc) return rc;
    std::memcpy(sl->h, host.data(), host.size());
    // work lists go up on the copy stream (behind the descriptor blobs they point to); every lane waits for them
    TB_CUDA(ctx, cudaMemcpyAsync(sl->d, sl->h, host.size(), cudaMemcpyHostToDevice, ctx->copy_stream));
    if (cacheable) {
        // keep the lists for the next contraction of this group: the pristine copy is taken on the copy stream before any
        // kernel of the group (they all wait for the event recorded behind it) starts counting inside the staging buffer
        while (!ctx->list_cache.empty() && (ctx->list_cache.size() >= tb_ctx::kListCacheMaxEntries ||
                                            ctx->list_cache_bytes + 2 * host.size() > tb_ctx::kListCacheMaxBytes)) {
            size_t old = 0;
            for (size_t q = 1; q < ctx->list_cache.size(); ++q)
                if (ctx->list_cache[q].used < ctx->list_cache[old].used) old = q;
            cudaFree(ctx->list_cache[old].d_pristine);  // (cudaFree waits for the work that may still read it)
            cudaFree(ctx->list_cache[old].d_work);
            ctx->list_cache_bytes -= 2 * ctx->list_cache[old].bytes;
            ctx->list_cache.erase(ctx->list_cache.begin() + (std::ptrdiff_t)old);
        }
        if (2 * host.size() <= tb_ctx::kListCacheMaxBytes) {
            tb_ctx::ListCacheEntry e;
            if (cudaMalloc(&e.d_pristine, host.size()) == cudaSuccess && cudaMalloc(&e.d_work, host.size()) == cudaSuccess) {
                make_key();  // (the arena may have grown since the lookup)
                e.key = key;
                e.launches = launches;
                e.first_solo_launch = first_solo_launch;
                e.bytes = host.size();
                e.lane_after = ctx->lane_cursor;
                e.used = ++ctx->list_cache_clock;
                TB_CUDA(ctx, cudaMemcpyAsync(e.d_pristine, sl->d, host.size(), cudaMemcpyDeviceToDevice, ctx->copy_stream));
                ctx->list_cache_bytes += 2 * host.size();
                ctx->list_cache.push_back(std::move(e));
            } else {
                cudaGetLastError();
                if (e.d_pristine) cudaFree(e.d_pristine);
                if (e.d_work) cudaFree(e.d_work);
            }
        }
    }
    rc = issue_launches(launches, first_solo_launch, (uint8_t*)sl->d, sl);
    if (rc) return rc;
    ctx->host_ms[3] += now_ms() - t_l0;
    ctx->h2d_bytes += (int64_t)host.size();
    return TB_OK;
}

// one batch of compiled plans: upload what is not resident, enqueue all launches (asynchronous)
int enqueue_batch(tb_ctx* ctx, tb_plan* const* plans, int64_t lo, int64_t hi, std::vector<int32_t>& status, bool single) {
    double t_u0 = now_ms();
    int rc = ensure_uploaded(ctx, plans + lo, hi - lo);
    if (rc) return rc;
    ctx->host_ms[1] += now_ms() - t_u0;
    std::vector<int64_t> groups[6];  // by tb_value_type
    for (int64_t i = lo; i < hi; ++i) {
        if (!plans[i]) continue;
        const int v = plans[i]->p.value_type;
        groups[(v >= 1 && v <= 5) ? v : TB_VALUE_F32].push_back(i);
    }
    for (int v : {(int)TB_VALUE_I32, (int)TB_VALUE_I16X2, (int)TB_VALUE_F32, (int)TB_VALUE_F64, (int)TB_VALUE_SIZE_CONFIG}) {
        rc = run_group(ctx, plans, groups[v], v, status, single);
        if (rc) return rc;
    }
    return TB_OK;
}

int begin_call(tb_ctx* ctx, int64_t n) {
    if (ctx->stream_open) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "a tb_stream is open on this context: finish it first");
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
#ifdef TB_KPROF
    {  // timeline of the most recent call only
        static TlRec* buf = nullptr;
        const unsigned cap = 4u << 20, zero = 0;
        if (!buf) {
            cudaMalloc(&buf, (size_t)cap * sizeof(TlRec));
            cudaMemcpyToSymbol(g_tl, &buf, sizeof buf);
            cudaMemcpyToSymbol(g_tl_cap, &cap, sizeof cap);
        }
        cudaDeviceSynchronize();
        cudaMemcpyToSymbol(g_tl_n, &zero, sizeof zero);
    }
#endif
    ctx->last_ms = 0;
    ctx->last_launches = 0;
    ctx->h2d_bytes = 0;
    ctx->d2h_bytes = 0;
    for (int q = 0; q < 4; ++q) {
        ctx->prof_ms[q] = 0;
        ctx->prof_launches[q] = 0;
        ctx->prof_union_ms[q] = 0;
    }
    ctx->prof_recs.clear();
    ctx->prof_used = 0;
    for (int q = 0; q < 6; ++q) ctx->host_ms[q] = 0;
    ctx->lane_cursor = 0;
    // descriptor chunks: when nothing is resident keep only the largest chunk (the device is idle between calls)
    {
        bool all_free = !ctx->chunks.empty();
        size_t best = 0;
        for (size_t q = 0; q < ctx->chunks.size(); ++q) {
            all_free = all_free && ctx->chunks[q].live == 0;
            if (ctx->chunks[q].cap > ctx->chunks[best].cap) best = q;
        }
        if (all_free && ctx->chunks.size() > 1) {
            BlobChunk keep = ctx->chunks[best];
            for (size_t q = 0; q < ctx->chunks.size(); ++q)
                if (q != best && ctx->chunks[q].d) cudaFree(ctx->chunks[q].d);
            keep.used = 0;
            ctx->chunks.assign(1, keep);
        }
    }
    int rc = ensure_results(ctx, (size_t)std::max<int64_t>(n, 1));
    if (rc) return rc;
    TB_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return TB_OK;
}

// wait for everything the call enqueued; device time and the per-launch profile of the call
int finish_sync(tb_ctx* ctx);

int finish_call(tb_ctx* ctx, tb_plan* const* plans, const double* r, int64_t n, const std::vector<int32_t>& status,
                double* out_values, int32_t* out_status, double* out_max, bool any) {
    if (any) {
        TB_CUDA(ctx, cudaMemcpyAsync(ctx->h_results, ctx->d_results, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d2h_bytes += (int64_t)n * 8;
    }
    int rc = finish_sync(ctx);
    if (rc) return rc;
    double mx = -std::numeric_limits<double>::infinity();
    int worst = TB_OK;
    for (int64_t i = 0; i < n; ++i) {
        double ri = r ? r[i] : 0.0;
        double v;
        if (!plans[i]) v = ri;  // empty graph: value is r (src/dynamic_ob.jl:39-40)
        else if (status[i] != TB_OK) {
            v = std::numeric_limits<double>::quiet_NaN();
            worst = status[i];
        } else v = ctx->h_results[i] + ri;
        out_values[i] = v;
        if (out_status) out_status[i] = status[i];
        if (status[i] == TB_OK && v > mx) mx = v;
    }
    if (out_max) *out_max = mx;
    if (worst != TB_OK) return set_err(ctx, worst, "one or more branches failed (see per-branch status): arena too small");
    return TB_OK;
}

int finish_sync(tb_ctx* ctx) {
    TB_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    int rc = sync_all_lanes(ctx);
    if (rc) return rc;
    float ms = 0;
    TB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->last_ms = ms;
    if (!ctx->prof_recs.empty()) {
        // per-launch CUDA events on the launching lanes: per kind the sum of the launch durations and the length of the
        // union of their [start, end] intervals (lanes overlap: the union is the time the kind was on the device)
        std::vector<std::pair<float, float>> iv[4];
        FILE* dump = nullptr;
        if (const char* path = getenv("TB_DUMP_LAUNCHES")) dump = fopen(path, "a");  // diagnostics: one record per launch
        for (const tb_ctx::ProfRec& pr : ctx->prof_recs) {
            float a = 0, b = 0;
            TB_CUDA(ctx, cudaEventElapsedTime(&a, ctx->ev0, pr.e0));
            TB_CUDA(ctx, cudaEventElapsedTime(&b, ctx->ev0, pr.e1));
            ctx->prof_ms[pr.kind] += b - a;
            ctx->prof_launches[pr.kind] += 1;
            iv[pr.kind].push_back({a, b});
            if (dump) fprintf(dump, "%d,%.4f,%.4f,%.6g,%.6g,%d,%u\n", pr.kind, a, b, pr.ops, pr.bytes, pr.n_insts, pr.grid);
        }
        if (dump) fclose(dump);
        for (int k = 0; k < 4; ++k) {
            std::sort(iv[k].begin(), iv[k].end());
            float end = -1, total = 0;
            for (const auto& x : iv[k]) {
                if (x.first > end) {
                    total += x.second - x.first;
                    end = x.second;
                } else if (x.second > end) {
                    total += x.second - end;
                    end = x.second;
                }
            }
            ctx->prof_union_ms[k] = total;
        }
        ctx->prof_recs.clear();
        ctx->prof_used = 0;
    }
    return TB_OK;
}

// Wave size of a call.  Heavy plans (cfg2: 2^27 ops each): waves of up to 128 plans, and a call with few plans (one
// rank's shard of a multi-GPU run) is cut into ~2 waves per lane (profiles/s03_wave_lane_sweep_cfg2.jsonl).  Light plans
// (cfg5: 2^22.5 ops each) are launch-bound: waves of up to 256, one per lane (profiles/s05_wave_sweep_cfg5_cfg2.jsonl:
// 1.78 ms instead of 2.16 ms for 1 024 plans).  mean_ops <= 0: unknown, treated as heavy.
int wave_for_call(tb_ctx* ctx, int64_t n, double mean_ops, bool resident = false) {
    const bool light = mean_ops > 0 && mean_ops < (double)(1 << 24);
    // dataflow: launch-bound calls -- light plans, or few plans that are not huge (a plan of >= 2^36 ops runs for
    // milliseconds per level: nothing to gain from dropping launches, and the level-synchronous GEMM instance is faster)
    ctx->dataflow = ctx->dataflow_mode == 1 ||
                    (ctx->dataflow_mode == 0 && (light || (n <= 2 * (int64_t)ctx->n_lanes && mean_ops < 68719476736.0)));
    const int64_t cfg = ctx->opts.max_wave > 0 ? ctx->opts.max_wave : (light ? 256 : 128);
    const int wpl = ctx->waves_per_lane > 0 ? ctx->waves_per_lane : (light ? 1 : 2);
    const int64_t parts = (int64_t)wpl * ctx->n_lanes;
    const int64_t per = (n + parts - 1) / parts;
    int64_t wave = std::min<int64_t>(cfg, std::max<int64_t>(16, per));
    // Resident light plans under the dataflow executor: their work lists are replayed from the device (no host pass to hide
    // behind other waves), and one persistent kernel over up to 1 024 plans keeps every CTA slot fed where four kernels
    // of 256 plans serialise pairwise (profiles/x4_listcache_wave_sweep.txt: cfg5 1.00 ms -> 0.91 ms)
    static const bool wave_forced = getenv("TB_WAVE") != nullptr || getenv("TB_WAVES_PER_LANE") != nullptr;
    if (resident && light && ctx->dataflow && ctx->opts.max_wave <= 0 && !wave_forced) wave = std::min<int64_t>(1024, std::max<int64_t>(16, n));
    ctx->call_lanes = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->n_lanes, (n + wave - 1) / wave));
    return (int)wave;
}

double mean_plan_ops(tb_plan* const* plans, int64_t lo, int64_t hi) {
    double sum = 0;
    int64_t cnt = 0;
    for (int64_t i = lo; i < hi; ++i)
        if (plans[i]) {
            sum += plans[i]->p.stats.ops;
            ++cnt;
        }
    return cnt ? sum / (double)cnt : 0.0;
}

int multi_contract_batch(tb_ctx* ctx, tb_plan* const* plans, const double* r, int64_t n, double* out_values, int32_t* out_status,
                         double* out_max);

// the asynchronous half of tb_contract_batch on ONE device (resident plans): upload what is missing, enqueue every launch
int batch_enqueue(tb_ctx* ctx, tb_plan* const* plans, int64_t n, int64_t n_results, std::vector<int32_t>& status, bool single) {
    int rc = begin_call(ctx, n_results);
    if (rc) return rc;
    if (ctx->prefill_results) {
        const unsigned blocks = (unsigned)((n_results + 255) / 256);
        k_fill_double<<<blocks, 256, 0, ctx->stream>>>(ctx->d_results, n_results, -std::numeric_limits<double>::infinity());
    }
    status.assign((size_t)n, TB_OK);
    const int64_t wave = wave_for_call(ctx, n, mean_plan_ops(plans, 0, n), true);
    ctx->call_wave = (int)wave;
    int64_t batch = single ? n : wave;
    for (int64_t lo = 0; lo < n && rc == TB_OK;) {
        const int64_t hi = std::min(n, lo + std::max<int64_t>(batch, 1));
        rc = enqueue_batch(ctx, plans, lo, hi, status, single);
        lo = hi;
        batch = std::min<int64_t>(batch * 2, wave * ctx->n_lanes * 2);
    }
    return rc;
}

int contract_impl(tb_ctx* ctx, tb_plan* const* plans, const double* r, int64_t n, double* out_values, int32_t* out_status,
                  double* out_max, bool single) {
    if (!ctx) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "ctx is NULL");
    if (n < 0 || (n > 0 && (!plans || !out_values))) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "bad arguments");
    if (!ctx->subs.empty()) {
        if (single) return contract_impl(ctx->subs[0], plans, r, n, out_values, out_status, out_max, true);
        return multi_contract_batch(ctx, plans, r, n, out_values, out_status, out_max);
    }
    std::vector<int32_t> status;
    bool any = false;
    for (int64_t i = 0; i < n; ++i) any = any || plans[i];
    int rc = batch_enqueue(ctx, plans, n, n, status, single);
    if (rc) {
        sync_all_lanes(ctx);
        return rc;
    }
    return finish_call(ctx, plans, r, n, status, out_values, out_status, out_max, any);
}

// a copy of `base` (compiled with fixed labels) for another assignment of the same labels: the descriptors are shared
// by value, only the leaf-pool words of the sliced leaves change; the copy is not resident anywhere
tb_plan* clone_for_assignment(const tb_plan* base, const uint8_t* values) {
    tb_plan* p = new tb_plan(*base);
    p->p.owner = nullptr;
    p->p.d_blob = nullptr;
    p->p.upload_uid = 0;
    p->p.res_prev = p->p.res_next = nullptr;
    p->p.blob_bytes = 0;
    p->p.assign(values);
    return p;
}

// the temporary plans of a *_networks / stream call all live in chunks of that call: release the device side once,
// free the host side off the caller's critical path (joined by the next call / tb_shutdown)
void release_temporary_plans(tb_ctx* ctx, std::vector<tb_plan*> plans, bool to_pool = false) {
    for (tb_plan* p : plans)
        if (p && p->p.d_blob) {
            for (auto& c : ctx->chunks)
                if ((uint8_t*)p->p.d_blob >= (uint8_t*)c.d && (uint8_t*)p->p.d_blob < (uint8_t*)c.d + c.cap) {
                    if (--c.live == 0) c.used = 0;
                    break;
                }
            unlink_resident(ctx, p->p);
            p->p.d_blob = nullptr;
            p->p.owner = nullptr;
        }
    if (to_pool) {  // compact temporaries (compile_temporary): the next call compiles into the same objects
        std::lock_guard<std::mutex> lk(ctx->pool_mu);
        size_t kept = 0;
        for (tb_plan*& p : plans) {
            if (!p) {
                ++kept;
                continue;
            }
            const size_t bytes = p->p.descriptor_capacity_bytes();
            if (ctx->plan_pool.size() < tb_ctx::kPlanPoolMax && ctx->plan_pool_bytes + bytes <= tb_ctx::kPlanPoolMaxBytes) {
                ctx->plan_pool.push_back(p);
                ctx->plan_pool_bytes += bytes;
                p = nullptr;
                ++kept;
            }
        }
        if (kept == plans.size()) return;
    }
    if (ctx->reaper.joinable()) ctx->reaper.join();
    ctx->reaper = std::thread([dead = std::move(plans)]() {
        for (tb_plan* p : dead) delete p;
    });
}

// 2^rank elements of the last single-plan contraction (tb_contract) at arena element offset `off` -> host doubles
int copy_tensor_to_host(tb_ctx* ctx, const Plan& P, int64_t off, int64_t n, double* out_data) {
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    double* d_tmp = nullptr;
    TB_CUDA(ctx, cudaMalloc(&d_tmp, (size_t)n * sizeof(double)));
    const uint8_t* src = (const uint8_t*)ctx->arena + (size_t)(ctx->last_plan_arena_base_elems + off) * (size_t)P.elem_size();
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (P.value_type == TB_VALUE_I32) k_to_double<int32_t><<<blocks, 256, 0, ctx->stream>>>((const int32_t*)src, d_tmp, n);
    else if (P.value_type == TB_VALUE_F64) k_to_double<double><<<blocks, 256, 0, ctx->stream>>>((const double*)src, d_tmp, n);
    else if (P.value_type == TB_VALUE_SIZE_CONFIG) k_to_double<long long><<<blocks, 256, 0, ctx->stream>>>((const long long*)src, d_tmp, n);
    else if (P.value_type == TB_VALUE_I16X2) k_to_double<int16_t><<<blocks, 256, 0, ctx->stream>>>((const int16_t*)src, d_tmp, n);
    else k_to_double<float><<<blocks, 256, 0, ctx->stream>>>((const float*)src, d_tmp, n);
    cudaError_t e = cudaMemcpyAsync(out_data, d_tmp, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_tmp);
    if (e != cudaSuccess) return set_err(ctx, TB_ERR_CUDA, cudaGetErrorString(e));
    return TB_OK;
}

}  // namespace

// streaming hand-off: branches are pushed as the host's slicer finishes them
struct tb_stream {
    tb_ctx* ctx = nullptr;
    int64_t capacity = 0;
    std::vector<tb_plan*> plans;
    std::vector<double> r;
    std::vector<int32_t> status;
    bool any = false;
    bool failed = false;
    double t0 = 0;
};

// no C++ exception crosses the ABI: every entry point that allocates is a function-try-block ending in TB_CATCH
int exception_to_status(tb_ctx* ctx) {
    try {
        throw;
    } catch (const std::bad_alloc&) {
        return set_err(ctx, TB_ERR_OUT_OF_MEMORY, "host memory allocation failed");
    } catch (const std::exception& e) {
        return set_err(ctx, TB_ERR_INTERNAL, std::string("C++ exception: ") + e.what());
    } catch (...) {
        return set_err(ctx, TB_ERR_INTERNAL, "unknown C++ exception");
    }
}
#define TB_CATCH(ctx_expr) \
    catch (...) { return exception_to_status(ctx_expr); }

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char* tb_version(void) { return "tbcuda 0.1.0 (sm_100a)"; }

const char* tb_last_error(const tb_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_tls_error.c_str(); }

int tb_init(const tb_options* opts, tb_ctx** out_ctx) try {
    if (!out_ctx) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "out_ctx is NULL");
    *out_ctx = nullptr;
    if (opts && opts->n_devices > 1) return tb_init_multi(opts->devices, opts->n_devices, opts, out_ctx);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(nullptr, TB_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libtbcuda has no CPU fallback");
    std::unique_ptr<tb_ctx> ctx(new tb_ctx());
    if (opts) ctx->opts = *opts;
    ctx->opts.devices = nullptr;  // read during init only
    ctx->device = ctx->opts.device;
    if (ctx->device < 0 || ctx->device >= ndev) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "device ordinal out of range");
    tb_ctx* c = ctx.get();
    TB_CUDA(nullptr, cudaSetDevice(c->device));
    cudaDeviceProp prop{};
    TB_CUDA(nullptr, cudaGetDeviceProperties(&prop, c->device));
    if (prop.major < 10)
        return set_err(nullptr, TB_ERR_CUDA, "device compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor) + " < 10.0: libtbcuda is built for sm_100a only");
    c->sm_count = prop.multiProcessorCount;
    TB_CUDA(nullptr, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    TB_CUDA(nullptr, cudaEventCreate(&c->ev0));
    TB_CUDA(nullptr, cudaEventCreate(&c->ev1));
    {
        const char* e1 = getenv("TB_GEMM_V1");
        c->gemm_v1 = e1 && e1[0] == '1';
        const char* e3 = getenv("TB_EPI_DIRECT");
        c->staged_epilogue = !(e3 && e3[0] == '1');
        const char* e4 = getenv("TB_WAVE");
        if (e4 && atoi(e4) >= 1 && c->opts.max_wave == 0) c->opts.max_wave = atoi(e4);
        const char* e5 = getenv("TB_WAVES_PER_LANE");
        if (e5 && atoi(e5) >= 1) c->waves_per_lane = atoi(e5);
        const char* e6 = getenv("TB_LEVEL_SYNC");  // A/B testing: one launch per dependency level and kernel kind (the round-1 executor)
        const char* e7 = getenv("TB_DATAFLOW");
        c->dataflow_mode = ((e6 && e6[0] == '1') || c->gemm_v1) ? 2 : ((e7 && e7[0] == '1') ? 1 : 0);
        if (c->opts.streams_per_device >= 1) c->n_lanes = std::min(c->opts.streams_per_device, (int)tb_ctx::kMaxLanes);
        const char* e2 = getenv("TB_LANES");
        if (e2 && atoi(e2) >= 1) c->n_lanes = std::min(atoi(e2), (int)tb_ctx::kMaxLanes);
        c->profile_mode = std::max(0, std::min(2, c->opts.timing));
    }
    TB_CUDA(nullptr, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    TB_CUDA(nullptr, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    TB_CUDA(nullptr, cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    for (int l = 1; l < c->n_lanes; ++l) {
        TB_CUDA(nullptr, cudaStreamCreateWithFlags(&c->side[l], cudaStreamNonBlocking));
        TB_CUDA(nullptr, cudaEventCreateWithFlags(&c->ev_join[l], cudaEventDisableTiming));
    }
    TB_CUDA(nullptr, cudaFuncSetAttribute(k_fused_subtrees<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    TB_CUDA(nullptr, cudaFuncSetAttribute(k_fused_subtrees<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    TB_CUDA(nullptr, cudaFuncSetAttribute(k_fused_subtrees<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    TB_CUDA(nullptr, cudaFuncSetAttribute(k_fused_subtrees<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    TB_CUDA(nullptr, cudaFuncSetAttribute(k_fused_subtrees<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFusedSmemMax));
    TB_CUDA(nullptr, cudaFuncSetAttribute(k_gemm<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    TB_CUDA(nullptr, cudaFuncSetAttribute(k_gemm<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    {
        const void* g2[] = {(const void*)k_gemm2<int32_t, false>, (const void*)k_gemm2<int32_t, true>, (const void*)k_gemm2<float, false>,
                            (const void*)k_gemm2<float, true>, (const void*)k_gemm2h<false>, (const void*)k_gemm2h<true>};
        for (const void* fn : g2) {
            TB_CUDA(nullptr, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
            TB_CUDA(nullptr, cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        }
    }
    {
        int nb = 0;
        TB_CUDA(nullptr, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gemm2<int32_t, true>, G2_THREADS, G2_SMEM_BYTES));
        c->gemm2_ctas_per_sm = nb;
        const char* eg = getenv("TB_GEMM_CTAS");  // persistent GEMM CTAs per SM and launch (experiments: 1 lets two lanes' GEMMs co-run)
        if (eg && atoi(eg) >= 1) c->gemm2_ctas_per_sm = std::min(nb, atoi(eg));
    }
    *out_ctx = ctx.release();
    return TB_OK;
} TB_CATCH(nullptr)

int tb_init_multi(const int32_t* devices, int32_t n_devices, const tb_options* opts, tb_ctx** out_ctx) try {
    if (!out_ctx) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "out_ctx is NULL");
    *out_ctx = nullptr;
    if (n_devices < 1 || n_devices > 64) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "n_devices must be in [1, 64]");
    std::unique_ptr<tb_ctx> ctx(new tb_ctx());
    if (opts) ctx->opts = *opts;
    ctx->opts.n_devices = n_devices;
    ctx->opts.devices = nullptr;  // read during init only
    std::vector<int> devs((size_t)n_devices);
    for (int d = 0; d < n_devices; ++d) devs[(size_t)d] = devices ? devices[d] : d;
    for (int d = 0; d < n_devices; ++d)
        for (int e = 0; e < d; ++e)
            if (devs[(size_t)d] == devs[(size_t)e]) ctx->host_combine = true;
    ctx->device = devs[0];
    const int total_threads = ctx->opts.host_threads > 0 ? ctx->opts.host_threads : std::max(1, (int)std::thread::hardware_concurrency());
    auto fail = [&](int rc) {
        for (tb_ctx* sub : ctx->subs) tb_shutdown(sub);
        ctx->subs.clear();
        return rc;
    };
    for (int d = 0; d < n_devices; ++d) {
        tb_options o = ctx->opts;
        o.device = devs[(size_t)d];
        o.n_devices = 0;
        o.devices = nullptr;
        o.host_threads = std::max(1, total_threads / n_devices);
        tb_ctx* sub = nullptr;
        int rc = tb_init(&o, &sub);
        if (rc) return fail(rc);
        sub->parent = ctx.get();
        ctx->subs.push_back(sub);
    }
    ctx->profile_mode = ctx->subs[0]->profile_mode;
    if (n_devices > 1 && !ctx->host_combine) {
        NcclApi* nc = nccl_api();
        if (!nc->lib || !nc->err.empty()) return fail(set_err(nullptr, TB_ERR_NCCL, nc->err.empty() ? "NCCL is not available" : nc->err));
        ctx->comms.assign((size_t)n_devices, nullptr);
        int e = nc->CommInitAll(ctx->comms.data(), n_devices, devs.data());
        if (e != 0) {
            ctx->comms.clear();
            return fail(set_err(nullptr, TB_ERR_NCCL, std::string("ncclCommInitAll: ") + nc->GetErrorString(e)));
        }
    }
    *out_ctx = ctx.release();
    return TB_OK;
} TB_CATCH(nullptr)

int tb_device_count(const tb_ctx* ctx) {
    if (!ctx) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "ctx is NULL");
    return ctx->subs.empty() ? 1 : (int)ctx->subs.size();
}

int tb_estimate(const tb_network* net, double* out_ops, double* out_sc) try {
    if (!net || !out_ops) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "net / out_ops is NULL");
    if (net->n_leaves == 0) {  // empty graph: nothing to contract
        *out_ops = 0;
        if (out_sc) *out_sc = 0;
        return TB_OK;
    }
    Plan P;
    std::string err;
    int rc = compile_plan(*net, TB_PLAN_ESTIMATE_ONLY, P, err);
    if (rc) return set_err(nullptr, rc, err);
    *out_ops = P.stats.ops;
    if (out_sc) *out_sc = P.stats.sc;
    return TB_OK;
} TB_CATCH(nullptr)

int tb_estimate_many(const tb_network* nets, int64_t n, int32_t threads, double* out_ops, double* out_sc) try {
    if (n < 0 || (n > 0 && (!nets || !out_ops))) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "nets / out_ops is NULL");
    int nt = threads > 0 ? threads : std::max(1, (int)std::thread::hardware_concurrency());
    nt = std::max(1, std::min<int>(nt, (int)std::max<int64_t>(1, n / 16)));
    std::atomic<int64_t> next{0};
    std::atomic<int> bad{TB_OK};
    std::mutex mu;
    std::string first_err;
    auto worker = [&] {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            double ops = 0, sc = 0;
            if (nets[i].n_leaves != 0) {
                Plan P;
                std::string err;
                const int rc = compile_guarded(nets[i], TB_PLAN_ESTIMATE_ONLY, P, err);
                if (rc) {
                    std::lock_guard<std::mutex> g(mu);
                    if (bad.load() == TB_OK) {
                        bad.store(rc);
                        first_err = "branch " + std::to_string(i) + ": " + err;
                    }
                    continue;
                }
                ops = P.stats.ops;
                sc = P.stats.sc;
            }
            out_ops[i] = ops;
            if (out_sc) out_sc[i] = sc;
        }
    };
    ThreadGroup th;
    for (int t = 0; t < nt - 1; ++t) th.spawn(worker);
    worker();
    th.join();
    if (bad.load() != TB_OK) return set_err(nullptr, bad.load(), first_err);
    return TB_OK;
} TB_CATCH(nullptr)

int tb_shutdown(tb_ctx* ctx) {
    if (!ctx) return TB_OK;
    if (!ctx->subs.empty()) {  // multi-GPU context: communicators first, then every device's own context
        if (!ctx->comms.empty()) {
            NcclApi* nc = nccl_api();
            for (void* c : ctx->comms)
                if (c) nc->CommDestroy(c);
        }
        for (tb_ctx* sub : ctx->subs) tb_shutdown(sub);
        delete ctx;
        return TB_OK;
    }
    if (ctx->reaper.joinable()) ctx->reaper.join();
    for (tb_plan* p : ctx->plan_pool) delete p;
    ctx->plan_pool.clear();
#ifdef TB_KPROF
    {
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        unsigned long long h[8] = {0};
        cudaMemcpyFromSymbol(h, g_kprof, sizeof h);
        if (const char* path = getenv("TB_TL_DUMP")) {
            unsigned n = 0, cap = 0;
            TlRec* buf = nullptr;
            cudaMemcpyFromSymbol(&n, g_tl_n, sizeof n);
            cudaMemcpyFromSymbol(&cap, g_tl_cap, sizeof cap);
            cudaMemcpyFromSymbol(&buf, g_tl, sizeof buf);
            n = std::min(n, cap);
            std::vector<TlRec> hrec(n);
            if (n) cudaMemcpy(hrec.data(), buf, (size_t)n * sizeof(TlRec), cudaMemcpyDeviceToHost);
            if (FILE* f = fopen(path, "wb")) {
                fwrite(hrec.data(), sizeof(TlRec), n, f);
                fclose(f);
            }
            fprintf(stderr, "KPROF timeline: %u CTA records -> %s\n", n, path);
        }
        fprintf(stderr, "KPROF k_gemm2h consumer cycles: wait_tile %.3e wait_data %.3e main %.3e epilogue %.3e lifetime %.3e tiles %llu\n",
                (double)h[0], (double)h[1], (double)h[2], (double)h[3], (double)h[4], h[5]);
    }
#endif
    cudaSetDevice(ctx->device);
    sync_all_lanes(ctx);
    // plans may outlive the context (a host language's GC decides when they are destroyed): detach them
    for (Plan* P = ctx->resident; P;) {
        Plan* nx = P->res_next;
        P->owner = nullptr;
        P->d_blob = nullptr;
        P->res_prev = P->res_next = nullptr;
        P = nx;
    }
    ctx->resident = nullptr;
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->permute_buf) cudaFree(ctx->permute_buf);
    for (auto& e : ctx->list_cache) {
        if (e.d_pristine) cudaFree(e.d_pristine);
        if (e.d_work) cudaFree(e.d_work);
    }
    if (ctx->table_buf) cudaFree(ctx->table_buf);
    if (ctx->table_out) cudaFree(ctx->table_out);
    for (auto& c : ctx->chunks)
        if (c.d) cudaFree(c.d);
    for (auto& sl : ctx->slots) {
        if (sl.h) cudaFreeHost(sl.h);
        if (sl.d) cudaFree(sl.d);
        if (sl.ev) cudaEventDestroy(sl.ev);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->d_results) cudaFree(ctx->d_results);
    if (ctx->h_results) cudaFreeHost(ctx->h_results);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (int l = 1; l < tb_ctx::kMaxLanes; ++l) {
        if (ctx->side[l]) cudaStreamDestroy(ctx->side[l]);
        if (ctx->ev_join[l]) cudaEventDestroy(ctx->ev_join[l]);
    }
    if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TB_OK;
}

int tb_plan_create(tb_ctx* ctx, const tb_network* net, tb_plan** out_plan) try {
    if (!net || !out_plan) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "net / out_plan is NULL");
    *out_plan = nullptr;
    std::unique_ptr<tb_plan> p(new tb_plan());
    std::string err;
    int rc = compile_plan(*net, ctx ? ctx->opts.plan_flags : 0u, p->p, err);
    if (rc) return set_err(ctx, rc, err);
    *out_plan = p.release();
    return TB_OK;
} TB_CATCH(ctx)

int tb_plan_destroy(tb_plan* plan) {
    if (!plan) return TB_OK;
    if (plan->p.owner && plan->p.d_blob) {
        tb_ctx* ctx = plan->p.owner;
        unlink_resident(ctx, plan->p);
        for (auto& c : ctx->chunks) {
            if ((uint8_t*)plan->p.d_blob >= (uint8_t*)c.d && (uint8_t*)plan->p.d_blob < (uint8_t*)c.d + c.cap) {
                if (--c.live == 0 && &c != &ctx->chunks.back()) {
                    cudaSetDevice(ctx->device);
                    cudaStreamSynchronize(ctx->stream);
                    cudaFree(c.d);
                    c.d = nullptr;
                    c.cap = c.used = 0;
                } else if (c.live == 0) {
                    c.used = 0;  // recycle the open chunk
                }
                break;
            }
        }
    }
    delete plan;
    return TB_OK;
}

int tb_plan_info(const tb_plan* plan, tb_plan_stats* out) {
    if (!plan || !out) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "plan / out is NULL");
    *out = plan->p.stats;
    return TB_OK;
}

int tb_plan_export(const tb_plan* plan, tb_step_info* out, int32_t cap) try {
    if (!plan) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "plan is NULL");
    int n = (int)plan->p.recs.size();
    if (out)
        for (int i = 0; i < n && i < cap; ++i) out[i] = plan->p.step_info((size_t)i);
    return n;
} TB_CATCH(nullptr)

int64_t tb_plan_export_raw(const tb_plan* plan, int32_t which, void* out, int64_t cap) try {
    if (!plan) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "plan is NULL");
    const Plan& P = plan->p;
    const void* src = nullptr;
    int64_t bytes = 0;
    int64_t header[4] = {P.arena_elems, P.root_off, P.n_levels, P.value_type};
    std::vector<uint8_t> pool_dev(P.pool_bytes() + 16);
    switch (which) {
        case 0: P.write_pool(pool_dev.data()); src = pool_dev.data(); bytes = (int64_t)P.pool_bytes(); break;
        case 1: src = P.sub_steps.data(); bytes = (int64_t)(P.sub_steps.size() * sizeof(SubStep)); break;
        case 2: src = P.subtrees.data(); bytes = (int64_t)(P.subtrees.size() * sizeof(SubTree)); break;
        case 3: src = P.big_steps.data(); bytes = (int64_t)(P.big_steps.size() * sizeof(BigStep)); break;
        case 4: src = P.big_level_begin.data(); bytes = (int64_t)P.big_level_begin.size() * 4; break;
        case 5: src = header; bytes = sizeof header; break;
        case 6: src = P.big_log2_ops.data(); bytes = (int64_t)P.big_log2_ops.size() * 4; break;
        case 7: src = P.big_bytes.data(); bytes = (int64_t)P.big_bytes.size() * 8; break;
        default: return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "unknown section");
    }
    if (out && cap > 0 && bytes > 0) std::memcpy(out, src, (size_t)std::min(cap, bytes));
    return bytes;
} TB_CATCH(nullptr)

int tb_contract(tb_ctx* ctx, tb_plan* plan, double* out_value) try {
    if (!plan || !out_value) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "plan / out_value is NULL");
    tb_plan* arr[1] = {plan};
    return contract_impl(ctx, arr, nullptr, 1, out_value, nullptr, nullptr, true);
} TB_CATCH(ctx)

int tb_contract_batch(tb_ctx* ctx, tb_plan* const* plans, const double* r, int64_t n, double* out_values, int32_t* out_status,
                      double* out_max) try {
    return contract_impl(ctx, plans, r, n, out_values, out_status, out_max, false);
} TB_CATCH(ctx)

}  // extern "C"

namespace {

// the asynchronous half of tb_contract_networks on ONE device: compile (worker threads), upload, enqueue every launch.
// n_results = length of the call's result vector (> n for a sub-context of a multi-GPU call, which scatters through
// ctx->index_map into a vector pre-filled with -inf).  The caller finishes with finish_call / finish_sync and
// release_temporary_plans(cs.plans).
struct NetworksCall {
    std::vector<tb_plan*> plans;
    std::vector<int32_t> status;
    bool any = false;
    double t_c0 = 0;
};

// TB_TRACE_CALL=1 (diagnostics): per pipeline batch, when the host had it compiled / enqueued and when the device finished
// it, printed to stderr after the call (device times from events on every lane, relative to the start of the call)
struct CallTrace {
    struct Row {
        int64_t lo, hi;
        double t_ready, t_enqueued;
        double upload_ms, lists_ms, launch_ms;  // cumulative over the call
        std::vector<cudaEvent_t> ev;
    };
    bool on = false;
    cudaEvent_t ev0 = nullptr;
    std::vector<Row> rows;
    static bool enabled() {
        static const bool e = getenv("TB_TRACE_CALL") != nullptr;
        return e;
    }
    void begin(tb_ctx* ctx) {
        on = enabled();
        ctx->trace_on = on;
        ctx->trace_t0 = now_ms();
        ctx->trace_notes.clear();
        if (!on) return;
        cudaEventCreate(&ev0);
        cudaEventRecord(ev0, ctx->stream);
    }
    void batch(tb_ctx* ctx, int64_t lo, int64_t hi, double t_ready, double t_enq) {
        if (!on) return;
        Row r{lo, hi, t_ready, t_enq, ctx->host_ms[1], ctx->host_ms[2], ctx->host_ms[3], {}};
        for (int l = 0; l < ctx->n_lanes; ++l) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, l == 0 ? ctx->stream : ctx->side[l]);
            r.ev.push_back(e);
        }
        rows.push_back(std::move(r));
    }
    void dump(tb_ctx* ctx, double t_host_done) {
        if (!on) return;
        cudaDeviceSynchronize();
        fprintf(stderr, "[tb trace] device %d: host done at %.3f ms\n", ctx->device, t_host_done);
        for (Row& r : rows) {
            float dev_done = 0;
            for (cudaEvent_t e : r.ev) {
                float ms = 0;
                cudaEventElapsedTime(&ms, ev0, e);
                dev_done = std::max(dev_done, ms);
                cudaEventDestroy(e);
            }
            fprintf(stderr, "[tb trace] plans %5lld..%5lld  compiled %7.3f  enqueued %7.3f  device done %7.3f ms   (cumulative upload %.2f lists %.2f launch %.2f)\n",
                    (long long)r.lo, (long long)r.hi, r.t_ready, r.t_enqueued, (double)dev_done, r.upload_ms, r.lists_ms, r.launch_ms);
        }
        for (const std::string& note : ctx->trace_notes) fprintf(stderr, "[tb trace] stall: %s\n", note.c_str());
        ctx->trace_notes.clear();
        ctx->trace_on = false;
        cudaEventDestroy(ev0);
        rows.clear();
        on = false;
    }
};

int networks_enqueue(tb_ctx* ctx, const tb_network* nets, int64_t n, int64_t n_results, NetworksCall& cs) {
    cs.t_c0 = now_ms();
    const double t_c0 = cs.t_c0;
    (void)t_c0;
    int rc = begin_call(ctx, n_results);
    if (rc) return rc;
    if (ctx->prefill_results) {  // multi-GPU call: entries of other devices (and of empty / failed branches) stay -inf for the max-reduce
        const unsigned blocks = (unsigned)((n_results + 255) / 256);
        k_fill_double<<<blocks, 256, 0, ctx->stream>>>(ctx->d_results, n_results, -std::numeric_limits<double>::infinity());
    }
    CallTrace trace;
    trace.begin(ctx);
    std::vector<tb_plan*>& plans = cs.plans;
    plans.assign((size_t)n, nullptr);
    std::vector<int> codes((size_t)n, TB_OK);
    std::vector<std::string> errs((size_t)n);
    std::unique_ptr<std::atomic<uint8_t>[]> done(new std::atomic<uint8_t>[(size_t)std::max<int64_t>(n, 1)]);
    for (int64_t i = 0; i < n; ++i) done[i].store(0, std::memory_order_relaxed);
    int nthreads = ctx->opts.host_threads > 0 ? ctx->opts.host_threads : (int)std::thread::hardware_concurrency();
    nthreads = std::max(1, std::min<int>(nthreads, (int)std::max<int64_t>(1, n / 8)));
    std::atomic<int64_t> next{0};
    std::atomic<bool> abort{false};
    const uint32_t flags = ctx->opts.plan_flags | TB_PLAN_TEMPORARY;
    // plan compilation runs on worker threads, in index order; the calling thread uploads and launches batch b
    // while the workers are already compiling batch b+1 (the GPU meanwhile executes batch b-1)
    std::atomic<bool> enqueued_all{false};
    auto worker = [&]() {
        for (;;) {
            while (ctx->helper.help_once()) {
            }
            int64_t i = next.fetch_add(1);
            if (i >= n) break;
            if (!abort.load(std::memory_order_relaxed) && nets[i].n_leaves != 0) {
                plans[i] = compile_temporary(ctx, nets[i], flags, codes[i], errs[i]);
                if (codes[i]) abort.store(true);
            }
            done[i].store(1, std::memory_order_release);
        }
        // nothing left to compile: keep helping the launching thread until the last batch is enqueued
        while (!enqueued_all.load(std::memory_order_acquire))
            if (!ctx->helper.help_once()) std::this_thread::yield();
    };
    ThreadGroup th;
    struct HelpersScope {  // the helper loop is live exactly while the workers run
        tb_ctx* c;
        std::atomic<bool>& all;
        ~HelpersScope() {
            all.store(true, std::memory_order_release);
            c->helper.has_helpers = false;
        }
    } helpers_scope{ctx, enqueued_all};
    ctx->helper.has_helpers = nthreads > 1;
    for (int t = 0; t < nthreads - 1; ++t) th.spawn(worker);
    ctx->call_wave = wave_for_call(ctx, n, 0.0);  // refined from the first compiled batch below
    // TB_BATCH_MAX (experiments): upper bound of a pipeline batch; smaller batches keep the GPU closer behind the
    // compiler threads (shorter tail after the last plan is compiled) at the price of smaller waves
    static const int64_t batch_cap = [] {
        const char* e = getenv("TB_BATCH_MAX");
        return e && atoll(e) >= 16 ? (int64_t)atoll(e) : (int64_t)0;
    }();
    // two waves per batch (profiles/s06_e2e_batch_sweep_cfg2.jsonl: 34.6 ms with 256-plan batches, 37.6 ms with 512)
    int64_t batch_max = batch_cap ? batch_cap : std::max<int64_t>(256, (int64_t)ctx->call_wave * ctx->n_lanes / 2);
    std::vector<int32_t>& status = cs.status;
    status.assign((size_t)n, TB_OK);
    bool& any = cs.any;
    double t_wait = 0;
    // small first batches: the GPU starts after ~64 compiled plans instead of a full batch (pipeline fill)
    int64_t batch = std::min<int64_t>(batch_max, 64);
    for (int64_t lo = 0, hi = 0; lo < n && rc == TB_OK; lo = hi, batch = std::min(batch_max, batch * 2)) {
        hi = std::min(n, lo + batch);
        const double tw0 = now_ms();
        if (nthreads == 1) {
            while (next.load() < hi && next.load() < n) {  // single-threaded: compile this batch inline
                int64_t i = next.fetch_add(1);
                if (i >= n) break;
                if (nets[i].n_leaves != 0) plans[i] = compile_temporary(ctx, nets[i], flags, codes[i], errs[i]);
                done[i].store(1, std::memory_order_release);
            }
        }
        for (int64_t i = lo; i < hi; ++i)
            while (!done[i].load(std::memory_order_acquire)) {
                // the launching thread compiles too while it would otherwise wait
                int64_t j = next.fetch_add(1);
                if (j < n) {
                    if (!abort.load(std::memory_order_relaxed) && nets[j].n_leaves != 0) {
                        plans[j] = compile_temporary(ctx, nets[j], flags, codes[j], errs[j]);
                        if (codes[j]) abort.store(true);
                    }
                    done[j].store(1, std::memory_order_release);
                } else {
                    std::this_thread::yield();
                }
            }
        t_wait += now_ms() - tw0;
        const double t_ready = now_ms() - t_c0;
        for (int64_t i = lo; i < hi; ++i) {
            if (codes[i]) {
                rc = set_err(ctx, codes[i], "branch " + std::to_string(i) + ": " + errs[i]);
                break;
            }
            any = any || plans[i];
        }
        if (rc == TB_OK && lo == 0) {  // the plans' weight is known now: light plans get larger waves
            ctx->call_wave = wave_for_call(ctx, n, mean_plan_ops(plans.data(), lo, hi));
            batch_max = batch_cap ? batch_cap : std::max<int64_t>(256, (int64_t)ctx->call_wave * ctx->n_lanes / 2);
        }
        if (rc == TB_OK) rc = enqueue_batch(ctx, plans.data(), lo, hi, status, false);
        trace.batch(ctx, lo, hi, t_ready, now_ms() - t_c0);
    }
    abort.store(rc != TB_OK);
    enqueued_all.store(true, std::memory_order_release);
    th.join();
    trace.dump(ctx, now_ms() - t_c0);
    ctx->host_ms[0] = t_wait;  // time the launching thread spent waiting for the compiler threads
    return rc;
}

int multi_contract_networks(tb_ctx* ctx, const tb_network* nets, const double* r, int64_t n, double* out_values,
                            int32_t* out_status, double* out_max);


// ================================================================================================
// Multi-GPU inside the library (component C1): ONE process, one sub-context per device with its own
// arena / streams / worker threads, branches dealt longest-first (LPT) by estimated cost, and ONE
// ncclAllReduce(ncclMax) over the per-branch result vector -- pre-filled with -inf on every device -- in place of
// maximum(res) (/root/reference/src/dynamic_ob.jl:27) and of any per-branch gather.  No tensor crosses NVLink.
// NCCL is resolved with dlopen at tb_init_multi time (libnccl.so.2), so single-GPU users need no NCCL at all.
// ================================================================================================
// longest-processing-time-first: units sorted by cost (descending, stable), each to the least-loaded device.
// `load` may carry work that is already pinned to a device.
void lpt_assign(const std::vector<double>& cost, const std::vector<int64_t>& units, std::vector<double>& load, std::vector<int>& owner) {
    std::vector<int64_t> ord(units);
    std::stable_sort(ord.begin(), ord.end(), [&](int64_t a, int64_t b) { return cost[(size_t)a] > cost[(size_t)b]; });
    for (int64_t i : ord) {
        int best = 0;
        for (int d = 1; d < (int)load.size(); ++d)
            if (load[(size_t)d] < load[(size_t)best]) best = d;
        owner[(size_t)i] = best;
        load[(size_t)best] += cost[(size_t)i];
    }
}

int host_threads_of(const tb_ctx* ctx) {
    return ctx->opts.host_threads > 0 ? ctx->opts.host_threads : std::max(1, (int)std::thread::hardware_concurrency());
}

// after every device thread has enqueued its share: combine the devices' result vectors (length n, -inf where a device
// has no entry) into the host vector of sub-context 0, then wait for all devices
int multi_combine(tb_ctx* ctx, int64_t n) {
    const int D = (int)ctx->subs.size();
    tb_ctx* s0 = ctx->subs[0];
    if (n > 0 && D > 1 && !ctx->host_combine) {
        NcclApi* nc = nccl_api();
        TB_NCCL(ctx, nc->GroupStart());
        for (int d = 0; d < D; ++d) {
            tb_ctx* sub = ctx->subs[(size_t)d];
            int e = nc->AllReduce(sub->d_results, sub->d_results, (size_t)n, kNcclFloat64, kNcclMax, ctx->comms[(size_t)d], sub->stream);
            if (e != 0) {
                nc->GroupEnd();
                return set_err(ctx, TB_ERR_NCCL, std::string("ncclAllReduce: ") + nc->GetErrorString(e));
            }
        }
        TB_NCCL(ctx, nc->GroupEnd());
    }
    const int n_read = (D > 1 && ctx->host_combine) ? D : 1;
    for (int d = 0; d < n_read && n > 0; ++d) {
        tb_ctx* sub = ctx->subs[(size_t)d];
        TB_CUDA(ctx, cudaSetDevice(sub->device));
        TB_CUDA(ctx, cudaMemcpyAsync(sub->h_results, sub->d_results, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, sub->stream));
        sub->d2h_bytes += n * 8;
    }
    int rc = TB_OK;
    for (int d = 0; d < D; ++d) {
        tb_ctx* sub = ctx->subs[(size_t)d];
        cudaSetDevice(sub->device);
        int r1 = finish_sync(sub);
        if (r1 && !rc) rc = set_err(ctx, r1, "device " + std::to_string(sub->device) + ": " + sub->last_error);
    }
    if (rc) return rc;
    for (int d = 1; d < n_read; ++d)  // testing configuration (repeated devices): max over the host copies
        for (int64_t i = 0; i < n; ++i) s0->h_results[i] = std::max(s0->h_results[i], ctx->subs[(size_t)d]->h_results[i]);
    return TB_OK;
}

void multi_aggregate(tb_ctx* ctx, double t0) {
    ctx->last_ms = 0;
    ctx->last_launches = 0;
    ctx->h2d_bytes = ctx->d2h_bytes = 0;
    for (int q = 0; q < 6; ++q) ctx->host_ms[q] = 0;
    for (int q = 0; q < 4; ++q) ctx->prof_ms[q] = ctx->prof_union_ms[q] = 0, ctx->prof_launches[q] = 0;
    for (tb_ctx* sub : ctx->subs) {
        ctx->last_ms = std::max(ctx->last_ms, sub->last_ms);
        ctx->last_launches += sub->last_launches;
        ctx->h2d_bytes += sub->h2d_bytes;
        ctx->d2h_bytes += sub->d2h_bytes;
        for (int q = 0; q < 5; ++q) ctx->host_ms[q] = std::max(ctx->host_ms[q], sub->host_ms[q]);
        for (int q = 0; q < 4; ++q) {
            ctx->prof_ms[q] += sub->prof_ms[q];
            ctx->prof_launches[q] += sub->prof_launches[q];
            ctx->prof_union_ms[q] = std::max(ctx->prof_union_ms[q], sub->prof_union_ms[q]);
        }
    }
    ctx->host_ms[5] = now_ms() - t0;
}

int contract_sliced_single(tb_ctx* ctx, const tb_network* net, const int32_t* sliced_labels, int32_t n_sliced, int64_t first,
                           int64_t count, double r, double* out_values, int32_t* out_status, double* out_max);

// ONE heavy branch over all devices: the 2^k assignments are dealt in contiguous ranges (slices of one branch cost the same)
int multi_contract_sliced(tb_ctx* ctx, const tb_network* net, const int32_t* sliced_labels, int32_t n_sliced, int64_t first,
                          int64_t count, double r, double* out_values, int32_t* out_status, double* out_max) {
    const int D = (int)ctx->subs.size();
    const double t0 = now_ms();
    std::vector<double> vals((size_t)std::max<int64_t>(count, 1));
    std::vector<int32_t> stat((size_t)std::max<int64_t>(count, 1), TB_OK);
    std::vector<int> rcs((size_t)D, TB_OK);
    std::vector<double> mxs((size_t)D, -std::numeric_limits<double>::infinity());
    ThreadGroup th;
    const int64_t base = count / D, extra = count % D;
    for (int d = 0; d < D; ++d) {
        const int64_t off = d * base + std::min<int64_t>(d, extra), cnt = base + (d < extra ? 1 : 0);
        if (cnt == 0) continue;
        th.spawn([&, d, off, cnt] {
            tb_ctx* sub = ctx->subs[(size_t)d];
            try {
                cudaSetDevice(sub->device);
                rcs[(size_t)d] = contract_sliced_single(sub, net, sliced_labels, n_sliced, first + off, cnt, r, vals.data() + off,
                                                        stat.data() + off, &mxs[(size_t)d]);
            } catch (...) {
                rcs[(size_t)d] = exception_to_status(sub);
            }
        });
    }
    th.join();
    int rc = TB_OK;
    for (int d = 0; d < D; ++d)
        if (rcs[(size_t)d] && !rc) rc = set_err(ctx, rcs[(size_t)d], "device " + std::to_string(ctx->subs[(size_t)d]->device) + ": " + ctx->subs[(size_t)d]->last_error);
    if (out_values) std::memcpy(out_values, vals.data(), (size_t)count * sizeof(double));
    if (out_status) std::memcpy(out_status, stat.data(), (size_t)count * sizeof(int32_t));
    if (out_max) *out_max = *std::max_element(mxs.begin(), mxs.end());
    multi_aggregate(ctx, t0);
    return rc;
}

int multi_contract_networks(tb_ctx* ctx, const tb_network* nets, const double* r, int64_t n, double* out_values,
                            int32_t* out_status, double* out_max) {
    const int D = (int)ctx->subs.size();
    const double t0 = now_ms();
    const double ninf = -std::numeric_limits<double>::infinity();
    // ---- fewer branches than devices can keep busy and an index-slicing budget: every branch becomes 2^k index slices
    // spread over all devices (BASELINE config 3 is ONE branch)
    int64_t n_live = 0;
    for (int64_t i = 0; i < n; ++i) n_live += nets[i].n_leaves != 0;
    if (ctx->opts.slice_budget > 0 && n_live > 0 && n_live < 2 * (int64_t)D) {
        int k = 0;
        while ((n_live << k) < 2 * (int64_t)D && k < ctx->opts.slice_budget) ++k;
        double mx = ninf;
        int rc = TB_OK;
        for (int64_t i = 0; i < n && rc == TB_OK; ++i) {
            const double ri = r ? r[i] : 0.0;
            if (out_status) out_status[i] = TB_OK;
            if (nets[i].n_leaves == 0) {
                out_values[i] = ri;
            } else {
                std::vector<int32_t> labels((size_t)k);
                std::string err;
                const int got = nets[i].n_fixed == 0 ? suggest_slices(nets[i], -1, k, labels.data(), nullptr, nullptr, err) : 0;
                if (got < 0) return set_err(ctx, got, err);
                double v = ninf;
                rc = multi_contract_sliced(ctx, &nets[i], labels.data(), got, 0, (int64_t)1 << got, ri, nullptr, nullptr, &v);
                out_values[i] = v;
            }
            mx = std::max(mx, out_values[i]);
        }
        if (out_max) *out_max = mx;
        return rc;
    }
    // ---- 1. cost of every branch (label-set pass only), all host threads
    std::vector<double> cost((size_t)n, 0.0);
    {
        std::atomic<int64_t> next{0};
        std::atomic<int> bad{TB_OK};
        std::vector<std::string> errs;
        std::mutex mu;
        auto worker = [&] {
            for (;;) {
                const int64_t i = next.fetch_add(1);
                if (i >= n) break;
                if (nets[i].n_leaves == 0) continue;
                Plan P;
                std::string err;
                const int rc = compile_guarded(nets[i], ctx->opts.plan_flags | TB_PLAN_ESTIMATE_ONLY, P, err);
                if (rc) {
                    std::lock_guard<std::mutex> g(mu);
                    if (bad.load() == TB_OK) {
                        bad.store(rc);
                        errs.push_back("branch " + std::to_string(i) + ": " + err);
                    }
                } else cost[(size_t)i] = P.stats.ops;
            }
        };
        const int nt = std::max(1, std::min<int>(host_threads_of(ctx), (int)std::max<int64_t>(1, n / 16)));
        ThreadGroup th;
        for (int t = 0; t < nt - 1; ++t) th.spawn(worker);
        worker();
        th.join();
        if (bad.load() != TB_OK) return set_err(ctx, bad.load(), errs.empty() ? "plan estimation failed" : errs[0]);
    }
    // ---- 2. longest-first assignment
    std::vector<int> owner((size_t)n, 0);
    {
        std::vector<int64_t> units;
        for (int64_t i = 0; i < n; ++i)
            if (nets[i].n_leaves != 0) units.push_back(i);
        std::vector<double> load((size_t)D, 0.0);
        lpt_assign(cost, units, load, owner);
    }
    // ---- 3. every device compiles, uploads and enqueues its share on its own thread
    struct Dev {
        std::vector<int64_t> idx;
        std::vector<tb_network> nets;
        NetworksCall cs;
        int rc = TB_OK;
    };
    std::vector<Dev> dev((size_t)D);
    for (int64_t i = 0; i < n; ++i)
        if (nets[i].n_leaves != 0) {
            dev[(size_t)owner[(size_t)i]].idx.push_back(i);
            dev[(size_t)owner[(size_t)i]].nets.push_back(nets[i]);
        }
    {
        ThreadGroup th;
        for (int d = 0; d < D; ++d)
            th.spawn([&, d] {
                tb_ctx* sub = ctx->subs[(size_t)d];
                Dev& dv = dev[(size_t)d];
                try {
                    cudaSetDevice(sub->device);
                    sub->index_map = dv.idx.data();
                    sub->prefill_results = true;
                    dv.rc = networks_enqueue(sub, dv.nets.data(), (int64_t)dv.nets.size(), std::max<int64_t>(n, 1), dv.cs);
                } catch (...) {
                    dv.rc = exception_to_status(sub);
                }
                sub->index_map = nullptr;
                sub->prefill_results = false;
            });
        th.join();
    }
    int rc = TB_OK;
    for (int d = 0; d < D; ++d)
        if (dev[(size_t)d].rc && !rc) rc = set_err(ctx, dev[(size_t)d].rc, "device " + std::to_string(ctx->subs[(size_t)d]->device) + ": " + ctx->subs[(size_t)d]->last_error);
    // ---- 4. one all-reduce(max) over the result vector, then wait
    if (rc == TB_OK) rc = multi_combine(ctx, n);
    else
        for (tb_ctx* sub : ctx->subs) {
            cudaSetDevice(sub->device);
            sync_all_lanes(sub);
        }
    // ---- 5. outputs
    if (rc == TB_OK) {
        std::vector<int32_t> status((size_t)n, TB_OK);
        for (int d = 0; d < D; ++d)
            for (size_t q = 0; q < dev[(size_t)d].idx.size(); ++q) status[(size_t)dev[(size_t)d].idx[q]] = dev[(size_t)d].cs.status[q];
        const double* h = ctx->subs[0]->h_results;
        double mx = ninf;
        int worst = TB_OK;
        for (int64_t i = 0; i < n; ++i) {
            const double ri = r ? r[i] : 0.0;
            double v;
            if (nets[i].n_leaves == 0) v = ri;
            else if (status[(size_t)i] != TB_OK) {
                v = std::numeric_limits<double>::quiet_NaN();
                worst = status[(size_t)i];
            } else v = h[i] + ri;
            out_values[i] = v;
            if (out_status) out_status[i] = status[(size_t)i];
            if (status[(size_t)i] == TB_OK && v > mx) mx = v;
        }
        if (out_max) *out_max = mx;
        if (worst != TB_OK) rc = set_err(ctx, worst, "one or more branches failed (see per-branch status): arena too small");
    }
    for (int d = 0; d < D; ++d) {
        cudaSetDevice(ctx->subs[(size_t)d]->device);
        release_temporary_plans(ctx->subs[(size_t)d], std::move(dev[(size_t)d].cs.plans), true);
    }
    multi_aggregate(ctx, t0);
    return rc;
}

// resident plans on a multi-GPU context: a plan stays on the device it first ran on; new plans are dealt longest-first
int multi_contract_batch(tb_ctx* ctx, tb_plan* const* plans, const double* r, int64_t n, double* out_values, int32_t* out_status,
                         double* out_max) {
    const int D = (int)ctx->subs.size();
    const double t0 = now_ms();
    std::vector<int> owner((size_t)n, -1);
    std::vector<double> cost((size_t)n, 0.0), load((size_t)D, 0.0);
    std::vector<int64_t> fresh;
    for (int64_t i = 0; i < n; ++i) {
        if (!plans[i]) continue;
        cost[(size_t)i] = plans[i]->p.stats.ops;
        if (plans[i]->p.d_blob && plans[i]->p.owner) {
            int d = 0;
            while (d < D && ctx->subs[(size_t)d] != plans[i]->p.owner) ++d;
            if (d == D) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "plan is resident on a different context");
            owner[(size_t)i] = d;
            load[(size_t)d] += cost[(size_t)i];
        } else fresh.push_back(i);
    }
    lpt_assign(cost, fresh, load, owner);
    struct Dev {
        std::vector<int64_t> idx;
        std::vector<tb_plan*> plans;
        std::vector<int32_t> status;
        int rc = TB_OK;
    };
    std::vector<Dev> dev((size_t)D);
    for (int64_t i = 0; i < n; ++i)
        if (plans[i]) {
            dev[(size_t)owner[(size_t)i]].idx.push_back(i);
            dev[(size_t)owner[(size_t)i]].plans.push_back(plans[i]);
        }
    {
        ThreadGroup th;
        for (int d = 0; d < D; ++d)
            th.spawn([&, d] {
                tb_ctx* sub = ctx->subs[(size_t)d];
                Dev& dv = dev[(size_t)d];
                try {
                    cudaSetDevice(sub->device);
                    sub->index_map = dv.idx.data();
                    sub->prefill_results = true;
                    dv.rc = batch_enqueue(sub, dv.plans.data(), (int64_t)dv.plans.size(), std::max<int64_t>(n, 1), dv.status, false);
                } catch (...) {
                    dv.rc = exception_to_status(sub);
                }
                sub->index_map = nullptr;
                sub->prefill_results = false;
            });
        th.join();
    }
    int rc = TB_OK;
    for (int d = 0; d < D; ++d)
        if (dev[(size_t)d].rc && !rc) rc = set_err(ctx, dev[(size_t)d].rc, "device " + std::to_string(ctx->subs[(size_t)d]->device) + ": " + ctx->subs[(size_t)d]->last_error);
    if (rc == TB_OK) rc = multi_combine(ctx, n);
    else
        for (tb_ctx* sub : ctx->subs) {
            cudaSetDevice(sub->device);
            sync_all_lanes(sub);
        }
    if (rc == TB_OK) {
        std::vector<int32_t> status((size_t)n, TB_OK);
        for (int d = 0; d < D; ++d)
            for (size_t q = 0; q < dev[(size_t)d].idx.size(); ++q) status[(size_t)dev[(size_t)d].idx[q]] = dev[(size_t)d].status[q];
        const double* h = ctx->subs[0]->h_results;
        double mx = -std::numeric_limits<double>::infinity();
        int worst = TB_OK;
        for (int64_t i = 0; i < n; ++i) {
            const double ri = r ? r[i] : 0.0;
            double v;
            if (!plans[i]) v = ri;
            else if (status[(size_t)i] != TB_OK) {
                v = std::numeric_limits<double>::quiet_NaN();
                worst = status[(size_t)i];
            } else v = h[i] + ri;
            out_values[i] = v;
            if (out_status) out_status[i] = status[(size_t)i];
            if (status[(size_t)i] == TB_OK && v > mx) mx = v;
        }
        if (out_max) *out_max = mx;
        if (worst != TB_OK) rc = set_err(ctx, worst, "one or more branches failed (see per-branch status): arena too small");
    }
    multi_aggregate(ctx, t0);
    return rc;
}

}  // namespace

extern "C" {

int tb_contract_networks(tb_ctx* ctx, const tb_network* nets, const double* r, int64_t n, double* out_values,
                         int32_t* out_status, double* out_max) try {
    if (!ctx) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "ctx is NULL");
    if (n < 0 || (n > 0 && (!nets || !out_values))) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "bad arguments");
    if (!ctx->subs.empty()) return multi_contract_networks(ctx, nets, r, n, out_values, out_status, out_max);
    NetworksCall cs;
    int rc = networks_enqueue(ctx, nets, n, n, cs);
    if (rc == TB_OK) rc = finish_call(ctx, cs.plans.data(), r, n, cs.status, out_values, out_status, out_max, cs.any);
    else sync_all_lanes(ctx);
    const double t_d0 = now_ms();
    release_temporary_plans(ctx, std::move(cs.plans), true);
    ctx->host_ms[4] = now_ms() - t_d0;
    ctx->host_ms[5] = now_ms() - cs.t_c0;
    return rc;
} TB_CATCH(ctx)

int tb_stream_begin(tb_ctx* ctx, int64_t capacity, tb_stream** out_stream) try {
    if (!ctx || !out_stream) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "ctx / out_stream is NULL");
    if (!ctx->subs.empty()) return tb_stream_begin(ctx->subs[0], capacity, out_stream);  // streams run on the first device
    *out_stream = nullptr;
    if (capacity < 1) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "capacity must be >= 1");
    int rc = begin_call(ctx, capacity);
    if (rc) return rc;
    std::unique_ptr<tb_stream> s(new tb_stream());
    s->ctx = ctx;
    s->capacity = capacity;
    s->t0 = now_ms();
    ctx->call_wave = wave_for_call(ctx, std::max<int64_t>(capacity, 1024), 0.0);
    ctx->stream_open = true;
    *out_stream = s.release();
    return TB_OK;
} TB_CATCH(ctx)

int tb_stream_push(tb_stream* s, const tb_network* nets, const double* r, int64_t n) try {
    if (!s) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "stream is NULL");
    tb_ctx* ctx = s->ctx;
    if (s->failed) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "the stream already failed: call tb_stream_finish");
    if (n < 0 || (n > 0 && !nets)) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "bad arguments");
    const int64_t lo = (int64_t)s->plans.size(), hi = lo + n;
    if (hi > s->capacity) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "more branches pushed than the capacity given to tb_stream_begin");
    if (n == 0) return TB_OK;
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    const double t_c0 = now_ms();
    s->plans.resize((size_t)hi, nullptr);
    s->status.resize((size_t)hi, TB_OK);
    for (int64_t i = 0; i < n; ++i) s->r.push_back(r ? r[i] : 0.0);
    std::vector<int> codes((size_t)n, TB_OK);
    std::vector<std::string> errs((size_t)n);
    std::atomic<int64_t> next{0};
    const uint32_t flags = ctx->opts.plan_flags | TB_PLAN_TEMPORARY;
    auto worker = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n) break;
            if (nets[i].n_leaves == 0) continue;
            s->plans[(size_t)(lo + i)] = compile_temporary(ctx, nets[i], flags, codes[i], errs[i]);
        }
    };
    int nthreads = ctx->opts.host_threads > 0 ? ctx->opts.host_threads : (int)std::thread::hardware_concurrency();
    nthreads = std::max(1, std::min<int>(nthreads, (int)std::max<int64_t>(1, n / 8)));
    ThreadGroup th;
    for (int t = 0; t < nthreads - 1; ++t) th.spawn(worker);
    worker();
    th.join();
    ctx->host_ms[0] += now_ms() - t_c0;
    int rc = TB_OK;
    for (int64_t i = 0; i < n && rc == TB_OK; ++i)
        if (codes[i]) rc = set_err(ctx, codes[i], "branch " + std::to_string(lo + i) + ": " + errs[i]);
    for (int64_t i = lo; i < hi; ++i) s->any = s->any || s->plans[(size_t)i];
    if (rc == TB_OK && lo == 0) ctx->call_wave = wave_for_call(ctx, std::max<int64_t>(s->capacity, 1024), mean_plan_ops(s->plans.data(), lo, hi));
    // the launches are asynchronous: the call returns while the GPU contracts, the host goes back to slicing
    if (rc == TB_OK) rc = enqueue_batch(ctx, s->plans.data(), lo, hi, s->status, false);
    if (rc) s->failed = true;
    return rc;
} TB_CATCH((s ? s->ctx : nullptr))

int tb_stream_finish(tb_stream* s, double* out_values, int32_t* out_status, int64_t cap, int64_t* out_n, double* out_max) try {
    if (!s) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "stream is NULL");
    tb_ctx* ctx = s->ctx;
    const int64_t n = (int64_t)s->plans.size();
    int rc = TB_OK;
    if (out_n) *out_n = n;
    if (s->failed) {
        sync_all_lanes(ctx);
        rc = set_err(ctx, TB_ERR_BAD_ARGUMENT, "the stream failed in tb_stream_push: " + ctx->last_error);
    } else if (n > 0 && (!out_values || cap < n)) {
        sync_all_lanes(ctx);
        rc = set_err(ctx, TB_ERR_BAD_ARGUMENT, "output buffer smaller than the number of pushed branches");
    } else {
        std::vector<double> vals((size_t)std::max<int64_t>(n, 1));
        rc = finish_call(ctx, s->plans.data(), s->r.data(), n, s->status, vals.data(), out_status, out_max, s->any);
        if (n > 0) std::memcpy(out_values, vals.data(), (size_t)n * sizeof(double));
    }
    release_temporary_plans(ctx, std::move(s->plans), true);
    ctx->host_ms[5] = now_ms() - s->t0;
    ctx->stream_open = false;
    delete s;
    return rc;
} TB_CATCH((s ? s->ctx : nullptr))

int tb_contract_sliced(tb_ctx* ctx, const tb_network* net, const int32_t* sliced_labels, int32_t n_sliced, int64_t first,
                       int64_t count, double r, double* out_values, int32_t* out_status, double* out_max) try {
    if (!ctx) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "ctx is NULL");
    if (!ctx->subs.empty()) {
        if (n_sliced < 0 || n_sliced > 40 || first < 0 || count < 0 || first + count > ((int64_t)1 << n_sliced))
            return set_err(ctx, TB_ERR_BAD_ARGUMENT, "assignment range outside [0, 2^n_sliced)");
        return multi_contract_sliced(ctx, net, sliced_labels, n_sliced, first, count, r, out_values, out_status, out_max);
    }
    return contract_sliced_single(ctx, net, sliced_labels, n_sliced, first, count, r, out_values, out_status, out_max);
} TB_CATCH(ctx)

}  // extern "C"

namespace {
int contract_sliced_single(tb_ctx* ctx, const tb_network* net, const int32_t* sliced_labels, int32_t n_sliced, int64_t first,
                           int64_t count, double r, double* out_values, int32_t* out_status, double* out_max) {
    if (!net || net->n_leaves < 1) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "net is NULL or empty");
    if (net->n_fixed != 0) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "the network of tb_contract_sliced must not carry fixed labels itself");
    if (n_sliced < 0 || n_sliced > 40 || (n_sliced > 0 && !sliced_labels)) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "bad sliced labels (0 <= n_sliced <= 40)");
    const int64_t n_assign = (int64_t)1 << n_sliced;
    if (first < 0 || count < 0 || first + count > n_assign) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "assignment range outside [0, 2^n_sliced)");
    if (count > ((int64_t)1 << 24)) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "more than 2^24 assignments in one call: split the range");
    // position of every sliced label; pairs of sliced labels joined by an edge leaf make an assignment infeasible
    std::vector<int32_t> pos((size_t)std::max(net->n_labels, 1), -1);
    for (int i = 0; i < n_sliced; ++i) {
        int32_t l = sliced_labels[i];
        if (l < 0 || l >= net->n_labels) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "sliced label out of range");
        if (pos[l] >= 0) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "sliced label repeated");
        pos[l] = i;
    }
    std::vector<uint64_t> conflicts;
    if (!net->leaf_off) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "leaf_off is NULL");
    for (int i = 0; i < net->n_leaves; ++i) {
        int b = net->leaf_off[i], e = net->leaf_off[i + 1];
        if (e - b != 2) continue;
        int32_t u = net->leaf_labels[b], v = net->leaf_labels[b + 1];
        if (u < 0 || v < 0 || u >= net->n_labels || v >= net->n_labels) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "leaf label out of range");
        if (u != v && pos[u] >= 0 && pos[v] >= 0) conflicts.push_back((1ull << pos[u]) | (1ull << pos[v]));
    }
    const double ninf = -std::numeric_limits<double>::infinity();
    std::vector<double> vals((size_t)count, ninf);
    std::vector<int32_t> stat((size_t)count, TB_OK);
    std::vector<int64_t> live;  // assignments that are contracted
    for (int64_t a = first; a < first + count; ++a) {
        bool ok = true;
        for (uint64_t c : conflicts)
            if (((uint64_t)a & c) == c) {
                ok = false;
                break;
            }
        if (ok) live.push_back(a);
    }
    int rc = TB_OK;
    // ONE compilation: every assignment is the same plan with other leaf-pool words (clone_for_assignment), contracted
    // as resident plans in chunks
    const double t_c0 = now_ms();
    double t_build = 0;
    const size_t kChunk = 4096;
    std::unique_ptr<tb_plan> base;
    std::vector<uint8_t> fv((size_t)std::max(n_sliced, 1));
    auto values_of = [&](int64_t a) {
        for (int i = 0; i < n_sliced; ++i) fv[(size_t)i] = (uint8_t)((a >> i) & 1);
        return fv.data();
    };
    for (size_t c0 = 0; c0 < live.size() && rc == TB_OK; c0 += kChunk) {
        const size_t c1 = std::min(live.size(), c0 + kChunk), nl = c1 - c0;
        const double tb0 = now_ms();
        if (!base) {
            tb_network n0 = *net;
            n0.n_fixed = n_sliced;
            n0.fixed_labels = sliced_labels;
            n0.fixed_values = values_of(live[0]);
            base.reset(new tb_plan());
            std::string err;
            rc = compile_plan(n0, ctx->opts.plan_flags | TB_PLAN_TEMPORARY, base->p, err);
            if (rc) return set_err(ctx, rc, err);
        }
        std::vector<tb_plan*> plans(nl, nullptr);
        for (size_t q = 0; q < nl; ++q) plans[q] = clone_for_assignment(base.get(), values_of(live[c0 + q]));
        t_build += now_ms() - tb0;
        std::vector<double> lv(nl), rr(nl, r);
        std::vector<int32_t> ls(nl, TB_OK);
        rc = contract_impl(ctx, plans.data(), rr.data(), (int64_t)nl, lv.data(), ls.data(), nullptr, false);
        for (size_t q = 0; q < nl; ++q) {
            vals[(size_t)(live[c0 + q] - first)] = lv[q];
            stat[(size_t)(live[c0 + q] - first)] = ls[q];
        }
        release_temporary_plans(ctx, std::move(plans));
    }
    ctx->host_ms[0] = t_build;
    ctx->host_ms[5] = now_ms() - t_c0;
    double mx = ninf;
    for (int64_t i = 0; i < count; ++i)
        if (stat[(size_t)i] == TB_OK && vals[(size_t)i] > mx) mx = vals[(size_t)i];
    if (out_values) std::memcpy(out_values, vals.data(), (size_t)count * sizeof(double));
    if (out_status) std::memcpy(out_status, stat.data(), (size_t)count * sizeof(int32_t));
    if (out_max) *out_max = mx;
    return rc;
}
}  // namespace

extern "C" {

int tb_plan_reassign(const tb_plan* base, const uint8_t* fixed_values, tb_plan** out_plan) try {
    if (!base || !fixed_values || !out_plan) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "base / fixed_values / out_plan is NULL");
    *out_plan = nullptr;
    if (base->p.n_fixed <= 0) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "the base plan was not created with fixed labels");
    for (int i = 0; i < base->p.n_fixed; ++i)
        if (fixed_values[i] > 1) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "fixed value must be 0 or 1");
    *out_plan = clone_for_assignment(base, fixed_values);
    return TB_OK;
} TB_CATCH(nullptr)

int tb_suggest_slices(tb_ctx* ctx, const tb_network* net, int32_t sc_target, int32_t max_sliced, int32_t* out_labels,
                      double* out_sc, double* out_tc) try {
    if (!net) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "net is NULL");
    std::string err;
    int rc = suggest_slices(*net, sc_target, max_sliced, out_labels, out_sc, out_tc, err);
    if (rc < 0) return set_err(ctx, rc, err);
    return rc;
} TB_CATCH(ctx)

int tb_plan_read_tensor(tb_ctx* ctx, tb_plan* plan, int32_t node, double* out_data, int64_t cap, int32_t* out_labels,
                        int32_t* out_rank) try {
    if (!ctx || !plan) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "ctx / plan is NULL");
    if (!ctx->subs.empty()) return tb_plan_read_tensor(ctx->subs[0], plan, node, out_data, cap, out_labels, out_rank);
    const Plan& P = plan->p;
    if (!(P.flags & TB_PLAN_KEEP_INTERMEDIATES)) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "plan was not created with TB_PLAN_KEEP_INTERMEDIATES");
    if (node < 0 || node >= P.n_tensors || P.loc[node] != LOC_ARENA)
        return set_err(ctx, TB_ERR_BAD_ARGUMENT, "node id is not an internal node");
    int rank = P.rank(node);
    if (out_rank) *out_rank = rank;
    if (out_labels)
        for (int i = 0; i < rank; ++i) out_labels[i] = P.layout(node)[i];
    int64_t n = (int64_t)1 << rank;
    if (!out_data) return TB_OK;
    if (cap < n) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "output buffer too small");
    return copy_tensor_to_host(ctx, P, P.off[node], n, out_data);
} TB_CATCH(ctx)

int tb_contract_tensor(tb_ctx* ctx, tb_plan* plan, double* out_data, int64_t cap, int32_t* out_labels, int32_t* out_rank) try {
    if (!ctx || !plan) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "ctx / plan is NULL");
    if (!ctx->subs.empty()) return tb_contract_tensor(ctx->subs[0], plan, out_data, cap, out_labels, out_rank);
    const Plan& P = plan->p;
    const int rank = P.rank(P.root_id);
    if (out_rank) *out_rank = rank;
    if (out_labels)
        for (int i = 0; i < rank; ++i) out_labels[i] = P.layout(P.root_id)[i];
    const int64_t n = (int64_t)1 << rank;
    if (!out_data) return TB_OK;  // size query
    if (cap < n) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "output buffer too small");
    double first = 0;
    tb_plan* arr[1] = {plan};
    int rc = contract_impl(ctx, arr, nullptr, 1, &first, nullptr, nullptr, true);
    if (rc) return rc;
    return copy_tensor_to_host(ctx, P, P.root_off, n, out_data);
} TB_CATCH(ctx)

int tb_contract_table(tb_ctx* ctx, tb_plan* plan, double* out_sizes, uint32_t* out_configs, int64_t cap, int32_t* out_labels,
                      int32_t* out_rank) try {
    if (!ctx || !plan) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "ctx / plan is NULL");
    if (!ctx->subs.empty()) return tb_contract_table(ctx->subs[0], plan, out_sizes, out_configs, cap, out_labels, out_rank);
    const Plan& P = plan->p;
    if (P.value_type != TB_VALUE_SIZE_CONFIG)
        return set_err(ctx, TB_ERR_BAD_ARGUMENT, "tb_contract_table needs a plan created with value_type TB_VALUE_SIZE_CONFIG");
    int rc = tb_contract_tensor(ctx, plan, out_sizes, cap, out_labels, out_rank);  // contracts; sizes of the root tensor
    if (rc || !out_sizes || !out_configs) return rc;
    const int64_t n = (int64_t)1 << P.rank(P.root_id);
    uint32_t* d_cfg = nullptr;
    TB_CUDA(ctx, cudaMalloc(&d_cfg, (size_t)n * sizeof(uint32_t)));
    const long long* src = (const long long*)ctx->arena + (ctx->last_plan_arena_base_elems + P.root_off);
    k_to_config<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(src, d_cfg, n);
    cudaError_t e = cudaMemcpyAsync(out_configs, d_cfg, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_cfg);
    if (e != cudaSuccess) return set_err(ctx, TB_ERR_CUDA, cudaGetErrorString(e));
    return TB_OK;
} TB_CATCH(ctx)

int tb_compactify_table(tb_ctx* ctx, int32_t rank, const double* sizes, uint8_t* out_keep) try {
    if (!ctx || !sizes || !out_keep) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "NULL argument");
    if (!ctx->subs.empty()) return tb_compactify_table(ctx->subs[0], rank, sizes, out_keep);
    if (rank < 0 || rank > 30) return set_err(ctx, TB_ERR_UNSUPPORTED, "rank must be in [0, 30]");
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = (int64_t)1 << rank;
    // device scratch: sizes | subset maxima | keep flags
    const size_t bytes = (size_t)n * (2 * sizeof(double) + 1);
    uint8_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(ctx, TB_ERR_OUT_OF_MEMORY, cudaGetErrorString(e));
    }
    double* d_sizes = (double*)d;
    double* d_z = d_sizes + n;
    uint8_t* d_keep = (uint8_t*)(d_z + n);
    e = cudaMemcpyAsync(d_sizes, sizes, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_z, d_sizes, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) {
        cudaEventRecord(ctx->ev0, ctx->stream);
        for (int bit = 0; bit < rank; ++bit)
            k_subset_max_stage<<<(unsigned)((n / 2 + 255) / 256), 256, 0, ctx->stream>>>(d_z, bit, n / 2);
        k_table_keep<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_sizes, d_z, rank, d_keep, n);
        cudaEventRecord(ctx->ev1, ctx->stream);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_keep, d_keep, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->last_ms = ms;
        ctx->last_launches = rank + 1;
    }
    cudaFree(d);
    if (e != cudaSuccess) return set_err(ctx, TB_ERR_CUDA, cudaGetErrorString(e));
    return TB_OK;
} TB_CATCH(ctx)

}  // extern "C"

namespace {
// One region of a batch: the region's graph read off the network's leaves (1 label = vertex tensor [0, w_v], 2 labels = edge
// tensor; tbcuda.h, tb_network -- the tree is not used) and its boundary vertices in the bit order of its rows.
int region_desc_of(tb_ctx* ctx, const tb_network* net, const int32_t* boundary_labels, int32_t rank, RegionDesc& R) {
    const int n = net->n_labels;
    if (n < 0 || n > 32) return set_err(ctx, TB_ERR_UNSUPPORTED, "a region has at most 32 vertices (configurations are 32-bit vertex masks)");
    if (rank < 0 || rank > n || rank > 24) return set_err(ctx, TB_ERR_UNSUPPORTED, "boundary rank must be in [0, min(n_labels, 24)]");
    if (rank > 0 && !boundary_labels) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "boundary_labels is NULL");
    if (net->n_fixed != 0) return set_err(ctx, TB_ERR_UNSUPPORTED, "the table calls do not take index-sliced networks");
    if (net->n_leaves < 0 || (net->n_leaves > 0 && (!net->leaf_off || !net->leaf_labels)))
        return set_err(ctx, TB_ERR_BAD_ARGUMENT, "leaf arrays are NULL");
    const int wd = net->weight_dtype;
    if (wd < TB_WEIGHT_UNIT || wd > TB_WEIGHT_F64) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "unknown weight_dtype");
    if (wd != TB_WEIGHT_UNIT && !net->weights) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "weights is NULL but weight_dtype is not UNIT");
    auto weight_of = [&](int v) -> double {
        switch (wd) {
            case TB_WEIGHT_I32: return (double)((const int32_t*)net->weights)[v];
            case TB_WEIGHT_I64: return (double)((const int64_t*)net->weights)[v];
            case TB_WEIGHT_F32: return (double)((const float*)net->weights)[v];
            case TB_WEIGHT_F64: return ((const double*)net->weights)[v];
            default: return 1.0;
        }
    };
    std::memset(&R, 0, sizeof(R));
    R.n = n;
    R.rank = rank;
    for (int32_t l = 0; l < net->n_leaves; ++l) {
        const int32_t b = net->leaf_off[l], e = net->leaf_off[l + 1];
        if (e - b < 1 || e - b > 2) return set_err(ctx, TB_ERR_UNSUPPORTED, "leaf tensors have 1 (vertex) or 2 (edge) labels");
        const int32_t u = net->leaf_labels[b], v = net->leaf_labels[e - 1];
        if (u < 0 || u >= n || v < 0 || v >= n) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "leaf label out of range");
        if (e - b == 1) R.w[u] += weight_of(u);
        else if (u == v) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "an edge tensor joins a vertex with itself");
        else {
            R.adj[u] |= 1u << v;
            R.adj[v] |= 1u << u;
        }
    }
    uint32_t bset = 0;
    for (int i = 0; i < rank; ++i) {
        const int32_t v = boundary_labels[i];
        if (v < 0 || v >= n || ((bset >> v) & 1)) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "boundary labels must be distinct labels of the network");
        bset |= 1u << v;
        R.bpos[i] = (uint8_t)v;
    }
    for (int v = 0; v < n; ++v)
        if (!((bset >> v) & 1)) R.ipos[R.n_int++] = (uint8_t)v;
    R.chunk_log2 = std::min(R.n_int, 12);
    return TB_OK;
}

// tb_table_configs (compactify == false: rows chosen by `keep`, NULL = all), tb_branching_table(s) (compactify == true: the
// dominated rows are dropped on the device between the optimum pass and the count pass, flags returned in out_keep), over a
// batch of n regions in the same launches.  Rows of region i: [row_base_i, row_base_i + 2^rank_i).
int table_configs_impl(tb_ctx* ctx, const tb_network* nets, const int32_t* boundary_off, const int32_t* boundary_labels, int64_t n,
                       const uint8_t* keep, bool compactify, uint8_t* out_keep, double* out_sizes, int64_t* out_row_off,
                       uint32_t* out_configs, int64_t cap, int64_t* out_total) {
    if (!ctx || !out_row_off) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "ctx / out_row_off is NULL");
    if (n < 0 || n > 0x7fffffffll || (n > 0 && (!nets || !boundary_off))) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "bad region list");
    if (!ctx->subs.empty())
        return table_configs_impl(ctx->subs[0], nets, boundary_off, boundary_labels, n, keep, compactify, out_keep, out_sizes,
                                  out_row_off, out_configs, cap, out_total);
    if (out_total) *out_total = 0;
    out_row_off[0] = 0;
    if (n == 0) return TB_OK;
    std::vector<RegionDesc> regs((size_t)n);
    int64_t n_rows = 0, n_cta = 0;
    int max_rank = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t rank = boundary_off[i + 1] - boundary_off[i];
        int rc = region_desc_of(ctx, &nets[i], boundary_labels ? boundary_labels + boundary_off[i] : nullptr, rank, regs[(size_t)i]);
        if (rc) return rc;
        RegionDesc& R = regs[(size_t)i];
        R.row_base = n_rows;
        R.cta_base = n_cta;
        n_rows += (int64_t)1 << rank;
        n_cta += (int64_t)1 << (rank + R.n_int - R.chunk_log2);
        max_rank = std::max(max_rank, (int)rank);
        if (n_cta > 0x7fffffffll || n_rows > ((int64_t)1 << 28)) return set_err(ctx, TB_ERR_UNSUPPORTED, "the batch of regions is too large for one call");
    }
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    auto grow = [&](void*& buf, size_t& have, size_t need) -> cudaError_t {
        if (have >= need) return cudaSuccess;
        if (buf) cudaFree(buf);
        buf = nullptr;
        have = 0;
        need = std::max<size_t>(need + need / 2, 1u << 20);
        cudaError_t e = cudaMalloc(&buf, need);
        if (e == cudaSuccess) have = need;
        else buf = nullptr;
        return e;
    };
    // device scratch: region descriptors | alpha keys | sizes | subset maxima | row offsets | chunk counts | chunk offsets | keep flags
    const size_t b_regs = (size_t)n * sizeof(RegionDesc), b_rows = (size_t)n_rows * 8, b_off = (size_t)(n_rows + 1) * 8,
                 b_cnt = (size_t)n_cta * 8, b_coff = (size_t)(n_cta + 1) * 8;
    TB_CUDA(ctx, grow(ctx->table_buf, ctx->table_cap, b_regs + 3 * b_rows + b_off + b_cnt + b_coff + (size_t)n_rows));
    uint8_t* d = (uint8_t*)ctx->table_buf;
    RegionDesc* d_regs = (RegionDesc*)d;
    unsigned long long* d_alpha = (unsigned long long*)(d + b_regs);
    double* d_sizes = (double*)(d + b_regs + b_rows);
    double* d_z = (double*)(d + b_regs + 2 * b_rows);
    int64_t* d_row_off = (int64_t*)(d + b_regs + 3 * b_rows);
    int64_t* d_cnt = (int64_t*)(d + b_regs + 3 * b_rows + b_off);
    int64_t* d_coff = (int64_t*)(d + b_regs + 3 * b_rows + b_off + b_cnt);
    uint8_t* d_keep = nullptr;
    if (keep || compactify) d_keep = d + b_regs + 3 * b_rows + b_off + b_cnt + b_coff;
    TB_CUDA(ctx, cudaMemcpyAsync(d_regs, regs.data(), b_regs, cudaMemcpyHostToDevice, st));  // (pageable source: staged before the call returns)
    if (keep && !compactify) TB_CUDA(ctx, cudaMemcpyAsync(d_keep, keep, (size_t)n_rows, cudaMemcpyHostToDevice, st));
    const unsigned g_rows = (unsigned)((n_rows + 255) / 256);
    const int nr = (int)n;
    int launches = 6;
    cudaEventRecord(ctx->ev0, st);
    k_region_init<<<g_rows, 256, 0, st>>>(d_alpha, n_rows);
    k_region_configs<0><<<(unsigned)n_cta, kRegionThreads, 0, st>>>(d_regs, nr, d_alpha, nullptr, nullptr, nullptr, nullptr);
    k_region_sizes<<<g_rows, 256, 0, st>>>(d_alpha, d_sizes, d_z, n_rows);
    if (compactify) {  // mis_compactify on the row optima of every region
        for (int bit = 0; bit < max_rank; ++bit) k_region_subset_max<<<g_rows, 256, 0, st>>>(d_regs, nr, d_z, bit, n_rows);
        k_region_keep<<<g_rows, 256, 0, st>>>(d_regs, nr, d_sizes, d_z, d_keep, n_rows);
        launches += max_rank + 1;
        if (out_keep) TB_CUDA(ctx, cudaMemcpyAsync(out_keep, d_keep, (size_t)n_rows, cudaMemcpyDeviceToHost, st));
    }
    k_region_configs<1><<<(unsigned)n_cta, kRegionThreads, 0, st>>>(d_regs, nr, d_alpha, d_keep, d_cnt, nullptr, nullptr);
    k_region_scan<<<1, 1024, 0, st>>>(d_cnt, d_coff, n_cta);
    k_region_row_off<<<(unsigned)((n_rows + 256) / 256), 256, 0, st>>>(d_regs, nr, d_coff, n_cta, d_row_off, n_rows);
    TB_CUDA(ctx, cudaGetLastError());
    TB_CUDA(ctx, cudaMemcpyAsync(out_row_off, d_row_off, b_off, cudaMemcpyDeviceToHost, st));
    if (out_sizes) TB_CUDA(ctx, cudaMemcpyAsync(out_sizes, d_sizes, b_rows, cudaMemcpyDeviceToHost, st));
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    const int64_t total = out_row_off[n_rows];
    if (out_total) *out_total = total;
    ctx->last_launches = launches;
    if (out_configs && total > 0) {
        if (cap < total) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "out_configs holds " + std::to_string(cap) + " configurations, the table has " + std::to_string(total));
        TB_CUDA(ctx, grow(ctx->table_out, ctx->table_out_cap, (size_t)total * sizeof(uint32_t)));
        uint32_t* d_out = (uint32_t*)ctx->table_out;
        k_region_configs<2><<<(unsigned)n_cta, kRegionThreads, 0, st>>>(d_regs, nr, d_alpha, d_keep, nullptr, d_coff, d_out);
        TB_CUDA(ctx, cudaGetLastError());
        TB_CUDA(ctx, cudaMemcpyAsync(out_configs, d_out, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        ctx->last_launches = launches + 1;
    }
    cudaEventRecord(ctx->ev1, st);
    TB_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->last_ms = ms;
    return TB_OK;
}
}  // namespace

extern "C" {

int tb_table_configs(tb_ctx* ctx, const tb_network* net, const int32_t* boundary_labels, int32_t rank, const uint8_t* keep,
                     double* out_sizes, int64_t* out_row_off, uint32_t* out_configs, int64_t cap, int64_t* out_total) try {
    if (!net) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "net is NULL");
    const int32_t off[2] = {0, rank};
    return table_configs_impl(ctx, net, off, boundary_labels, 1, keep, false, nullptr, out_sizes, out_row_off, out_configs, cap, out_total);
} TB_CATCH(ctx)

int tb_branching_table(tb_ctx* ctx, const tb_network* net, const int32_t* boundary_labels, int32_t rank, uint8_t* out_keep,
                       double* out_sizes, int64_t* out_row_off, uint32_t* out_configs, int64_t cap, int64_t* out_total) try {
    if (!net) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "net is NULL");
    const int32_t off[2] = {0, rank};
    return table_configs_impl(ctx, net, off, boundary_labels, 1, nullptr, true, out_keep, out_sizes, out_row_off, out_configs, cap, out_total);
} TB_CATCH(ctx)

int tb_branching_tables(tb_ctx* ctx, const tb_network* nets, const int32_t* boundary_off, const int32_t* boundary_labels, int64_t n,
                        uint8_t* out_keep, double* out_sizes, int64_t* out_row_off, uint32_t* out_configs, int64_t cap,
                        int64_t* out_total) try {
    return table_configs_impl(ctx, nets, boundary_off, boundary_labels, n, nullptr, true, out_keep, out_sizes, out_row_off, out_configs, cap, out_total);
} TB_CATCH(ctx)

int tb_last_timing(const tb_ctx* ctx, double* out_device_ms, int64_t* out_launches) {
    if (!ctx) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "ctx is NULL");
    if (out_device_ms) *out_device_ms = ctx->last_ms;
    if (out_launches) *out_launches = ctx->last_launches;
    return TB_OK;
}

int tb_set_stream(tb_ctx* ctx, void* cuda_stream) {
    if (!ctx) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "ctx is NULL");
    if (!ctx->subs.empty()) return set_err(ctx, TB_ERR_UNSUPPORTED, "tb_set_stream needs a single-device context (a stream belongs to one device)");
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return TB_OK;
}

int tb_profile(tb_ctx* ctx, int enable) {
    if (!ctx) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "ctx is NULL");
    ctx->profile_mode = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
    for (tb_ctx* sub : ctx->subs) sub->profile_mode = ctx->profile_mode;
    return TB_OK;
}

int tb_last_profile_union(const tb_ctx* ctx, double* ms_by_kind) {
    if (!ctx || !ms_by_kind) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "NULL argument");
    for (int q = 0; q < 4; ++q) ms_by_kind[q] = ctx->prof_union_ms[q];
    return TB_OK;
}

int tb_last_profile(const tb_ctx* ctx, double* ms_by_kind, int64_t* launches_by_kind) {
    if (!ctx) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "ctx is NULL");
    for (int q = 0; q < 4; ++q) {
        if (ms_by_kind) ms_by_kind[q] = ctx->prof_ms[q];
        if (launches_by_kind) launches_by_kind[q] = ctx->prof_launches[q];
    }
    return TB_OK;
}

int tb_last_host_breakdown(const tb_ctx* ctx, double* ms6) {
    if (!ctx || !ms6) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "NULL argument");
    for (int q = 0; q < 6; ++q) ms6[q] = ctx->host_ms[q];
    return TB_OK;
}

int tb_last_transfers(const tb_ctx* ctx, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (!ctx) return set_err(nullptr, TB_ERR_BAD_ARGUMENT, "ctx is NULL");
    if (h2d_bytes) *h2d_bytes = ctx->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = ctx->d2h_bytes;
    return TB_OK;
}

int tb_permute_bits(tb_ctx* ctx, const void* in, void* out, int32_t rank, const int32_t* perm) try {
    if (!ctx || !in || !out || !perm) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "NULL argument");
    if (!ctx->subs.empty()) return tb_permute_bits(ctx->subs[0], in, out, rank, perm);
    if (rank < 0 || rank > 31) return set_err(ctx, TB_ERR_UNSUPPORTED, "rank must be in [0, 31]");
    std::vector<int> dst_of_src(rank, -1);
    for (int i = 0; i < rank; ++i) {
        if (perm[i] < 0 || perm[i] >= rank || dst_of_src[perm[i]] != -1) return set_err(ctx, TB_ERR_BAD_ARGUMENT, "perm is not a permutation");
        dst_of_src[perm[i]] = i;
    }
    TB_CUDA(ctx, cudaSetDevice(ctx->device));
    PermuteDesc d{};
    d.rank = (uint8_t)rank;
    // tile = the 6 low source bits (256-byte reads) + the source bits that feed the 6 low destination bits (256-byte
    // writes), filled up with the lowest remaining bits to 2^12 elements: a permutation that leaves the low bits alone
    // (the identity at the extreme) still moves 16 KB per CTA instead of 128 bytes
    const int Lb = std::min(rank, 6);
    std::vector<int> in_tile(rank, 0);
    int n_tile = 0;
    for (int i = 0; i < Lb; ++i) {
        n_tile += !in_tile[i];
        in_tile[i] = 1;  // low source bits
        n_tile += !in_tile[perm[i]];
        in_tile[perm[i]] = 1;  // source bits feeding the low destination bits
    }
    for (int b = 0; b < rank && n_tile < 12; ++b)
        if (!in_tile[b]) {
            in_tile[b] = 1;
            ++n_tile;
        }
    std::vector<int> tsrc, gsrc;
    for (int b = 0; b < rank; ++b) (in_tile[b] ? tsrc : gsrc).push_back(b);
    d.u = (uint8_t)tsrc.size();
    d.ng = (uint8_t)gsrc.size();
    std::vector<int> tdst;  // destination bits of the tile, ascending
    for (int b : tsrc) tdst.push_back(dst_of_src[b]);
    std::sort(tdst.begin(), tdst.end());
    for (size_t j = 0; j < tsrc.size(); ++j) d.tile_src_bit[j] = (uint8_t)tsrc[j];
    for (size_t j = 0; j < tdst.size(); ++j) {
        d.tile_dst_bit[j] = (uint8_t)tdst[j];
        int sb = perm[tdst[j]];
        d.dst_to_src_tile[j] = (uint8_t)(std::find(tsrc.begin(), tsrc.end(), sb) - tsrc.begin());
    }
    for (size_t j = 0; j < gsrc.size(); ++j) {
        d.grid_src_bit[j] = (uint8_t)gsrc[j];
        d.grid_dst_bit[j] = (uint8_t)dst_of_src[gsrc[j]];
    }
    size_t bytes = ((size_t)1 << rank) * 4;
    const size_t half_cap = (bytes + 255) / 256 * 256;
    if (ctx->permute_cap < 2 * half_cap) {  // the scratch pair is kept: repeated calls do not allocate
        if (ctx->permute_buf) cudaFree(ctx->permute_buf);
        ctx->permute_buf = nullptr;
        ctx->permute_cap = 0;
        cudaError_t ea = cudaMalloc(&ctx->permute_buf, 2 * half_cap);
        if (ea != cudaSuccess) {
            cudaGetLastError();
            return set_err(ctx, TB_ERR_OUT_OF_MEMORY, cudaGetErrorString(ea));
        }
        ctx->permute_cap = 2 * half_cap;
    }
    void *din = ctx->permute_buf, *dout = (uint8_t*)ctx->permute_buf + half_cap;
    cudaError_t e = cudaMemcpyAsync(din, in, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        cudaEventRecord(ctx->ev0, ctx->stream);
        k_permute_bits<<<(unsigned)(1u << d.ng), 256, 0, ctx->stream>>>((const uint32_t*)din, (uint32_t*)dout, d);
        cudaEventRecord(ctx->ev1, ctx->stream);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->last_ms = ms;
        ctx->last_launches = 1;
    }
    if (e != cudaSuccess) return set_err(ctx, TB_ERR_CUDA, cudaGetErrorString(e));
    return TB_OK;
} TB_CATCH(ctx)

}  // extern "C"
