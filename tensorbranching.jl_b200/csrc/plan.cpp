// plan.cpp -- compile one branch network (leaf label lists + binary contraction tree) into a flat,
// layout-resolved step list.  Replaces, for the hot path, what the reference does on every
// solve_slice call: uncompress -> parse_eincode -> decorate (/root/reference/src/types.jl:75-79) and
// OMEinsum's per-node label analysis + permutedims decisions [upstream].  Pure host C++.
//
// Passes: validate -> label sets (bottom-up) -> split-K rewrite (long reductions become a partial step
// that keeps some reduced labels + a unary max step) -> layouts (top-down, consumer dictates operand
// layout) -> kinds (fused subtree / generic / tiled GEMM) -> levels -> arena lifetimes -> descriptors.
//
// contract_slices compiles thousands of these per call, so everything is flat arrays + O(1) stamp
// tests; no per-node heap allocation.
#include "plan.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#ifdef TB_PLAN_PROFILE
#include <chrono>
#include <cstdio>
static double g_pt[16];
static const char* g_pn[16];
#define PT(i, name) { auto now_ = std::chrono::steady_clock::now(); g_pt[i] += std::chrono::duration<double, std::micro>(now_ - last_).count(); g_pn[i] = name; last_ = now_; }
void tb_plan_profile_dump(int reps) { for (int i = 0; i < 16; ++i) if (g_pn[i]) printf("%-12s %8.2f us\n", g_pn[i], g_pt[i] / reps); }
#else
#define PT(i, name)
#endif

namespace tb {
namespace {

struct NodeCls {  // label classes of one contraction, labels stored in cls_data at off in this order
    int32_t off = 0;
    uint8_t nm = 0, nn = 0, nb = 0, nk = 0, nka = 0, nkb = 0, tm = 0, tn = 0;
    uint8_t kfirst = 0;  // packed int16 GEMM operands: [K0 | X_lo | K_rest | X_hi | Bt]
};

struct FreeList {
    std::vector<std::pair<int64_t, int64_t>> blocks;  // sorted, non-adjacent free blocks below `top`
    int64_t top = 0;
    int64_t alloc(int64_t size) {
        for (size_t i = 0; i < blocks.size(); ++i) {
            if (blocks[i].second >= size) {
                int64_t o = blocks[i].first;
                blocks[i].first += size;
                blocks[i].second -= size;
                if (blocks[i].second == 0) blocks.erase(blocks.begin() + i);
                return o;
            }
        }
        if (!blocks.empty() && blocks.back().first + blocks.back().second == top) {
            int64_t o = blocks.back().first;
            top = o + size;
            blocks.pop_back();
            return o;
        }
        int64_t o = top;
        top += size;
        return o;
    }
    void release(int64_t o, int64_t size) {
        auto it = std::lower_bound(blocks.begin(), blocks.end(), std::make_pair(o, (int64_t)0));
        it = blocks.insert(it, {o, size});
        auto nx = it + 1;
        if (nx != blocks.end() && it->first + it->second == nx->first) {
            it->second += nx->second;
            blocks.erase(nx);
        }
        if (it != blocks.begin()) {
            auto pv = it - 1;
            if (pv->first + pv->second == it->first) {
                pv->second += it->second;
                blocks.erase(it);
            }
        }
    }
};

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct Lcg {
    uint64_t s;
    explicit Lcg(uint64_t seed) : s(seed * 6364136223846793005ull + 1442695040888963407ull) {}
    uint32_t next() {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        return (uint32_t)(s >> 33);
    }
    void shuffle(int32_t* v, int n) {
        for (int i = n; i > 1; --i) std::swap(v[i - 1], v[next() % i]);
    }
};

// block-level split-k of the generic kernel: ks thread-bits per output, po output-bits per CTA
inline void generic_split(int rc, int nkt, int& ks, int& po) {
    int ks0 = std::min(std::max(nkt - 6, 0), 8);
    ks = (rc < 8 - ks0) ? std::min(nkt, 8 - rc) : ks0;
    po = std::min(rc, 8 - ks);
}

// sort labels v[0..n) by key (small n): insertion sort on (key, label)
inline void sort_by_key(int32_t* v, const int32_t* key, int n) {
    for (int i = 1; i < n; ++i) {
        int32_t x = v[i];
        int32_t kx = key[x];
        int j = i - 1;
        while (j >= 0 && (key[v[j]] > kx || (key[v[j]] == kx && v[j] > x))) {
            v[j + 1] = v[j];
            --j;
        }
        v[j + 1] = x;
    }
}

}  // namespace

namespace {

// The compiler's working arrays.  contract_slices compiles thousands of plans per call on every worker thread, so
// each thread keeps one set and a compilation borrows it: after the first few plans no pass allocates memory.
struct CompilerArrays {
    // flat per-tensor storage: tree, sorted label sets
    std::vector<int32_t> lch, rch, parent, lab_off, lab_data;
    std::vector<uint8_t> leaf, unary, lab_n;
    std::vector<int8_t> fixed;        // index slicing: -1 (free) or the value the label is fixed to
    std::vector<int32_t> fixed_idx;   // label -> its position in net.fixed_labels
    std::vector<int32_t> leaf_vertex, leaf_fa, leaf_fb;
    std::vector<int32_t> lo, hi, minpos, maxpos;  // leaf positions (depth-first) covered by a subtree / holding a label
    std::vector<uint8_t> is_open;
    std::vector<int32_t> stA, stB, stC;  // stamp arrays for O(1) membership
    std::vector<uint8_t> folded, kept;
    std::vector<int32_t> post0, topo;  // internal nodes, children before parents: the given tree / after the split-K rewrite
    std::vector<NodeCls> cls;
    std::vector<int32_t> cls_data;
    std::vector<int32_t> leaf_pool_off;
    std::vector<uint8_t> fus;
    std::vector<int64_t> peak;
    std::vector<int8_t> kind;
    std::vector<uint8_t> posA, posB, posCc;
    // pass-local arrays
    std::vector<int32_t> nint, first;                      // value_type_and_positions
    std::vector<int32_t> key, posC, batA, batB, secA, secB;  // layouts
    std::vector<uint8_t> tiny_shift;                       // layouts -> emit_fused: SubStep::a_shift | b_shift of tiny nodes
    std::vector<uint8_t> is_tiny;
    std::vector<int32_t> order, stack, where;              // emit_fused
    std::vector<int32_t> big_order, big_index;             // emit_big
};

// One compilation: the passes below run in order and share the arrays above.
struct PlanCompiler : CompilerArrays {
    const tb_network& net;
    uint32_t extra_flags;
    Plan& P;
    std::string& err;
    static CompilerArrays& thread_arrays() {
        static thread_local CompilerArrays a;
        return a;
    }
    PlanCompiler(const tb_network& n, uint32_t f, Plan& p, std::string& e) : CompilerArrays(std::move(thread_arrays())), net(n), extra_flags(f), P(p), err(e) {}
    ~PlanCompiler() { thread_arrays() = std::move(static_cast<CompilerArrays&>(*this)); }
    PlanCompiler(const PlanCompiler&) = delete;
    PlanCompiler& operator=(const PlanCompiler&) = delete;

    // ---- state shared by the passes
    bool temporary = false, estimate_only = false, synth = false, finished = false;
    int nL = 0, nN = 0, nT0 = 0, nTmax = 0, nT = 0, NLAB = 1, root = 0;
    int vt = 0, wd = 0;
    bool wide = false, half = false;
    static constexpr int TILE_M_MAX = GEMM_TILE_MAX;  // tile bits of the M side
    static constexpr int MT_LOG = 3;                  // log2 of a thread's microtile extent (8 x 8 outputs, both value widths)
    int STAGE_ELEMS = 0;
    double est_ops = 0, est_all = 0, est_peak = 0;
    int est_sc = 0;
    int stamp = 0;
    size_t lab_total = 0, lay_top = 0, cls_top = 0;
    int n_fused_roots = 0, n_big = 0;
    double ops_f = 0, ops_g = 0, ops_m = 0, bytes = 0, bytes_m = 0, sc = 0;

    int fail(int code, const std::string& m) {
        err = m;
        return code;
    }
    int32_t* labp(int t) { return lab_data.data() + lab_off[t]; }
    double weight_of(int v) const {
        switch (wd) {
            case TB_WEIGHT_UNIT: return 1.0;
            case TB_WEIGHT_I32: return (double)((const int32_t*)net.weights)[v];
            case TB_WEIGHT_I64: return (double)((const int64_t*)net.weights)[v];
            case TB_WEIGHT_F32: return (double)((const float*)net.weights)[v];
            default: return ((const double*)net.weights)[v];
        }
    }
    int rank_of(int t) const { return (int)P.lay_n[t]; }
    int64_t size_of(int t) const { return (int64_t)1 << P.lay_n[t]; }
    const int32_t* layp(int t) const { return P.lay_data.data() + P.lay_off[t]; }
    void set_pos(std::vector<uint8_t>& pos, int t, bool on) {
        const int32_t* v = layp(t);
        for (int i = 0; i < rank_of(t); ++i) pos[v[i]] = on ? (uint8_t)i : NO_BIT;
    }
    Plan::StepRec rec_of(int t, int knd, int lvl) const {
        const NodeCls& c = cls[t];
        Plan::StepRec r{};
        r.node = t;
        r.left = lch[t];
        r.right = rch[t];
        r.kind = (int8_t)knd;
        r.level = lvl;
        r.nm = c.nm;
        r.nn = c.nn;
        r.nb = c.nb;
        r.nk = c.nk;
        r.nka = c.nka;
        r.nkb = c.nkb;
        r.tm = c.tm;
        r.tn = c.tn;
        return r;
    }
    // algorithmic work of one step (ops by kernel kind, bytes the reference's tensors would move)
    void account(int t, int knd) {
        const NodeCls& c = cls[t];
        auto p2 = [](int e) { return (double)(1ull << e); };
        int tc = rank_of(t) + c.nk + c.nka + c.nkb - folded[t];
        const double eb = (double)Plan::elem_size_of(vt);
        if (unary[t]) {  // second half of a split node: engine overhead, not algorithmic work
            bytes += eb * p2(rank_of(t));
            return;
        }
        (knd == KIND_FUSED ? ops_f : knd == KIND_GENERIC ? ops_g : ops_m) += p2(tc);
        double cb = (t >= nT0) ? 0.0 : p2(rank_of(t));  // a partial node's output is not algorithmic traffic
        const double nb_ = eb * (p2(rank_of(lch[t])) + p2(rank_of(rch[t])) + p2(rank_of(t)));
        bytes += eb * (p2(rank_of(lch[t])) + p2(rank_of(rch[t])) + cb);
        if (knd == KIND_GEMM) bytes_m += nb_;  // what the kernel itself moves (a partial output included)
    }

    // validate the network, copy the tree into flat per-tensor arrays, apply index slicing to the leaves
    int load_network() {
        if (net.n_leaves < 1) return fail(TB_ERR_BAD_ARGUMENT, "network has no leaves (pass a NULL plan for an empty graph)");
        if (net.n_labels < 0) return fail(TB_ERR_BAD_ARGUMENT, "negative n_labels");
        if (!net.leaf_off || (net.leaf_off[net.n_leaves] > 0 && !net.leaf_labels))
            return fail(TB_ERR_BAD_ARGUMENT, "leaf_off / leaf_labels is NULL");
        if (net.n_leaves > 1 && (!net.node_left || !net.node_right))
            return fail(TB_ERR_BAD_ARGUMENT, "node_left / node_right is NULL");
        if (net.n_open < 0 || (net.n_open > 0 && !net.open_labels)) return fail(TB_ERR_BAD_ARGUMENT, "bad open labels");

        temporary = (extra_flags & TB_PLAN_TEMPORARY) != 0;
        P.temporary = temporary;
        estimate_only = (extra_flags & TB_PLAN_ESTIMATE_ONLY) != 0;
        extra_flags &= ~(TB_PLAN_TEMPORARY | TB_PLAN_ESTIMATE_ONLY);
        P.flags = (net.flags & ~(TB_PLAN_TEMPORARY | TB_PLAN_ESTIMATE_ONLY)) | extra_flags;
        if (P.flags & TB_PLAN_KEEP_INTERMEDIATES) P.flags |= TB_PLAN_NO_FUSED_SUBTREES;
        P.n_labels = net.n_labels;
        synth = (net.n_leaves == 1);
        nL = net.n_leaves + (synth ? 1 : 0);
        nN = nL - 1;
        nT0 = nL + nN;      // tensors of the given tree
        nTmax = nT0 + 2 * nN;  // + (partial node, unit leaf) per split
        P.n_leaves = nL;
        P.n_nodes = nN;
        NLAB = std::max(net.n_labels, 1);

        // ---- flat per-tensor storage
        lch.assign(nTmax, -1);
        rch.assign(nTmax, -1);
        parent.assign(nTmax, -1);
        leaf.assign(nTmax, 0);
        unary.assign(nTmax, 0);
        lab_off.assign(nTmax, 0);
        lab_n.assign(nTmax, 0);
        lab_data.clear();
        lab_data.reserve((size_t)nT0 * 8);

        // ---- index slicing: fixed[l] = -1 (free) or the value label l is fixed to
        fixed.clear();
        fixed_idx.clear();
        if (net.n_fixed < 0 || (net.n_fixed > 0 && (!net.fixed_labels || !net.fixed_values)))
            return fail(TB_ERR_BAD_ARGUMENT, "bad fixed labels");
        if (net.n_fixed > 0) {
            fixed.assign(NLAB, -1);
            fixed_idx.assign(NLAB, -1);
            for (int i = 0; i < net.n_fixed; ++i) {
                int32_t l = net.fixed_labels[i];
                if (l < 0 || l >= net.n_labels) return fail(TB_ERR_BAD_ARGUMENT, "fixed label out of range");
                if (fixed[l] >= 0) return fail(TB_ERR_BAD_ARGUMENT, "fixed label repeated");
                if (net.fixed_values[i] > 1) return fail(TB_ERR_BAD_ARGUMENT, "fixed value must be 0 or 1");
                fixed[l] = (int8_t)net.fixed_values[i];
                fixed_idx[l] = i;
            }
            for (int i = 0; i < net.n_open; ++i)
                if (net.open_labels[i] >= 0 && net.open_labels[i] < net.n_labels && fixed[net.open_labels[i]] >= 0)
                    return fail(TB_ERR_BAD_ARGUMENT, "a label cannot be both open and fixed");
        }

        // ---- leaves.  leaf_vertex[i] = vertex of a vertex leaf (-1: edge / unit leaf).  A leaf that lost labels to
        //      index slicing reads a slice of its tensor from a pool slot of its own, filled by Plan::assign from the
        //      values of its fixed labels (leaf_fa / leaf_fb = their positions in net.fixed_labels): all 2^k assignments
        //      of the same labels share every descriptor of the plan and differ only in these pool words.
        leaf_vertex.assign(nL, -1);
        leaf_fa.assign(nL, -1);
        leaf_fb.assign(nL, -1);
        for (int i = 0; i < nL; ++i) leaf[i] = 1;
        for (int i = 0; i < net.n_leaves; ++i) {
            int b = net.leaf_off[i], e = net.leaf_off[i + 1];
            if (e < b) return fail(TB_ERR_BAD_ARGUMENT, "leaf_off not monotone");
            int r = e - b;
            if (r < 1 || r > 2)
                return fail(TB_ERR_UNSUPPORTED, "leaf " + std::to_string(i) + " has " + std::to_string(r) +
                                                    " labels; IndependentSet leaves have 1 (vertex) or 2 (edge)");
            lab_off[i] = (int32_t)lab_data.size();
            int nfix = 0;  // fixed labels of this leaf
            for (int q = b; q < e; ++q) {
                int32_t l = net.leaf_labels[q];
                if (l < 0 || l >= net.n_labels) return fail(TB_ERR_BAD_ARGUMENT, "leaf label out of range");
                if (!fixed.empty() && fixed[l] >= 0) {
                    (nfix++ ? leaf_fb[i] : leaf_fa[i]) = fixed_idx[l];
                } else {
                    lab_data.push_back(l);
                }
            }
            if (r == 2 && net.leaf_labels[b] == net.leaf_labels[b + 1])
                return fail(TB_ERR_UNSUPPORTED, "edge tensor with a repeated label (self loop)");
            if (r == 1) leaf_vertex[i] = net.leaf_labels[b];
            lab_n[i] = (uint8_t)(r - nfix);
            if (lab_n[i] == 2) {
                int32_t* v = lab_data.data() + lab_off[i];
                if (v[0] > v[1]) std::swap(v[0], v[1]);
            }
        }
        if (synth) lab_off[1] = (int32_t)lab_data.size();

        // ---- tree
        if (synth) {
            lch[nL] = 0;
            rch[nL] = 1;
        } else {
            for (int j = 0; j < nN; ++j) {
                lch[nL + j] = net.node_left[j];
                rch[nL + j] = net.node_right[j];
            }
        }
        for (int j = 0; j < nN; ++j) {
            int id = nL + j;
            for (int c : {lch[id], rch[id]}) {
                if (c < 0 || c >= id) return fail(TB_ERR_NOT_BINARY_TREE, "node " + std::to_string(j) + ": child id must be in [0, own id)");
                if (parent[c] != -1) return fail(TB_ERR_NOT_BINARY_TREE, "tensor " + std::to_string(c) + " is used twice");
                parent[c] = id;
            }
            if (lch[id] == rch[id]) return fail(TB_ERR_NOT_BINARY_TREE, "node contracts a tensor with itself");
        }
        for (int t = 0; t < nT0 - 1; ++t)
            if (parent[t] == -1) return fail(TB_ERR_NOT_BINARY_TREE, "tensor " + std::to_string(t) + " is never contracted (forest, not a tree)");
        root = nT0 - 1;
        P.root_id = root;
        return TB_OK;
    }

    // resolve the value type (auto -> packed int16 when the weights allow), number the leaves depth-first
    int value_type_and_positions() {
        // ---- weights / value type
        vt = net.value_type;
        wd = net.weight_dtype;
        if (wd < TB_WEIGHT_UNIT || wd > TB_WEIGHT_F64) return fail(TB_ERR_BAD_ARGUMENT, "unknown weight_dtype");
        if (wd != TB_WEIGHT_UNIT && !net.weights) return fail(TB_ERR_BAD_ARGUMENT, "weights is NULL but weight_dtype is not UNIT");
        const bool int_weights = !(wd == TB_WEIGHT_F32 || wd == TB_WEIGHT_F64);
        bool auto_i16 = false;
        if (vt == TB_VALUE_AUTO) {
            vt = int_weights ? TB_VALUE_I32 : TB_VALUE_F32;
            auto_i16 = int_weights && !(P.flags & TB_PLAN_NO_I16);  // falls back to int32 below if the weights do not fit
        }
        if (vt != TB_VALUE_I32 && vt != TB_VALUE_F32 && vt != TB_VALUE_I16X2 && vt != TB_VALUE_F64 && vt != TB_VALUE_SIZE_CONFIG)
            return fail(TB_ERR_BAD_ARGUMENT, "unknown value_type");
        wide = (vt == TB_VALUE_F64 || vt == TB_VALUE_SIZE_CONFIG);  // 8-byte values: generic + fused kernels only
        if (wide) P.flags |= TB_PLAN_NO_GEMM;
        if (vt == TB_VALUE_SIZE_CONFIG && net.n_labels > 32)
            return fail(TB_ERR_UNSUPPORTED, "value type size+configuration keeps the chosen vertices in a 32-bit mask: at most 32 labels "
                                            "(regions of the branching tables have <= n_max = 20 vertices, src/types.jl:10)");
        if (vt == TB_VALUE_I16X2 || auto_i16) {
            // packed int16 needs every partial sum < 2^13: check sum |w| over the vertex leaves
            double sum_abs = 0;
            bool integral = true;
            for (int i = 0; i < net.n_leaves; ++i)
                if (leaf_vertex[i] >= 0) {
                    double w = weight_of(leaf_vertex[i]);
                    integral = integral && (w == std::floor(w));
                    sum_abs += std::fabs(w);
                }
            const bool fits = integral && sum_abs < 8192.0;
            if (vt == TB_VALUE_I16X2 && !fits) return fail(TB_ERR_UNSUPPORTED, "value type i16x2 needs integer weights with sum |w| < 8192");
            if (auto_i16 && fits) vt = TB_VALUE_I16X2;
        }
        P.value_type = vt;
        half = (vt == TB_VALUE_I16X2);
        STAGE_ELEMS = GEMM_STAGE_ELEMS * (half ? 2 : 1);

        // ---- leaf positions (DFS order) and subtree ranges
        lo.assign(nT0, 0);
        hi.assign(nT0, 0);
        // No traversal: ids are topological (children < parent, checked by load_network), so subtree sizes come from one
        // ascending sweep and the depth-first numbers (left operand first) from one descending sweep.  post0 = the
        // internal nodes in depth-first post-order: node t is emitted after the internal nodes of both subtrees.
        {
            nint.assign(nT0, 0);   // internal nodes in the subtree
            first.assign(nT0, 0);  // post-order index of the subtree's first internal node
            for (int i = 0; i < nL; ++i) hi[i] = 1;              // hi holds the leaf COUNT of the subtree during the sweeps
            for (int t = nL; t < nT0; ++t) {
                hi[t] = hi[lch[t]] + hi[rch[t]];
                nint[t] = nint[lch[t]] + nint[rch[t]] + 1;
            }
            post0.assign(nN, 0);
            for (int t = nT0 - 1; t >= nL; --t) {
                const int l = lch[t], r = rch[t];
                lo[l] = lo[t];
                lo[r] = lo[t] + hi[l];
                first[l] = first[t];
                first[r] = first[t] + nint[l];
                post0[first[t] + nint[t] - 1] = t;
            }
            for (int t = 0; t < nT0; ++t) hi[t] = lo[t] + hi[t] - 1;
        }
        minpos.assign(NLAB, std::numeric_limits<int32_t>::max());
        maxpos.assign(NLAB, -1);
        is_open.assign(NLAB, 0);
        for (int i = 0; i < nL; ++i)
            for (int q = 0; q < lab_n[i]; ++q) {
                int32_t l = labp(i)[q];
                minpos[l] = std::min(minpos[l], lo[i]);
                maxpos[l] = std::max(maxpos[l], lo[i]);
            }
        for (int i = 0; i < net.n_open; ++i) {
            int32_t l = net.open_labels[i];
            if (l < 0 || l >= net.n_labels || maxpos[l] < 0) return fail(TB_ERR_BAD_ARGUMENT, "open label does not occur in any leaf");
            if (is_open[l]) return fail(TB_ERR_BAD_ARGUMENT, "open label repeated");
            is_open[l] = 1;
        }
        return TB_OK;
    }

    // label set of every node, bottom-up; a label is reduced at the lowest node that covers all its leaves
    int label_sets() {
        // ---- label sets bottom-up (ids of the given tree are already topological); sets are sorted
        est_ops = 0;  // sum over nodes of 2^(labels involved): the reference's 2^tc (src/types.jl:120)
        est_sc = 0;
        for (int i = 0; i < nL; ++i) est_sc = std::max(est_sc, (int)lab_n[i]);
        size_t lab_top = lab_data.size();  // lab_data is grown in large steps; its size is set to lab_top at the end
        for (int t = nL; t < nT0; ++t) {
            const int A = lch[t], B = rch[t];
            const int na = lab_n[A], nb = lab_n[B];
            int32_t u[80];
            int nu = 0;
            {
                // union of two sorted sets; the comparison results feed the cursors, not branches
                const int32_t *a = labp(A), *b = labp(B);
                int i = 0, j = 0;
                while (i < na && j < nb) {
                    const int32_t x = a[i], y = b[j];
                    u[nu++] = x < y ? x : y;
                    i += x <= y;
                    j += y <= x;
                }
                while (i < na) u[nu++] = a[i++];
                while (j < nb) u[nu++] = b[j++];
            }
            if (nu > 62) return fail(TB_ERR_UNSUPPORTED, "a contraction involves more than 62 labels");
            const size_t at = lab_top;
            lab_off[t] = (int32_t)at;
            if (lab_data.size() < at + 64) lab_data.resize(std::max(2 * lab_data.size(), at + 1024));  // room for one more set
            int no = 0;
            {
                // a label is reduced here iff it is not open and all its leaves lie inside this subtree
                int32_t* w = lab_data.data() + at;
                const int32_t lo_t = lo[t], hi_t = hi[t];
                for (int q = 0; q < nu; ++q) {
                    const int32_t l = u[q];
                    const bool closed = !is_open[l] & (minpos[l] >= lo_t) & (maxpos[l] <= hi_t);
                    w[no] = l;
                    no += !closed;
                }
            }
            lab_top = at + (size_t)no;
            if (no > MAX_RANK) return fail(TB_ERR_UNSUPPORTED, "intermediate tensor of rank " + std::to_string(no) + " > 31");
            if (nu - no > 30) return fail(TB_ERR_UNSUPPORTED, "a contraction reduces more than 30 labels");
            lab_n[t] = (uint8_t)no;
            est_ops += (double)(1ull << nu);
            est_sc = std::max(est_sc, no);
        }
        lab_data.resize(lab_top);
        if (estimate_only) {
            P.stats = tb_plan_stats{};
            P.stats.ops = est_ops;
            P.stats.tc = est_ops > 0 ? std::log2(est_ops) : 0;
            P.stats.sc = est_sc;
            P.stats.n_nodes = nN;
            P.stats.value_type = vt;
            finished = true;
            return TB_OK;
        }

        // ---- the reference's memory estimators on the GIVEN tree (before any rewrite), in elements, all label sizes 2:
        //      contraction_all_memory = log2(sum of the sizes of all intermediates)          (src/utils.jl:222-229)
        //      contraction_peak_memory = log2(max of the running total)                       (src/utils.jl:197-219):
        //      the walk is depth-first, left operand first; a node adds its result and releases the operands of its
        //      CHILD nodes (its own operands are released one step later, by its parent) -- restated as the reference has it
        est_all = 0;
        est_peak = 0;
        return TB_OK;
    }

    // the reference's contraction_peak_memory / contraction_all_memory on the given tree
    int reference_estimators() {
        if (!temporary) {
            // the depth-first order (left operand first) is the order in which the leaf positions lo[] were numbered above:
            // a node is finished when its last leaf has been seen, so sorting is not needed -- walk the nodes by the stack
            double p2[33];
            for (int i = 0; i <= 32; ++i) p2[i] = (double)(1ull << i);
            double cur = 0;
            for (int i = 0; i < net.n_leaves; ++i) cur += p2[lab_n[i]];
            est_peak = cur;
            std::vector<double> freed_later(nT0, 0.0);  // sum of a node's operand sizes
            std::vector<int32_t> stack;
            stack.reserve(128);
            stack.push_back(root);
            while (!stack.empty()) {
                const int t = stack.back();
                if (t < 0) {  // second visit of node ~t: both operands are done
                    stack.pop_back();
                    const int x = ~t, A = lch[x], B = rch[x];
                    const double alloc = p2[lab_n[x]];
                    cur += alloc - (freed_later[A] + freed_later[B]);
                    freed_later[x] = p2[lab_n[A]] + p2[lab_n[B]];
                    if (cur > est_peak) est_peak = cur;
                    est_all += alloc;
                    continue;
                }
                stack.pop_back();
                if (leaf[t]) continue;
                stack.push_back(~t);
                stack.push_back(rch[t]);
                stack.push_back(lch[t]);
            }
        }
        return TB_OK;
    }

    // split long reductions into a partial step + a max over the kept labels (or fold it into the consumer)
    int split_k() {
        // stamp arrays for O(1) membership
        stA.assign(NLAB, -1);
        stB.assign(NLAB, -1);
        stC.assign(NLAB, -1);
        stamp = 0;

        // ---- split-K rewrite: node t = contract(A, B) with a long reduction becomes
        //      u = contract(A, B) keeping `sk` of the shared reduced labels, t = max over those labels of u
        //      (t = contract(u, unit scalar)).  Balances CTA run times inside a level launch.
        nT = nT0;
        folded.assign(nTmax, 0);  // split labels a node reduces on behalf of its children (not algorithmic work)
        kept.assign(nTmax, 0);    // split labels a node keeps as extra output labels for its consumer
        if (!(P.flags & TB_PLAN_NO_SPLIT_K)) {
            for (int t = nL; t < nT0; ++t) {
                const int A = lch[t], B = rch[t];
                if ((int)lab_n[A] + (int)lab_n[B] < 16) continue;  // cannot have a long reduction
                ++stamp;
                for (int q = 0; q < lab_n[B]; ++q) stB[labp(B)[q]] = stamp;
                for (int q = 0; q < lab_n[t]; ++q) stC[labp(t)[q]] = stamp;
                int nm = 0, nn = 0, nk = 0, nka = 0, nkb = 0;
                int32_t Ksh[40];
                for (int q = 0; q < lab_n[A]; ++q) {
                    int32_t l = labp(A)[q];
                    stA[l] = stamp;
                    bool inB = stB[l] == stamp, inC = stC[l] == stamp;
                    if (inB && !inC) Ksh[nk++] = l;
                    else if (!inB && inC) ++nm;
                    else if (!inB && !inC) ++nka;
                }
                for (int q = 0; q < lab_n[B]; ++q) {
                    int32_t l = labp(B)[q];
                    if (stA[l] != stamp) (stC[l] == stamp ? nn : nkb)++;
                }
                const int rc = lab_n[t];
                const int gm = nm, gn = nn;
                const bool gemm_like = !(P.flags & TB_PLAN_NO_GEMM) && gm >= MT_LOG && gn >= 3 &&
                                       std::min(gm, TILE_M_MAX) + std::min(gn, GEMM_TILE_MAX) >= MT_LOG + 6 && nk >= 1 && nka == 0 &&
                                       nkb == 0 && !leaf[A] && !leaf[B];
                int serial_log, limit;
                if (gemm_like) {
                    serial_log = nk;
                    limit = 8;
                } else {
                    int ks, po;
                    generic_split(rc, nk + nka + nkb, ks, po);
                    serial_log = nk + nka + nkb - ks;
                    limit = 7;
                }
                int sk = (serial_log > limit && nk > 0) ? std::min(serial_log - limit, gemm_like ? nk - 1 : nk) : 0;
                if (gemm_like) {
                    // parallelism: a node with few output tiles and a long reduction would occupy only a few CTAs of its
                    // level launch for a long time; split k until it has ~32 tiles (keeping >= 32 k-steps per tile).  Its
                    // output is small by construction, so the extra unary max pass is cheap.
                    const int tile_log = 14;
                    const int t_log = std::max(0, rc - tile_log);
                    static const int target_log = [] {
                        const char* e = getenv("TB_SPLIT_TARGET");  // log2 of the tiles a node should have; 0 disables
                        return e ? atoi(e) : 0;  // measured on cfg2 (profiles/s02_split_target_sweep.jsonl): with 4 lanes in flight extra levels cost more than idle CTA slots
                    }();
                    const int sk_par = std::min(std::max(0, target_log - t_log), std::max(0, nk - 5));
                    sk = std::max(sk, sk_par);
                }
                sk = std::min(sk, MAX_RANK - rc);
                if (sk <= 0) continue;
                // If the consumer of t is itself a reduction over (almost) all of t, it can reduce the split labels too:
                // t keeps them as output labels and they become labels private to one operand of the consumer.  The
                // consumer then reads the partial results once - exactly what the separate max pass would have read.
                static const bool no_fold = getenv("TB_NO_FOLD") != nullptr;  // diagnostics: keep the separate max pass
                if (const int pr = parent[t]; pr >= 0 && !no_fold) {
                    const int sib = lch[pr] == t ? rch[pr] : lch[pr];
                    ++stamp;
                    for (int q = 0; q < rc; ++q) stA[labp(t)[q]] = stamp;
                    for (int q = 0; q < lab_n[pr]; ++q) stC[labp(pr)[q]] = stamp;
                    int n_sib_out = 0;  // output labels of the consumer that only the sibling carries
                    for (int q = 0; q < lab_n[sib]; ++q) {
                        const int32_t l = labp(sib)[q];
                        if (stA[l] != stamp && stC[l] == stamp) ++n_sib_out;
                    }
                    if (n_sib_out <= 1 && (int)folded[pr] + sk <= 8) {
                        int32_t tmp[48];
                        int n = 0;
                        for (int q = 0; q < rc; ++q) tmp[n++] = labp(t)[q];
                        for (int q = nk - sk; q < nk; ++q) tmp[n++] = Ksh[q];
                        std::sort(tmp, tmp + n);
                        lab_off[t] = (int32_t)lab_data.size();
                        lab_n[t] = (uint8_t)n;
                        lab_data.insert(lab_data.end(), tmp, tmp + n);
                        folded[pr] = (uint8_t)(folded[pr] + sk);
                        kept[t] = (uint8_t)sk;
                        continue;
                    }
                }
                const int u = nT++, ul = nT++;
                lch[u] = A;
                rch[u] = B;
                parent[u] = t;
                folded[u] = folded[t];
                folded[t] = 0;
                // lab[u] = lab[t] + the last sk shared reduced labels, sorted
                {
                    int32_t tmp[48];
                    int n = 0;
                    for (int q = 0; q < rc; ++q) tmp[n++] = labp(t)[q];
                    for (int q = nk - sk; q < nk; ++q) tmp[n++] = Ksh[q];
                    std::sort(tmp, tmp + n);
                    lab_off[u] = (int32_t)lab_data.size();
                    lab_n[u] = (uint8_t)n;
                    lab_data.insert(lab_data.end(), tmp, tmp + n);
                }
                leaf[ul] = 1;
                parent[ul] = t;
                lab_off[ul] = (int32_t)lab_data.size();
                lab_n[ul] = 0;
                parent[A] = parent[B] = u;
                lch[t] = u;
                rch[t] = ul;
                unary[t] = 1;
            }
        }
        P.n_tensors = nT;
        return TB_OK;
    }

    // children before parents, depth-first (left operand first)
    int topological_order() {
        // ---- topological order of internal nodes (children before parents)
        topo.clear();
        topo.reserve(nT);
        // the split-K rewrite turns node t into t = max(u), u = contract(A, B): u directly precedes t
        for (int t : post0) {
            if (unary[t]) topo.push_back(lch[t]);
            topo.push_back(t);
        }
        return TB_OK;
    }

    // top-down: the consumer dictates the label order of both operands, so no step ever permutes
    int layouts() {
        // ---- layouts top-down (the consumer dictates the layout of both of its operands)
        P.lay_off.assign(nT, 0);
        P.lay_n.assign(nT, 0);
        lab_total = 0;
        for (int t = 0; t < nT; ++t) lab_total += lab_n[t];
        P.lay_data.assign(lab_total + (size_t)net.n_open + 64, 0);
        lay_top = 0;
        cls.assign(nT, NodeCls{});
        cls_data.assign(lab_total + 64, 0);  // every operand label falls in exactly one class of its consumer
        cls_top = 0;
        {
            if (net.n_open) {
                ++stamp;
                for (int q = 0; q < lab_n[root]; ++q) stC[labp(root)[q]] = stamp;
                if (net.n_open != lab_n[root]) return fail(TB_ERR_INTERNAL, "open labels do not match the root label set");
                for (int i = 0; i < net.n_open; ++i)
                    if (stC[net.open_labels[i]] != stamp) return fail(TB_ERR_INTERNAL, "open labels do not match the root label set");
                P.lay_off[root] = 0;
                P.lay_n[root] = (uint8_t)net.n_open;
                for (int i = 0; i < net.n_open; ++i) P.lay_data[lay_top++] = net.open_labels[i];
            }
            key.assign(NLAB, 0);  // sort key per label (valid for the node being processed)
            is_tiny.assign(nT, 0);
            tiny_shift.resize((size_t)nT * 32);
            // posC[l] = (stamp of the node being processed) * 64 + position of l in the node's output layout: a stale
            // stamp means "not an output label", so nothing has to be reset between nodes
            posC.assign(NLAB, -1);
            batA.assign(NLAB, -1);  // stamps: label is a batch label inside child A / B
            batB.assign(NLAB, -1);
            secA.assign(NLAB, -1);  // stamps: label belongs only to the child's SECOND operand
            secB.assign(NLAB, -1);
            const bool scramble = (P.flags & TB_PLAN_SCRAMBLE_LAYOUT) != 0;
            for (auto it = topo.rbegin(); it != topo.rend(); ++it) {
                const int t = *it;
                const int sN = ++stamp;  // stamp of this node (stA, stB, posC, batA, batB)
                const int32_t *inA = stA.data(), *inB = stB.data();  // inA[l] == sN: label l belongs to operand A
                {
                    const int L = lch[t], R = rch[t];
                    const int32_t *pl = labp(L), *pr = labp(R);
                    for (int q = 0, n = lab_n[L]; q < n; ++q) stA[pl[q]] = sN;
                    for (int q = 0, n = lab_n[R]; q < n; ++q) stB[pr[q]] = sN;
                    if (P.lay_n[t] > 0 && !leaf[L] && !leaf[R]) {
                        // orientation: the operand that owns the label at C bit 0 becomes the M side (contract(A,B) == contract(B,A)),
                        // so that the low output bits are m tile bits 0,1,.. and the epilogue can move whole 16-byte vectors
                        const int32_t l0 = P.lay_data[P.lay_off[t]];
                        if (stB[l0] == sN && stA[l0] != sN) {
                            std::swap(lch[t], rch[t]);
                            std::swap(inA, inB);
                        }
                    }
                }
                const int A = lch[t], B = rch[t];
                const int32_t* lc = P.lay_data.data() + P.lay_off[t];
                const int rc = P.lay_n[t];
                for (int i = 0; i < rc; ++i) posC[lc[i]] = sN * 64 + i;
                auto in_c = [&](int32_t l) { return (posC[l] >> 6) == sN; };
                auto pos_c = [&](int32_t l) { return posC[l] & 63; };
                // nodes whose tensors are all tiny end up inside fused subtrees (shared memory): label order is
                // irrelevant there, so skip the ordering analysis (90 % of all nodes)
                const bool tiny = rc <= 6 && lab_n[A] <= 6 && lab_n[B] <= 6;
                if (tiny && !scramble) {
                    // Fast path (no sorting, no data-dependent branches).  A tiny node is never a GEMM step and all its
                    // M / N labels are tile labels:  A = [M | K | KA | Bt],  B = [N | K | KB | Bt],  each class in label
                    // order.  One pass over B classifies its labels and yields every class size (every output label
                    // belongs to A or B, so nm = rc - nn - nb); A and B are then written straight into place.
                    static_assert(GEMM_TILE_MAX >= 6, "tiny nodes keep all M / N labels inside the tile");
                    const int32_t *pa = labp(A), *pb = labp(B);
                    const int ra = lab_n[A], rb = lab_n[B];
                    // While the labels are placed, their positions are recorded as the step's shift tables (position in A / B
                    // of the label at output bit p; SubStep::a_shift / b_shift); slot 63 swallows the labels that are reduced.
                    uint8_t kb[8];               // class of B's q-th label: 0 KB, 1 N, 2 K, 3 Bt
                    uint8_t pcb[8];              // its output bit (63: none)
                    uint8_t sa[64], sb[64];
                    std::memset(sa, NO_BIT, 16);
                    std::memset(sb, NO_BIT, 16);
                    int cnt[4] = {0, 0, 0, 0};
                    for (int q = 0; q < rb; ++q) {
                        const int32_t l = pb[q];
                        const int32_t pv = posC[l];
                        const int inC = (pv >> 6) == sN;
                        const int k = ((inA[l] == sN) << 1) | inC;
                        kb[q] = (uint8_t)k;
                        pcb[q] = (uint8_t)(inC ? (pv & 63) : 63);
                        ++cnt[k];
                    }
                    const int nkb = cnt[0], nn = cnt[1], nk = cnt[2], nb = cnt[3];
                    const int nm = rc - nn - nb, nka = ra - nm - nk - nb;
                    NodeCls& c = cls[t];
                    c.off = 0;  // the label classes of a tiny node are never read back (only GEMM steps do)
                    c.nm = (uint8_t)nm;
                    c.nn = (uint8_t)nn;
                    c.nb = (uint8_t)nb;
                    c.nk = (uint8_t)nk;
                    c.nka = (uint8_t)nka;
                    c.nkb = (uint8_t)nkb;
                    c.tm = (uint8_t)nm;
                    c.tn = (uint8_t)nn;
                    c.kfirst = 0;
                    P.lay_off[A] = (int32_t)lay_top;
                    P.lay_n[A] = (uint8_t)ra;
                    P.lay_off[B] = (int32_t)(lay_top + ra);
                    P.lay_n[B] = (uint8_t)rb;
                    int32_t* wa = P.lay_data.data() + lay_top;
                    int32_t* wb = wa + ra;
                    int32_t sink;
                    // write cursors by class (0 KA, 1 M, 2 K, 3 Bt) inside A, and of the shared classes inside B
                    int32_t* ca[4] = {wa + nm + nk, wa, wa + nm, wa + nm + nk + nka};
                    int32_t* cs[4] = {&sink, &sink, wb + nn, wb + nn + nk + nkb};
                    const int step[4] = {0, 0, 1, 1};
                    for (int q = 0; q < ra; ++q) {
                        const int32_t l = pa[q];
                        const int32_t pv = posC[l];
                        const int inC = (pv >> 6) == sN;
                        const int pc = inC ? (pv & 63) : 63;
                        const int k = ((inB[l] == sN) << 1) | inC;
                        int32_t* da = ca[k]++;
                        *da = l;
                        sa[pc] = (uint8_t)(da - wa);
                        int32_t* db = cs[k];
                        *db = l;
                        cs[k] += step[k];
                        sb[k == 3 ? pc : 63] = (uint8_t)(db - wb);
                    }
                    int32_t* cbw[4] = {wb + nn + nk, wb, &sink, &sink};  // KB, N; shared labels were written from A
                    const int stepb[4] = {1, 1, 0, 0};
                    for (int q = 0; q < rb; ++q) {
                        const int k = kb[q];
                        int32_t* db = cbw[k];
                        *db = pb[q];
                        cbw[k] += stepb[k];
                        sb[k == 1 ? pcb[q] : 63] = (uint8_t)(db - wb);
                    }
                    is_tiny[t] = 1;
                    std::memcpy(tiny_shift.data() + (size_t)t * 32, sa, 16);
                    std::memcpy(tiny_shift.data() + (size_t)t * 32 + 16, sb, 16);
                    lay_top += (size_t)(ra + rb);
                    continue;
                }
                // batch labels of the children (labels shared by a child's own operands)
                for (int side = 0; side < 2 && !tiny; ++side) {
                    const int ch = side ? B : A;
                    if (leaf[ch]) continue;
                    auto& bat = side ? batB : batA;
                    auto& sec = side ? secB : secA;
                    const int c1 = lch[ch], c2 = rch[ch];
                    const int sS = ++stamp;
                    for (int q = 0; q < lab_n[c1]; ++q) stC[labp(c1)[q]] = sS;
                    for (int q = 0; q < lab_n[c2]; ++q) {
                        const int32_t l = labp(c2)[q];
                        if (stC[l] == sS) bat[l] = sN;
                        else sec[l] = sN;
                    }
                }
                // label classes; the class index is computed from the membership bits, so no branch depends on the data
                int32_t cl[8][40];  // 1 M, 2 N, 3 Bt (output labels);  4 KA, 5 K, 6 KB (reduced);  0, 7: dropped
                int cn[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                // output labels in the order of their positions in C (every output label belongs to A or B) ...
                for (int i = 0; i < rc; ++i) {
                    const int32_t l = lc[i];
                    const int k = (int)(inA[l] == sN) | ((int)(inB[l] == sN) << 1);
                    cl[k][cn[k]++] = l;
                }
                // ... reduced labels in label order
                for (int q = 0, n = lab_n[A]; q < n; ++q) {
                    const int32_t l = labp(A)[q];
                    const int k = in_c(l) ? 7 : 4 + (int)(inB[l] == sN);
                    cl[k][cn[k]++] = l;
                }
                cn[7] = 0;
                for (int q = 0, n = lab_n[B]; q < n; ++q) {
                    const int32_t l = labp(B)[q];
                    const int k = (in_c(l) | (inA[l] == sN)) ? 7 : 6;
                    cl[k][cn[k]++] = l;
                }
                int32_t *M = cl[1], *N = cl[2], *Bt = cl[3], *K = cl[5], *KA = cl[4], *KB = cl[6];
                const int nm = cn[1], nn = cn[2], nb = cn[3], nk = cn[5], nka = cn[4], nkb = cn[6];
                // M / N: group the labels by their class inside the producing child so that the low address bits of
                // the operand form a run of the child's own tile labels (coalesced stores in the child): the child's
                // larger output-only class first, then its other one, batch labels of the child last; then by position in C
                auto class_key = [&](int32_t* v, int n, const std::vector<int32_t>& bat, const std::vector<int32_t>& sec) {
                    int n_first = 0, n_sec = 0;
                    for (int i = 0; i < n; ++i) {
                        if (bat[v[i]] == sN) continue;
                        (sec[v[i]] == sN ? n_sec : n_first)++;
                    }
                    const int k_first = n_first >= n_sec ? 0 : 1024, k_sec = n_first >= n_sec ? 1024 : 0;
                    for (int i = 0; i < n; ++i)
                        key[v[i]] = (bat[v[i]] == sN ? 2048 : (sec[v[i]] == sN ? k_sec : k_first)) + pos_c(v[i]);
                    sort_by_key(v, key.data(), n);
                };
                // tile set = the labels with the LOWEST positions in C (this node's own stores are enumerated in C order);
                // inside the tile set the order follows the child's classes (the child's stores), except that the label
                // with the highest C position goes last: it selects the epilogue round, so it must not be a low C bit
                auto order_side = [&](int32_t* v, int n, int tmax, const std::vector<int32_t>& bat, const std::vector<int32_t>& sec) {
                    const int tl = std::min(n, tmax);  // v arrives sorted by position in C
                    // the lowest `nlow` tile labels stay in C order (then a 16-byte output vector is contiguous in the
                    // staging buffer too); the others follow the producing child's classes
                    // labels that are batch labels of the producing child cannot be low output bits of that child: last
                    for (int i = 0; i < tl; ++i) key[v[i]] = (bat[v[i]] == sN ? 1024 : 0) + pos_c(v[i]);
                    sort_by_key(v, key.data(), tl);
                    int n_free = 0;
                    while (n_free < tl && bat[v[n_free]] != sN) ++n_free;
                    const int nlow = std::min(n_free, half ? 3 : 2);
                    if (n_free > nlow) class_key(v + nlow, n_free - nlow, bat, sec);
                    if (tl >= 2) {
                        int top = 0;
                        for (int i = 1; i < tl; ++i)
                            if (pos_c(v[i]) > pos_c(v[top])) top = i;
                        const int32_t lt = v[top];
                        for (int i = top; i + 1 < tl; ++i) v[i] = v[i + 1];
                        v[tl - 1] = lt;
                    }
                    if (n > tl) class_key(v + tl, n - tl, bat, sec);
                };
                if (!tiny) {
                    order_side(M, nm, TILE_M_MAX, batA, secA);
                    order_side(N, nn, GEMM_TILE_MAX, batB, secB);
                    for (int i = 0; i < nk; ++i) key[K[i]] = (batA[K[i]] == sN) + (batB[K[i]] == sN);
                    sort_by_key(K, key.data(), nk);
                }
                if (scramble) {
                    Lcg g((uint64_t)t * 977 + 13);
                    g.shuffle(M, nm);
                    g.shuffle(N, nn);
                    g.shuffle(Bt, nb);
                    g.shuffle(K, nk);
                    g.shuffle(KA, nka);
                    g.shuffle(KB, nkb);
                }
                NodeCls& c = cls[t];
                c.off = (int32_t)cls_top;
                c.nm = (uint8_t)nm;
                c.nn = (uint8_t)nn;
                c.nb = (uint8_t)nb;
                c.nk = (uint8_t)nk;
                c.nka = (uint8_t)nka;
                c.nkb = (uint8_t)nkb;
                c.tm = (uint8_t)std::min(nm, TILE_M_MAX);
                c.tn = (uint8_t)std::min(nn, GEMM_TILE_MAX);
                {
                    int32_t* w = cls_data.data() + cls_top;
                    for (int i = 0; i < nm; ++i) *w++ = M[i];
                    for (int i = 0; i < nn; ++i) *w++ = N[i];
                    for (int i = 0; i < nb; ++i) *w++ = Bt[i];
                    for (int i = 0; i < nk; ++i) *w++ = K[i];
                    for (int i = 0; i < nka; ++i) *w++ = KA[i];
                    for (int i = 0; i < nkb; ++i) *w++ = KB[i];
                    cls_top = (size_t)(w - cls_data.data());
                }
                // packed int16 GEMM: the two halves of a word are two consecutive k, so one shared reduced label becomes
                // address bit 0 of BOTH operands:  A = [K0 | M_lo | K_rest | M_hi | Bt],  B = [K0 | N_lo | K_rest | N_hi | Bt]
                c.kfirst = (half && !(P.flags & TB_PLAN_NO_GEMM) && nm >= MT_LOG && nn >= 3 && c.tm + c.tn >= MT_LOG + 6 && nk >= 1 &&
                            nka == 0 && nkb == 0 && !leaf[A] && !leaf[B])
                               ? 1
                               : 0;
                if (c.kfirst) {
                    const int ra = nm + nk + nb, rb = nn + nk + nb;
                    P.lay_off[A] = (int32_t)lay_top;
                    P.lay_n[A] = (uint8_t)ra;
                    P.lay_off[B] = (int32_t)(lay_top + ra);
                    P.lay_n[B] = (uint8_t)rb;
                    int32_t* w = P.lay_data.data() + lay_top;
                    *w++ = K[0];
                    for (int i = 0; i < c.tm; ++i) *w++ = M[i];
                    for (int i = 1; i < nk; ++i) *w++ = K[i];
                    for (int i = c.tm; i < nm; ++i) *w++ = M[i];
                    for (int i = 0; i < nb; ++i) *w++ = Bt[i];
                    *w++ = K[0];
                    for (int i = 0; i < c.tn; ++i) *w++ = N[i];
                    for (int i = 1; i < nk; ++i) *w++ = K[i];
                    for (int i = c.tn; i < nn; ++i) *w++ = N[i];
                    for (int i = 0; i < nb; ++i) *w++ = Bt[i];
                    lay_top += (size_t)(ra + rb);
                } else
                // A = [M_lo | K | KA | M_hi | Bt],  B = [N_lo | K | KB | N_hi | Bt]
                {
                    const int ra = nm + nk + nka + nb, rb = nn + nk + nkb + nb;
                    P.lay_off[A] = (int32_t)lay_top;
                    P.lay_n[A] = (uint8_t)ra;
                    P.lay_off[B] = (int32_t)(lay_top + ra);
                    P.lay_n[B] = (uint8_t)rb;
                    int32_t* w = P.lay_data.data() + lay_top;
                    for (int i = 0; i < c.tm; ++i) *w++ = M[i];
                    for (int i = 0; i < nk; ++i) *w++ = K[i];
                    for (int i = 0; i < nka; ++i) *w++ = KA[i];
                    for (int i = c.tm; i < nm; ++i) *w++ = M[i];
                    for (int i = 0; i < nb; ++i) *w++ = Bt[i];
                    for (int i = 0; i < c.tn; ++i) *w++ = N[i];
                    for (int i = 0; i < nk; ++i) *w++ = K[i];
                    for (int i = 0; i < nkb; ++i) *w++ = KB[i];
                    for (int i = c.tn; i < nn; ++i) *w++ = N[i];
                    for (int i = 0; i < nb; ++i) *w++ = Bt[i];
                    lay_top += (size_t)(ra + rb);
                }
            }
        }
        return TB_OK;
    }

    // leaf tensors (edge / unit / vertex values, slots of sliced leaves)
    int leaf_pool() {
        // ---- pool (leaf tensors)
        leaf_pool_off.assign(nT, 0);
        {
            auto push_val = [&](double x, bool neg_inf, int cfg_bit = -1) { P.pool.push_back(Plan::encode_value(vt, x, neg_inf, cfg_bit)); };
            const bool int_values = (vt == TB_VALUE_I32 || vt == TB_VALUE_I16X2 || vt == TB_VALUE_SIZE_CONFIG);
            P.pool.reserve(8 + 2 * (size_t)nL + 4);
            push_val(0, false); push_val(0, false); push_val(0, false); push_val(0, true);  // edge
            push_val(0, false);                                                            // unit
            push_val(0, false); push_val(0, false); push_val(0, false);                    // pad
            double sum_abs = 0;
            for (int i = 0; i < nT; ++i) {
                if (!leaf[i]) continue;
                const int vtx = i < nL ? leaf_vertex[i] : -1;
                if (i < nL && leaf_fa[i] >= 0) {
                    // sliced leaf: a 2-element slot of its own.  vertex [0, w][x] -> scalar; edge with one end fixed -> row x of
                    // the (symmetric) edge tensor, (0, 0) or (0, -inf); both ends fixed -> scalar, -inf iff both are 1
                    double w = 0;
                    if (vtx >= 0) {
                        w = weight_of(vtx);
                        if (std::isnan(w)) return fail(TB_ERR_BAD_ARGUMENT, "NaN weight");
                        if (int_values) {
                            if (w != std::floor(w)) return fail(TB_ERR_UNSUPPORTED, "integer value types need integer weights");
                            sum_abs += std::fabs(w);
                            if (sum_abs >= (double)(1 << 29)) return fail(TB_ERR_UNSUPPORTED, "sum of |weights| >= 2^29 overflows the i32 sentinel scheme");
                        }
                    }
                    Plan::PoolPatch pp{};
                    pp.vtx = vtx;
                    pp.off = (int32_t)P.pool.size();
                    pp.kind = (uint8_t)(vtx >= 0 ? 0 : (leaf_fb[i] < 0 ? 1 : 2));
                    pp.fa = leaf_fa[i];
                    pp.fb = leaf_fb[i];
                    pp.w = w;
                    P.patches.push_back(pp);
                    leaf_pool_off[i] = pp.off;
                    push_val(0, false);
                    push_val(0, false);
                } else if (vtx < 0 && lab_n[i] == 0) {
                    leaf_pool_off[i] = POOL_UNIT;
                } else if (vtx < 0) {
                    leaf_pool_off[i] = POOL_EDGE;
                } else {
                    double w = weight_of(vtx);
                    if (std::isnan(w)) return fail(TB_ERR_BAD_ARGUMENT, "NaN weight");
                    if (int_values) {
                        if (w != std::floor(w)) return fail(TB_ERR_UNSUPPORTED, "integer value types need integer weights");
                        sum_abs += std::fabs(w);
                        if (sum_abs >= (double)(1 << 29)) return fail(TB_ERR_UNSUPPORTED, "sum of |weights| >= 2^29 overflows the i32 sentinel scheme");
                    }
                    leaf_pool_off[i] = (int32_t)P.pool.size();
                    push_val(0, false);
                    push_val(w, false, vtx);  // size+configuration: choosing the vertex also sets its bit of the mask
                }
            }
            while (P.pool.size() % 4) P.pool.push_back(0);
            P.n_fixed = net.n_fixed;
            if (net.n_fixed > 0) P.assign(net.fixed_values);
            if (P.pool.size() > 65535) P.flags |= TB_PLAN_NO_FUSED_SUBTREES;  // fused steps address the pool with 16 bits
        }
        return TB_OK;
    }

    // fused subtree / generic / tiled GEMM per node, dependency levels of the big steps
    int kinds_and_levels() {
        // ---- kinds: fused subtrees / generic / gemm
        fus.assign(nT, 0);
        peak.assign(nT, 0);
        const bool allow_fused = !(P.flags & TB_PLAN_NO_FUSED_SUBTREES);
        for (int t : topo) {
            const int A = lch[t], B = rch[t];
            const NodeCls& c = cls[t];
            int tc = rank_of(t) + c.nk + c.nka + c.nkb;
            const bool ok = allow_fused & (rank_of(t) <= FUSED_MAX_RANK) & (rank_of(A) <= FUSED_MAX_RANK) &
                            (rank_of(B) <= FUSED_MAX_RANK) & (tc <= FUSED_MAX_TC) & ((leaf[A] | fus[A]) != 0) & ((leaf[B] | fus[B]) != 0);
            int64_t pA = leaf[A] ? 0 : peak[A], sA = leaf[A] ? 0 : size_of(A);
            int64_t pB = leaf[B] ? 0 : peak[B], sB = leaf[B] ? 0 : size_of(B);
            int64_t pk = size_of(t) + (pA >= pB ? std::max(pA, sA + pB) : std::max(pB, sB + pA));
            peak[t] = pk;
            fus[t] = ok & (pk <= (wide ? FUSED_SMEM_ELEMS / 2 : FUSED_SMEM_ELEMS));  // 32 KB of values either way
        }
        P.loc.assign(nT, LOC_ARENA);
        P.off.assign(nT, 0);
        P.level.assign(nT, -1);
        for (int i = 0; i < nT; ++i)
            if (leaf[i]) {
                P.loc[i] = LOC_POOL;
                P.off[i] = leaf_pool_off[i];
            }

        // ---- levels
        kind.assign(nT, -1);
        n_fused_roots = 0;
        n_big = 0;
        for (int t : topo) {
            if (fus[t]) {
                kind[t] = KIND_FUSED;
                bool is_sub_root = (t == root) || !fus[parent[t]];
                P.level[t] = is_sub_root ? 0 : -1;
                n_fused_roots += is_sub_root;
            } else {
                int lv = 1;
                for (int c : {lch[t], rch[t]})
                    if (!leaf[c]) lv = std::max(lv, P.level[c] + 1);
                P.level[t] = lv;
                const NodeCls& c = cls[t];
                static const int min_nk = [] {
                    const char* e = getenv("TB_GEMM_MIN_NK");  // experiments: shorter reductions run as generic steps
                    return e ? std::max(1, atoi(e)) : 1;
                }();
                bool gemm = !(P.flags & TB_PLAN_NO_GEMM) && c.nm >= MT_LOG && c.nn >= 3 && c.tm + c.tn >= MT_LOG + 6 && c.nk >= min_nk &&
                            c.nka == 0 && c.nkb == 0 && !leaf[lch[t]] && !leaf[rch[t]];
                kind[t] = gemm ? KIND_GEMM : KIND_GENERIC;
                P.n_levels = std::max(P.n_levels, lv);
                ++n_big;
            }
        }
        return TB_OK;
    }

    // arena offsets, safe for the dataflow executor (a node reuses only the dead interior of its own subtree)
    int arena() {
        // ---- arena allocation, safe for the dataflow executor: the only ordering between steps is operand -> consumer, so a
        //      node's output may overlap ONLY tensors that are dead once the node's operands are complete: the interior of its
        //      own subtree (everything below its two operands).  Sibling subtrees run concurrently and never share memory;
        //      fused-subtree roots (level 0, all produced by one launch) get fresh blocks.  Post-order walk; every tensor hands
        //      the free blocks of its subtree up to its consumer.
        {
            typedef std::vector<std::pair<int64_t, int64_t>> Blocks;  // (offset, size), sorted by offset, coalesced
            std::vector<Blocks> after(nT);  // free blocks inside the subtree of t once t is complete (t's own block excluded)
            const bool keep = (P.flags & TB_PLAN_KEEP_INTERMEDIATES) != 0;
            int64_t top = 0;
            auto insert_block = [](Blocks& bl, int64_t o, int64_t sz) {
                auto it = std::lower_bound(bl.begin(), bl.end(), std::make_pair(o, (int64_t)0));
                it = bl.insert(it, {o, sz});
                auto nx = it + 1;
                if (nx != bl.end() && it->first + it->second == nx->first) {
                    it->second += nx->second;
                    bl.erase(nx);
                }
                if (it != bl.begin()) {
                    auto pv = it - 1;
                    if (pv->first + pv->second == it->first) {
                        pv->second += it->second;
                        bl.erase(it);
                    }
                }
            };
            for (int t : topo) {
                if (P.level[t] < 0) continue;  // interior of a fused subtree: shared memory
                const int64_t need = align_up(size_of(t), 64);
                Blocks pool;
                if (!keep && kind[t] != KIND_FUSED)
                    for (int c : {lch[t], rch[t]})
                        if (!leaf[c] && P.level[c] >= 0) {
                            for (const auto& bk : after[c]) insert_block(pool, bk.first, bk.second);
                            Blocks().swap(after[c]);
                        }
                // best fit inside the subtree's dead interior, else a fresh block at the top of the arena
                int best = -1;
                for (int i = 0; i < (int)pool.size(); ++i)
                    if (pool[(size_t)i].second >= need && (best < 0 || pool[(size_t)i].second < pool[(size_t)best].second)) best = i;
                if (best >= 0) {
                    P.off[t] = pool[(size_t)best].first;
                    pool[(size_t)best].first += need;
                    pool[(size_t)best].second -= need;
                    if (pool[(size_t)best].second == 0) pool.erase(pool.begin() + best);
                } else {
                    P.off[t] = top;
                    top += need;
                }
                if (!keep && kind[t] != KIND_FUSED)
                    for (int c : {lch[t], rch[t]})
                        if (!leaf[c] && P.level[c] >= 0) insert_block(pool, P.off[c], align_up(size_of(c), 64));
                after[t].swap(pool);
            }
            P.arena_elems = top;
            P.root_off = P.off[root];
        }
        return TB_OK;
    }

    // descriptors of the fused subtrees (shared-memory stack allocation inside a subtree)
    int emit_fused() {
        // ---- emit steps
        posA.assign(NLAB, NO_BIT);
        posB.assign(NLAB, NO_BIT);
        posCc.assign(NLAB, NO_BIT);
        ops_f = ops_g = ops_m = bytes = bytes_m = sc = 0;
        for (int t = 0; t < nT0; ++t) sc = std::max(sc, (double)((int)lab_n[t] - (int)kept[t]));
        if (!temporary) P.recs.reserve(topo.size());
        P.sub_steps.reserve(topo.size() - n_big);
        P.subtrees.reserve(n_fused_roots);
        P.big_steps.reserve(n_big);

        // Fused subtrees.  Shared memory is a stack: a node's result sits at the bottom, the operand with the larger peak
        // is computed first, directly above it, the other operand above that one.  The offsets follow top-down from
        // that rule; the steps are the post-order (first operand's subtree, second operand's subtree, node), obtained
        // by reversing a pre-order walk that visits the second operand first.
        where.assign(NLAB, -1);  // where[l] = (stamp of the step) * 16 + position of l in the step's output
        for (int t : topo) {
            if (kind[t] != KIND_FUSED || P.level[t] != 0) continue;
            SubTree st{};
            st.first_step = (uint32_t)P.sub_steps.size();
            st.out_off = P.off[t];
            int64_t max_top = 0;
            order.clear();
            stack.clear();
            stack.push_back(t);
            while (!stack.empty()) {
                const int x = stack.back();
                stack.pop_back();
                order.push_back(x);
                const int A = lch[x], B = rch[x];
                const int64_t pA = leaf[A] ? 0 : peak[A], pB = leaf[B] ? 0 : peak[B];
                const int first = pA >= pB ? A : B, second = pA >= pB ? B : A;
                int64_t cur = x == t ? 0 : P.off[x] + size_of(x);  // the subtree's root is written to the arena
                if (!leaf[first]) {
                    P.loc[first] = LOC_SMEM;
                    P.off[first] = cur;
                    cur += size_of(first);
                    stack.push_back(first);
                }
                if (!leaf[second]) {
                    P.loc[second] = LOC_SMEM;
                    P.off[second] = cur;
                    cur += size_of(second);
                    stack.push_back(second);
                }
                max_top = std::max(max_top, cur);
            }
            for (size_t h = order.size(); h-- > 0;) {
                const int x = order[h], A = lch[x], B = rch[x];
                const NodeCls& c = cls[x];
                const bool is_root = x == t;
                SubStep s{};
                s.a_off = (uint16_t)P.off[A];
                s.b_off = (uint16_t)P.off[B];
                s.c_off = is_root ? 0 : (uint16_t)P.off[x];
                s.a_loc = (uint8_t)P.loc[A];
                s.b_loc = (uint8_t)P.loc[B];
                s.c_loc = is_root ? LOC_ARENA : LOC_SMEM;
                s.rc = (uint8_t)rank_of(x);
                s.nk = c.nk;
                s.nka = c.nka;
                s.nkb = c.nkb;
                s.sa = c.tm;
                s.sb = c.tn;
                s.pad = c.kfirst;  // reduction bit 0 is address bit 0 of both operands, the other K bits start at sa+1 / sb+1
                if (is_tiny[x]) {  // recorded while the layouts were written
                    std::memcpy(s.a_shift, tiny_shift.data() + (size_t)x * 32, 32);
                } else {
                    // position of each output label in A / B: stamp the (short) output, then walk A and B once; labels
                    // that are reduced land in a spare slot instead of taking a branch
                    uint8_t sh[2][32];
                    std::memset(sh, NO_BIT, sizeof sh);
                    const int sx = ++stamp;
                    const int32_t* lx = layp(x);
                    for (int i = 0, rx = rank_of(x); i < rx; ++i) where[lx[i]] = sx * 16 + i;
                    const int32_t* la_ = layp(A);
                    for (int i = 0, ra_ = rank_of(A); i < ra_; ++i) {
                        const int32_t wv = where[la_[i]];
                        sh[0][(wv >> 4) == sx ? (wv & 15) : 31] = (uint8_t)i;
                    }
                    const int32_t* lb_ = layp(B);
                    for (int i = 0, rb_ = rank_of(B); i < rb_; ++i) {
                        const int32_t wv = where[lb_[i]];
                        sh[1][(wv >> 4) == sx ? (wv & 15) : 31] = (uint8_t)i;
                    }
                    std::memcpy(s.a_shift, sh[0], sizeof s.a_shift);
                    std::memcpy(s.b_shift, sh[1], sizeof s.b_shift);
                }
                P.sub_steps.push_back(s);
                if (!temporary) P.recs.push_back(rec_of(x, KIND_FUSED, 0));
                account(x, KIND_FUSED);
            }
            st.n_steps = (uint32_t)P.sub_steps.size() - st.first_step;
            st.smem_elems = (uint32_t)max_top;
            P.subtrees.push_back(st);
        }
        return TB_OK;
    }

    // descriptors of the generic and GEMM steps, ordered by level, with their dependencies
    int emit_big() {
        // big steps by level
        {
            // stable counting sort of the big steps by level: big_level_begin[lv] = first step of level lv (levels from 1)
            std::vector<int32_t>& order = big_order;
            std::vector<int32_t>& cursor = stack;
            P.big_level_begin.assign(P.n_levels + 2, 0);
            for (int t : topo)
                if (kind[t] == KIND_GENERIC || kind[t] == KIND_GEMM) ++P.big_level_begin[P.level[t] + 1];
            for (int lv = 1; lv <= P.n_levels; ++lv) P.big_level_begin[lv + 1] += P.big_level_begin[lv];
            cursor.assign(P.big_level_begin.begin(), P.big_level_begin.end());
            order.assign(n_big, 0);
            for (int t : topo)
                if (kind[t] == KIND_GENERIC || kind[t] == KIND_GEMM) order[cursor[P.level[t]]++] = t;
            big_index.assign(nT, -1);  // node -> its position in big_steps (dependencies of the dataflow executor)
            P.big_dep_a.reserve(n_big);
            P.big_dep_b.reserve(n_big);
            for (int t : order) {
                const NodeCls& c = cls[t];
                const int32_t* cd = cls_data.data() + c.off;
                const int32_t *cM = cd, *cN = cd + c.nm, *cBt = cd + c.nm + c.nn;
                const int A = lch[t], B = rch[t];
                BigStep s{};
                s.a_off = P.off[A];
                s.b_off = P.off[B];
                s.c_off = P.off[t];
                s.a_loc = (uint8_t)P.loc[A];
                s.b_loc = (uint8_t)P.loc[B];
                s.kind = (uint8_t)kind[t];
                s.rc = (uint8_t)rank_of(t);
                s.nk = c.nk;
                s.nka = c.nka;
                s.nkb = c.nkb;
                s.sa = c.tm;
                s.sb = c.tn;
                s.tm = c.tm;
                s.tn = c.tn;
                std::memset(s.a_shift, NO_BIT, 32);
                std::memset(s.b_shift, NO_BIT, 32);
                std::memset(s.c_shift, NO_BIT, 32);
                if (kind[t] == KIND_GENERIC) {
                    set_pos(posA, A, true);
                    set_pos(posB, B, true);
                    const int32_t* lt = layp(t);
                    for (int i = 0; i < rank_of(t); ++i) {
                        s.a_shift[i] = posA[lt[i]];
                        s.b_shift[i] = posB[lt[i]];
                        s.c_shift[i] = (uint8_t)i;
                    }
                    set_pos(posA, A, false);
                    set_pos(posB, B, false);
                    s.store_mode = c.kfirst;  // generic steps reuse this byte: K0-first operand layouts
                    int ks, po;
                    generic_split(s.rc, s.nk + s.nka + s.nkb, ks, po);
                    if (ks == 0 && s.rc >= 10 && !wide) {  // streaming node: 4 consecutive outputs per thread and iteration
                        s.vec4 = 1;
                        po = s.rc >= 14 ? 12 : 10;  // 4096 outputs per CTA (4 iterations) amortise the per-CTA set-up
                    }
                    s.ks = (uint8_t)ks;
                    s.po = (uint8_t)po;
                    s.n_tiles = 1u << (s.rc - po);
                } else {
                    set_pos(posCc, t, true);
                    int q = 0;
                    for (int i = 0; i < c.tm; ++i) s.c_shift[q++] = posCc[cM[i]];
                    for (int i = 0; i < c.tn; ++i) s.c_shift[q++] = posCc[cN[i]];
                    for (int i = c.tm; i < c.nm; ++i) s.c_shift[q++] = posCc[cM[i]];
                    for (int i = c.tn; i < c.nn; ++i) s.c_shift[q++] = posCc[cN[i]];
                    for (int i = 0; i < c.nb; ++i) s.c_shift[q++] = posCc[cBt[i]];
                    set_pos(posCc, t, false);
                    s.n_mhi = (uint8_t)(c.nm - c.tm);
                    s.n_nhi = (uint8_t)(c.nn - c.tn);
                    s.ng = (uint8_t)(s.rc - c.tm - c.tn);
                    int tps_log = (c.tm - MT_LOG) + (c.tn - 3);  // threads per sub-tile
                    int s_log = 8 - tps_log;                     // sub-tiles per CTA (256 threads)
                    int64_t per_k = ((int64_t)1 << s_log) * (((int64_t)1 << c.tm) + ((int64_t)1 << c.tn));
                    int kc = 0;
                    while (kc + 1 <= s.nk && (per_k << (kc + 1)) <= STAGE_ELEMS) ++kc;
                    s.kc = (uint8_t)kc;
                    int64_t groups = (int64_t)1 << s.ng;
                    int64_t S = (int64_t)1 << s_log;
                    s.n_tiles = (uint32_t)((groups + S - 1) / S);
                    s.store_mode = STORE_SCALAR;
                    if (s.c_shift[0] == 0 && s.c_shift[1] == 1) s.store_mode = STORE_VEC_M;
                    else if (s.c_shift[c.tm] == 0 && s.c_shift[c.tm + 1] == 1) s.store_mode = STORE_VEC_N;
                    // staged epilogue tables (k_gemm2): the tile is written in 4 rounds (the top m tile bit and the top n
                    // tile bit select the round); inside a round the (tm-1)+(tn-1) remaining tile bits are enumerated in
                    // C-address order.  a_shift[i] = position of the i-th such bit in the shared-memory staging index
                    // ([m bits 0..tm-2 | n bits 0..tn-2]), b_shift[i] = its C shift.  a_shift[30/31] = C shift of the top
                    // m / n tile bit, b_shift[31] = 1 if the two lowest bits are C bits 0,1 (128-bit stores).
                    // packed int16 only: a consumer thread holds the m tile bits {0, mp, tm-1}; mp (a_shift[29], default
                    // 1) is chosen so that the m labels on C bits 0..2 are thread-local, and b_shift[29] = 1 swaps the
                    // roles of m0 and m_mp (the thread pairs outputs along m_mp).  The tables use LOGICAL m positions:
                    // [first local bit, second local bit, the other in-round m bits in ascending order].
                    {
                        const int nbr = c.tm + c.tn - 2;
                        int ent_cs[16], ent_sp[16], ne = 0;
                        auto build = [&](int mp, int mswap) -> int {
                            int lm[8], nl = 0;  // logical position of the in-round m bit q
                            lm[mswap ? mp : 0] = nl++;
                            if (c.tm >= 3) lm[mswap ? 0 : mp] = nl++;
                            for (int q = 1; q < c.tm - 1; ++q)
                                if (q != mp) lm[q] = nl++;
                            ne = 0;
                            for (int i = 0; i < c.tm - 1; ++i) { ent_cs[ne] = s.c_shift[i]; ent_sp[ne++] = lm[i]; }
                            for (int i = 0; i < c.tn - 1; ++i) { ent_cs[ne] = s.c_shift[c.tm + i]; ent_sp[ne++] = (c.tm - 1) + i; }
                            for (int i = 1; i < ne; ++i) {
                                int cs = ent_cs[i], sp = ent_sp[i], j = i - 1;
                                while (j >= 0 && ent_cs[j] > cs) { ent_cs[j + 1] = ent_cs[j]; ent_sp[j + 1] = ent_sp[j]; --j; }
                                ent_cs[j + 1] = cs; ent_sp[j + 1] = sp;
                            }
                            const bool evec = half ? (ent_cs[0] == 0 && ent_cs[1] == 1 && ent_cs[2] == 2) : (ent_cs[0] == 0 && ent_cs[1] == 1);
                            s.b_shift[31] = evec ? 1 : 0;
                            // ecase != 0: the elements of one 16-byte output vector are also contiguous in the staging
                            // buffer (one LDS.128 instead of 4 / 8 scalar loads).  int32: 1 = C bits 0,1 are m0,m1.
                            // packed int16: the thread holds (u, v) x (n0, n1) of a round (u, v = its local m bits), so it
                            // can write any of the orders C bits (0,1,2) = 1: (u,v,m2)  2: (u,v,n0)  3: (u,n0,v)
                            // 4: (u,n0,n1) as 16-byte vectors; for 2..4 the staging index is [the three C bits | the fourth
                            // thread-local bit | m2.. | n2..] and the table holds positions in THAT layout.
                            int ecase = 0;
                            if (half) {
                                const int n0 = c.tm - 1, n1 = c.tm;  // layout-1 staging positions of the n bits 0,1
                                if (evec && ent_sp[0] == 0) {
                                    if (ent_sp[1] == 1 && ent_sp[2] == 2) ecase = 1;
                                    else if (ent_sp[1] == 1 && ent_sp[2] == n0) ecase = 2;
                                    else if (ent_sp[1] == n0 && ent_sp[2] == 1) ecase = 3;
                                    else if (ent_sp[1] == n0 && ent_sp[2] == n1) ecase = 4;
                                }
                                if (ecase >= 2) {
                                    static const int low[5][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 1, 2, 3}, {0, 2, 1, 3}, {0, 3, 1, 2}};
                                    for (int i = 0; i < ne; ++i) {  // {u, v, n0, n1} -> low[ecase], m q>=2 -> q+2, n unchanged
                                        const int sp = ent_sp[i];
                                        if (sp == 0) ent_sp[i] = low[ecase][0];
                                        else if (sp == 1) ent_sp[i] = low[ecase][1];
                                        else if (sp == n0) ent_sp[i] = low[ecase][2];
                                        else if (sp == n1) ent_sp[i] = low[ecase][3];
                                        else if (sp < n0) ent_sp[i] = sp + 2;
                                    }
                                }
                            } else {
                                ecase = (ent_sp[0] == 0 && ent_sp[1] == 1) ? 1 : 0;
                            }
                            return ecase;
                        };
                        int mp = 1, mswap = 0;
                        if (half && c.tm >= 4) {
                            int q_by_cs[3] = {-1, -1, -1};
                            for (int q = 0; q < c.tm - 1; ++q)
                                if (s.c_shift[q] < 3) q_by_cs[s.c_shift[q]] = q;
                            const int q0 = q_by_cs[0];
                            if (q0 > 0) {
                                mp = q0;
                                mswap = 1;
                            } else if (q0 == 0) {
                                const int q1 = q_by_cs[1] > 0 ? q_by_cs[1] : q_by_cs[2];
                                if (q1 > 1) mp = q1;
                            }
                        }
                        int ecase = build(mp, mswap);
                        if (ecase == 0 && (mp != 1 || mswap)) {
                            mp = 1;
                            mswap = 0;
                            ecase = build(mp, mswap);
                        }
                        s.a_shift[29] = (uint8_t)mp;
                        s.b_shift[29] = (uint8_t)mswap;
                        s.a_shift[30] = s.c_shift[c.tm - 1];
                        s.a_shift[31] = s.c_shift[c.tm + c.tn - 1];
                        s.b_shift[30] = (uint8_t)ecase;
                        for (int i = 0; i < nbr; ++i) { s.a_shift[i] = (uint8_t)ent_sp[i]; s.b_shift[i] = (uint8_t)ent_cs[i]; }
                    }
                    // lanes of a warp should write neighbouring addresses: put the tile dimension that owns C's bit 2
                    // (bit 0 if stores are scalar) on the low lane bits
                    {
                        const int probe = s.store_mode == STORE_SCALAR ? 0 : 2;
                        bool n_owns = false;
                        for (int i = 0; i < c.tn; ++i)
                            if (s.c_shift[c.tm + i] == probe) n_owns = true;
                        s.lane_n_first = n_owns ? 1 : 0;
                    }
                }
                big_index[t] = (int32_t)P.big_steps.size();
                P.big_log2_ops.push_back((float)(rank_of(t) + c.nk + c.nka + c.nkb));
                P.big_bytes.push_back((double)Plan::elem_size_of(vt) * ((double)(1ull << rank_of(A)) + (double)(1ull << rank_of(B)) + (double)(1ull << rank_of(t))));
                P.big_dep_a.push_back(big_index[A]);  // -1 for leaves and fused subtrees
                P.big_dep_b.push_back(big_index[B]);
                P.big_steps.push_back(s);
                if (!temporary) P.recs.push_back(rec_of(t, kind[t], P.level[t]));
                account(t, kind[t]);
            }
        }
        return TB_OK;
    }

    // statistics of the compiled plan (tb_plan_info)
    int finish_stats() {
        tb_plan_stats& S = P.stats;
        S.sc = sc;
        S.ops = ops_f + ops_g + ops_m;
        S.tc = S.ops > 0 ? std::log2(S.ops) : 0;
        S.algo_bytes = bytes;
        S.arena_elems = P.arena_elems;
        S.n_nodes = nN;
        S.n_levels = P.n_levels;
        S.n_fused_subtrees = (int)P.subtrees.size();
        S.n_fused_steps = (int)P.sub_steps.size();
        S.n_gemm_steps = 0;
        S.n_generic_steps = 0;
        for (auto& b : P.big_steps) (b.kind == KIND_GEMM ? S.n_gemm_steps : S.n_generic_steps)++;
        S.value_type = vt;
        S.root_rank = rank_of(root);
        S.gemm_ops = ops_m;
        S.fused_ops = ops_f;
        S.generic_ops = ops_g;
        S.gemm_bytes = bytes_m;
        S.peak_memory_log2 = est_peak > 0 ? std::log2(est_peak) : 0;
        S.all_memory_log2 = est_all > 0 ? std::log2(est_all) : 0;
        return TB_OK;
    }

    int run() {
#ifdef TB_PLAN_PROFILE
        auto last_ = std::chrono::steady_clock::now();
#endif
        int rc;
        if ((rc = load_network()) != TB_OK || finished) return rc;
        PT(0, "validate")
        if ((rc = value_type_and_positions()) != TB_OK || finished) return rc;
        PT(1, "positions")
        if ((rc = label_sets()) != TB_OK || finished) return rc;
        PT(2, "labelsets")
        if ((rc = reference_estimators()) != TB_OK || finished) return rc;
        PT(11, "estimators")
        if ((rc = split_k()) != TB_OK || finished) return rc;
        PT(3, "splitk")
        if ((rc = topological_order()) != TB_OK || finished) return rc;
        PT(4, "topo")
        if ((rc = layouts()) != TB_OK || finished) return rc;
        PT(5, "layouts")
        if ((rc = leaf_pool()) != TB_OK || finished) return rc;
        PT(6, "pool")
        if ((rc = kinds_and_levels()) != TB_OK || finished) return rc;
        PT(7, "kinds")
        if ((rc = arena()) != TB_OK || finished) return rc;
        PT(8, "arena")
        if ((rc = emit_fused()) != TB_OK || finished) return rc;
        PT(9, "fused_emit")
        if ((rc = emit_big()) != TB_OK || finished) return rc;
        PT(10, "big_emit")
        return finish_stats();
    }
};

}  // namespace

int compile_plan(const tb_network& net, uint32_t extra_flags, Plan& P, std::string& err) {
    return PlanCompiler(net, extra_flags, P, err).run();
}

// Greedy choice of the labels to slice (tb_suggest_slices).  Works on the label sets of the given tree only:
// for every node the labels involved (union of the operands) and the labels it keeps; removing label l
// shrinks every set that holds it by one.  Each pick minimises (sc, number of tensors of rank sc, ops).
int suggest_slices(const tb_network& net, int sc_target, int max_sliced, int32_t* out_labels, double* out_sc,
                   double* out_tc, std::string& err) {
    auto fail = [&](int code, const std::string& m) {
        err = m;
        return code;
    };
    if (net.n_leaves < 1 || !net.leaf_off) return fail(TB_ERR_BAD_ARGUMENT, "network has no leaves");
    if (net.n_fixed != 0) return fail(TB_ERR_BAD_ARGUMENT, "network already carries fixed labels");
    if (max_sliced < 0 || (max_sliced > 0 && !out_labels)) return fail(TB_ERR_BAD_ARGUMENT, "bad max_sliced / out_labels");
    const int nL = net.n_leaves, nN = nL - 1, nT = nL + nN;
    if (nN > 0 && (!net.node_left || !net.node_right)) return fail(TB_ERR_BAD_ARGUMENT, "node_left / node_right is NULL");
    const int NLAB = std::max(net.n_labels, 1);
    // occurrences of each label over the leaves, and per subtree (small-to-large merging would be faster; trees here
    // have a few thousand nodes and the sets are short, so sorted-vector merges are enough)
    std::vector<int32_t> total(NLAB, 0);
    std::vector<std::vector<std::pair<int32_t, int32_t>>> cnt(nT);  // (label, occurrences in the subtree), open labels only
    std::vector<uint8_t> is_open(NLAB, 0);
    for (int i = 0; i < net.n_open; ++i) {
        int32_t l = net.open_labels[i];
        if (l < 0 || l >= net.n_labels) return fail(TB_ERR_BAD_ARGUMENT, "open label out of range");
        is_open[l] = 1;
    }
    for (int i = 0; i < nL; ++i) {
        int b = net.leaf_off[i], e = net.leaf_off[i + 1];
        if (e < b || e - b > 2) return fail(TB_ERR_UNSUPPORTED, "leaf with more than 2 labels");
        for (int q = b; q < e; ++q) {
            int32_t l = net.leaf_labels[q];
            if (l < 0 || l >= net.n_labels) return fail(TB_ERR_BAD_ARGUMENT, "leaf label out of range");
            ++total[l];
            cnt[i].push_back({l, 1});
        }
        std::sort(cnt[i].begin(), cnt[i].end());
    }
    // per node: involved labels (set_in) and kept labels (set_out), flattened
    std::vector<int32_t> in_off(nN + 1, 0), out_off(nN + 1, 0), in_data, out_data;
    std::vector<uint8_t> used(nT, 0);
    for (int j = 0; j < nN; ++j) {
        const int a = net.node_left[j], b = net.node_right[j], t = nL + j;
        if (a < 0 || a >= t || b < 0 || b >= t || a == b || used[a] || used[b])
            return fail(TB_ERR_NOT_BINARY_TREE, "tree arrays do not describe a binary tree over the leaves");
        used[a] = used[b] = 1;
        auto &ca = cnt[a], &cb = cnt[b];
        auto& ct = cnt[t];
        ct.reserve(ca.size() + cb.size());
        size_t x = 0, y = 0;
        while (x < ca.size() || y < cb.size()) {
            std::pair<int32_t, int32_t> m;
            if (y >= cb.size() || (x < ca.size() && ca[x].first < cb[y].first)) m = ca[x++];
            else if (x >= ca.size() || cb[y].first < ca[x].first) m = cb[y++];
            else {
                m = {ca[x].first, ca[x].second + cb[y].second};
                ++x;
                ++y;
            }
            in_data.push_back(m.first);
            if (m.second < total[m.first] || is_open[m.first]) {
                ct.push_back(m);
                out_data.push_back(m.first);
            }
        }
        in_off[j + 1] = (int32_t)in_data.size();
        out_off[j + 1] = (int32_t)out_data.size();
        std::vector<std::pair<int32_t, int32_t>>().swap(ca);
        std::vector<std::pair<int32_t, int32_t>>().swap(cb);
    }
    std::vector<uint8_t> removed(NLAB, 0);
    std::vector<int32_t> rank_out(std::max(nN, 1), 0), rank_in(std::max(nN, 1), 0);
    std::vector<int32_t> n_top(NLAB, 0);
    std::vector<double> gain(NLAB, 0.0);
    int n_picked = 0;
    double sc = 0, ops = 0;
    for (;;) {
        sc = 0;
        ops = 0;
        for (int i = 0; i < nL; ++i) {  // leaves count towards sc too (types.jl:121)
            int r = 0;
            for (int q = net.leaf_off[i]; q < net.leaf_off[i + 1]; ++q) r += !removed[net.leaf_labels[q]];
            sc = std::max(sc, (double)r);
        }
        for (int j = 0; j < nN; ++j) {
            int ro = 0, ri = 0;
            for (int q = out_off[j]; q < out_off[j + 1]; ++q) ro += !removed[out_data[q]];
            for (int q = in_off[j]; q < in_off[j + 1]; ++q) ri += !removed[in_data[q]];
            rank_out[j] = ro;
            rank_in[j] = ri;
            sc = std::max(sc, (double)ro);
            ops += std::ldexp(1.0, ri);
        }
        if (n_picked >= max_sliced || (sc_target >= 0 && sc <= sc_target)) break;
        std::fill(n_top.begin(), n_top.end(), 0);
        std::fill(gain.begin(), gain.end(), 0.0);
        int tops = 0;
        for (int j = 0; j < nN; ++j) {
            if (rank_out[j] == (int)sc) {
                ++tops;
                for (int q = out_off[j]; q < out_off[j + 1]; ++q)
                    if (!removed[out_data[q]]) ++n_top[out_data[q]];
            }
            const double half = std::ldexp(1.0, rank_in[j] - 1);
            for (int q = in_off[j]; q < in_off[j + 1]; ++q)
                if (!removed[in_data[q]]) gain[in_data[q]] += half;
        }
        int best = -1;
        for (int l = 0; l < net.n_labels; ++l) {
            if (removed[l] || is_open[l] || total[l] == 0) continue;
            // memory-driven (sc_target given): the label in most tensors of the top rank, ties by ops removed;
            // parallelism-driven (sc_target < 0): the label that removes most ops (least total overhead)
            const bool better = best < 0 || (sc_target >= 0
                                                 ? (n_top[l] > n_top[best] || (n_top[l] == n_top[best] && gain[l] > gain[best]))
                                                 : (gain[l] > gain[best] || (gain[l] == gain[best] && n_top[l] > n_top[best])));
            if (better) best = l;
        }
        if (best < 0 || (tops > 0 && n_top[best] == 0 && gain[best] == 0.0)) break;
        removed[best] = 1;
        out_labels[n_picked++] = best;
    }
    if (out_sc) *out_sc = sc;
    if (out_tc) *out_tc = ops > 0 ? std::log2(ops) : 0.0;
    return n_picked;
}

tb_step_info Plan::step_info(size_t i) const {
    const StepRec& r = recs[i];
    tb_step_info s{};
    s.node = r.node;
    s.left = r.left;
    s.right = r.right;
    s.kind = r.kind;
    s.level = r.level;
    s.rank_a = rank(r.left);
    s.rank_b = rank(r.right);
    s.rank_c = rank(r.node);
    s.n_m = r.nm;
    s.n_n = r.nn;
    s.n_b = r.nb;
    s.n_k = r.nk;
    s.n_ka = r.nka;
    s.n_kb = r.nkb;
    s.tile_m = r.tm;
    s.tile_n = r.tn;
    s.c_offset = off[r.node];
    for (int q = 0; q < 32; ++q) s.labels_a[q] = s.labels_b[q] = s.labels_c[q] = -1;
    for (int q = 0; q < s.rank_a; ++q) s.labels_a[q] = layout(r.left)[q];
    for (int q = 0; q < s.rank_b; ++q) s.labels_b[q] = layout(r.right)[q];
    for (int q = 0; q < s.rank_c; ++q) s.labels_c[q] = layout(r.node)[q];
    return s;
}

void Plan::recycle() {
    Plan fresh;
    // an array that is not listed here just loses its capacity
#define TB_KEEP(v) \
    v.clear();     \
    fresh.v.swap(v);
    TB_KEEP(lay_off) TB_KEEP(lay_n) TB_KEEP(lay_data) TB_KEEP(loc) TB_KEEP(off) TB_KEEP(level) TB_KEEP(pool) TB_KEEP(patches)
    TB_KEEP(sub_steps) TB_KEEP(subtrees) TB_KEEP(big_steps) TB_KEEP(big_level_begin) TB_KEEP(big_log2_ops) TB_KEEP(big_bytes)
    TB_KEEP(big_dep_a) TB_KEEP(big_dep_b) TB_KEEP(recs)
#undef TB_KEEP
    *this = std::move(fresh);
}

void Plan::copy_descriptors_to(Plan& dst) const {
    dst.n_labels = n_labels;
    dst.n_leaves = n_leaves;
    dst.n_nodes = n_nodes;
    dst.flags = flags;
    dst.temporary = temporary;
    dst.value_type = value_type;
    dst.pool = pool;
    dst.patches = patches;
    dst.n_fixed = n_fixed;
    dst.sub_steps = sub_steps;
    dst.subtrees = subtrees;
    dst.big_steps = big_steps;
    dst.big_level_begin = big_level_begin;
    dst.big_log2_ops = big_log2_ops;
    dst.big_bytes = big_bytes;
    dst.big_dep_a = big_dep_a;
    dst.big_dep_b = big_dep_b;
    dst.n_levels = n_levels;
    dst.arena_elems = arena_elems;
    dst.root_off = root_off;
    dst.root_id = root_id;
    dst.stats = stats;
    dst.n_tensors = 0;  // no tensor table: tb_read_tensor / tb_contract_tensor refuse such a plan
}

uint64_t Plan::encode_value(int vt, double x, bool neg_inf, int config_bit) {
    uint64_t bits = 0;
    if (vt == TB_VALUE_I32 || vt == TB_VALUE_I16X2) {
        int32_t v = neg_inf ? (vt == TB_VALUE_I16X2 ? Tropical<int16_t>::kNegInf : Tropical<int32_t>::kNegInf) : (int32_t)x;
        uint32_t b32;
        std::memcpy(&b32, &v, 4);
        bits = b32;
    } else if (vt == TB_VALUE_F32) {
        float v = neg_inf ? -std::numeric_limits<float>::infinity() : (float)x;
        uint32_t b32;
        std::memcpy(&b32, &v, 4);
        bits = b32;
    } else if (vt == TB_VALUE_F64) {
        double v = neg_inf ? -std::numeric_limits<double>::infinity() : x;
        std::memcpy(&bits, &v, 8);
    } else {  // TB_VALUE_SIZE_CONFIG: size << 32 | vertex mask; tropical zero = -2^62 (size -2^30, empty mask)
        int64_t v = neg_inf ? -((int64_t)1 << 62) : (int64_t)((uint64_t)(int64_t)x << 32);
        if (!neg_inf && config_bit >= 0) v |= (int64_t)1 << config_bit;
        std::memcpy(&bits, &v, 8);
    }
    return bits;
}

void Plan::assign(const uint8_t* values) {
    for (const PoolPatch& pp : patches) {
        const bool a = values[pp.fa] != 0, b = pp.fb >= 0 && values[pp.fb] != 0;
        switch (pp.kind) {
            case 0: pool[pp.off] = encode_value(value_type, a ? pp.w : 0.0, false, a ? pp.vtx : -1); break;
            case 1:
                pool[pp.off] = encode_value(value_type, 0.0, false);
                pool[pp.off + 1] = encode_value(value_type, 0.0, a);
                break;
            default: pool[pp.off] = encode_value(value_type, 0.0, a && b); break;
        }
    }
}

void Plan::write_pool(uint8_t* dst) const {
    const int es = elem_size();
    if (es == 8) {
        if (!pool.empty()) std::memcpy(dst, pool.data(), pool.size() * 8);
    } else if (es == 4) {
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        for (size_t i = 0; i < pool.size(); ++i) d[i] = (uint32_t)pool[i];
    } else {
        int16_t* d = reinterpret_cast<int16_t*>(dst);
        for (size_t i = 0; i < pool.size(); ++i) d[i] = (int16_t)(int32_t)(uint32_t)pool[i];
    }
}

void build_blob(Plan& P, std::vector<uint8_t>& blob) {
    auto a16 = [](size_t x) { return (x + 15) / 16 * 16; };
    size_t pool_b = a16(P.pool_bytes());
    size_t sub_b = a16(P.sub_steps.size() * sizeof(SubStep));
    size_t big_b = a16(P.big_steps.size() * sizeof(BigStep));
    P.sub_blob_off = pool_b;
    P.big_blob_off = pool_b + sub_b;
    P.blob_bytes = pool_b + sub_b + big_b;
    blob.assign(P.blob_bytes, 0);
    P.write_pool(blob.data());
    if (!P.sub_steps.empty()) std::memcpy(blob.data() + P.sub_blob_off, P.sub_steps.data(), P.sub_steps.size() * sizeof(SubStep));
    if (!P.big_steps.empty()) std::memcpy(blob.data() + P.big_blob_off, P.big_steps.data(), P.big_steps.size() * sizeof(BigStep));
}

}  // namespace tb
