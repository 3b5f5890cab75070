// plan.cpp -- compile one branch network (leaf label lists + binary contraction tree) into a flat,
// layout-resolved step list.  Replaces, for the hot path, what the reference does on every
// solve_slice call: uncompress -> parse_eincode -> decorate (/root/reference/src/types.jl:75-79) and
// OMEinsum's per-node label analysis + permutedims decisions [upstream].  Pure host C++.
#include "plan.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>

namespace tb {
namespace {

struct Classes {
    std::vector<int32_t> M, N, Bt, K, KA, KB;
    int tm = 0, tn = 0;
};

inline bool contains_sorted(const std::vector<int32_t>& v, int32_t x) {
    return std::binary_search(v.begin(), v.end(), x);
}

struct FreeList {
    // sorted, non-adjacent free blocks (offset, size) below `top`
    std::vector<std::pair<int64_t, int64_t>> blocks;
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
        // extend: if the last free block touches top, grow from it
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
        // merge with next
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
    template <typename V> void shuffle(V& v) {
        for (size_t i = v.size(); i > 1; --i) std::swap(v[i - 1], v[next() % i]);
    }
};

}  // namespace

int compile_plan(const tb_network& net, uint32_t extra_flags, Plan& P, std::string& err) {
    auto fail = [&](int code, const std::string& m) {
        err = m;
        return code;
    };
    if (net.n_leaves < 1) return fail(TB_ERR_BAD_ARGUMENT, "network has no leaves (pass a NULL plan for an empty graph)");
    if (net.n_labels < 0) return fail(TB_ERR_BAD_ARGUMENT, "negative n_labels");
    if (!net.leaf_off || (net.leaf_off[net.n_leaves] > 0 && !net.leaf_labels))
        return fail(TB_ERR_BAD_ARGUMENT, "leaf_off / leaf_labels is NULL");
    if (net.n_leaves > 1 && (!net.node_left || !net.node_right))
        return fail(TB_ERR_BAD_ARGUMENT, "node_left / node_right is NULL");
    if (net.n_open < 0 || (net.n_open > 0 && !net.open_labels)) return fail(TB_ERR_BAD_ARGUMENT, "bad open labels");

    P.flags = net.flags | extra_flags;
    if (P.flags & TB_PLAN_KEEP_INTERMEDIATES) P.flags |= TB_PLAN_NO_FUSED_SUBTREES;
    P.n_labels = net.n_labels;
    const bool synth = (net.n_leaves == 1);
    const int nL = net.n_leaves + (synth ? 1 : 0);
    const int nN = nL - 1;
    const int nT = nL + nN;
    P.n_leaves = nL;
    P.n_nodes = nN;

    // ---- leaves
    std::vector<std::vector<int32_t>> lab(nT);  // sorted label sets
    for (int i = 0; i < net.n_leaves; ++i) {
        int b = net.leaf_off[i], e = net.leaf_off[i + 1];
        if (e < b) return fail(TB_ERR_BAD_ARGUMENT, "leaf_off not monotone");
        int r = e - b;
        if (r < 1 || r > 2)
            return fail(TB_ERR_UNSUPPORTED, "leaf " + std::to_string(i) + " has " + std::to_string(r) +
                                                " labels; IndependentSet leaves have 1 (vertex) or 2 (edge)");
        for (int q = b; q < e; ++q) {
            int32_t l = net.leaf_labels[q];
            if (l < 0 || l >= net.n_labels) return fail(TB_ERR_BAD_ARGUMENT, "leaf label out of range");
            lab[i].push_back(l);
        }
        if (r == 2 && lab[i][0] == lab[i][1]) return fail(TB_ERR_UNSUPPORTED, "edge tensor with a repeated label (self loop)");
        std::sort(lab[i].begin(), lab[i].end());
    }

    // ---- tree
    std::vector<int32_t> left(nN), right(nN), parent(nT, -1);
    if (synth) {
        left[0] = 0;
        right[0] = 1;
    } else {
        for (int j = 0; j < nN; ++j) {
            left[j] = net.node_left[j];
            right[j] = net.node_right[j];
        }
    }
    for (int j = 0; j < nN; ++j) {
        int id = nL + j;
        for (int c : {left[j], right[j]}) {
            if (c < 0 || c >= id) return fail(TB_ERR_NOT_BINARY_TREE, "node " + std::to_string(j) + ": child id must be in [0, own id)");
            if (parent[c] != -1) return fail(TB_ERR_NOT_BINARY_TREE, "tensor " + std::to_string(c) + " is used twice");
            parent[c] = id;
        }
        if (left[j] == right[j]) return fail(TB_ERR_NOT_BINARY_TREE, "node contracts a tensor with itself");
    }
    for (int t = 0; t < nT - 1; ++t)
        if (parent[t] == -1) return fail(TB_ERR_NOT_BINARY_TREE, "tensor " + std::to_string(t) + " is never contracted (forest, not a tree)");
    const int root = nT - 1;
    P.root_id = root;
    auto is_leaf = [&](int t) { return t < nL; };
    auto L = [&](int t) { return left[t - nL]; };
    auto R = [&](int t) { return right[t - nL]; };

    // ---- weights / value type
    int vt = net.value_type;
    const int wd = net.weight_dtype;
    if (wd != TB_WEIGHT_UNIT && !net.weights) return fail(TB_ERR_BAD_ARGUMENT, "weights is NULL but weight_dtype is not UNIT");
    auto weight_of = [&](int v) -> double {
        switch (wd) {
            case TB_WEIGHT_UNIT: return 1.0;
            case TB_WEIGHT_I32: return (double)((const int32_t*)net.weights)[v];
            case TB_WEIGHT_I64: return (double)((const int64_t*)net.weights)[v];
            case TB_WEIGHT_F32: return (double)((const float*)net.weights)[v];
            case TB_WEIGHT_F64: return ((const double*)net.weights)[v];
            default: return std::numeric_limits<double>::quiet_NaN();
        }
    };
    if (wd < TB_WEIGHT_UNIT || wd > TB_WEIGHT_F64) return fail(TB_ERR_BAD_ARGUMENT, "unknown weight_dtype");
    if (vt == TB_VALUE_AUTO) vt = (wd == TB_WEIGHT_F32 || wd == TB_WEIGHT_F64) ? TB_VALUE_F32 : TB_VALUE_I32;
    if (vt == TB_VALUE_I16X2) return fail(TB_ERR_UNSUPPORTED, "value type i16x2 is reserved (not implemented)");
    if (vt != TB_VALUE_I32 && vt != TB_VALUE_F32) return fail(TB_ERR_BAD_ARGUMENT, "unknown value_type");
    P.value_type = vt;

    // ---- leaf positions (DFS order) and subtree ranges
    std::vector<int32_t> lo(nT, 0), hi(nT, 0);
    {
        std::vector<int32_t> stack{root};
        int pos = 0;
        std::vector<int32_t> leafpos(nL, -1);
        while (!stack.empty()) {
            int t = stack.back();
            stack.pop_back();
            if (is_leaf(t)) {
                leafpos[t] = pos++;
            } else {
                stack.push_back(R(t));
                stack.push_back(L(t));
            }
        }
        for (int t = 0; t < nL; ++t) lo[t] = hi[t] = leafpos[t];
        for (int t = nL; t < nT; ++t) {
            lo[t] = std::min(lo[L(t)], lo[R(t)]);
            hi[t] = std::max(hi[L(t)], hi[R(t)]);
        }
    }
    std::vector<int32_t> minpos(std::max(net.n_labels, 1), std::numeric_limits<int32_t>::max());
    std::vector<int32_t> maxpos(std::max(net.n_labels, 1), -1);
    std::vector<uint8_t> is_open(std::max(net.n_labels, 1), 0);
    for (int i = 0; i < nL; ++i)
        for (int32_t l : lab[i]) {
            minpos[l] = std::min(minpos[l], lo[i]);
            maxpos[l] = std::max(maxpos[l], lo[i]);
        }
    for (int i = 0; i < net.n_open; ++i) {
        int32_t l = net.open_labels[i];
        if (l < 0 || l >= net.n_labels || maxpos[l] < 0) return fail(TB_ERR_BAD_ARGUMENT, "open label does not occur in any leaf");
        if (is_open[l]) return fail(TB_ERR_BAD_ARGUMENT, "open label repeated");
        is_open[l] = 1;
    }

    // ---- label sets bottom-up
    for (int t = nL; t < nT; ++t) {
        const auto &a = lab[L(t)], &b = lab[R(t)];
        std::vector<int32_t> u;
        u.reserve(a.size() + b.size());
        std::set_union(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(u));
        if ((int)u.size() > 40) return fail(TB_ERR_UNSUPPORTED, "a contraction involves more than 40 labels");
        auto& o = lab[t];
        for (int32_t l : u) {
            bool closed = !is_open[l] && minpos[l] >= lo[t] && maxpos[l] <= hi[t];
            if (!closed) o.push_back(l);
        }
        if ((int)o.size() > MAX_RANK) return fail(TB_ERR_UNSUPPORTED, "intermediate tensor of rank " + std::to_string(o.size()) + " > 31");
        if ((int)(u.size() - o.size()) > 30) return fail(TB_ERR_UNSUPPORTED, "a contraction reduces more than 30 labels");
    }

    // ---- layouts top-down
    P.layout.assign(nT, {});
    std::vector<Classes> cls(nT);
    {
        if (net.n_open) {
            std::vector<int32_t> o(net.open_labels, net.open_labels + net.n_open);
            std::vector<int32_t> s = o;
            std::sort(s.begin(), s.end());
            if (s != lab[root]) return fail(TB_ERR_INTERNAL, "open labels do not match the root label set");
            P.layout[root] = o;
        }
        std::vector<int32_t> posC(std::max(net.n_labels, 1), -1);
        auto batch_in = [&](int child, int32_t l) -> int {
            if (is_leaf(child)) return 0;
            return contains_sorted(lab[L(child)], l) && contains_sorted(lab[R(child)], l);
        };
        for (int t = nT - 1; t >= nL; --t) {
            const int A = L(t), B = R(t);
            const auto& lc = P.layout[t];
            for (size_t i = 0; i < lc.size(); ++i) posC[lc[i]] = (int)i;
            Classes& c = cls[t];
            for (int32_t l : lab[A]) {
                bool inB = contains_sorted(lab[B], l), inC = posC[l] >= 0;
                if (inB) (inC ? c.Bt : c.K).push_back(l);
                else (inC ? c.M : c.KA).push_back(l);
            }
            for (int32_t l : lab[B]) {
                if (contains_sorted(lab[A], l)) continue;
                (posC[l] >= 0 ? c.N : c.KB).push_back(l);
            }
            auto key_sort = [&](std::vector<int32_t>& v, int child) {
                std::sort(v.begin(), v.end(), [&](int32_t x, int32_t y) {
                    int bx = child >= 0 ? batch_in(child, x) : 0, by = child >= 0 ? batch_in(child, y) : 0;
                    if (bx != by) return bx < by;
                    return posC[x] < posC[y];
                });
            };
            key_sort(c.M, A);
            key_sort(c.N, B);
            key_sort(c.Bt, -1);
            std::sort(c.K.begin(), c.K.end(), [&](int32_t x, int32_t y) {
                int bx = batch_in(A, x) + batch_in(B, x), by = batch_in(A, y) + batch_in(B, y);
                if (bx != by) return bx < by;
                return x < y;
            });
            if (P.flags & TB_PLAN_SCRAMBLE_LAYOUT) {
                Lcg g((uint64_t)t * 977 + 13);
                g.shuffle(c.M);
                g.shuffle(c.N);
                g.shuffle(c.Bt);
                g.shuffle(c.K);
                g.shuffle(c.KA);
                g.shuffle(c.KB);
            }
            c.tm = std::min<int>((int)c.M.size(), GEMM_TILE_MAX);
            c.tn = std::min<int>((int)c.N.size(), GEMM_TILE_MAX);
            auto& la = P.layout[A];
            la.assign(c.M.begin(), c.M.begin() + c.tm);
            la.insert(la.end(), c.K.begin(), c.K.end());
            la.insert(la.end(), c.KA.begin(), c.KA.end());
            la.insert(la.end(), c.M.begin() + c.tm, c.M.end());
            la.insert(la.end(), c.Bt.begin(), c.Bt.end());
            auto& lb = P.layout[B];
            lb.assign(c.N.begin(), c.N.begin() + c.tn);
            lb.insert(lb.end(), c.K.begin(), c.K.end());
            lb.insert(lb.end(), c.KB.begin(), c.KB.end());
            lb.insert(lb.end(), c.N.begin() + c.tn, c.N.end());
            lb.insert(lb.end(), c.Bt.begin(), c.Bt.end());
            for (int32_t l : lc) posC[l] = -1;
        }
    }
    auto rank_of = [&](int t) { return (int)P.layout[t].size(); };
    auto size_of = [&](int t) { return (int64_t)1 << rank_of(t); };

    // ---- pool (leaf tensors)
    std::vector<int64_t> leaf_pool_off(nL, 0);
    {
        auto push_val = [&](double x, bool neg_inf) {
            uint32_t bits;
            if (vt == TB_VALUE_I32) {
                int32_t v = neg_inf ? Tropical<int32_t>::kNegInf : (int32_t)x;
                std::memcpy(&bits, &v, 4);
            } else {
                float v = neg_inf ? -std::numeric_limits<float>::infinity() : (float)x;
                std::memcpy(&bits, &v, 4);
            }
            P.pool.push_back(bits);
        };
        push_val(0, false); push_val(0, false); push_val(0, false); push_val(0, true);  // edge
        push_val(0, false);                                                            // unit
        push_val(0, false); push_val(0, false); push_val(0, false);                    // pad
        double sum_abs = 0;
        for (int i = 0; i < nL; ++i) {
            if (synth && i == 1) {
                leaf_pool_off[i] = POOL_UNIT;
            } else if (lab[i].size() == 2) {
                leaf_pool_off[i] = POOL_EDGE;
            } else {
                double w = weight_of(lab[i][0]);
                if (std::isnan(w)) return fail(TB_ERR_BAD_ARGUMENT, "NaN weight");
                if (vt == TB_VALUE_I32) {
                    if (w != std::floor(w)) return fail(TB_ERR_UNSUPPORTED, "value type i32 needs integer weights");
                    sum_abs += std::fabs(w);
                    if (sum_abs >= (double)(1 << 29)) return fail(TB_ERR_UNSUPPORTED, "sum of |weights| >= 2^29 overflows the i32 sentinel scheme");
                }
                leaf_pool_off[i] = (int64_t)P.pool.size();
                push_val(0, false);
                push_val(w, false);
            }
        }
        while (P.pool.size() % 4) P.pool.push_back(0);
        if (P.pool.size() > 65535) {
            // fused steps address the pool with 16 bits; larger pools simply disable fusion
            P.flags |= TB_PLAN_NO_FUSED_SUBTREES;
        }
    }

    // ---- kinds: fused subtrees / generic / gemm
    std::vector<uint8_t> fus(nT, 0);
    std::vector<int64_t> peak(nT, 0);
    const bool allow_fused = !(P.flags & TB_PLAN_NO_FUSED_SUBTREES);
    for (int t = nL; t < nT; ++t) {
        const int A = L(t), B = R(t);
        const Classes& c = cls[t];
        int tc = rank_of(t) + (int)(c.K.size() + c.KA.size() + c.KB.size());
        bool ok = allow_fused && rank_of(t) <= FUSED_MAX_RANK && rank_of(A) <= FUSED_MAX_RANK &&
                  rank_of(B) <= FUSED_MAX_RANK && tc <= FUSED_MAX_TC && (is_leaf(A) || fus[A]) && (is_leaf(B) || fus[B]);
        int64_t pA = is_leaf(A) ? 0 : peak[A], sA = is_leaf(A) ? 0 : size_of(A);
        int64_t pB = is_leaf(B) ? 0 : peak[B], sB = is_leaf(B) ? 0 : size_of(B);
        int64_t pk = size_of(t) + (pA >= pB ? std::max(pA, sA + pB) : std::max(pB, sB + pA));
        peak[t] = pk;
        fus[t] = ok && pk <= FUSED_SMEM_ELEMS;
    }
    P.loc.assign(nT, LOC_ARENA);
    P.off.assign(nT, 0);
    P.level.assign(nT, -1);
    for (int i = 0; i < nL; ++i) {
        P.loc[i] = LOC_POOL;
        P.off[i] = leaf_pool_off[i];
    }

    // ---- levels
    std::vector<int> kind(nT, -1);
    for (int t = nL; t < nT; ++t) {
        if (fus[t]) {
            kind[t] = KIND_FUSED;
            bool is_sub_root = (t == root) || !fus[parent[t]];
            P.level[t] = is_sub_root ? 0 : -1;
        } else {
            int lv = 1;
            for (int c : {L(t), R(t)})
                if (!is_leaf(c)) lv = std::max(lv, P.level[c] + 1);
            P.level[t] = lv;
            const Classes& c = cls[t];
            bool gemm = !(P.flags & TB_PLAN_NO_GEMM) && c.M.size() >= 3 && c.N.size() >= 3 && c.tm + c.tn >= 9 &&
                        c.K.size() >= 1 && c.KA.empty() && c.KB.empty() && !is_leaf(L(t)) && !is_leaf(R(t));
            kind[t] = gemm ? KIND_GEMM : KIND_GENERIC;
            P.n_levels = std::max(P.n_levels, lv);
        }
    }

    // ---- arena allocation by level intervals
    {
        std::vector<std::vector<int>> born(P.n_levels + 1), dies(P.n_levels + 2);
        for (int t = nL; t < nT; ++t) {
            if (P.level[t] < 0) continue;
            born[P.level[t]].push_back(t);
            if (t != root) dies[P.level[parent[t]]].push_back(t);
        }
        FreeList fl;
        const bool keep = (P.flags & TB_PLAN_KEEP_INTERMEDIATES) != 0;
        for (int lv = 0; lv <= P.n_levels; ++lv) {
            auto& bs = born[lv];
            std::sort(bs.begin(), bs.end(), [&](int x, int y) { return size_of(x) != size_of(y) ? size_of(x) > size_of(y) : x < y; });
            for (int t : bs) P.off[t] = fl.alloc(align_up(size_of(t), 64));
            if (!keep)
                for (int t : dies[lv]) fl.release(P.off[t], align_up(size_of(t), 64));
        }
        P.arena_elems = fl.top;
        P.root_off = P.off[root];
    }

    // ---- emit steps
    auto fill_info = [&](int t, int knd, int lvl) {
        tb_step_info s{};
        const Classes& c = cls[t];
        s.node = t;
        s.left = L(t);
        s.right = R(t);
        s.kind = knd;
        s.level = lvl;
        s.rank_a = rank_of(L(t));
        s.rank_b = rank_of(R(t));
        s.rank_c = rank_of(t);
        s.n_m = (int)c.M.size();
        s.n_n = (int)c.N.size();
        s.n_b = (int)c.Bt.size();
        s.n_k = (int)c.K.size();
        s.n_ka = (int)c.KA.size();
        s.n_kb = (int)c.KB.size();
        s.tile_m = c.tm;
        s.tile_n = c.tn;
        s.c_offset = P.off[t];
        for (int i = 0; i < 32; ++i) s.labels_a[i] = s.labels_b[i] = s.labels_c[i] = -1;
        for (int i = 0; i < s.rank_a; ++i) s.labels_a[i] = P.layout[L(t)][i];
        for (int i = 0; i < s.rank_b; ++i) s.labels_b[i] = P.layout[R(t)][i];
        for (int i = 0; i < s.rank_c; ++i) s.labels_c[i] = P.layout[t][i];
        return s;
    };
    // position of a label in a layout, NO_BIT if absent
    auto pos_in = [&](int t, int32_t l) -> uint8_t {
        const auto& v = P.layout[t];
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i] == l) return (uint8_t)i;
        return NO_BIT;
    };
    double ops_f = 0, ops_g = 0, ops_m = 0, bytes = 0, sc = 0;
    for (int t = 0; t < nT; ++t) sc = std::max(sc, (double)rank_of(t));
    auto account = [&](int t, int knd) {
        const Classes& c = cls[t];
        int tc = rank_of(t) + (int)(c.K.size() + c.KA.size() + c.KB.size());
        double o = std::ldexp(1.0, tc);
        (knd == KIND_FUSED ? ops_f : knd == KIND_GENERIC ? ops_g : ops_m) += o;
        bytes += 4.0 * (std::ldexp(1.0, rank_of(L(t))) + std::ldexp(1.0, rank_of(R(t))) + std::ldexp(1.0, rank_of(t)));
    };

    // fused subtrees: post-order with "reserve C, then children above it" shared-memory stack
    for (int t = nL; t < nT; ++t) {
        if (kind[t] != KIND_FUSED || P.level[t] != 0) continue;
        SubTree st{};
        st.first_step = (uint32_t)P.sub_steps.size();
        st.out_off = P.off[t];
        int64_t max_top = 0;
        std::function<void(int, int64_t, bool)> emit = [&](int x, int64_t base, bool is_root) {
            const int A = L(x), B = R(x);
            int64_t above = base;
            if (!is_root) {
                P.loc[x] = LOC_SMEM;
                P.off[x] = base;
                above = base + size_of(x);
            }
            max_top = std::max(max_top, above);
            int64_t pA = is_leaf(A) ? 0 : peak[A], pB = is_leaf(B) ? 0 : peak[B];
            int first = pA >= pB ? A : B, second = pA >= pB ? B : A;
            int64_t cur = above;
            for (int ch : {first, second}) {
                if (is_leaf(ch)) continue;
                emit(ch, cur, false);
                cur += size_of(ch);
                max_top = std::max(max_top, cur);
            }
            const Classes& c = cls[x];
            SubStep s{};
            s.a_off = (uint16_t)P.off[A];
            s.b_off = (uint16_t)P.off[B];
            s.c_off = is_root ? 0 : (uint16_t)P.off[x];
            s.a_loc = (uint8_t)P.loc[A];
            s.b_loc = (uint8_t)P.loc[B];
            s.c_loc = is_root ? LOC_ARENA : LOC_SMEM;
            s.rc = (uint8_t)rank_of(x);
            s.nk = (uint8_t)c.K.size();
            s.nka = (uint8_t)c.KA.size();
            s.nkb = (uint8_t)c.KB.size();
            s.sa = (uint8_t)c.tm;
            s.sb = (uint8_t)c.tn;
            std::memset(s.a_shift, NO_BIT, sizeof s.a_shift);
            std::memset(s.b_shift, NO_BIT, sizeof s.b_shift);
            for (int i = 0; i < rank_of(x); ++i) {
                s.a_shift[i] = pos_in(A, P.layout[x][i]);
                s.b_shift[i] = pos_in(B, P.layout[x][i]);
            }
            P.sub_steps.push_back(s);
            P.info.push_back(fill_info(x, KIND_FUSED, 0));
            account(x, KIND_FUSED);
        };
        emit(t, 0, true);
        st.n_steps = (uint32_t)P.sub_steps.size() - st.first_step;
        st.smem_elems = (uint32_t)max_top;
        P.subtrees.push_back(st);
    }

    // big steps by level
    {
        std::vector<int> order;
        for (int t = nL; t < nT; ++t)
            if (kind[t] == KIND_GENERIC || kind[t] == KIND_GEMM) order.push_back(t);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return P.level[x] < P.level[y]; });
        P.big_level_begin.assign(P.n_levels + 2, 0);
        for (int t : order) {
            const Classes& c = cls[t];
            const int A = L(t), B = R(t);
            BigStep s{};
            s.a_off = P.off[A];
            s.b_off = P.off[B];
            s.c_off = P.off[t];
            s.a_loc = (uint8_t)P.loc[A];
            s.b_loc = (uint8_t)P.loc[B];
            s.kind = (uint8_t)kind[t];
            s.rc = (uint8_t)rank_of(t);
            s.nk = (uint8_t)c.K.size();
            s.nka = (uint8_t)c.KA.size();
            s.nkb = (uint8_t)c.KB.size();
            s.sa = (uint8_t)c.tm;
            s.sb = (uint8_t)c.tn;
            s.tm = (uint8_t)c.tm;
            s.tn = (uint8_t)c.tn;
            std::memset(s.a_shift, NO_BIT, 32);
            std::memset(s.b_shift, NO_BIT, 32);
            std::memset(s.c_shift, NO_BIT, 32);
            if (kind[t] == KIND_GENERIC) {
                for (int i = 0; i < rank_of(t); ++i) {
                    s.a_shift[i] = pos_in(A, P.layout[t][i]);
                    s.b_shift[i] = pos_in(B, P.layout[t][i]);
                    s.c_shift[i] = (uint8_t)i;
                }
                int nkt = s.nk + s.nka + s.nkb;
                if (s.rc >= 8) {
                    s.ks = 0;
                    s.n_tiles = 1u << (s.rc - 8);
                } else {
                    s.ks = (uint8_t)std::min(8 - s.rc, nkt);
                    s.n_tiles = 1;
                }
            } else {
                int q = 0;
                for (int i = 0; i < c.tm; ++i) s.c_shift[q++] = pos_in(t, c.M[i]);
                for (int i = 0; i < c.tn; ++i) s.c_shift[q++] = pos_in(t, c.N[i]);
                for (size_t i = c.tm; i < c.M.size(); ++i) s.c_shift[q++] = pos_in(t, c.M[i]);
                for (size_t i = c.tn; i < c.N.size(); ++i) s.c_shift[q++] = pos_in(t, c.N[i]);
                for (int32_t l : c.Bt) s.c_shift[q++] = pos_in(t, l);
                s.n_mhi = (uint8_t)(c.M.size() - c.tm);
                s.n_nhi = (uint8_t)(c.N.size() - c.tn);
                s.ng = (uint8_t)(s.rc - c.tm - c.tn);
                int tps_log = (c.tm - 3) + (c.tn - 3);          // threads per sub-tile
                int s_log = 8 - tps_log;                        // sub-tiles per CTA (256 threads)
                int64_t per_k = ((int64_t)1 << s_log) * (((int64_t)1 << c.tm) + ((int64_t)1 << c.tn));
                int kc = 0;
                while (kc + 1 <= s.nk && (per_k << (kc + 1)) <= GEMM_STAGE_ELEMS) ++kc;
                s.kc = (uint8_t)kc;
                int64_t groups = (int64_t)1 << s.ng;
                int64_t S = (int64_t)1 << s_log;
                s.n_tiles = (uint32_t)((groups + S - 1) / S);
                s.store_mode = STORE_SCALAR;
                if (s.c_shift[0] == 0 && s.c_shift[1] == 1) s.store_mode = STORE_VEC_M;
                else if (s.c_shift[c.tm] == 0 && s.c_shift[c.tm + 1] == 1) s.store_mode = STORE_VEC_N;
            }
            P.big_steps.push_back(s);
            P.info.push_back(fill_info(t, kind[t], P.level[t]));
            account(t, kind[t]);
        }
        // level offsets
        int idx = 0;
        for (int lv = 1; lv <= P.n_levels + 1; ++lv) {
            while (idx < (int)order.size() && P.level[order[idx]] < lv) ++idx;
            P.big_level_begin[lv] = idx;
        }
        P.big_level_begin[0] = 0;
    }

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
    return TB_OK;
}

void build_blob(Plan& P, std::vector<uint8_t>& blob) {
    auto a16 = [](size_t x) { return (x + 15) / 16 * 16; };
    size_t pool_b = a16(P.pool.size() * 4);
    size_t sub_b = a16(P.sub_steps.size() * sizeof(SubStep));
    size_t big_b = a16(P.big_steps.size() * sizeof(BigStep));
    P.sub_blob_off = pool_b;
    P.big_blob_off = pool_b + sub_b;
    P.blob_bytes = pool_b + sub_b + big_b;
    blob.assign(P.blob_bytes, 0);
    if (!P.pool.empty()) std::memcpy(blob.data(), P.pool.data(), P.pool.size() * 4);
    if (!P.sub_steps.empty()) std::memcpy(blob.data() + P.sub_blob_off, P.sub_steps.data(), P.sub_steps.size() * sizeof(SubStep));
    if (!P.big_steps.empty()) std::memcpy(blob.data() + P.big_blob_off, P.big_steps.data(), P.big_steps.size() * sizeof(BigStep));
}

}  // namespace tb
