// kernels.cuh -- sm_100a device code of libtbcuda: the max-plus contraction kernels.
//
//   k_fused_subtrees  K4/K6: one CTA walks a whole small subtree of the contraction tree with every
//                     intermediate in shared memory (leaf tensors come from the plan's pool).
//   k_generic         any single contraction, one thread per output element (block-level split-k for
//                     tiny outputs); memory-bound nodes, skinny nodes, nodes touching leaves.
//   k_gemm            K1/K3: tiled batched max-plus GEMM, 8x8 register microtile, cp.async double
//                     buffered contiguous operand panels, DPX VIADDMNMX (int32) / FADD+FMNMX (f32).
//   k_permute_bits    K5: bit-permutation transpose through shared memory, 128-bit global accesses.
//   k_finalize        root scalars -> per-branch result vector (the input of maximum(res),
//                     /root/reference/src/dynamic_ob.jl:27).
//
// Semantics of one step (OMEinsum binary rule + TropicalGEMM mul! [upstream]):
//   C[M,N,Bt] = max_{K,KA,KB} A[M,K,KA,Bt] + B[N,K,KB,Bt]      with (+) = max, (x) = +.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "desc.h"
#include <type_traits>

namespace tb {

template <typename T> struct Ops;
template <> struct Ops<int32_t> {
    typedef int4 vec4;
    static __device__ __forceinline__ int32_t neg_inf() { return -(1 << 30); }
    // DPX: one VIADDMNMX = max(a + b, c)
    static __device__ __forceinline__ int32_t addmax(int32_t a, int32_t b, int32_t c) { return __viaddmax_s32(a, b, c); }
    static __device__ __forceinline__ int32_t vmax(int32_t a, int32_t b) { return max(a, b); }
    static __device__ __forceinline__ double to_double(int32_t v) { return v <= -(1 << 29) ? -INFINITY : (double)v; }
};
template <> struct Ops<int16_t> {
    typedef short4 vec4;  // 4 elements = 8 bytes
    static __device__ __forceinline__ int16_t neg_inf() { return (int16_t)(-(1 << 14)); }
    static __device__ __forceinline__ int16_t addmax(int16_t a, int16_t b, int16_t c) {
        return (int16_t)__viaddmax_s32((int)a, (int)b, (int)c);  // values stay in [-2^15, 2^13): no wrap
    }
    static __device__ __forceinline__ int16_t vmax(int16_t a, int16_t b) { return a > b ? a : b; }
    static __device__ __forceinline__ double to_double(int16_t v) { return v <= -(1 << 13) ? -INFINITY : (double)v; }
};
template <> struct Ops<float> {
    typedef float4 vec4;
    static __device__ __forceinline__ float neg_inf() { return -INFINITY; }
    static __device__ __forceinline__ float addmax(float a, float b, float c) { return fmaxf(__fadd_rn(a, b), c); }
    static __device__ __forceinline__ float vmax(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ double to_double(float v) { return (double)v; }
};

// 8-byte value types (no tiled GEMM kernel: generic + fused kernels only; small problems by construction)
struct alignas(16) Double4 { double x, y, z, w; };
struct alignas(16) Long4 { long long x, y, z, w; };
template <> struct Ops<double> {  // Tropical{Float64}
    typedef Double4 vec4;
    static __device__ __forceinline__ double neg_inf() { return -INFINITY; }
    static __device__ __forceinline__ double addmax(double a, double b, double c) { return fmax(__dadd_rn(a, b), c); }
    static __device__ __forceinline__ double vmax(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double to_double(double v) { return v; }
};
// "size + one optimal configuration" (the SingleConfigMax element of the branching tables) as ONE int64:
// high 32 bits = size (int32, -2^30 = tropical zero), low 32 bits = bit mask of the chosen vertices.  Every vertex leaf is
// used exactly once in a contraction tree, so the masks of two operands of a node are disjoint: a (x) b = a + b (the masks
// OR, no carry into the size), a (+) b = max(a, b) (size first, ties by mask): plain max-plus on int64.
template <> struct Ops<long long> {
    typedef Long4 vec4;
    static __device__ __forceinline__ long long neg_inf() { return -(1ll << 62); }
    static __device__ __forceinline__ long long addmax(long long a, long long b, long long c) { return max(a + b, c); }
    static __device__ __forceinline__ long long vmax(long long a, long long b) { return max(a, b); }
    static __device__ __forceinline__ double to_double(long long v) { return (v >> 32) <= -(1ll << 29) ? -INFINITY : (double)(v >> 32); }
};

template <typename T> __device__ __forceinline__ T shfl_xor_t(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ int16_t shfl_xor_t<int16_t>(int16_t v, int m) {
    return (int16_t)__shfl_xor_sync(0xffffffffu, (int)v, m);
}

// acc = max(acc, max_i(a[i] + b[i])) over the 16 bytes at a and b (both 16-byte aligned)
template <typename T> __device__ __forceinline__ T dot16(const T* a, const T* b, T acc);
template <> __device__ __forceinline__ int32_t dot16<int32_t>(const int32_t* a, const int32_t* b, int32_t acc) {
    const int4 x = *reinterpret_cast<const int4*>(a), y = *reinterpret_cast<const int4*>(b);
    acc = __viaddmax_s32(x.x, y.x, acc);
    acc = __viaddmax_s32(x.y, y.y, acc);
    acc = __viaddmax_s32(x.z, y.z, acc);
    return __viaddmax_s32(x.w, y.w, acc);
}
template <> __device__ __forceinline__ float dot16<float>(const float* a, const float* b, float acc) {
    const float4 x = *reinterpret_cast<const float4*>(a), y = *reinterpret_cast<const float4*>(b);
    acc = fmaxf(__fadd_rn(x.x, y.x), acc);
    acc = fmaxf(__fadd_rn(x.y, y.y), acc);
    acc = fmaxf(__fadd_rn(x.z, y.z), acc);
    return fmaxf(__fadd_rn(x.w, y.w), acc);
}
template <> __device__ __forceinline__ double dot16<double>(const double* a, const double* b, double acc) {
    acc = fmax(__dadd_rn(a[0], b[0]), acc);
    return fmax(__dadd_rn(a[1], b[1]), acc);
}
template <> __device__ __forceinline__ long long dot16<long long>(const long long* a, const long long* b, long long acc) {
    acc = max(a[0] + b[0], acc);
    return max(a[1] + b[1], acc);
}
template <> __device__ __forceinline__ int16_t dot16<int16_t>(const int16_t* a, const int16_t* b, int16_t acc) {
    const uint4 x = *reinterpret_cast<const uint4*>(a), y = *reinterpret_cast<const uint4*>(b);
    unsigned p = ((unsigned)(uint16_t)acc) * 0x10001u;  // packed pair (acc, acc)
    p = __viaddmax_s16x2(x.x, y.x, p);
    p = __viaddmax_s16x2(x.y, y.y, p);
    p = __viaddmax_s16x2(x.z, y.z, p);
    p = __viaddmax_s16x2(x.w, y.w, p);
    const int16_t lo = (int16_t)(p & 0xffffu), hi = (int16_t)(p >> 16);
    return lo > hi ? lo : hi;
}

__device__ __forceinline__ uint32_t scatter_bits(uint32_t x, const uint8_t* sh, int n) {
    uint32_t off = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t s = sh[i];
        if (s != NO_BIT) off |= ((x >> i) & 1u) << s;
    }
    return off;
}

#ifdef TB_KPROF
// diagnostics build: one record per CTA (which launch, which SM, when) for an offline occupancy timeline
struct TlRec {
    unsigned long long key, t0, t1;
    unsigned smid, kind;
};
__device__ TlRec* g_tl;
__device__ unsigned g_tl_n, g_tl_cap;
__device__ __forceinline__ unsigned long long tl_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tl_log(const void* key, unsigned kind, unsigned long long t0) {
    const unsigned i = atomicAdd(&g_tl_n, 1u);
    if (i < g_tl_cap) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        g_tl[i] = TlRec{(unsigned long long)key, t0, tl_now(), smid, kind};
    }
}
#define TL_BEGIN const unsigned long long tl_t0 = tl_now();
#define TL_END(key, kind, cond) if (cond) tl_log(key, kind, tl_t0);
#else
#define TL_BEGIN
#define TL_END(key, kind, cond)
#endif

// largest idx with starts[idx] <= b (starts[0] == 0).  Called by all 32 lanes of a converged warp: a 32-ary search,
// 2-3 dependent loads instead of the ~12 of a binary search (the search sits at the head of every CTA / tile).
__device__ __forceinline__ int find_inst(const uint32_t* __restrict__ starts, int n, uint32_t b) {
    const int lane = threadIdx.x & 31;
    int lo = 0, len = n;
    while (len > 1) {
        const int stride = (len + 31) >> 5;
        const int pos = lo + lane * stride;
        const bool le = pos < lo + len && __ldg(starts + pos) <= b;
        const unsigned m = __ballot_sync(0xffffffffu, le);  // monotone; lane 0 is always set
        const int j = 31 - __clz((int)m);
        const int hi = lo + len;
        lo += j * stride;
        len = min(stride, hi - lo);
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------
// K4: fused small subtrees.  grid = one CTA per (branch, subtree), 128 threads.
// ------------------------------------------------------------------------------------------------
constexpr int FUSED_THREADS = 128;

// Dynamic shared memory of one CTA: [all step descriptors of the subtree | a copy of the branch's leaf pool | data].
// Descriptors and pool arrive with one round of wide loads, after that a step only touches shared memory (the
// subtrees are ~25 steps of 4..32 outputs each: latency, not throughput, is what a step costs).
__host__ __device__ inline uint32_t fused_desc_bytes(uint32_t n_steps) { return n_steps * (uint32_t)sizeof(SubStep); }
__host__ __device__ inline uint32_t fused_pool_bytes(uint32_t n_pool, uint32_t elem) { return (n_pool * elem + 15u) & ~15u; }

template <typename T>
__global__ void __launch_bounds__(FUSED_THREADS) k_fused_subtrees(const SubInst* __restrict__ insts, int n_insts) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    if ((int)blockIdx.x >= n_insts) return;
    TL_BEGIN
    const SubInst inst = insts[blockIdx.x];
    const int tid = threadIdx.x;
    {
        const uint4* g = reinterpret_cast<const uint4*>(inst.steps);
        uint4* sdst = reinterpret_cast<uint4*>(dyn_smem);
        const uint32_t nv = inst.n_steps * (uint32_t)(sizeof(SubStep) / 16);
        for (uint32_t i = tid; i < nv; i += FUSED_THREADS) sdst[i] = __ldg(g + i);
    }
    const uint32_t dbytes = fused_desc_bytes(inst.n_steps);
    const T* pool = reinterpret_cast<const T*>(inst.pool);
    if (inst.n_pool) {  // pool slots come in multiples of 4: 8-byte granules for every value type
        const uint2* g = reinterpret_cast<const uint2*>(inst.pool);
        uint2* sdst = reinterpret_cast<uint2*>(dyn_smem + dbytes);
        const uint32_t nv = inst.n_pool * (uint32_t)sizeof(T) / 8u;
        for (uint32_t i = tid; i < nv; i += FUSED_THREADS) sdst[i] = __ldg(g + i);
        pool = reinterpret_cast<const T*>(dyn_smem + dbytes);
    }
    T* smem = reinterpret_cast<T*>(dyn_smem + dbytes + fused_pool_bytes(inst.n_pool, sizeof(T)));
    T* out = reinterpret_cast<T*>(inst.out);
    __syncthreads();
    for (uint32_t st = 0; st < inst.n_steps; ++st) {
        const SubStep& d = reinterpret_cast<const SubStep*>(dyn_smem)[st];
        const T* A = (d.a_loc == LOC_SMEM ? smem : pool) + d.a_off;
        const T* B = (d.b_loc == LOC_SMEM ? smem : pool) + d.b_off;
        T* C = (d.c_loc == LOC_SMEM) ? smem + d.c_off : out;
        const int rc = d.rc, nk = d.nk, nka = d.nka;
        const int nkt = nk + nka + d.nkb;
        int ks = 7 - rc;  // lanes cooperating on one output: 2^ks consecutive lanes (within a warp)
        ks = ks < 0 ? 0 : ks;
        ks = ks > 5 ? 5 : ks;
        ks = ks > nkt ? nkt : ks;
        const uint32_t W = 1u << (rc + ks);
        const uint32_t kmask = (1u << nk) - 1u, amask = (1u << (nk + nka)) - 1u;
        const uint32_t n_red = 1u << nkt;
        for (uint32_t w0 = 0; w0 < W; w0 += FUSED_THREADS) {
            const uint32_t w = w0 + tid;
            const bool active = w < W;
            const uint32_t c = w >> ks, kp = w & ((1u << ks) - 1u);
            T acc = Ops<T>::neg_inf();
            if (active) {
                const uint32_t offA = scatter_bits(c, d.a_shift, rc);
                const uint32_t offB = scatter_bits(c, d.b_shift, rc);
                if (d.pad) {  // K0-first operands (packed int16 GEMM layout executed by the fused path)
                    for (uint32_t r = kp; r < n_red; r += (1u << ks))
                        acc = Ops<T>::addmax(A[offA + ((r & 1u) | ((r >> 1) << (d.sa + 1)))],
                                             B[offB + ((r & 1u) | ((r >> 1) << (d.sb + 1)))], acc);
                } else {
                    for (uint32_t r = kp; r < n_red; r += (1u << ks)) {
                        const uint32_t ra = (r & amask) << d.sa;
                        const uint32_t rb = ((r & kmask) | ((r >> (nk + nka)) << nk)) << d.sb;
                        acc = Ops<T>::addmax(A[offA + ra], B[offB + rb], acc);
                    }
                }
            }
            for (int s = 1; s < (1 << ks); s <<= 1) acc = Ops<T>::vmax(acc, shfl_xor_t(acc, s));
            if (active && kp == 0) C[c] = acc;
        }
        __syncthreads();
    }
    TL_END(insts, 0u, tid == 0)
}

// ------------------------------------------------------------------------------------------------
// generic single contraction.  grid = sum of n_tiles over the launch, 256 threads.
// ------------------------------------------------------------------------------------------------
constexpr int BIG_THREADS = 256;

// One CTA-sized piece ("tile") of a generic step: 256 threads.  Shared by the stand-alone kernel (level-synchronous
// launches) and by the consumer warps of the persistent GEMM kernels (dataflow launches, PERSIST: the 256 consumer
// threads synchronise on named barrier 1 instead of __syncthreads).
template <bool PERSIST> __device__ __forceinline__ void generic_sync() {
    if (PERSIST) asm volatile("bar.sync 1, 256;\n" ::: "memory");
    else __syncthreads();
}

template <typename T, bool PERSIST>
__device__ __forceinline__ void generic_tile(const BigStep& sd, const void* poolp, void* arenap, uint32_t tile, int tid, T* s_red) {
    const T* pool = reinterpret_cast<const T*>(poolp);
    T* arena = reinterpret_cast<T*>(arenap);
    // no __restrict__ / read-only qualifiers on the operands: inside a dataflow launch they were written by other CTAs of
    // the SAME kernel, so the non-coherent load path (LDG.CONSTANT) must never be used for them
    const T* A = (sd.a_loc == LOC_POOL ? pool : arena) + sd.a_off;
    const T* B = (sd.b_loc == LOC_POOL ? pool : arena) + sd.b_off;
    T* C = arena + sd.c_off;
    const int rc = sd.rc, nk = sd.nk, nka = sd.nka, ks = sd.ks, po = sd.po;
    const int nkt = nk + nka + sd.nkb;
    if (sd.vec4) {
        // streaming mode: this thread owns outputs c4 .. c4+3 (C address bits 0,1), 128-bit store
        typedef typename Ops<T>::vec4 vec4;
        const uint32_t a0 = sd.a_shift[0], a1 = sd.a_shift[1], b0 = sd.b_shift[0], b1 = sd.b_shift[1];
        const int modeA = (a0 == NO_BIT && a1 == NO_BIT) ? OPV_BCAST : ((a0 == 0 && a1 == 1) ? OPV_VEC : OPV_GATHER);
        const int modeB = (b0 == NO_BIT && b1 == NO_BIT) ? OPV_BCAST : ((b0 == 0 && b1 == 1) ? OPV_VEC : OPV_GATHER);
        const uint32_t dA1 = a0 == NO_BIT ? 0u : (1u << a0), dA2 = a1 == NO_BIT ? 0u : (1u << a1);
        const uint32_t dB1 = b0 == NO_BIT ? 0u : (1u << b0), dB2 = b1 == NO_BIT ? 0u : (1u << b1);
        const uint32_t kmask = (1u << nk) - 1u, amask = (1u << (nk + nka)) - 1u;
        const uint32_t n_red = 1u << nkt;
        const int sa = sd.sa, sb = sd.sb;
        const bool kfirst = sd.store_mode != 0;  // generic steps: K0-first operand layouts
        // NV vectors per thread; the reduction loop is outermost so that the loads of all NV vectors are in flight
        // together (these nodes stream: memory-level parallelism is what bounds them)
        auto vectors = [&](auto nv_tag) {
            constexpr int NV = decltype(nv_tag)::value;
            uint32_t c4[NV], offA[NV], offB[NV];
            T acc[NV][4];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                c4[v] = ((((tile * NV) + (uint32_t)v) << 8) | (uint32_t)tid) << 2;
                offA[v] = scatter_bits(c4[v], sd.a_shift, rc);
                offB[v] = scatter_bits(c4[v], sd.b_shift, rc);
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[v][q] = Ops<T>::neg_inf();
            }
            for (uint32_t r = 0; r < n_red; ++r) {
                const uint32_t ra = kfirst ? ((r & 1u) | ((r >> 1) << (sa + 1))) : ((r & amask) << sa);
                const uint32_t rb = kfirst ? ((r & 1u) | ((r >> 1) << (sb + 1))) : (((r & kmask) | ((r >> (nk + nka)) << nk)) << sb);
                T av[NV][4], bv[NV][4];
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const T* pa = A + offA[v] + ra;
                    const T* pb = B + offB[v] + rb;
                    if (modeA == OPV_VEC) {
                        const vec4 x = *reinterpret_cast<const vec4*>(pa);
                        av[v][0] = x.x; av[v][1] = x.y; av[v][2] = x.z; av[v][3] = x.w;
                    } else if (modeA == OPV_BCAST) {
                        av[v][0] = av[v][1] = av[v][2] = av[v][3] = pa[0];
                    } else {
                        av[v][0] = pa[0]; av[v][1] = pa[dA1]; av[v][2] = pa[dA2]; av[v][3] = pa[dA1 + dA2];
                    }
                    if (modeB == OPV_VEC) {
                        const vec4 x = *reinterpret_cast<const vec4*>(pb);
                        bv[v][0] = x.x; bv[v][1] = x.y; bv[v][2] = x.z; bv[v][3] = x.w;
                    } else if (modeB == OPV_BCAST) {
                        bv[v][0] = bv[v][1] = bv[v][2] = bv[v][3] = pb[0];
                    } else {
                        bv[v][0] = pb[0]; bv[v][1] = pb[dB1]; bv[v][2] = pb[dB2]; bv[v][3] = pb[dB1 + dB2];
                    }
                }
#pragma unroll
                for (int v = 0; v < NV; ++v)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[v][q] = Ops<T>::addmax(av[v][q], bv[v][q], acc[v][q]);
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                vec4 o;
                o.x = acc[v][0]; o.y = acc[v][1]; o.z = acc[v][2]; o.w = acc[v][3];
                *reinterpret_cast<vec4*>(C + c4[v]) = o;
            }
        };
        if (po == 12) vectors(std::integral_constant<int, 4>{});
        else vectors(std::integral_constant<int, 1>{});
        return;
    }
    // thread -> (output, k-part): 2^po outputs per CTA, 2^ks threads share one output
    const uint32_t c = (tile << po) | (tid & ((1u << po) - 1u));
    const uint32_t kp = (uint32_t)tid >> po;
    const bool active = kp < (1u << ks);
    T acc = Ops<T>::neg_inf();
    if (active) {
        const uint32_t offA = scatter_bits(c, sd.a_shift, rc);
        const uint32_t offB = scatter_bits(c, sd.b_shift, rc);
        const uint32_t kmask = (1u << nk) - 1u, amask = (1u << (nk + nka)) - 1u;
        const uint32_t n_red = 1u << nkt;
        const int sa = sd.sa, sb = sd.sb;
        constexpr uint32_t V = 16 / sizeof(T);  // elements per 16-byte vector
        if (sd.store_mode != 0) {  // K0-first operand layouts
            for (uint32_t r = kp; r < n_red; r += (1u << ks))
                acc = Ops<T>::addmax(A[offA + ((r & 1u) | ((r >> 1) << (sa + 1)))], B[offB + ((r & 1u) | ((r >> 1) << (sb + 1)))], acc);
        } else if (nka == 0 && sd.nkb == 0 && sa == 0 && sb == 0 && (n_red >> ks) >= V) {
            // both operands carry the reduced labels in their lowest address bits (full reductions, e.g. the root
            // of every tree): 128-bit loads, consecutive threads take consecutive vectors
#pragma unroll 2
            for (uint32_t r = kp * V; r < n_red; r += (V << ks)) acc = dot16<T>(A + offA + r, B + offB + r, acc);
        } else if (sa == 0 && sb == 0 && (1u << nk) >= V && (n_red >> ks) >= V) {
            // the same with labels private to one operand (split-K labels folded into this reduction): a vector of V
            // consecutive r shares the private index, so both operands still read 16 contiguous bytes
#pragma unroll 2
            for (uint32_t r = kp * V; r < n_red; r += (V << ks))
                acc = dot16<T>(A + offA + (r & amask), B + offB + ((r & kmask) | ((r >> (nk + nka)) << nk)), acc);
        } else if (nka == 0 && sd.nkb == 0) {
#pragma unroll 4
            for (uint32_t r = kp; r < n_red; r += (1u << ks))
                acc = Ops<T>::addmax(A[offA + (r << sa)], B[offB + (r << sb)], acc);
        } else {
            for (uint32_t r = kp; r < n_red; r += (1u << ks)) {
                const uint32_t ra = (r & amask) << sa;
                const uint32_t rb = ((r & kmask) | ((r >> (nk + nka)) << nk)) << sb;
                acc = Ops<T>::addmax(A[offA + ra], B[offB + rb], acc);
            }
        }
    }
    if (ks == 0) {
        if (active) C[c] = acc;
    } else {
        // tree reduction over the k-parts (uniform trip count: ks is per-CTA)
        if (PERSIST) generic_sync<true>();  // s_red is epilogue staging memory: the previous GEMM tile's readers must be done
        s_red[tid] = acc;
        generic_sync<PERSIST>();
        for (int s = ks - 1; s >= 0; --s) {
            if (kp < (1u << s)) {
                acc = Ops<T>::vmax(acc, s_red[tid + ((1u << s) << po)]);
                s_red[tid] = acc;
            }
            generic_sync<PERSIST>();
        }
        if (kp == 0) C[c] = acc;
    }
}

// The persistent kernels call the generic tile through a real function call: inlined, its (divergent, table-driven) body
// made ptxas drop the warp-uniform address arithmetic of the GEMM main loop in the same kernel (12 % on BASELINE config 4).
// No accumulator is live between tiles, so the call costs nothing that matters.
template <typename T>
__device__ __noinline__ void generic_tile_call(const BigStep* sd, const void* poolp, void* arenap, uint32_t tile, int tid, T* s_red) {
    generic_tile<T, true>(*sd, poolp, arenap, tile, tid, s_red);
}

template <typename T>
__global__ void __launch_bounds__(BIG_THREADS, 4) k_generic(const BigInst* __restrict__ insts,
                                                         const uint32_t* __restrict__ tile_starts, int n_insts) {
    __shared__ BigStep sd;
    __shared__ T s_red[BIG_THREADS];
    TL_BEGIN
    const int tid = threadIdx.x;
    const int idx = find_inst(tile_starts, n_insts, blockIdx.x);
    const BigInst inst = insts[idx];
    const uint32_t tile = blockIdx.x - __ldg(tile_starts + idx);
    if (tid < (int)(sizeof(BigStep) / 4)) reinterpret_cast<uint32_t*>(&sd)[tid] = __ldg(reinterpret_cast<const uint32_t*>(inst.step) + tid);
    __syncthreads();
    generic_tile<T, false>(sd, inst.pool, inst.arena, tile, tid, s_red);
    TL_END(insts, 1u, tid == 0)
}

// ------------------------------------------------------------------------------------------------
// K1/K3: tiled batched max-plus GEMM.  256 threads, each an 8x8 microtile; a CTA holds S = 2^s_log
// independent (2^tm x 2^tn) sub-tiles (consecutive grid indices) so that small tiles still fill it.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

constexpr int GEMM_SMEM_BYTES = 2 * GEMM_STAGE_ELEMS * 4;

template <typename T>
__global__ void __launch_bounds__(BIG_THREADS, 2) k_gemm(const BigInst* __restrict__ insts,
                                                         const uint32_t* __restrict__ tile_starts, int n_insts) {
    typedef typename Ops<T>::vec4 vec4;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    T* stage_mem = reinterpret_cast<T*>(dyn_smem);
    __shared__ BigStep sd;
    __shared__ long long s_abase[32], s_bbase[32], s_cbase[32];
    const int tid = threadIdx.x;
    const int idx = find_inst(tile_starts, n_insts, blockIdx.x);
    const BigInst inst = insts[idx];
    const uint32_t tile = blockIdx.x - __ldg(tile_starts + idx);
    if (tid < (int)(sizeof(BigStep) / 4)) reinterpret_cast<uint32_t*>(&sd)[tid] = __ldg(reinterpret_cast<const uint32_t*>(inst.step) + tid);
    __syncthreads();
    const int tm = sd.tm, tn = sd.tn, nk = sd.nk, kc = sd.kc, ng = sd.ng;
    const int tps_log = tm + tn - 6, s_log = 8 - tps_log, S = 1 << s_log;
    T* arena = reinterpret_cast<T*>(inst.arena);
    const T* __restrict__ Ag = arena + sd.a_off;
    const T* __restrict__ Bg = arena + sd.b_off;
    T* __restrict__ Cg = arena + sd.c_off;
    if (tid < S) {
        const unsigned long long g = (unsigned long long)tile * S + tid;
        long long ab = -1, bb = -1, cb = -1;
        if (g < (1ull << ng)) {
            const int n_mhi = sd.n_mhi, n_nhi = sd.n_nhi;
            const unsigned long long gm = g & ((1ull << n_mhi) - 1ull);
            const unsigned long long gn = (g >> n_mhi) & ((1ull << n_nhi) - 1ull);
            const unsigned long long gb = g >> (n_mhi + n_nhi);
            ab = (long long)((gm | (gb << n_mhi)) << (tm + nk));
            bb = (long long)((gn | (gb << n_nhi)) << (tn + nk));
            cb = (long long)scatter_bits((uint32_t)g, sd.c_shift + tm + tn, ng);
        }
        s_abase[tid] = ab;
        s_bbase[tid] = bb;
        s_cbase[tid] = cb;
    }
    __syncthreads();

    const int la = kc + tm, lb = kc + tn;  // log2 elements of one sub-tile's A / B chunk
    T* stageB_off = nullptr;
    (void)stageB_off;
    auto load_stage = [&](int stage, int chunk) {
        T* sA = stage_mem + stage * GEMM_STAGE_ELEMS;
        T* sB = sA + ((size_t)S << la);
        const int nvA = 1 << (s_log + la - 2), nvB = 1 << (s_log + lb - 2);
        for (int v = tid; v < nvA; v += BIG_THREADS) {
            const int s = v >> (la - 2), w = v & ((1 << (la - 2)) - 1);
            const long long base = s_abase[s];
            if (base >= 0) cp_async16(sA + ((size_t)s << la) + w * 4, Ag + base + ((long long)chunk << la) + w * 4);
        }
        for (int v = tid; v < nvB; v += BIG_THREADS) {
            const int s = v >> (lb - 2), w = v & ((1 << (lb - 2)) - 1);
            const long long base = s_bbase[s];
            if (base >= 0) cp_async16(sB + ((size_t)s << lb) + w * 4, Bg + base + ((long long)chunk << lb) + w * 4);
        }
    };

    const int sub = tid >> tps_log, lt = tid & ((1 << tps_log) - 1);
    const int tmh = lt & ((1 << (tm - 3)) - 1), tnh = lt >> (tm - 3);
    const int m_lo = tmh * 4, m_hi = (1 << (tm - 1)) + tmh * 4;
    const int n_lo = tnh * 4, n_hi = (1 << (tn - 1)) + tnh * 4;

    T acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = Ops<T>::neg_inf();

    const int nchunks = 1 << (nk - kc);
    load_stage(0, 0);
    cp_async_commit();
    for (int ch = 0; ch < nchunks; ++ch) {
        if (ch + 1 < nchunks) load_stage((ch + 1) & 1, ch + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const T* sA = stage_mem + (ch & 1) * GEMM_STAGE_ELEMS + ((size_t)sub << la);
        const T* sB = stage_mem + (ch & 1) * GEMM_STAGE_ELEMS + ((size_t)S << la) + ((size_t)sub << lb);
        const int KC = 1 << kc;
#pragma unroll 2
        for (int kk = 0; kk < KC; ++kk) {
            const T* ar = sA + (kk << tm);
            const T* br = sB + (kk << tn);
            const vec4 a0 = *reinterpret_cast<const vec4*>(ar + m_lo);
            const vec4 a1 = *reinterpret_cast<const vec4*>(ar + m_hi);
            const vec4 b0 = *reinterpret_cast<const vec4*>(br + n_lo);
            const vec4 b1 = *reinterpret_cast<const vec4*>(br + n_hi);
            const T a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const T b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = Ops<T>::addmax(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    const long long cbase = s_cbase[sub];
    if (cbase < 0) return;
    T* __restrict__ Ct = Cg + cbase;
    uint32_t moff[8], noff[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t mi = (i < 4) ? (uint32_t)(m_lo + i) : (uint32_t)(m_hi + i - 4);
        const uint32_t ni = (i < 4) ? (uint32_t)(n_lo + i) : (uint32_t)(n_hi + i - 4);
        moff[i] = scatter_bits(mi, sd.c_shift, tm);
        noff[i] = scatter_bits(ni, sd.c_shift + tm, tn);
    }
    if (sd.store_mode == STORE_VEC_M) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int ih = 0; ih < 2; ++ih) {
                vec4 v;
                v.x = acc[ih * 4 + 0][j];
                v.y = acc[ih * 4 + 1][j];
                v.z = acc[ih * 4 + 2][j];
                v.w = acc[ih * 4 + 3][j];
                *reinterpret_cast<vec4*>(Ct + moff[ih * 4] + noff[j]) = v;
            }
    } else if (sd.store_mode == STORE_VEC_N) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int jh = 0; jh < 2; ++jh) {
                vec4 v;
                v.x = acc[i][jh * 4 + 0];
                v.y = acc[i][jh * 4 + 1];
                v.z = acc[i][jh * 4 + 2];
                v.w = acc[i][jh * 4 + 3];
                *reinterpret_cast<vec4*>(Ct + moff[i] + noff[jh * 4]) = v;
            }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) Ct[moff[i] + noff[j]] = acc[i][j];
    }
}

// ------------------------------------------------------------------------------------------------
// K1/K3 v2: persistent, warp-specialised tiled max-plus GEMM.
//   warp 8 (producer): fetches tiles from a global counter, decodes the step descriptor, publishes a
//     TileInfo record and streams the contiguous A / B panels into a 4-stage shared-memory ring with
//     TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx), running ahead across tile boundaries.
//   warps 0-7 (consumers): wait on the stage's "full" mbarrier, run the 8x8 DPX microtile loop, release
//     the stage, and scatter-store the tile when its k loop ends.
// No __syncthreads in steady state; prologue / epilogue of one tile overlap the k loop of the next.
// ------------------------------------------------------------------------------------------------
constexpr int G2_STAGES = 4;
constexpr int G2_CONSUMERS = 256;
constexpr int G2_PRODUCERS = 128;  // one warpgroup (setmaxnreg is per warpgroup); only its first warp works
constexpr int G2_THREADS = G2_PRODUCERS + G2_CONSUMERS;
constexpr int G2_RING_BYTES = G2_STAGES * GEMM_STAGE_ELEMS * 4;
constexpr int G2_STG_ELEMS = 4096;                         // one quarter of a CTA tile
constexpr int G2_SMEM_BYTES = G2_RING_BYTES + 2 * G2_STG_ELEMS * 4;  // ring + double-buffered epilogue staging

struct TileInfo {
    long long cbase[32];  // per sub-tile element offset into C, -1 = inactive
    void* C;
    int tm, tn, kc, nchunks, store_mode, valid, lane_n_first;
    unsigned char c_shift[16];
    unsigned char e_spos[12], e_cs[12];  // staged epilogue: sorted in-round tile bits -> staging position / C shift
    unsigned char e_cs_mtop, e_cs_ntop, e_vec, e_fast;
    // dataflow launches (see TileInfoH)
    int kind;
    uint32_t tile;
    const void* pool;
    void* arena;
    unsigned int* done;
};

// bank swizzle of the epilogue staging index: permutes 16-byte chunks inside a 128-byte row
__device__ __forceinline__ uint32_t stg_swz(uint32_t x) { return x ^ ((((x >> 5) ^ (x >> 8) ^ (x >> 11)) & 7u) << 2); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
// spin limit of every in-kernel wait (~20 s at 1.97 GHz): in a dataflow launch a wait may legitimately last as long as the
// producing node runs (BASELINE config 4: ~0.1 s); anything beyond the limit is a lost arrival and traps instead of hanging
constexpr long long kSpinTimeoutClocks = 40000000000ll;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done = 0;
    long long t0 = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done) {  // a lost arrival must fail loudly, never hang the GPU
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kSpinTimeoutClocks) __trap();
        }
    }
}

// ---- dataflow launches: completion counters in global memory.  done[i] starts at (tiles of instance i) x (consumer warps
// per CTA) and every consumer warp decrements it once per finished tile (release); the producer warp of a CTA that has
// been handed a tile of a dependent instance spins until the counters of both operands' producers read 0 (acquire).
// Tiles are handed out in topological order by ONE atomic counter, so whatever a spinning CTA waits for has already
// been handed to a CTA that is running (or finished): no co-residency assumption, no deadlock.
__device__ __forceinline__ void dep_wait(const unsigned int* p) {
    long long t0 = 0;
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
        if (v == 0) return;
        __nanosleep(64);
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > kSpinTimeoutClocks) __trap();
    }
}
// Called by the SIGNAL warp of a CTA once all consumer warps have arrived on the tile's bar_sig mbarrier (their global
// stores of the tile happen-before the arrival, CTA scope): the gpu-scope fence is cumulative over them, then one relaxed
// decrement by the number of consumer warps publishes the tile.  The consumer warps themselves never wait for their
// stores to be acknowledged -- they are already in the next tile's main loop.
__device__ __forceinline__ void dep_signal(unsigned int* p, unsigned n) {
    asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(0u - n) : "memory");
}
// the producer warp reads, with TMA (async proxy), global memory that other CTAs of the same kernel wrote with ordinary
// stores (generic proxy): cross-proxy fence between the acquire above and the bulk copies below
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
// the same on a precomputed shared-memory address (hot loops: no generic -> shared conversion per wait)
__device__ __forceinline__ void mbar_wait_addr(unsigned addr, unsigned parity) {
    unsigned done = 0;
    long long t0 = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kSpinTimeoutClocks) __trap();
        }
    }
}
__device__ __forceinline__ void mbar_arrive_addr(unsigned addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// Register budget: 2 CTAs/SM x 384 threads are launched with 80 registers/thread; setmaxnreg moves registers
// INSIDE a CTA's own pool (SASS: USETMAXREG...CTAPOOL), so 128*P + 256*C <= 384*80 must hold or the .inc never
// succeeds (a silent hang).  P = 24 (producer warpgroup), C = 104 (consumer warpgroups): 3072 + 26624 = 29696.
static_assert(128 * 24 + 256 * 104 <= 384 * 80, "setmaxnreg budget exceeds the CTA pool");
static_assert(128 * 32 + 256 * 104 <= 384 * 80, "setmaxnreg budget of k_gemm2h exceeds the CTA pool");
template <typename T, bool DF>
__global__ void __launch_bounds__(G2_THREADS, 2) k_gemm2(const BigInst* __restrict__ insts, const uint32_t* __restrict__ tile_starts,
                                                         int n_insts, uint32_t total_tiles, unsigned int* __restrict__ counter,
                                                         int staged_epilogue, unsigned int* done) {
    typedef typename Ops<T>::vec4 vec4;
    if (!DF) done = nullptr;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    T* stage_mem = reinterpret_cast<T*>(dyn_smem);
    T* stg_mem = reinterpret_cast<T*>(dyn_smem + G2_RING_BYTES);
    __shared__ __align__(8) uint64_t bar_full[G2_STAGES], bar_empty[G2_STAGES], bar_tfull[2], bar_tempty[2], bar_sig[2];
    __shared__ TileInfo tinfo[2];
    __shared__ __align__(16) BigStep s_gstep[2];  // generic tiles: the descriptor the consumers execute
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < G2_STAGES; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], G2_CONSUMERS / 32);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_tfull[i], 1);
            mbar_init(&bar_tempty[i], 1);                  // the signal warp hands the slot back
            mbar_init(&bar_sig[i], G2_CONSUMERS / 32);     // every consumer warp is done with the tile
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (tid < G2_PRODUCERS) {
        // ------------------------------------------------------------------ producer warpgroup
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;\n" ::);
        if (tid >= 64) return;
        if (tid >= 32) {
            // ------------------------------------------------------------------ signal warp: publishes finished tiles
            if (tid != 32) return;
            for (unsigned tcount = 0;; ++tcount) {
                const int slot = tcount & 1;
                mbar_wait(&bar_tfull[slot], (tcount >> 1) & 1);
                if (!tinfo[slot].valid) break;
                unsigned int* const dp = tinfo[slot].done;
                mbar_wait(&bar_sig[slot], (tcount >> 1) & 1);
                if (dp) dep_signal(dp, G2_CONSUMERS / 32);
                mbar_arrive(&bar_tempty[slot]);
            }
            return;
        }
        const int lane = tid;
        unsigned it = 0;  // global chunk counter (ring position)
        int waited_idx = -1;  // dataflow: the instance whose operands this warp has already waited for
        for (unsigned tcount = 0;; ++tcount) {
            unsigned tile_g = 0;
            if (lane == 0) tile_g = atomicAdd(counter, 1u);
            tile_g = __shfl_sync(0xffffffffu, tile_g, 0);
            const int slot = tcount & 1;
            mbar_wait(&bar_tempty[slot], ((tcount >> 1) & 1) ^ 1);
            if (tile_g >= total_tiles) {
                if (lane == 0) {
                    tinfo[slot].valid = 0;
                    mbar_arrive(&bar_tfull[slot]);
                }
                break;
            }
            const int idx = find_inst(tile_starts, n_insts, tile_g);
            const BigInst inst = insts[idx];
            const uint32_t tile = tile_g - __ldg(tile_starts + idx);
            const BigStep* __restrict__ d = inst.step;
            if (done && idx != waited_idx) {  // dataflow launch: both operands must be complete before anything of this instance is read
                if (lane < 2) {  // both operands' counters are polled concurrently
                    const int dep = lane ? inst.dep_b : inst.dep_a;
                    if (dep >= 0) dep_wait(done + dep);
                }
                __syncwarp();
                fence_proxy_async();
                waited_idx = idx;
            }
            if (DF && d->kind != KIND_GEMM) {  // a generic step's tile: hand the descriptor to the consumer warps
                TileInfo& tg = tinfo[slot];
                uint32_t* gd = reinterpret_cast<uint32_t*>(&s_gstep[slot]);
                const uint32_t* gs = reinterpret_cast<const uint32_t*>(d);
                for (int w = lane; w < (int)(sizeof(BigStep) / 4); w += 32) gd[w] = __ldg(gs + w);
                if (lane == 0) {
                    tg.kind = KIND_GENERIC;
                    tg.tile = tile;
                    tg.pool = inst.pool;
                    tg.arena = inst.arena;
                    tg.done = done ? done + idx : nullptr;
                    tg.valid = 1;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_tfull[slot]);
                continue;
            }
            const int tm = d->tm, tn = d->tn, nk = d->nk, kc = d->kc, ng = d->ng;
            const int s_log = 8 - (tm + tn - 6), S = 1 << s_log;
            const T* Ag = reinterpret_cast<const T*>(inst.arena) + d->a_off;
            const T* Bg = reinterpret_cast<const T*>(inst.arena) + d->b_off;
            long long ab = -1, bb = -1, cb = -1;
            if (lane < S) {
                const unsigned long long g = (unsigned long long)tile * S + lane;
                if (g < (1ull << ng)) {
                    const int n_mhi = d->n_mhi, n_nhi = d->n_nhi;
                    const unsigned long long gm = g & ((1ull << n_mhi) - 1ull);
                    const unsigned long long gn = (g >> n_mhi) & ((1ull << n_nhi) - 1ull);
                    const unsigned long long gb = g >> (n_mhi + n_nhi);
                    ab = (long long)((gm | (gb << n_mhi)) << (tm + nk));
                    bb = (long long)((gn | (gb << n_nhi)) << (tn + nk));
                    cb = (long long)scatter_bits((uint32_t)g, d->c_shift + tm + tn, ng);
                }
            }
            TileInfo& ti = tinfo[slot];
            ti.cbase[lane] = cb;
            if (lane < 16) ti.c_shift[lane] = (lane < tm + tn) ? d->c_shift[lane] : NO_BIT;
            if (lane < 12) {
                ti.e_spos[lane] = d->a_shift[lane];
                ti.e_cs[lane] = d->b_shift[lane];
            }
            if (lane == 0) {
                ti.C = reinterpret_cast<T*>(inst.arena) + d->c_off;
                ti.tm = tm;
                ti.tn = tn;
                ti.kc = kc;
                ti.nchunks = 1 << (nk - kc);
                ti.store_mode = d->store_mode;
                ti.lane_n_first = d->lane_n_first;
                ti.e_cs_mtop = d->a_shift[30];
                ti.e_cs_ntop = d->a_shift[31];
                ti.e_vec = d->b_shift[31];
                ti.e_fast = d->b_shift[30];
                ti.kind = KIND_GEMM;
                ti.done = done ? done + idx : nullptr;
                ti.valid = 1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tfull[slot]);
            const int la = kc + tm, lb = kc + tn;
            const unsigned n_active = __popc(__ballot_sync(0xffffffffu, ab >= 0));
            const unsigned stage_bytes = n_active * ((1u << la) + (1u << lb)) * 4u;
            const int nchunks = 1 << (nk - kc);
            for (int ch = 0; ch < nchunks; ++ch, ++it) {
                const int stage = it % G2_STAGES;
                mbar_wait(&bar_empty[stage], ((it / G2_STAGES) & 1) ^ 1);
                if (lane == 0) mbar_arrive_expect_tx(&bar_full[stage], stage_bytes);
                __syncwarp();
                if (ab >= 0) {
                    T* sA = stage_mem + stage * GEMM_STAGE_ELEMS + ((size_t)lane << la);
                    T* sB = stage_mem + stage * GEMM_STAGE_ELEMS + ((size_t)S << la) + ((size_t)lane << lb);
                    bulk_g2s(sA, Ag + ab + ((long long)ch << la), (1u << la) * 4u, &bar_full[stage]);
                    bulk_g2s(sB, Bg + bb + ((long long)ch << lb), (1u << lb) * 4u, &bar_full[stage]);
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;\n" ::);
    const int ctid = tid - G2_PRODUCERS;
    const int lane = tid & 31;
    unsigned it = 0;
    for (unsigned tcount = 0;; ++tcount) {
        const int slot = tcount & 1;
        mbar_wait(&bar_tfull[slot], (tcount >> 1) & 1);
        const TileInfo& ti = tinfo[slot];
        if (!ti.valid) break;
        if constexpr (DF) {
            if (ti.kind != KIND_GEMM) {
                generic_tile_call<T>(&s_gstep[slot], ti.pool, ti.arena, ti.tile, ctid, stg_mem);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_sig[slot]);
                continue;
            }
        }
        const int tm = ti.tm, tn = ti.tn, kc = ti.kc, nchunks = ti.nchunks;
        const int tps_log = tm + tn - 6, S = 1 << (8 - tps_log);
        const int sub = ctid >> tps_log, lt = ctid & ((1 << tps_log) - 1);
        const int tmh = ti.lane_n_first ? (lt >> (tn - 3)) : (lt & ((1 << (tm - 3)) - 1));
        const int tnh = ti.lane_n_first ? (lt & ((1 << (tn - 3)) - 1)) : (lt >> (tm - 3));
        const int m_lo = tmh * 4, m_hi = (1 << (tm - 1)) + tmh * 4;
        const int n_lo = tnh * 4, n_hi = (1 << (tn - 1)) + tnh * 4;
        const int la = kc + tm, lb = kc + tn;
        T acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = Ops<T>::neg_inf();
        for (int ch = 0; ch < nchunks; ++ch, ++it) {
            const int stage = it % G2_STAGES;
            mbar_wait(&bar_full[stage], (it / G2_STAGES) & 1);
            const T* sA = stage_mem + stage * GEMM_STAGE_ELEMS + ((size_t)sub << la);
            const T* sB = stage_mem + stage * GEMM_STAGE_ELEMS + ((size_t)S << la) + ((size_t)sub << lb);
            const int KC = 1 << kc;
#pragma unroll 2
            for (int kk = 0; kk < KC; ++kk) {
                const T* ar = sA + (kk << tm);
                const T* br = sB + (kk << tn);
                const vec4 a0 = *reinterpret_cast<const vec4*>(ar + m_lo);
                const vec4 a1 = *reinterpret_cast<const vec4*>(ar + m_hi);
                const vec4 b0 = *reinterpret_cast<const vec4*>(br + n_lo);
                const vec4 b1 = *reinterpret_cast<const vec4*>(br + n_hi);
                const T a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const T b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = Ops<T>::addmax(a[i], b[j], acc[i][j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[stage]);
        }
        if (staged_epilogue) {
            // ---- staged epilogue: 4 rounds of (registers -> shared memory in tile order -> global in C-address order)
            const int nbr = tm + tn - 2;
            const uint32_t qbase = ((uint32_t)sub << nbr) | ((uint32_t)tmh << 2) | ((uint32_t)tnh << (tm + 1));
            uint32_t ts = 0, tc = 0;  // contribution of this thread's fixed element bits (2..9) that are tile bits
#pragma unroll
            for (int b = 2; b < 10; ++b) {
                const uint32_t bit = ((uint32_t)ctid >> (b - 2)) & 1u;
                if (b < nbr) {
                    ts |= bit << ti.e_spos[b];
                    tc |= bit << ti.e_cs[b];
                }
            }
            const uint32_t ds1 = 1u << ti.e_spos[0], ds2 = 1u << ti.e_spos[1];
            const bool evec = ti.e_vec != 0, efast = ti.e_fast != 0;
            uint32_t ts1 = 0, tc1 = 0;  // element-granular mapping (fallback): element bits 0..7 come from the thread id
            if (!evec) {
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const uint32_t bit = ((uint32_t)ctid >> b) & 1u;
                    if (b < nbr) {
                        ts1 |= bit << ti.e_spos[b];
                        tc1 |= bit << ti.e_cs[b];
                    }
                }
            }
            T* __restrict__ Cb = reinterpret_cast<T*>(ti.C);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int ih = r & 1, jh = r >> 1;
                T* buf = stg_mem + (r & 1) * G2_STG_ELEMS;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    vec4 v;
                    v.x = acc[ih * 4 + 0][jh * 4 + j];
                    v.y = acc[ih * 4 + 1][jh * 4 + j];
                    v.z = acc[ih * 4 + 2][jh * 4 + j];
                    v.w = acc[ih * 4 + 3][jh * 4 + j];
                    *reinterpret_cast<vec4*>(buf + stg_swz(qbase | ((uint32_t)j << (tm - 1)))) = v;
                }
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
                const uint32_t roff = ((uint32_t)ih << ti.e_cs_mtop) | ((uint32_t)jh << ti.e_cs_ntop);
                if (evec) {
#pragma unroll
                    for (int itr = 0; itr < 4; ++itr) {
                        const uint32_t e4 = ((uint32_t)itr << 10) | ((uint32_t)ctid << 2);
                        const uint32_t sub_e = e4 >> nbr;
                        uint32_t so = ts, co = tc;
#pragma unroll
                        for (int b = 10; b < 12; ++b) {
                            const uint32_t bit = ((uint32_t)itr >> (b - 10)) & 1u;
                            if (b < nbr) {
                                so |= bit << ti.e_spos[b];
                                co |= bit << ti.e_cs[b];
                            }
                        }
                        const long long cb = ti.cbase[sub_e];
                        if (cb >= 0) {
                            so |= sub_e << nbr;
                            vec4 o;
                            if (efast) {
                                o = *reinterpret_cast<const vec4*>(buf + stg_swz(so));
                            } else {
                                o.x = buf[stg_swz(so)];
                                o.y = buf[stg_swz(so | ds1)];
                                o.z = buf[stg_swz(so | ds2)];
                                o.w = buf[stg_swz(so | ds1 | ds2)];
                            }
                            *reinterpret_cast<vec4*>(Cb + cb + roff + co) = o;
                        }
                    }
                } else {
                    // C bits 0,1 are not both tile bits: consecutive LANES take consecutive tile elements in C order,
                    // so a warp still writes the densest address set this layout allows
#pragma unroll 4
                    for (int itr = 0; itr < 16; ++itr) {
                        const uint32_t e1 = ((uint32_t)itr << 8) | (uint32_t)ctid;
                        const uint32_t sub_e = e1 >> nbr;
                        uint32_t so = ts1, co = tc1;
#pragma unroll
                        for (int b = 8; b < 12; ++b) {
                            const uint32_t bit = ((uint32_t)itr >> (b - 8)) & 1u;
                            if (b < nbr) {
                                so |= bit << ti.e_spos[b];
                                co |= bit << ti.e_cs[b];
                            }
                        }
                        const long long cb = ti.cbase[sub_e];
                        if (cb >= 0) Cb[cb + roff + co] = buf[stg_swz(so | (sub_e << nbr))];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_sig[slot]);
            continue;
        }
        const long long cbase = ti.cbase[sub];
        if (cbase >= 0) {
            T* __restrict__ Ct = reinterpret_cast<T*>(ti.C) + cbase;
            uint32_t moff[8], noff[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t mi = (i < 4) ? (uint32_t)(m_lo + i) : (uint32_t)(m_hi + i - 4);
                const uint32_t ni = (i < 4) ? (uint32_t)(n_lo + i) : (uint32_t)(n_hi + i - 4);
                moff[i] = scatter_bits(mi, ti.c_shift, tm);
                noff[i] = scatter_bits(ni, ti.c_shift + tm, tn);
            }
            const int store_mode = ti.store_mode;
            if (store_mode == STORE_VEC_M) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int ih = 0; ih < 2; ++ih) {
                        vec4 v;
                        v.x = acc[ih * 4 + 0][j];
                        v.y = acc[ih * 4 + 1][j];
                        v.z = acc[ih * 4 + 2][j];
                        v.w = acc[ih * 4 + 3][j];
                        *reinterpret_cast<vec4*>(Ct + moff[ih * 4] + noff[j]) = v;
                    }
            } else if (store_mode == STORE_VEC_N) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int jh = 0; jh < 2; ++jh) {
                        vec4 v;
                        v.x = acc[i][jh * 4 + 0];
                        v.y = acc[i][jh * 4 + 1];
                        v.z = acc[i][jh * 4 + 2];
                        v.w = acc[i][jh * 4 + 3];
                        *reinterpret_cast<vec4*>(Ct + moff[i] + noff[jh * 4]) = v;
                    }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) Ct[moff[i] + noff[j]] = acc[i][j];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_sig[slot]);
    }
}

// ------------------------------------------------------------------------------------------------
// K2: packed int16x2 variant of k_gemm2.  Same producer / mbarrier / TMA structure.  Operands are stored
// [K0 | tile labels | K_rest | ...], so one 32-bit word holds the SAME (m) or (n) element for two consecutive k.
// A consumer thread owns 8 x 8 outputs as 8 x 8 packed accumulators whose halves collect the even-k and the odd-k
// partial maxima: per k-pair 2 x LDS.128 (A: 8 words) + 2 x LDS.128 (B: 8 words) and 64 VIADDMNMX.S16x2 = 128
// tropical ops, no operand duplication.  The epilogue takes max(lo, hi) per accumulator and stores int16.
// ------------------------------------------------------------------------------------------------
#ifdef TB_KPROF
// diagnostics build: SM-cycle accounting of k_gemm2h consumers (warp 0 of the consumers of every CTA)
// [0] waiting for a tile descriptor  [1] waiting for operand stages  [2] main loop  [3] epilogue  [4] CTA lifetime  [5] tiles
__device__ unsigned long long g_kprof[8];
#endif
constexpr int G2H_TSLOTS = 4;  // tile descriptors the producer may publish ahead of the consumers
constexpr int G2H_STG_ELEMS = 4096;  // int16 elements of one staged quarter of a 128 x 128 tile (8 KB)
struct TileInfoH {
    long long cbase[32];
    void* C;
    int tm, tn, kc, nchunks, valid, lane_n_first;
    unsigned char e_spos[14], e_cs[14];
    unsigned char e_cs_mtop, e_cs_ntop, e_vec, e_fast, mp, mswap;
    int inst;  // index of the (branch, step) instance in this launch: consumers cache per-step values on it
    // dataflow launches
    int kind;               // KIND_GEMM, or KIND_GENERIC: the consumer warps run generic_tile on s_gstep[slot]
    uint32_t tile;          // generic: index of this tile inside its instance
    const void* pool;       // generic: leaf pool / arena of the branch
    void* arena;
    unsigned int* done;     // completion counter of the instance (nullptr: level-synchronous launch)
};
__device__ __forceinline__ uint32_t stg_swz_h(uint32_t x) { return x ^ ((((x >> 6) ^ (x >> 9) ^ (x >> 12)) & 7u) << 3); }

// The k loop of one pipeline stage for a tile shape known at compile time: every shared-memory offset (k-pair row pitch,
// upper half of the m / n microtile) is an immediate of the load instruction, the only address arithmetic left is three
// adds per two k-pair rows (128 VIADDMNMX).  ua / ua_p = byte addresses of this thread's first A words (tile bit m0 / m_mp),
// ub = of its first B words, all in the stage being consumed.  Same software pipeline as the generic loop below: B double
// buffered in registers, the two halves of A reloaded in place right after their last use.
template <int TM, int TN>
__device__ __forceinline__ void g2h_stage_loop(uint32_t ua, uint32_t ua_p, uint32_t ub, int kp_rows, uint32_t (&acc)[8][8]) {
    constexpr int ROWA = 4 << TM, ROWB = 4 << TN, A_HI = 2 << TM, B_HI = 4 << (TN - 1);
    uint32_t a[8], b0[8], b1[8];
#define G2T_LDA_LO(OFF)                                                                                                  \
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(a[0]), "=r"(a[1]) : "r"(ua), "n"(OFF));                     \
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(a[2]), "=r"(a[3]) : "r"(ua_p), "n"(OFF));
#define G2T_LDA_HI(OFF)                                                                                                  \
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(a[4]), "=r"(a[5]) : "r"(ua), "n"((OFF) + A_HI));            \
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(a[6]), "=r"(a[7]) : "r"(ua_p), "n"((OFF) + A_HI));
#define G2T_LDB(bb, OFF)                                                                                                 \
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(bb[0]), "=r"(bb[1]), "=r"(bb[2]), "=r"(bb[3]) : "r"(ub), "n"(OFF)); \
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(bb[4]), "=r"(bb[5]), "=r"(bb[6]), "=r"(bb[7]) : "r"(ub), "n"((OFF) + B_HI));
#define G2T_MATH(i0, bb)                                                                                                 \
    _Pragma("unroll") for (int i = i0; i < i0 + 4; ++i)                                                                   \
        _Pragma("unroll") for (int j = 0; j < 8; ++j) acc[i][j] = __viaddmax_s16x2(a[i], bb[j], acc[i][j]);
    G2T_LDA_LO(0)
    G2T_LDB(b0, 0)
    int kk = 0;
#pragma unroll 1
    for (int n2 = kp_rows >> 1; n2 > 0; --n2) {  // two k-pair rows (128 VIADDMNMX) per trip: 3 address adds + a countdown
        G2T_LDA_HI(0)
        G2T_LDB(b1, ROWB)
        G2T_MATH(0, b0)
        G2T_LDA_LO(ROWA)
        G2T_MATH(4, b0)
        G2T_LDA_HI(ROWA)
        G2T_LDB(b0, 2 * ROWB)  // the last one reads one row past the chunk (still shared memory): discarded
        G2T_MATH(0, b1)
        G2T_LDA_LO(2 * ROWA)
        G2T_MATH(4, b1)
        ua += 2 * ROWA;
        ua_p += 2 * ROWA;
        ub += 2 * ROWB;
    }
    kk = kp_rows & ~1;
    if (kk < kp_rows) {  // a single k-pair row (kc == 1)
        G2T_LDA_HI(0)
        G2T_MATH(0, b0)
        G2T_MATH(4, b0)
    }
#undef G2T_LDA_LO
#undef G2T_LDA_HI
#undef G2T_LDB
#undef G2T_MATH
}

// All pipeline stages of one tile for a compile-time tile shape: wait for the stage, run its k loop, hand it back.
// a_thr / b_thr = this thread's A / B byte addresses in stage 0; full0 / empty0 = shared addresses of bar_full[0] / bar_empty[0].
template <int TM, int TN>
__device__ __forceinline__ void g2h_tile_mainloop(uint32_t a_thr, uint32_t a_p, uint32_t b_thr, int kp_rows, int nchunks, unsigned& it,
                                                  uint32_t full0, uint32_t empty0, int lane, uint32_t (&acc)[8][8]) {
    constexpr uint32_t STAGE_BYTES = (uint32_t)GEMM_STAGE_ELEMS * 4u;  // 2 * GEMM_STAGE_ELEMS int16 elements
#pragma unroll 1
    for (int ch = 0; ch < nchunks; ++ch, ++it) {
        const unsigned stage = it % G2_STAGES;
        mbar_wait_addr(full0 + stage * 8u, (it / G2_STAGES) & 1u);
        const uint32_t ua = a_thr + stage * STAGE_BYTES;
        g2h_stage_loop<TM, TN>(ua, ua + a_p, b_thr + stage * STAGE_BYTES, kp_rows, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive_addr(empty0 + stage * 8u);
    }
}

// DF = dataflow launch (all levels of a wave, generic tiles included, completion counters).  The level-synchronous
// instance (DF = false) contains no generic-tile code at all: with it in the same function ptxas no longer proves the
// main loop's address arithmetic warp-uniform (no uniform-datapath instructions), which costs 7 % on DPX-bound workloads.
template <bool DF>
__global__ void __launch_bounds__(G2_THREADS, 2) k_gemm2h(const BigInst* __restrict__ insts, const uint32_t* __restrict__ tile_starts,
                                                          int n_insts, uint32_t total_tiles, unsigned int* __restrict__ counter,
                                                          unsigned int* done) {
    typedef int16_t T;
    if (!DF) done = nullptr;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    T* stage_mem = reinterpret_cast<T*>(dyn_smem);                       // G2_STAGES x 16 KB
    T* stg_mem = reinterpret_cast<T*>(dyn_smem + G2_RING_BYTES);         // 2 x 16 KB
    constexpr int STAGE_ELEMS = GEMM_STAGE_ELEMS * 2;                    // int16 elements per stage
    __shared__ __align__(8) uint64_t bar_full[G2_STAGES], bar_empty[G2_STAGES], bar_tfull[G2H_TSLOTS], bar_tempty[G2H_TSLOTS],
        bar_sig[G2H_TSLOTS];
    __shared__ TileInfoH tinfo[G2H_TSLOTS];
    __shared__ __align__(16) BigStep s_step;  // the producer's copy of the current step descriptor
    __shared__ __align__(16) BigStep s_gstep[G2H_TSLOTS];  // generic tiles: the descriptor the consumers execute
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < G2_STAGES; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], G2_CONSUMERS / 32);
        }
        for (int i = 0; i < G2H_TSLOTS; ++i) {
            mbar_init(&bar_tfull[i], 1);
            mbar_init(&bar_tempty[i], 1);                  // the signal warp hands the slot back
            mbar_init(&bar_sig[i], G2_CONSUMERS / 32);     // every consumer warp is done with the tile
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    if (tid < G2_PRODUCERS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n" ::);
        if (tid >= 64) return;
        if (tid >= 32) {
            // ------------------------------------------------------------------ signal warp: publishes finished tiles
            if (tid != 32) return;
            for (unsigned tcount = 0;; ++tcount) {
                const int slot = tcount % G2H_TSLOTS;
                const unsigned par = (tcount / G2H_TSLOTS) & 1;
                mbar_wait(&bar_tfull[slot], par);
                if (!tinfo[slot].valid) break;
                unsigned int* const dp = tinfo[slot].done;
                mbar_wait(&bar_sig[slot], par);
                if (dp) dep_signal(dp, G2_CONSUMERS / 32);
                mbar_arrive(&bar_tempty[slot]);
            }
            return;
        }
        const int lane = tid;
        unsigned it = 0;
        // the tile counter is read one tile ahead (the atomic's latency overlaps the current tile's set-up), and the
        // step descriptor is decoded once per (branch, step) instance into shared memory: consecutive tiles mostly
        // belong to the same instance
        unsigned tile_next = blockIdx.x;  // first tile: no atomic on the launch's critical path (the counter hands out gridDim.x + ...)
        uint32_t cur_start = 1, cur_end = 0;  // tile range of the cached instance (empty)
        int cur_idx = -1;
        const unsigned char* cur_arena = nullptr;
        const void* cur_pool = nullptr;
        for (unsigned tcount = 0;; ++tcount) {
            const unsigned tile_g = __shfl_sync(0xffffffffu, tile_next, 0);
            if (lane == 0 && tile_g < total_tiles) tile_next = atomicAdd(counter, 1u) + gridDim.x;
            const int slot = tcount % G2H_TSLOTS;
            mbar_wait(&bar_tempty[slot], ((tcount / G2H_TSLOTS) & 1) ^ 1);
            if (tile_g >= total_tiles) {
                if (lane == 0) {
                    tinfo[slot].valid = 0;
                    mbar_arrive(&bar_tfull[slot]);
                }
                break;
            }
            if (tile_g < cur_start || tile_g >= cur_end) {
                cur_idx = find_inst(tile_starts, n_insts, tile_g);
                const BigInst inst = insts[cur_idx];
                cur_start = __ldg(tile_starts + cur_idx);
                cur_arena = reinterpret_cast<const unsigned char*>(inst.arena);
                cur_pool = inst.pool;
                const uint32_t* gs = reinterpret_cast<const uint32_t*>(inst.step);
                uint32_t* ss = reinterpret_cast<uint32_t*>(&s_step);
                for (int w = lane; w < (int)(sizeof(BigStep) / 4); w += 32) ss[w] = __ldg(gs + w);
                __syncwarp();
                cur_end = cur_start + s_step.n_tiles;
                if (done) {  // dataflow launch: both operands must be complete before anything of this instance is read
                    if (lane < 2) {  // both operands' counters are polled concurrently
                        const int dep = lane ? inst.dep_b : inst.dep_a;
                        if (dep >= 0) dep_wait(done + dep);
                    }
                    __syncwarp();
                    fence_proxy_async();
                }
            }
            const int idx = cur_idx;
            const uint32_t tile = tile_g - cur_start;
            const BigStep* d = &s_step;
            const unsigned char* arena = cur_arena;
            if (DF && d->kind != KIND_GEMM) {  // a generic step's tile: hand the descriptor to the consumer warps, nothing to stage
                TileInfoH& tg = tinfo[slot];
                uint32_t* gd = reinterpret_cast<uint32_t*>(&s_gstep[slot]);
                const uint32_t* ss = reinterpret_cast<const uint32_t*>(&s_step);
                for (int w = lane; w < (int)(sizeof(BigStep) / 4); w += 32) gd[w] = ss[w];
                if (lane == 0) {
                    tg.kind = KIND_GENERIC;
                    tg.tile = tile;
                    tg.pool = cur_pool;
                    tg.arena = const_cast<unsigned char*>(arena);
                    tg.done = done ? done + idx : nullptr;
                    tg.inst = idx;
                    tg.valid = 1;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_tfull[slot]);
                continue;
            }
            const int tm = d->tm, tn = d->tn, nk = d->nk, kc = d->kc, ng = d->ng;
            const int s_log = 8 - (tm + tn - 6), S = 1 << s_log;
            const T* Ag = reinterpret_cast<const T*>(arena) + d->a_off;
            const T* Bg = reinterpret_cast<const T*>(arena) + d->b_off;
            long long ab = -1, bb = -1, cb = -1;
            if (lane < S) {
                const unsigned long long g = (unsigned long long)tile * S + lane;
                if (g < (1ull << ng)) {
                    const int n_mhi = d->n_mhi, n_nhi = d->n_nhi;
                    const unsigned long long gm = g & ((1ull << n_mhi) - 1ull);
                    const unsigned long long gn = (g >> n_mhi) & ((1ull << n_nhi) - 1ull);
                    const unsigned long long gb = g >> (n_mhi + n_nhi);
                    ab = (long long)((gm | (gb << n_mhi)) << (tm + nk));
                    bb = (long long)((gn | (gb << n_nhi)) << (tn + nk));
                    cb = (long long)scatter_bits((uint32_t)g, d->c_shift + tm + tn, ng);
                }
            }
            TileInfoH& ti = tinfo[slot];
            ti.cbase[lane] = cb;
            if (lane < 14) {
                ti.e_spos[lane] = d->a_shift[lane];
                ti.e_cs[lane] = d->b_shift[lane];
            }
            if (lane == 0) {
                ti.C = const_cast<T*>(reinterpret_cast<const T*>(arena)) + d->c_off;
                ti.tm = tm;
                ti.tn = tn;
                ti.kc = kc;
                ti.nchunks = 1 << (nk - kc);
                ti.lane_n_first = d->lane_n_first;
                ti.e_cs_mtop = d->a_shift[30];
                ti.e_cs_ntop = d->a_shift[31];
                ti.e_vec = d->b_shift[31];
                ti.e_fast = d->b_shift[30];
                ti.mp = d->a_shift[29];
                ti.mswap = d->b_shift[29];
                ti.inst = idx;
                ti.kind = KIND_GEMM;
                ti.done = done ? done + idx : nullptr;
                ti.valid = 1;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tfull[slot]);
            const int la = kc + tm, lb = kc + tn;
            const unsigned n_active = __popc(__ballot_sync(0xffffffffu, ab >= 0));
            const unsigned stage_bytes = n_active * ((1u << la) + (1u << lb)) * 2u;
            const int nchunks = 1 << (nk - kc);
            for (int ch = 0; ch < nchunks; ++ch, ++it) {
                const int stage = it % G2_STAGES;
                mbar_wait(&bar_empty[stage], ((it / G2_STAGES) & 1) ^ 1);
                if (lane == 0) mbar_arrive_expect_tx(&bar_full[stage], stage_bytes);
                __syncwarp();
                if (ab >= 0) {
                    T* sA = stage_mem + stage * STAGE_ELEMS + ((size_t)lane << la);
                    T* sB = stage_mem + stage * STAGE_ELEMS + ((size_t)S << la) + ((size_t)lane << lb);
                    bulk_g2s(sA, Ag + ab + ((long long)ch << la), (1u << la) * 2u, &bar_full[stage]);
                    bulk_g2s(sB, Bg + bb + ((long long)ch << lb), (1u << lb) * 2u, &bar_full[stage]);
                }
            }
        }
        return;
    }

    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;\n" ::);
    const int ctid = tid - G2_PRODUCERS;
    const int lane = tid & 31;
    unsigned it = 0;
    int ep_inst = -1;  // the step whose epilogue offsets (ts, tc) this thread holds
    uint32_t ts = 0, tc = 0;
#ifdef TB_KPROF
    const unsigned long long tl_t0 = tl_now();
    long long kp_t0 = clock64(), kp_last = kp_t0, kp_acc[4] = {0, 0, 0, 0};
    unsigned kp_tiles = 0;
#define KP(i) { const long long n_ = clock64(); kp_acc[i] += n_ - kp_last; kp_last = n_; }
#else
#define KP(i)
#endif
    for (unsigned tcount = 0;; ++tcount) {
        const int slot = tcount % G2H_TSLOTS;
        mbar_wait(&bar_tfull[slot], (tcount / G2H_TSLOTS) & 1);
        KP(0)
        const TileInfoH& ti = tinfo[slot];
        if (!ti.valid) break;
        if constexpr (DF) {
            if (ti.kind != KIND_GEMM) {
                generic_tile_call<T>(&s_gstep[slot], ti.pool, ti.arena, ti.tile, ctid, stg_mem);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_sig[slot]);
                KP(3)
                continue;
            }
        }
        const int tm = ti.tm, tn = ti.tn, kc = ti.kc;
        int nchunks = ti.nchunks;
        const int tps_log = tm + tn - 6, S = 1 << (8 - tps_log);
        const int sub = ctid >> tps_log, lt = ctid & ((1 << tps_log) - 1);
        const int tmh = ti.lane_n_first ? (lt >> (tn - 3)) : (lt & ((1 << (tm - 3)) - 1));
        const int tnh = ti.lane_n_first ? (lt & ((1 << (tn - 3)) - 1)) : (lt >> (tm - 3));
        // a k-pair row of A is 2^tm words (2^(tm+1) int16): word m holds (A[m, k even], A[m, k odd]).  This thread
        // owns the m tile bits {0, mp, tm-1}; the bits of tmh fill the other positions in ascending order.
        const int mp = ti.mp;
        const int tmw = ((tmh & ((1 << (mp - 1)) - 1)) << 1) | ((tmh >> (mp - 1)) << (mp + 1));  // word offset
        const int m_lo = tmw * 2, m_p = 2 << mp, m_hi = 2 << (tm - 1);  // int16 offsets inside a row
        const int n_lo = tnh * 8;
        const int la = kc + tm, lb = kc + tn;
        uint32_t acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0xC000C000u;  // (-2^14, -2^14)
        // per-thread shared-memory byte addresses of its A / B words in stage 0, k-pair row 0; everything added inside
        // the loops is warp-uniform
        const uint32_t stage_base = (uint32_t)__cvta_generic_to_shared(stage_mem);
        const uint32_t a_thr = stage_base + ((((uint32_t)sub << la) + (uint32_t)m_lo) << 1);
        const uint32_t b_thr = stage_base + ((((uint32_t)S << la) + ((uint32_t)sub << lb) + (uint32_t)n_lo) << 1);
        const uint32_t a_p = (uint32_t)m_p << 1, a_hi = (uint32_t)m_hi << 1, b_hi = 4u << (tn - 1);
        const uint32_t rowA = 4u << tm, rowB = 4u << tn;  // bytes per k-pair row
        // dataflow instance: the common tile shapes run the whole main loop with compile-time offsets (g2h_tile_mainloop)
        if constexpr (DF) {
            const uint32_t full0 = (uint32_t)__cvta_generic_to_shared(&bar_full[0]);
            const uint32_t empty0 = (uint32_t)__cvta_generic_to_shared(&bar_empty[0]);
            const int kpr = 1 << (kc - 1);
            bool fast = true;
            switch (tm * 8 + tn) {
                case 7 * 8 + 7: g2h_tile_mainloop<7, 7>(a_thr, a_p, b_thr, kpr, nchunks, it, full0, empty0, lane, acc); break;
                case 7 * 8 + 6: g2h_tile_mainloop<7, 6>(a_thr, a_p, b_thr, kpr, nchunks, it, full0, empty0, lane, acc); break;
                case 6 * 8 + 7: g2h_tile_mainloop<6, 7>(a_thr, a_p, b_thr, kpr, nchunks, it, full0, empty0, lane, acc); break;
                default: fast = false;
            }
            if (fast) nchunks = 0;  // nothing left for the generic loop below
            KP(2)
        }
        for (int ch = 0; ch < nchunks; ++ch, ++it) {
            const int stage = it % G2_STAGES;
            mbar_wait(&bar_full[stage], (it / G2_STAGES) & 1);
            KP(1)
            const int KP_ROWS = 1 << (kc - 1);  // k-pair rows in this chunk
            // Software pipeline over the k-pair rows: B is double-buffered in registers, the first half of A
            // (rows 0..3 of the microtile) is reloaded in place as soon as its last use has issued, the second half at
            // the top of the step whose second half needs it: no load is followed directly by its consumers, so the
            // two warps a lone CTA has per scheduler can keep the VIADDMNMX pipe busy.
            uint32_t ua = a_thr + (uint32_t)stage * (uint32_t)(STAGE_ELEMS * 2);
            uint32_t ub = b_thr + (uint32_t)stage * (uint32_t)(STAGE_ELEMS * 2);
            uint32_t a[8], b0[8], b1[8];
#define G2H_LDA_LO(addr)                                                                                            \
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a[0]), "=r"(a[1]) : "r"(addr));                           \
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a[2]), "=r"(a[3]) : "r"((addr) + a_p));
#define G2H_LDA_HI(addr)                                                                                            \
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a[4]), "=r"(a[5]) : "r"((addr) + a_hi));                  \
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a[6]), "=r"(a[7]) : "r"((addr) + a_hi + a_p));
#define G2H_LDB(bb, addr)                                                                                           \
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(bb[0]), "=r"(bb[1]), "=r"(bb[2]), "=r"(bb[3]) : "r"(addr)); \
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(bb[4]), "=r"(bb[5]), "=r"(bb[6]), "=r"(bb[7]) : "r"((addr) + b_hi));
#define G2H_MATH(i0, bb)                                                                                            \
    _Pragma("unroll") for (int i = i0; i < i0 + 4; ++i)                                                              \
        _Pragma("unroll") for (int j = 0; j < 8; ++j) acc[i][j] = __viaddmax_s16x2(a[i], bb[j], acc[i][j]);
            G2H_LDA_LO(ua)
            G2H_LDB(b0, ub)
            int kk = 0;
#pragma unroll 1
            for (; kk + 2 <= KP_ROWS; kk += 2) {
                G2H_LDA_HI(ua)
                G2H_LDB(b1, ub + rowB)
                G2H_MATH(0, b0)
                G2H_LDA_LO(ua + rowA)
                G2H_MATH(4, b0)
                ua += rowA;
                ub += rowB;
                G2H_LDA_HI(ua)
                G2H_LDB(b0, ub + rowB)  // the last one reads one row past the chunk (still shared memory): discarded
                G2H_MATH(0, b1)
                G2H_LDA_LO(ua + rowA)
                G2H_MATH(4, b1)
                ua += rowA;
                ub += rowB;
            }
            if (kk < KP_ROWS) {  // a single k-pair row (kc == 1)
                G2H_LDA_HI(ua)
                G2H_MATH(0, b0)
                G2H_MATH(4, b0)
            }
#undef G2H_LDA_LO
#undef G2H_LDA_HI
#undef G2H_LDB
#undef G2H_MATH
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[stage]);
            KP(2)
        }
        // ---- staged epilogue: out = max(even-k half, odd-k half); int16 elements, 8 per 16-byte global vector.
        // Everything that does not depend on the round is computed once per tile (the ALU pipe that executes these
        // instructions is the one the main loop's VIADDMNMX need).
        {
            const int nbr = tm + tn - 2;
            const uint32_t qbase = ((uint32_t)sub << nbr) | ((uint32_t)tmh << 2) | ((uint32_t)tnh << (tm + 1));
            if (ti.inst != ep_inst) {  // consecutive tiles of a persistent CTA mostly belong to the same step
                ep_inst = ti.inst;
                ts = tc = 0;
#pragma unroll
                for (int b = 3; b < 11; ++b) {
                    const uint32_t bit = ((uint32_t)ctid >> (b - 3)) & 1u;
                    if (b < nbr) {
                        ts |= bit << ti.e_spos[b];
                        tc |= bit << ti.e_cs[b];
                    }
                }
            }
            const bool evec = ti.e_vec != 0;
            const int ecase = ti.e_fast;  // staging layout (see plan.cpp): 0/1 m-major, 2..4 16-byte vectors per thread
            const bool mswap = ti.mswap != 0;  // the thread pairs its outputs along m_mp instead of m0
            T* __restrict__ Cb = reinterpret_cast<T*>(ti.C);
            const uint32_t stg_base = (uint32_t)__cvta_generic_to_shared(stg_mem);
            // staging write addresses (bytes, buffer 0)
            uint32_t w_addr[4];
            if (ecase <= 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) w_addr[j] = stg_base + (stg_swz_h(qbase | ((uint32_t)j << (tm - 1))) << 1);
            } else {
                const uint32_t qbase2 = ((uint32_t)sub << nbr) | ((uint32_t)tmh << 4) | ((uint32_t)tnh << (tm + 1));
                w_addr[0] = stg_base + (stg_swz_h(qbase2) << 1);
                w_addr[1] = w_addr[0] ^ 16u;  // staging index bit 3 (the swizzle never reads it)
                w_addr[2] = w_addr[3] = 0;
            }
            // read side of the 128-bit path: two vectors per thread and round
            uint32_t r_addr[2] = {0, 0};
            T* g_ptr[2] = {nullptr, nullptr};
            if (evec && ecase != 0) {
#pragma unroll
                for (int itr = 0; itr < 2; ++itr) {
                    const uint32_t e8 = ((uint32_t)itr << 11) | ((uint32_t)ctid << 3);
                    const uint32_t sub_e = e8 >> nbr;
                    uint32_t so = ts, co = tc;
                    if (11 < nbr) {
                        so |= (uint32_t)itr << ti.e_spos[11];
                        co |= (uint32_t)itr << ti.e_cs[11];
                    }
                    const long long cb = ti.cbase[sub_e];
                    r_addr[itr] = stg_base + (stg_swz_h(so | (sub_e << nbr)) << 1);
                    g_ptr[itr] = cb >= 0 ? Cb + cb + co : nullptr;
                }
            }
            uint32_t ts1 = 0, tc1 = 0;
            if (!evec) {
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    const uint32_t bit = ((uint32_t)ctid >> b) & 1u;
                    if (b < nbr) {
                        ts1 |= bit << ti.e_spos[b];
                        tc1 |= bit << ti.e_cs[b];
                    }
                }
            }
            // The four quarters of the tile (top m bit x top n bit) have a staging buffer each (4 x 8 KB = the whole 128 x 128
            // int16 tile), so ONE barrier separates all shared-memory writes from all reads; the barrier in front only
            // orders this tile's writes behind the previous tile's reads (every warp passed those a main loop ago).
            asm volatile("bar.sync 1, 256;\n" ::: "memory");
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int ih = r & 1, jh = r >> 1;
                const uint32_t boff = (uint32_t)r * (uint32_t)(G2H_STG_ELEMS * 2);  // bytes
                uint32_t W[2][4];  // W[second local m bit][n0 + 2 n1] = outputs (first local m bit = 0, 1) as one packed word
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i1 = 0; i1 < 2; ++i1) {
                        // accumulator rows: bit 0 = m0, bit 1 = m_mp, bit 2 = top m bit
                        const uint32_t w0 = mswap ? acc[ih * 4 + i1][jh * 4 + j] : acc[ih * 4 + 2 * i1][jh * 4 + j];
                        const uint32_t w1 = mswap ? acc[ih * 4 + i1 + 2][jh * 4 + j] : acc[ih * 4 + 2 * i1 + 1][jh * 4 + j];
                        // (lo0, lo1) vs (hi0, hi1): max of the even-k and odd-k partial maxima of both outputs at once
                        W[i1][j] = __vmaxs2(__byte_perm(w0, w1, 0x5410), __byte_perm(w0, w1, 0x7632));
                    }
                if (ecase <= 1) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(w_addr[j] + boff), "r"(W[0][j]), "r"(W[1][j]) : "memory");
                } else {
                    uint4 v0, v1;
                    if (ecase == 2) {  // word order (m1, n0), second vector n1 = 1
                        v0 = make_uint4(W[0][0], W[1][0], W[0][1], W[1][1]);
                        v1 = make_uint4(W[0][2], W[1][2], W[0][3], W[1][3]);
                    } else if (ecase == 3) {  // (n0, m1), second vector n1 = 1
                        v0 = make_uint4(W[0][0], W[0][1], W[1][0], W[1][1]);
                        v1 = make_uint4(W[0][2], W[0][3], W[1][2], W[1][3]);
                    } else {  // (n0, n1), second vector m1 = 1
                        v0 = make_uint4(W[0][0], W[0][1], W[0][2], W[0][3]);
                        v1 = make_uint4(W[1][0], W[1][1], W[1][2], W[1][3]);
                    }
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(w_addr[0] + boff), "r"(v0.x), "r"(v0.y), "r"(v0.z), "r"(v0.w) : "memory");
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(w_addr[1] + boff), "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w) : "memory");
                }
            }
            asm volatile("bar.sync 1, 256;\n" ::: "memory");
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int ih = r & 1, jh = r >> 1;
                const uint32_t boff = (uint32_t)r * (uint32_t)(G2H_STG_ELEMS * 2);  // bytes
                T* buf = stg_mem + r * G2H_STG_ELEMS;
                const uint32_t roff = ((uint32_t)ih << ti.e_cs_mtop) | ((uint32_t)jh << ti.e_cs_ntop);
                if (evec && ecase != 0) {
#pragma unroll
                    for (int itr = 0; itr < 2; ++itr) {
                        if (g_ptr[itr]) {
                            uint4 o;
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"(r_addr[itr] + boff) : "memory");
                            *reinterpret_cast<uint4*>(g_ptr[itr] + roff) = o;
                        }
                    }
                } else if (evec) {  // C bits 0..2 are tile bits but not contiguous in the staging buffer: gather
#pragma unroll
                    for (int itr = 0; itr < 2; ++itr) {
                        const uint32_t e8 = ((uint32_t)itr << 11) | ((uint32_t)ctid << 3);
                        const uint32_t sub_e = e8 >> nbr;
                        uint32_t so = ts, co = tc;
                        if (11 < nbr) {
                            so |= (uint32_t)itr << ti.e_spos[11];
                            co |= (uint32_t)itr << ti.e_cs[11];
                        }
                        const long long cb = ti.cbase[sub_e];
                        if (cb >= 0) {
                            so |= sub_e << nbr;
                            T v[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const uint32_t dq = ((q & 1) << ti.e_spos[0]) | (((q >> 1) & 1) << ti.e_spos[1]) | (((q >> 2) & 1) << ti.e_spos[2]);
                                v[q] = buf[stg_swz_h(so | dq)];
                            }
                            uint4 o;
                            o.x = (uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
                            o.y = (uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16);
                            o.z = (uint32_t)(uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16);
                            o.w = (uint32_t)(uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16);
                            *reinterpret_cast<uint4*>(Cb + cb + roff + co) = o;
                        }
                    }
                } else {
#pragma unroll 4
                    for (int itr = 0; itr < 16; ++itr) {
                        const uint32_t e1 = ((uint32_t)itr << 8) | (uint32_t)ctid;
                        const uint32_t sub_e = e1 >> nbr;
                        uint32_t so = ts1, co = tc1;
#pragma unroll
                        for (int b = 8; b < 12; ++b) {
                            const uint32_t bit = ((uint32_t)itr >> (b - 8)) & 1u;
                            if (b < nbr) {
                                so |= bit << ti.e_spos[b];
                                co |= bit << ti.e_cs[b];
                            }
                        }
                        const long long cb = ti.cbase[sub_e];
                        if (cb >= 0) Cb[cb + roff + co] = buf[stg_swz_h(so | (sub_e << nbr))];
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_sig[slot]);
        KP(3)
#ifdef TB_KPROF
        ++kp_tiles;
#endif
    }
#ifdef TB_KPROF
    if (ctid == 0) {
        tl_log(insts, 2u, tl_t0);
        for (int q = 0; q < 4; ++q) atomicAdd(&g_kprof[q], (unsigned long long)kp_acc[q]);
        atomicAdd(&g_kprof[4], (unsigned long long)(clock64() - kp_t0));
        atomicAdd(&g_kprof[5], (unsigned long long)kp_tiles);
    }
#endif
#undef KP
}

// ------------------------------------------------------------------------------------------------
// root scalars -> result vector
// ------------------------------------------------------------------------------------------------
struct FinalInst {
    const void* src;
    int64_t out_index;
};

template <typename T>
__global__ void k_finalize(const FinalInst* __restrict__ f, int n, double* __restrict__ results) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) results[f[i].out_index] = Ops<T>::to_double(*reinterpret_cast<const T*>(f[i].src));
}

// multi-GPU calls: every device's result vector starts as -inf so that ONE all-reduce(max) assembles the per-branch vector
__global__ void k_fill_double(double* __restrict__ dst, int64_t n, double v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

// low 32 bits of the size + configuration elements: the chosen vertices (0 where the element is tropical zero)
__global__ void k_to_config(const long long* __restrict__ src, uint32_t* __restrict__ dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (src[i] >> 32) <= -(1ll << 29) ? 0u : (uint32_t)(src[i] & 0xffffffffll);
}

// Branching-table reduction (mis_compactify of the reference's table solver): one stage of the subset-max transform.
// After the stages of all bits, z[a] = max over b subset-of a of the input.  A thread owns the pair (a, a | 1 << bit).
__global__ void k_subset_max_stage(double* __restrict__ z, int bit, int64_t n_pairs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const int64_t low = i & (((int64_t)1 << bit) - 1);
    const int64_t a = ((i >> bit) << (bit + 1)) | low;
    const int64_t b = a | ((int64_t)1 << bit);
    z[b] = fmax(z[b], z[a]);
}

// keep[a] = 1 iff entry a is feasible and no entry b that chooses a STRICT subset of a's boundary vertices is at least as
// large: the strict subsets of a are the subsets of a without vertex i, for the vertices i of a.
__global__ void k_table_keep(const double* __restrict__ sizes, const double* __restrict__ z, int rank, uint8_t* __restrict__ keep, int64_t n) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const double v = sizes[a];
    double best = -INFINITY;
    for (int i = 0; i < rank; ++i)
        if ((a >> i) & 1) best = fmax(best, z[a ^ ((int64_t)1 << i)]);
    keep[a] = (v > -INFINITY && !(best >= v)) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// ALL optimal configurations of branching tables (the ConfigsMax element type of the reference's table solver,
// reached from src/branch.jl:79), for a BATCH of regions in the same launches.  A region has at most 32 vertices
// (n_max = 20 by default, src/types.jl:10), so the set algebra of the reference (unions and cartesian products of
// configuration sets carried through the contraction) is replaced by a filter over the 2^n vertex sets of the region: a
// set belongs to the row of its boundary configuration a iff it is independent and its weight equals the row's optimum.
// Index space: (region, boundary configuration a, chunk of the 2^n_int interior configurations); one CTA per triple (the
// region is found by bisection over the regions' first-CTA numbers), a thread walks the chunk with stride blockDim.
//   pass 0: alpha[row] = max weight (atomicMax on an order-preserving integer image of the double)
//   (mis_compactify of every region's alpha: subset-max stages + keep flags, batched over all rows)
//   pass 1: chunk_count[cta] = number of optimal sets in the chunk (rows with keep == 0 count nothing)
//   pass 2: the sets are written at chunk_off[cta] in ascending interior index (block-wide ordered compaction), so the
//           output is deterministic: regions in order, rows in boundary-configuration order, each row sorted.
// ------------------------------------------------------------------------------------------------
struct RegionDesc {
    double w[32];           // weight of vertex v (sum of its vertex tensors' weights)
    uint32_t adj[32];       // neighbours of vertex v inside the region
    uint8_t bpos[32];       // boundary bit i of a row index  -> vertex
    uint8_t ipos[32];       // interior bit i of a chunk index -> vertex
    int32_t n, rank, n_int; // vertices, boundary vertices, interior vertices
    int32_t chunk_log2;     // interior configurations per CTA = 2^chunk_log2
    int64_t row_base;       // first row of the region in the batch's row arrays (sum of 2^rank of the regions before it)
    int64_t cta_base;       // first CTA of the region
};
static_assert(sizeof(RegionDesc) % 8 == 0, "RegionDesc is copied as 8-byte words");
static constexpr int kRegionThreads = 256;

__device__ __forceinline__ unsigned long long region_key(double x) {  // order-preserving: x < y <=> key(x) < key(y)
    const long long b = __double_as_longlong(x);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}
__device__ __forceinline__ double region_unkey(unsigned long long k) {
    const long long b = (long long)k;
    return __longlong_as_double(b < 0 ? (b ^ (long long)0x8000000000000000ull) : ~b);
}
__device__ __forceinline__ uint32_t region_deposit(uint64_t x, const uint8_t* pos) {
    uint32_t m = 0;
    while (x) {
        m |= 1u << pos[__ffsll((long long)x) - 1];
        x &= x - 1;
    }
    return m;
}
// weight of the vertex set `cfg`, -inf when it is not independent (vertices added in ascending order: the same order in
// every pass, so equality with the row optimum is exact for real weights too)
__device__ __forceinline__ double region_weight(const RegionDesc& R, uint32_t cfg) {
    double s = 0;
    for (uint32_t x = cfg; x; x &= x - 1) {
        const int v = __ffs((int)x) - 1;
        if (R.adj[v] & cfg) return -INFINITY;
        s += R.w[v];
    }
    return s;
}
// the region that owns element x of an array laid out region by region: largest r with first(r) <= x
template <typename F>
__device__ __forceinline__ int region_of(int n_regions, int64_t x, F first) {
    int lo = 0, hi = n_regions - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (first(mid) <= x) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

template <int PASS>
__global__ void __launch_bounds__(kRegionThreads)
k_region_configs(const RegionDesc* __restrict__ regs, int n_regions, unsigned long long* __restrict__ alpha_key,
                 const uint8_t* __restrict__ keep, int64_t* __restrict__ chunk_count, const int64_t* __restrict__ chunk_off,
                 uint32_t* __restrict__ out_configs) {
    __shared__ RegionDesc R;
    __shared__ uint32_t warp_tot[kRegionThreads / 32];
    __shared__ unsigned long long red[kRegionThreads / 32];
    const int64_t cta = blockIdx.x;
    {
        const int r = region_of(n_regions, cta, [&](int q) { return regs[q].cta_base; });
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(regs + r);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&R);
        for (int i = threadIdx.x; i < (int)(sizeof(RegionDesc) / 8); i += kRegionThreads) dst[i] = src[i];
    }
    __syncthreads();
    const int chunks_log2 = R.n_int - R.chunk_log2;
    const uint64_t local = (uint64_t)(cta - R.cta_base);
    const uint64_t a = local >> chunks_log2, chunk = local & (((uint64_t)1 << chunks_log2) - 1);
    const int64_t row = R.row_base + (int64_t)a;
    const uint32_t bmask = region_deposit(a, R.bpos);
    const uint64_t i0 = chunk << R.chunk_log2, len = (uint64_t)1 << R.chunk_log2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double row_best = -INFINITY;
    if (PASS > 0) {
        row_best = region_unkey(alpha_key[row]);
        if (row_best == -INFINITY || (keep && !keep[row])) {  // infeasible or dropped row: nothing to count or write
            if (PASS == 1 && threadIdx.x == 0) chunk_count[cta] = 0;
            return;
        }
    }
    if (PASS == 0 && region_weight(R, bmask) == -INFINITY) return;  // two adjacent boundary vertices chosen: the row stays -inf
    double best = -INFINITY;
    int64_t written = 0;
    uint32_t mine = 0;
    for (uint64_t j = 0; j < len; j += kRegionThreads) {  // uniform trip count: the block-wide scan below needs every thread
        const uint64_t i = j + threadIdx.x;
        bool hit = false;
        uint32_t cfg = 0;
        if (i < len) {
            cfg = bmask | region_deposit(i0 + i, R.ipos);
            const double wgt = region_weight(R, cfg);
            if (PASS == 0) best = fmax(best, wgt);
            else hit = wgt == row_best;
        }
        if (PASS == 1) mine += hit;
        if (PASS == 2) {
            const uint32_t ball = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) warp_tot[warp] = __popc(ball);
            __syncthreads();
            uint32_t before = 0, total = 0;
            for (int q = 0; q < kRegionThreads / 32; ++q) {
                before += q < warp ? warp_tot[q] : 0;
                total += warp_tot[q];
            }
            if (hit) out_configs[chunk_off[cta] + written + before + __popc(ball & ((1u << lane) - 1))] = cfg;
            written += total;
            __syncthreads();
        }
    }
    if (PASS == 0) {
        unsigned long long k = region_key(best);
        for (int o = 16; o; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
            k = other > k ? other : k;
        }
        if (lane == 0) red[warp] = k;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 1; q < kRegionThreads / 32; ++q) k = red[q] > k ? red[q] : k;
            atomicMax(&alpha_key[row], k);
        }
    }
    if (PASS == 1) {
        for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0) warp_tot[warp] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int q = 0; q < kRegionThreads / 32; ++q) t += warp_tot[q];
            chunk_count[cta] = t;
        }
    }
}

// exclusive prefix sum of the chunk counts (one CTA: a contiguous segment per thread, then a scan of the 1024 segment sums);
// off[n] = total
__global__ void __launch_bounds__(1024) k_region_scan(const int64_t* __restrict__ count, int64_t* __restrict__ off, int64_t n) {
    __shared__ int64_t seg[1024];
    const int64_t per = (n + 1023) / 1024, lo = min(n, per * threadIdx.x), hi = min(n, lo + per);
    int64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += count[i];
    seg[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int q = 0; q < 1024; ++q) {
            const int64_t t = seg[q];
            seg[q] = run;
            run += t;
        }
        off[n] = run;
    }
    __syncthreads();
    s = seg[threadIdx.x];
    for (int64_t i = lo; i < hi; ++i) {
        off[i] = s;
        s += count[i];
    }
}

// per row of the batch: its offset into the configuration array = the offset of its first chunk; row_off[n_rows] = total
__global__ void k_region_row_off(const RegionDesc* __restrict__ regs, int n_regions, const int64_t* __restrict__ chunk_off,
                                 int64_t n_cta, int64_t* __restrict__ row_off, int64_t n_rows) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_rows) return;
    if (g == n_rows) {
        row_off[g] = chunk_off[n_cta];
        return;
    }
    const int r = region_of(n_regions, g, [&](int q) { return regs[q].row_base; });
    const int chunks_log2 = regs[r].n_int - regs[r].chunk_log2;
    row_off[g] = chunk_off[regs[r].cta_base + ((g - regs[r].row_base) << chunks_log2)];
}

// mis_compactify of every region of the batch: one stage of the subset-max transform (rows whose configuration has `bit`
// set take the max with the row that lacks it; regions of rank <= bit are left alone) ...
__global__ void k_region_subset_max(const RegionDesc* __restrict__ regs, int n_regions, double* __restrict__ z, int bit, int64_t n_rows) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_rows) return;
    const int r = region_of(n_regions, g, [&](int q) { return regs[q].row_base; });
    const int64_t a = g - regs[r].row_base;
    if (bit < regs[r].rank && ((a >> bit) & 1)) z[g] = fmax(z[g], z[g - ((int64_t)1 << bit)]);
}
// ... and the keep flags (as k_table_keep, per region)
__global__ void k_region_keep(const RegionDesc* __restrict__ regs, int n_regions, const double* __restrict__ sizes,
                              const double* __restrict__ z, uint8_t* __restrict__ keep, int64_t n_rows) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_rows) return;
    const int r = region_of(n_regions, g, [&](int q) { return regs[q].row_base; });
    const int64_t a = g - regs[r].row_base;
    const double v = sizes[g];
    double best = -INFINITY;
    for (int i = 0; i < regs[r].rank; ++i)
        if ((a >> i) & 1) best = fmax(best, z[g - ((int64_t)1 << i)]);
    keep[g] = (v > -INFINITY && !(best >= v)) ? 1 : 0;
}

__global__ void k_region_sizes(const unsigned long long* __restrict__ alpha_key, double* __restrict__ sizes, double* __restrict__ z, int64_t n) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n) sizes[a] = z[a] = region_unkey(alpha_key[a]);
}
__global__ void k_region_init(unsigned long long* __restrict__ alpha_key, int64_t n) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a < n) alpha_key[a] = region_key(-INFINITY);
}

template <typename T>
__global__ void k_to_double(const T* __restrict__ src, double* __restrict__ dst, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = Ops<T>::to_double(src[i]);
}

// ------------------------------------------------------------------------------------------------
// K5: bit-permutation transpose of a 2^rank array of 4-byte elements: bit i of the destination
// address is bit perm[i] of the source address.  A CTA moves the tile spanned by the union of the low
// L source bits and the source bits feeding the low L destination bits (<= 2^10 elements) through
// shared memory: 128-bit coalesced global reads along the source order, 128-bit coalesced global
// writes along the destination order.
// ------------------------------------------------------------------------------------------------
struct PermuteDesc {
    uint8_t rank, u, ng, pad;
    uint8_t tile_src_bit[12];   // tile element bit j (source order)  -> source address bit
    uint8_t tile_dst_bit[12];   // tile element bit j (dest order)    -> destination address bit
    uint8_t dst_to_src_tile[12];// tile element bit j (dest order)    -> tile element bit (source order)
    uint8_t grid_src_bit[32];   // grid bit j -> source address bit
    uint8_t grid_dst_bit[32];   // grid bit j -> destination address bit
};

__global__ void __launch_bounds__(256) k_permute_bits(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, PermuteDesc d) {
    __shared__ uint32_t tile[4096 + 128];  // 2^u <= 4096 elements, one pad word per 32
    const int tid = threadIdx.x;
    const uint64_t g = blockIdx.x;
    uint64_t sbase = 0, dbase = 0;
    for (int j = 0; j < d.ng; ++j) {
        const uint64_t bit = (g >> j) & 1ull;
        sbase |= bit << d.grid_src_bit[j];
        dbase |= bit << d.grid_dst_bit[j];
    }
    const int n = 1 << d.u;
    if (d.u >= 2 && d.tile_src_bit[0] == 0 && d.tile_src_bit[1] == 1 && d.tile_dst_bit[0] == 0 && d.tile_dst_bit[1] == 1) {
        // 128-bit path: 4 consecutive tile elements are 4 consecutive addresses on both sides
        for (int e = tid * 4; e < n; e += 256 * 4) {
            uint64_t so = 0;
            for (int j = 2; j < d.u; ++j) so |= (uint64_t)((e >> j) & 1) << d.tile_src_bit[j];
            const uint4 v = *reinterpret_cast<const uint4*>(in + sbase + so);
            tile[(e + 0) + ((e + 0) >> 5)] = v.x;
            tile[(e + 1) + ((e + 1) >> 5)] = v.y;
            tile[(e + 2) + ((e + 2) >> 5)] = v.z;
            tile[(e + 3) + ((e + 3) >> 5)] = v.w;
        }
        __syncthreads();
        for (int f = tid * 4; f < n; f += 256 * 4) {
            uint64_t dof = 0;
            uint32_t e0 = 0;
            for (int j = 2; j < d.u; ++j) {
                const uint32_t bit = (f >> j) & 1;
                dof |= (uint64_t)bit << d.tile_dst_bit[j];
                e0 |= bit << d.dst_to_src_tile[j];
            }
            uint32_t r[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t e = e0 | ((q & 1) << d.dst_to_src_tile[0]) | (((q >> 1) & 1) << d.dst_to_src_tile[1]);
                r[q] = tile[e + (e >> 5)];
            }
            *reinterpret_cast<uint4*>(out + dbase + dof) = make_uint4(r[0], r[1], r[2], r[3]);
        }
    } else {
        for (int e = tid; e < n; e += 256) {
            uint64_t so = 0;
            for (int j = 0; j < d.u; ++j) so |= (uint64_t)((e >> j) & 1) << d.tile_src_bit[j];
            tile[e + (e >> 5)] = in[sbase + so];
        }
        __syncthreads();
        for (int f = tid; f < n; f += 256) {
            uint64_t dof = 0;
            uint32_t e = 0;
            for (int j = 0; j < d.u; ++j) {
                const uint32_t bit = (f >> j) & 1;
                dof |= (uint64_t)bit << d.tile_dst_bit[j];
                e |= bit << d.dst_to_src_tile[j];
            }
            out[dbase + dof] = tile[e + (e >> 5)];
        }
    }
}

}  // namespace tb
