/*
 * tropical_ref.c -- CPU ORACLE, plain C + OpenMP (test infrastructure / CPU baseline, NOT product code).
 *
 * PARITY STATUS: vector-level parity unpinned (the reference has no golden vectors for this path and
 * cannot run here); pinned by property against exact MIS solvers through tests/test_oracle.py, and
 * cross-checked against the numpy restatement oracle/tropical_oracle.py.
 *
 * Restates, for one branch, what the reference executes on the CPU:
 *   solve_slice                    /root/reference/src/dynamic_ob.jl:30-34
 *   contract_slices (the loop)     /root/reference/src/dynamic_ob.jl:36-48   (tref_contract_batch)
 *   leaf tensors, binary rule      tropical_ref_impl.inc (value-type generic, included twice below)
 *   TropicalGEMM.jl's role         the register-tiled SIMD max-plus micro-kernels in this file
 *
 * Value types:
 *   f32  Tropical{Float32}, the reference's default element_type (src/dynamic_ob.jl:6)
 *   i16  int16 with the sentinel -2^14 for -inf: exact for integer weights with sum |w| < 8192 (every unit-weight
 *        config of BASELINE.json); twice the SIMD lanes of f32 -- the best this port can do for those workloads,
 *        so that the GPU / CPU ratio of bench.py is not taken against a soft baseline.
 * ISA: the micro-kernels exist as AVX2 and AVX-512(BW) functions (GCC target attributes) and are picked at run
 * time with __builtin_cpu_supports, so the library built in one container uses the best ISA of the box it runs on
 * (tref_simd() reports which).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * max-plus GEMM:  C[n, m, b] = max_k A[k, m, b] + B[n, k, b]      (first index fastest everywhere)
 * Register tile: 4 rows (m) x 2 vectors (n), k innermost and blocked so that a B panel stays in L2.
 * Work items = (batch, block of 4 rows); spread over OpenMP threads when the caller is not already
 * inside the branch-parallel loop (single heavy branches).
 * ---------------------------------------------------------------------------------------------- */
#define KBLOCK 2048

#define DEF_GEMM(NAME, TARGET, T, VT, VL, SET1, LOADU, STOREU, ADD, MAX, NEGINF)                                        \
    TARGET static inline void NAME##_tile2(const T* Ab, const T* Bb, T* Cb, size_t m0, size_t n0, size_t N, size_t K,    \
                                           size_t k0, size_t k1) {                                                       \
        T* c = Cb + m0 * N + n0;                                                                                         \
        VT c00, c01, c10, c11, c20, c21, c30, c31;                                                                       \
        if (k0 == 0) {                                                                                                   \
            c00 = c01 = c10 = c11 = c20 = c21 = c30 = c31 = SET1(NEGINF);                                                \
        } else {                                                                                                         \
            c00 = LOADU(c); c01 = LOADU(c + VL); c10 = LOADU(c + N); c11 = LOADU(c + N + VL);                            \
            c20 = LOADU(c + 2 * N); c21 = LOADU(c + 2 * N + VL); c30 = LOADU(c + 3 * N); c31 = LOADU(c + 3 * N + VL);    \
        }                                                                                                                \
        const T *a0 = Ab + (m0 + 0) * K, *a1 = Ab + (m0 + 1) * K, *a2 = Ab + (m0 + 2) * K, *a3 = Ab + (m0 + 3) * K;      \
        const T* b = Bb + n0 + k0 * N;                                                                                   \
        for (size_t k = k0; k < k1; ++k, b += N) {                                                                       \
            const VT b0 = LOADU(b), b1 = LOADU(b + VL);                                                                  \
            VT a = SET1(a0[k]);                                                                                          \
            c00 = MAX(c00, ADD(a, b0)); c01 = MAX(c01, ADD(a, b1));                                                      \
            a = SET1(a1[k]);                                                                                             \
            c10 = MAX(c10, ADD(a, b0)); c11 = MAX(c11, ADD(a, b1));                                                      \
            a = SET1(a2[k]);                                                                                             \
            c20 = MAX(c20, ADD(a, b0)); c21 = MAX(c21, ADD(a, b1));                                                      \
            a = SET1(a3[k]);                                                                                             \
            c30 = MAX(c30, ADD(a, b0)); c31 = MAX(c31, ADD(a, b1));                                                      \
        }                                                                                                                \
        STOREU(c, c00); STOREU(c + VL, c01); STOREU(c + N, c10); STOREU(c + N + VL, c11);                                \
        STOREU(c + 2 * N, c20); STOREU(c + 2 * N + VL, c21); STOREU(c + 3 * N, c30); STOREU(c + 3 * N + VL, c31);        \
    }                                                                                                                    \
    TARGET static inline void NAME##_tile1(const T* Ab, const T* Bb, T* Cb, size_t m0, size_t n0, size_t N, size_t K,    \
                                           size_t k0, size_t k1) {                                                       \
        T* c = Cb + m0 * N + n0;                                                                                         \
        VT c00, c10, c20, c30;                                                                                           \
        if (k0 == 0) {                                                                                                   \
            c00 = c10 = c20 = c30 = SET1(NEGINF);                                                                        \
        } else {                                                                                                         \
            c00 = LOADU(c); c10 = LOADU(c + N); c20 = LOADU(c + 2 * N); c30 = LOADU(c + 3 * N);                          \
        }                                                                                                                \
        const T *a0 = Ab + (m0 + 0) * K, *a1 = Ab + (m0 + 1) * K, *a2 = Ab + (m0 + 2) * K, *a3 = Ab + (m0 + 3) * K;      \
        const T* b = Bb + n0 + k0 * N;                                                                                   \
        for (size_t k = k0; k < k1; ++k, b += N) {                                                                       \
            const VT b0 = LOADU(b);                                                                                      \
            c00 = MAX(c00, ADD(SET1(a0[k]), b0));                                                                        \
            c10 = MAX(c10, ADD(SET1(a1[k]), b0));                                                                        \
            c20 = MAX(c20, ADD(SET1(a2[k]), b0));                                                                        \
            c30 = MAX(c30, ADD(SET1(a3[k]), b0));                                                                        \
        }                                                                                                                \
        STOREU(c, c00); STOREU(c + N, c10); STOREU(c + 2 * N, c20); STOREU(c + 3 * N, c30);                              \
    }                                                                                                                    \
    TARGET static void NAME(const T* A, const T* B, T* C, size_t M, size_t N, size_t K, size_t Bn) {                     \
        const size_t mblocks = M / 4, items = Bn * mblocks;                                                              \
        const int par = items >= 64 && (double)M * (double)N * (double)K * (double)Bn >= 1e7;                            \
        _Pragma("omp parallel for schedule(static) if (par)")                                                            \
        for (size_t it = 0; it < items; ++it) {                                                                          \
            const size_t b = it / mblocks, m0 = (it % mblocks) * 4;                                                      \
            const T* Ab = A + b * M * K;                                                                                 \
            const T* Bb = B + b * N * K;                                                                                 \
            T* Cb = C + b * M * N;                                                                                       \
            for (size_t k0 = 0; k0 < K; k0 += KBLOCK) {                                                                  \
                const size_t k1 = k0 + KBLOCK < K ? k0 + KBLOCK : K;                                                     \
                size_t n0 = 0;                                                                                           \
                for (; n0 + 2 * VL <= N; n0 += 2 * VL) NAME##_tile2(Ab, Bb, Cb, m0, n0, N, K, k0, k1);                   \
                for (; n0 + VL <= N; n0 += VL) NAME##_tile1(Ab, Bb, Cb, m0, n0, N, K, k0, k1);                           \
            }                                                                                                            \
        }                                                                                                                \
    }

#define TGT_AVX2 __attribute__((target("avx2,fma")))
#define TGT_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx2,fma")))

DEF_GEMM(gemm_f32_avx2, TGT_AVX2, float, __m256, 8, _mm256_set1_ps, _mm256_loadu_ps, _mm256_storeu_ps, _mm256_add_ps,
         _mm256_max_ps, -INFINITY)
DEF_GEMM(gemm_f32_avx512, TGT_AVX512, float, __m512, 16, _mm512_set1_ps, _mm512_loadu_ps, _mm512_storeu_ps, _mm512_add_ps,
         _mm512_max_ps, -INFINITY)
#define LD256I(p) _mm256_loadu_si256((const __m256i*)(p))
#define ST256I(p, v) _mm256_storeu_si256((__m256i*)(p), v)
#define LD512I(p) _mm512_loadu_si512((const void*)(p))
#define ST512I(p, v) _mm512_storeu_si512((void*)(p), v)
DEF_GEMM(gemm_i16_avx2, TGT_AVX2, int16_t, __m256i, 16, _mm256_set1_epi16, LD256I, ST256I, _mm256_add_epi16,
         _mm256_max_epi16, (int16_t)-16384)
DEF_GEMM(gemm_i16_avx512, TGT_AVX512, int16_t, __m512i, 32, _mm512_set1_epi16, LD512I, ST512I, _mm512_add_epi16,
         _mm512_max_epi16, (int16_t)-16384)

/* shapes too small for a register tile */
#define DEF_PLAIN(NAME, T, NEGINF)                                                                              \
    static void NAME(const T* Ab, const T* Bb, T* Cb, size_t M, size_t N, size_t K) {                           \
        for (size_t m = 0; m < M; ++m) {                                                                        \
            T* c = Cb + m * N;                                                                                  \
            for (size_t n = 0; n < N; ++n) c[n] = NEGINF;                                                       \
            for (size_t k = 0; k < K; ++k) {                                                                    \
                const T a = Ab[m * K + k];                                                                      \
                const T* __restrict brow = Bb + k * N;                                                          \
                _Pragma("omp simd")                                                                             \
                for (size_t n = 0; n < N; ++n) {                                                                \
                    T v = (T)(a + brow[n]);                                                                     \
                    c[n] = v > c[n] ? v : c[n];                                                                 \
                }                                                                                               \
            }                                                                                                   \
        }                                                                                                       \
    }
DEF_PLAIN(gemm_f32_plain, float, -INFINITY)
DEF_PLAIN(gemm_i16_plain, int16_t, (int16_t)-16384)

static int g_isa = -1; /* 0 scalar, 1 avx2, 2 avx512 */
static int g_isa_cap = 2; /* TREF_ISA=avx2 / scalar caps the dispatch (A/B runs) */
static int isa(void) {
    if (g_isa < 0) {
        __builtin_cpu_init();
        int v = 0;
        if (__builtin_cpu_supports("avx2")) v = 1;
        if (v == 1 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")) v = 2;
        const char* e = getenv("TREF_ISA");
        if (e && !strcmp(e, "avx2")) g_isa_cap = 1;
        if (e && !strcmp(e, "scalar")) g_isa_cap = 0;
        g_isa = v < g_isa_cap ? v : g_isa_cap;
    }
    return g_isa;
}
const char* tref_simd(void) { return isa() == 2 ? "avx512" : isa() == 1 ? "avx2" : "scalar"; }

static void tropical_gemm_f32(const float* A, const float* B, float* C, int lm, int ln, int lk, int lb) {
    const size_t M = (size_t)1 << lm, N = (size_t)1 << ln, K = (size_t)1 << lk, Bn = (size_t)1 << lb;
    if (M >= 4 && isa() == 2 && N >= 16) { gemm_f32_avx512(A, B, C, M, N, K, Bn); return; }
    if (M >= 4 && isa() >= 1 && N >= 8) { gemm_f32_avx2(A, B, C, M, N, K, Bn); return; }
    for (size_t b = 0; b < Bn; ++b) gemm_f32_plain(A + b * M * K, B + b * N * K, C + b * M * N, M, N, K);
}
static void tropical_gemm_i16(const int16_t* A, const int16_t* B, int16_t* C, int lm, int ln, int lk, int lb) {
    const size_t M = (size_t)1 << lm, N = (size_t)1 << ln, K = (size_t)1 << lk, Bn = (size_t)1 << lb;
    if (M >= 4 && isa() == 2 && N >= 32) { gemm_i16_avx512(A, B, C, M, N, K, Bn); return; }
    if (M >= 4 && isa() >= 1 && N >= 16) { gemm_i16_avx2(A, B, C, M, N, K, Bn); return; }
    for (size_t b = 0; b < Bn; ++b) gemm_i16_plain(A + b * M * K, B + b * N * K, C + b * M * N, M, N, K);
}

/* ------------------------------------------------------------------------------------------------ f32 instance */
#define T float
#define FN(name) name##_f32
#define T_NEG_INF (-INFINITY)
#define T_FROM_DOUBLE(x) ((float)(x))
#define T_TO_DOUBLE(x) ((double)(x))
#define T_GEMM tropical_gemm_f32
#include "tropical_ref_impl.inc"
#undef T
#undef FN
#undef T_NEG_INF
#undef T_FROM_DOUBLE
#undef T_TO_DOUBLE
#undef T_GEMM
/* ------------------------------------------------------------------------------------------------ int16 instance */
#define T int16_t
#define FN(name) name##_i16
#define T_NEG_INF ((int16_t)-16384)
#define T_FROM_DOUBLE(x) ((int16_t)(x))
#define T_TO_DOUBLE(x) ((x) <= -8192 ? -INFINITY : (double)(x))
#define T_GEMM tropical_gemm_i16
#include "tropical_ref_impl.inc"
#undef T
#undef FN
#undef T_NEG_INF
#undef T_FROM_DOUBLE
#undef T_TO_DOUBLE
#undef T_GEMM

/* value_type: 0 = f32, 1 = i16 (caller guarantees integer weights with sum |w| < 8192), 2 = auto (i16 when legal) */
static int i16_legal(int n_labels, int n_leaves, const int* leaf_off, const int* leaf_labels, const double* weights) {
    double s = 0;
    for (int i = 0; i < n_leaves; ++i) {
        if (leaf_off[i + 1] - leaf_off[i] != 1) continue;
        double w = weights ? weights[leaf_labels[leaf_off[i]]] : 1.0;
        if (w != floor(w)) return 0;
        s += fabs(w);
    }
    (void)n_labels;
    return s < 8192.0;
}

int tref_contract_fixed_vt(int n_labels, int n_leaves, const int* leaf_off, const int* leaf_labels, const int* left,
                           const int* right, const double* weights, const signed char* fixed, int value_type,
                           double* out_value, double* out_ops) {
    int use16 = value_type == 1 || (value_type == 2 && i16_legal(n_labels, n_leaves, leaf_off, leaf_labels, weights));
    if (use16) return contract_fixed_i16(n_labels, n_leaves, leaf_off, leaf_labels, left, right, weights, fixed, out_value, out_ops);
    return contract_fixed_f32(n_labels, n_leaves, leaf_off, leaf_labels, left, right, weights, fixed, out_value, out_ops);
}

int tref_contract_fixed(int n_labels, int n_leaves, const int* leaf_off, const int* leaf_labels, const int* left,
                        const int* right, const double* weights, const signed char* fixed, double* out_value,
                        double* out_ops) {
    return tref_contract_fixed_vt(n_labels, n_leaves, leaf_off, leaf_labels, left, right, weights, fixed, 0, out_value, out_ops);
}

int tref_contract(int n_labels, int n_leaves, const int* leaf_off, const int* leaf_labels, const int* left,
                  const int* right, const double* weights, double* out_value, double* out_ops) {
    return tref_contract_fixed_vt(n_labels, n_leaves, leaf_off, leaf_labels, left, right, weights, NULL, 0, out_value, out_ops);
}

void tref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* index slices of ONE network: slice i holds the labels with fixed[i * n_labels + l] >= 0 at that value */
int tref_contract_slices_of_vt(int n, int n_labels, int n_leaves, const int* leaf_off, const int* leaf_labels,
                               const int* left, const int* right, const double* weights, const signed char* fixed,
                               int value_type, double* out_values, double* out_ops) {
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 1) if (n >= nthreads)
    for (int i = 0; i < n; ++i) {
        double ops = 0;
        tref_contract_fixed_vt(n_labels, n_leaves, leaf_off, leaf_labels, left, right, weights,
                               fixed + (size_t)i * (size_t)n_labels, value_type, &out_values[i], &ops);
        if (out_ops) out_ops[i] = ops;
    }
    return nthreads;
}
int tref_contract_slices_of(int n, int n_labels, int n_leaves, const int* leaf_off, const int* leaf_labels,
                            const int* left, const int* right, const double* weights, const signed char* fixed,
                            double* out_values, double* out_ops) {
    return tref_contract_slices_of_vt(n, n_labels, n_leaves, leaf_off, leaf_labels, left, right, weights, fixed, 0, out_values, out_ops);
}

/* the loop of contract_slices, branches distributed over OpenMP threads (the most favourable CPU
 * arrangement: every core runs its own branch end to end).  Network i is described by the i-th
 * entries of the pointer arrays.  Returns the number of threads used.
 * many branches: one branch per thread; fewer branches than threads: the branches run one after the other and the
 * threads share each GEMM / permute instead */
int tref_contract_batch_vt(int n, const int* n_labels, const int* n_leaves, const int* const* leaf_off,
                           const int* const* leaf_labels, const int* const* left, const int* const* right,
                           const double* const* weights, int value_type, double* out_values, double* out_ops) {
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 1) if (n >= nthreads)
    for (int i = 0; i < n; ++i) {
        double ops = 0;
        if (n_leaves[i] == 0) {
            out_values[i] = 0;
        } else {
            tref_contract_fixed_vt(n_labels[i], n_leaves[i], leaf_off[i], leaf_labels[i], left[i], right[i],
                                   weights ? weights[i] : NULL, NULL, value_type, &out_values[i], &ops);
        }
        if (out_ops) out_ops[i] = ops;
    }
    return nthreads;
}
int tref_contract_batch(int n, const int* n_labels, const int* n_leaves, const int* const* leaf_off,
                        const int* const* leaf_labels, const int* const* left, const int* const* right,
                        const double* const* weights, double* out_values, double* out_ops) {
    return tref_contract_batch_vt(n, n_labels, n_leaves, leaf_off, leaf_labels, left, right, weights, 0, out_values, out_ops);
}
