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
 *   leaf tensors                   generate_tensors of IndependentSet [upstream GenericTensorNetworks]
 *   per node                       OMEinsum binary rule [upstream]: classify labels, permutedims both
 *                                  operands to matrix form, batched tropical GEMM (TropicalGEMM.jl's
 *                                  role: C = max_k A + B), result left in [n | m | batch] order.
 * Values are Tropical{Float32} (the reference's default element_type, src/dynamic_ob.jl:6).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int rank;
    int labels[40]; /* bit 0 (fastest) first */
    float* data;
} tens;

static int find_label(const tens* t, int l) {
    for (int i = 0; i < t->rank; ++i)
        if (t->labels[i] == l) return i;
    return -1;
}

/* dst has labels `order` (n_order of them, all present in src); plain permutedims */
static float* permute(const tens* src, const int* order, int n_order) {
    size_t n = (size_t)1 << n_order;
    float* out = (float*)malloc(n * sizeof(float));
    int sh[40];
    for (int i = 0; i < n_order; ++i) sh[i] = find_label(src, order[i]);
    int identity = (n_order == src->rank);
    for (int i = 0; i < n_order && identity; ++i) identity = (sh[i] == i);
    if (identity) {
        memcpy(out, src->data, n * sizeof(float));
        return out;
    }
    /* byte-wise lookup tables of the bit scatter */
    size_t lut[5][256];
    int nbytes = (n_order + 7) / 8;
    for (int by = 0; by < nbytes; ++by)
        for (int v = 0; v < 256; ++v) {
            size_t s = 0;
            for (int i = 0; i < 8 && by * 8 + i < n_order; ++i) s |= ((size_t)((v >> i) & 1)) << sh[by * 8 + i];
            lut[by][v] = s;
        }
    const float* sd = src->data;
    /* threads only when the caller is not already inside the branch-parallel loop (nested regions are serialised) */
#pragma omp parallel for schedule(static) if (n >= ((size_t)1 << 20))
    for (size_t hi = 0; hi < n; hi += 256) {
        size_t sb = 0;
        for (int by = 1; by < nbytes; ++by) sb |= lut[by][(hi >> (8 * by)) & 255];
        size_t lim = n - hi < 256 ? n - hi : 256;
        for (size_t lo = 0; lo < lim; ++lo) out[hi + lo] = sd[sb | lut[0][lo]];
    }
    return out;
}

/* unary max over one label */
static void reduce_label(tens* t, int pos) {
    size_t n = (size_t)1 << (t->rank - 1);
    float* out = (float*)malloc(n * sizeof(float));
    size_t lowmask = ((size_t)1 << pos) - 1;
    for (size_t d = 0; d < n; ++d) {
        size_t s0 = (d & lowmask) | ((d & ~lowmask) << 1);
        float a = t->data[s0], b = t->data[s0 | ((size_t)1 << pos)];
        out[d] = a > b ? a : b;
    }
    free(t->data);
    t->data = out;
    for (int i = pos; i < t->rank - 1; ++i) t->labels[i] = t->labels[i + 1];
    t->rank--;
}

/* C[n, m, b] = max_k A[k, m, b] + B[n, k, b]   (n fastest everywhere).
 * TropicalGEMM.jl's role [upstream]: a register-tiled SIMD max-plus micro-kernel (it uses LoopVectorization; here
 * AVX2 / AVX-512 intrinsics, 4 rows x 2 vectors of accumulators, k innermost) so that the CPU baseline is a fair
 * one; the plain loop below covers the shapes too small for a tile.  Row blocks are spread over OpenMP threads when
 * the caller is not already inside a parallel region (single heavy branches). */
#if defined(__AVX512F__)
#include <immintrin.h>
typedef __m512 vf;
#define VL 16
#define vf_set1(x) _mm512_set1_ps(x)
#define vf_loadu(p) _mm512_loadu_ps(p)
#define vf_storeu(p, v) _mm512_storeu_ps(p, v)
#define vf_add(a, b) _mm512_add_ps(a, b)
#define vf_max(a, b) _mm512_max_ps(a, b)
#define TREF_SIMD "avx512"
#elif defined(__AVX2__)
#include <immintrin.h>
typedef __m256 vf;
#define VL 8
#define vf_set1(x) _mm256_set1_ps(x)
#define vf_loadu(p) _mm256_loadu_ps(p)
#define vf_storeu(p, v) _mm256_storeu_ps(p, v)
#define vf_add(a, b) _mm256_add_ps(a, b)
#define vf_max(a, b) _mm256_max_ps(a, b)
#define TREF_SIMD "avx2"
#else
#define TREF_SIMD "scalar"
#endif

const char* tref_simd(void) { return TREF_SIMD; }

static void tropical_gemm_plain(const float* Ab, const float* Bb, float* Cb, size_t M, size_t N, size_t K) {
    for (size_t m = 0; m < M; ++m) {
        float* c = Cb + m * N;
        for (size_t n = 0; n < N; ++n) c[n] = -INFINITY;
        for (size_t k = 0; k < K; ++k) {
            const float a = Ab[m * K + k];
            const float* __restrict brow = Bb + k * N;
#pragma omp simd
            for (size_t n = 0; n < N; ++n) {
                float v = a + brow[n];
                c[n] = v > c[n] ? v : c[n];
            }
        }
    }
}

#ifdef VL
/* one 4 x (2 VL) tile of C: rows m0..m0+3, columns n0..n0+2VL-1 */
static inline void tile_4x2(const float* Ab, const float* Bb, float* Cb, size_t m0, size_t n0, size_t N, size_t K) {
    const vf ninf = vf_set1(-INFINITY);
    vf c00 = ninf, c01 = ninf, c10 = ninf, c11 = ninf, c20 = ninf, c21 = ninf, c30 = ninf, c31 = ninf;
    const float *a0 = Ab + (m0 + 0) * K, *a1 = Ab + (m0 + 1) * K, *a2 = Ab + (m0 + 2) * K, *a3 = Ab + (m0 + 3) * K;
    const float* b = Bb + n0;
    for (size_t k = 0; k < K; ++k, b += N) {
        const vf b0 = vf_loadu(b), b1 = vf_loadu(b + VL);
        vf a = vf_set1(a0[k]);
        c00 = vf_max(c00, vf_add(a, b0)); c01 = vf_max(c01, vf_add(a, b1));
        a = vf_set1(a1[k]);
        c10 = vf_max(c10, vf_add(a, b0)); c11 = vf_max(c11, vf_add(a, b1));
        a = vf_set1(a2[k]);
        c20 = vf_max(c20, vf_add(a, b0)); c21 = vf_max(c21, vf_add(a, b1));
        a = vf_set1(a3[k]);
        c30 = vf_max(c30, vf_add(a, b0)); c31 = vf_max(c31, vf_add(a, b1));
    }
    float* c = Cb + m0 * N + n0;
    vf_storeu(c, c00); vf_storeu(c + VL, c01);
    vf_storeu(c + N, c10); vf_storeu(c + N + VL, c11);
    vf_storeu(c + 2 * N, c20); vf_storeu(c + 2 * N + VL, c21);
    vf_storeu(c + 3 * N, c30); vf_storeu(c + 3 * N + VL, c31);
}
#endif

static void tropical_gemm(const float* A, const float* B, float* C, int lm, int ln, int lk, int lb) {
    const size_t M = (size_t)1 << lm, N = (size_t)1 << ln, K = (size_t)1 << lk, Bn = (size_t)1 << lb;
#ifdef VL
    if (M >= 4 && N >= 2 * VL) {
        /* work items = (batch, block of 4 rows); a B column panel (K x 2VL) is reused by consecutive row blocks */
        const size_t mblocks = M / 4, items = Bn * mblocks;
        const int par = items >= 64 && (double)M * (double)N * (double)K * (double)Bn >= 1e7;
#pragma omp parallel for schedule(static) if (par)
        for (size_t it = 0; it < items; ++it) {
            const size_t b = it / mblocks, m0 = (it % mblocks) * 4;
            const float* Ab = A + b * M * K;
            const float* Bb = B + b * N * K;
            float* Cb = C + b * M * N;
            for (size_t n0 = 0; n0 < N; n0 += 2 * VL) tile_4x2(Ab, Bb, Cb, m0, n0, N, K);
        }
        return;
    }
#endif
    for (size_t b = 0; b < Bn; ++b) tropical_gemm_plain(A + b * M * K, B + b * N * K, C + b * M * N, M, N, K);
}

static int in_list(const int* v, int n, int x) {
    for (int i = 0; i < n; ++i)
        if (v[i] == x) return 1;
    return 0;
}

/* returns 0 on success */
/* fixed (may be NULL): per label -1 = free, 0 / 1 = index slicing, the label is held at that value: every leaf
 * carrying it is restricted to that index and the label leaves the network. */
int tref_contract_fixed(int n_labels, int n_leaves, const int* leaf_off, const int* leaf_labels, const int* left,
                        const int* right, const double* weights, const signed char* fixed, double* out_value,
                        double* out_ops) {
    int n_nodes = n_leaves - 1, n_t = n_leaves + (n_nodes > 0 ? n_nodes : 0);
    tens* T = (tens*)calloc((size_t)n_t, sizeof(tens));
    int* total = (int*)calloc((size_t)(n_labels > 0 ? n_labels : 1), sizeof(int));
    /* per-tensor count of leaves containing each label is tracked sparsely: cnt[t][i] for labels[i] */
    int(*cnt)[40] = (int(*)[40])calloc((size_t)n_t, sizeof(int[40]));
    double ops = 0;
    for (int i = 0; i < n_leaves; ++i) {
        int r = leaf_off[i + 1] - leaf_off[i];
        T[i].rank = r;
        T[i].data = (float*)malloc(sizeof(float) * ((size_t)1 << r));
        for (int q = 0; q < r; ++q) {
            T[i].labels[q] = leaf_labels[leaf_off[i] + q];
            cnt[i][q] = 1;
            total[T[i].labels[q]]++;
        }
        if (r == 1) {
            T[i].data[0] = 0.0f;
            T[i].data[1] = weights ? (float)weights[T[i].labels[0]] : 1.0f;
        } else if (r == 2) {
            T[i].data[0] = T[i].data[1] = T[i].data[2] = 0.0f;
            T[i].data[3] = -INFINITY;
        } else {
            free(T); free(total); free(cnt);
            return -3;
        }
        if (fixed) {
            for (int q = r - 1; q >= 0; --q) {
                int l = T[i].labels[q];
                if (fixed[l] < 0) continue;
                /* keep index fixed[l] of position q */
                int rk = T[i].rank;
                size_t lo = (size_t)1 << q, n_out = (size_t)1 << (rk - 1);
                for (size_t o = 0; o < n_out; ++o) {
                    size_t src = (o & (lo - 1)) | ((o & ~(lo - 1)) << 1) | ((size_t)fixed[l] << q);
                    T[i].data[o] = T[i].data[src];
                }
                total[l]--;
                for (int z = q; z + 1 < rk; ++z) { T[i].labels[z] = T[i].labels[z + 1]; cnt[i][z] = cnt[i][z + 1]; }
                T[i].rank = rk - 1;
            }
        }
    }
    for (int j = 0; j < n_nodes; ++j) {
        tens* A = &T[left[j]];
        tens* B = &T[right[j]];
        int* ca = cnt[left[j]];
        int* cb = cnt[right[j]];
        tens* Cn = &T[n_leaves + j];
        int* cc = cnt[n_leaves + j];
        /* classify */
        int M[40], N[40], Bt[40], K[40], nm = 0, nn = 0, nb = 0, nk = 0;
        int cM[40], cN[40], cB[40];
        int ua = A->rank, ub = B->rank;
        {
            int union_n = 0;
            for (int i = 0; i < A->rank; ++i) union_n++;
            for (int i = 0; i < B->rank; ++i)
                if (find_label(A, B->labels[i]) < 0) union_n++;
            ops += ldexp(1.0, union_n);
        }
        /* labels private to one operand that close here: unary max first */
        for (int i = A->rank - 1; i >= 0; --i) {
            int l = A->labels[i];
            if (find_label(B, l) < 0 && ca[i] >= total[l]) {
                reduce_label(A, i);
                for (int q = i; q < A->rank; ++q) ca[q] = ca[q + 1];
            }
        }
        for (int i = B->rank - 1; i >= 0; --i) {
            int l = B->labels[i];
            if (find_label(A, l) < 0 && cb[i] >= total[l]) {
                reduce_label(B, i);
                for (int q = i; q < B->rank; ++q) cb[q] = cb[q + 1];
            }
        }
        (void)ua; (void)ub;
        for (int i = 0; i < A->rank; ++i) {
            int l = A->labels[i];
            int pb = find_label(B, l);
            if (pb < 0) { cM[nm] = ca[i]; M[nm++] = l; }
            else if (ca[i] + cb[pb] >= total[l]) K[nk++] = l;
            else { cB[nb] = ca[i] + cb[pb]; Bt[nb++] = l; }
        }
        for (int i = 0; i < B->rank; ++i) {
            int l = B->labels[i];
            if (find_label(A, l) < 0) { cN[nn] = cb[i]; N[nn++] = l; }
        }
        if (nn < nm) { /* tropical GEMM is symmetric: make the vectorised inner dimension the larger one */
            tens* tt = A; A = B; B = tt;
            int tmp[40];
            memcpy(tmp, M, sizeof tmp); memcpy(M, N, sizeof tmp); memcpy(N, tmp, sizeof tmp);
            memcpy(tmp, cM, sizeof tmp); memcpy(cM, cN, sizeof tmp); memcpy(cN, tmp, sizeof tmp);
            int t2 = nm; nm = nn; nn = t2;
        }
        /* matrix forms: A -> [k | m | b], B -> [n | k | b] */
        int ordA[40], ordB[40], na = 0, nbb = 0;
        for (int i = 0; i < nk; ++i) ordA[na++] = K[i];
        for (int i = 0; i < nm; ++i) ordA[na++] = M[i];
        for (int i = 0; i < nb; ++i) ordA[na++] = Bt[i];
        for (int i = 0; i < nn; ++i) ordB[nbb++] = N[i];
        for (int i = 0; i < nk; ++i) ordB[nbb++] = K[i];
        for (int i = 0; i < nb; ++i) ordB[nbb++] = Bt[i];
        float* Am = permute(A, ordA, na);
        float* Bm = permute(B, ordB, nbb);
        Cn->rank = nn + nm + nb;
        Cn->data = (float*)malloc(sizeof(float) * ((size_t)1 << Cn->rank));
        int q = 0;
        for (int i = 0; i < nn; ++i) { cc[q] = cN[i]; Cn->labels[q++] = N[i]; }
        for (int i = 0; i < nm; ++i) { cc[q] = cM[i]; Cn->labels[q++] = M[i]; }
        for (int i = 0; i < nb; ++i) { cc[q] = cB[i]; Cn->labels[q++] = Bt[i]; }
        tropical_gemm(Am, Bm, Cn->data, nm, nn, nk, nb);
        free(Am);
        free(Bm);
        free(A->data); A->data = NULL;
        free(B->data); B->data = NULL;
        (void)in_list;
    }
    /* root: drop whatever labels remain (single-leaf networks) */
    tens* Rt = &T[n_t - 1];
    float best = -INFINITY;
    for (size_t i = 0; i < ((size_t)1 << Rt->rank); ++i) best = Rt->data[i] > best ? Rt->data[i] : best;
    *out_value = (double)best;
    if (out_ops) *out_ops = ops;
    free(Rt->data);
    free(T);
    free(total);
    free(cnt);
    return 0;
}

int tref_contract(int n_labels, int n_leaves, const int* leaf_off, const int* leaf_labels, const int* left,
                  const int* right, const double* weights, double* out_value, double* out_ops) {
    return tref_contract_fixed(n_labels, n_leaves, leaf_off, leaf_labels, left, right, weights, NULL, out_value, out_ops);
}

/* the loop of contract_slices, branches distributed over OpenMP threads (the most favourable CPU
 * arrangement: every core runs its own branch end to end).  Network i is described by the i-th
 * entries of the pointer arrays.  Returns the number of threads used. */
void tref_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* index slices of ONE network: slice i holds the labels with fixed[i * n_labels + l] >= 0 at that value */
int tref_contract_slices_of(int n, int n_labels, int n_leaves, const int* leaf_off, const int* leaf_labels,
                            const int* left, const int* right, const double* weights, const signed char* fixed,
                            double* out_values, double* out_ops) {
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 1) if (n >= nthreads)
    for (int i = 0; i < n; ++i) {
        double ops = 0;
        tref_contract_fixed(n_labels, n_leaves, leaf_off, leaf_labels, left, right, weights,
                            fixed + (size_t)i * (size_t)n_labels, &out_values[i], &ops);
        if (out_ops) out_ops[i] = ops;
    }
    return nthreads;
}

int tref_contract_batch(int n, const int* n_labels, const int* n_leaves, const int* const* leaf_off,
                        const int* const* leaf_labels, const int* const* left, const int* const* right,
                        const double* const* weights, double* out_values, double* out_ops) {
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    /* many branches: one branch per thread (the most favourable CPU arrangement); fewer branches than threads: the
     * branches run one after the other and the threads share each GEMM / permute instead */
#pragma omp parallel for schedule(dynamic, 1) if (n >= nthreads)
    for (int i = 0; i < n; ++i) {
        double ops = 0;
        if (n_leaves[i] == 0) {
            out_values[i] = 0;
        } else {
            tref_contract(n_labels[i], n_leaves[i], leaf_off[i], leaf_labels[i], left[i], right[i],
                          weights ? weights[i] : NULL, &out_values[i], &ops);
        }
        if (out_ops) out_ops[i] = ops;
    }
    return nthreads;
}
