// dpx_peak.cu -- measures the chip's register-resident tropical-op issue rates (the missing roofline
// denominator): VIADDMNMX (s32), VIADDMNMX.S16x2, FADD+FMNMX (f32), IADD+IMNMX,
// and the shared-memory-fed 8x8 microtile inner loop.  Prints one JSON object.
//   tropical op = one a+b followed by one max.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                        \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

constexpr int ACC = 32;      // independent accumulators per thread
constexpr int ITERS = 4096;  // each iteration: ACC ops

__global__ void __launch_bounds__(256) k_s32(int* out, int a0, int b0) {
    int acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = threadIdx.x + i;
    int a = a0 + threadIdx.x, b = b0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) acc[i] = __viaddmax_s32(acc[i], a, b + i);
        a ^= it;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the form the GEMM uses: acc = max(a + b, acc)
__global__ void __launch_bounds__(256) k_s32_gemmform(int* out, int a0, int b0) {
    int acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = -(1 << 30);
    int a[4] = {a0 + (int)threadIdx.x, a0 + 1, a0 + 2, a0 + 3};
    int b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = b0 + j * (int)threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i * 8 + j] = __viaddmax_s32(a[i], b[j], acc[i * 8 + j]);
        a[it & 3] += it;
        b[it & 7] ^= it;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_s16x2(unsigned* out, unsigned a0, unsigned b0) {
    unsigned acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = 0;
    unsigned a[4] = {a0 + threadIdx.x, a0 + 1, a0 + 2, a0 + 3};
    unsigned b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = b0 + j * threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i * 8 + j] = __viaddmax_s16x2(a[i], b[j], acc[i * 8 + j]);
        a[it & 3] += it;
        b[it & 7] ^= it;
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_f32(float* out, float a0, float b0) {
    float acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = -1e30f;
    float a[4] = {a0 + threadIdx.x, a0 + 1, a0 + 2, a0 + 3};
    float b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = b0 + j * threadIdx.x;
    // every operand changes every iteration through COMPILE-TIME indices: with a run-time index into a register array
    // (the first version) each update costs a predicated copy per element, and those ISETP / FADD took issue slots from
    // the measured instructions (it reported 6.9 Top/s, less than k_gemm2<float> achieves); with operands that stay
    // unchanged the compiler drops the repeated max of an unchanged sum.  12 extra FADD per 32 ops go to the FMA pipe,
    // the FMNMX of the measured pairs to the ALU pipe that bounds them.
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i * 8 + j] = fmaxf(__fadd_rn(a[i], b[j]), acc[i * 8 + j]);
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] += 1.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] += 0.5f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// plain IADD + IMNMX (what a non-DPX part would issue)
__global__ void __launch_bounds__(256) k_s32_2op(int* out, int a0, int b0) {
    int acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = -(1 << 30);
    int a[4] = {a0 + (int)threadIdx.x, a0 + 1, a0 + 2, a0 + 3};
    int b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = b0 + j * (int)threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int s;
                asm volatile("add.s32 %0, %1, %2;" : "=r"(s) : "r"(a[i]), "r"(b[j]));
                asm volatile("max.s32 %0, %1, %2;" : "=r"(acc[i * 8 + j]) : "r"(s), "r"(acc[i * 8 + j]));
            }
        a[it & 3] += it;
        b[it & 7] ^= it;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8x8 microtile fed from shared memory exactly like k_gemm's inner loop (128x128 tile, KC = 16)
__global__ void __launch_bounds__(256, 2) k_smem_tile(int* out, int seed) {
    __shared__ __align__(16) int sA[16 * 128];
    __shared__ __align__(16) int sB[16 * 128];
    for (int i = threadIdx.x; i < 16 * 128; i += 256) {
        sA[i] = (i * 7 + seed) & 1023;
        sB[i] = (i * 13 + seed) & 1023;
    }
    __syncthreads();
    int acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = -(1 << 30);
    const int tmh = threadIdx.x & 15, tnh = threadIdx.x >> 4;
    for (int it = 0; it < ITERS / 32; ++it) {
#pragma unroll 2
        for (int kk = 0; kk < 16; ++kk) {
            const int4 a0 = *reinterpret_cast<const int4*>(sA + kk * 128 + tmh * 4);
            const int4 a1 = *reinterpret_cast<const int4*>(sA + kk * 128 + 64 + tmh * 4);
            const int4 b0 = *reinterpret_cast<const int4*>(sB + kk * 128 + tnh * 4);
            const int4 b1 = *reinterpret_cast<const int4*>(sB + kk * 128 + 64 + tnh * 4);
            const int a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const int b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __viaddmax_s32(a[i], b[j], acc[i][j]);
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s ^= acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> static double time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    const int blocks = sms * 8;
    void* out;
    CK(cudaMalloc(&out, (size_t)blocks * 256 * 4));
    const double ops = (double)blocks * 256 * ITERS * ACC;
    double t;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_max\": %d", p.name, sms, p.clockRate / 1000);
    t = time_ms([&] { k_s32<<<blocks, 256>>>((int*)out, 1, 2); });
    printf(", \"viaddmax_s32_chain_Gops\": %.1f", ops / t * 1e-6);
    t = time_ms([&] { k_s32_gemmform<<<blocks, 256>>>((int*)out, 1, 2); });
    printf(", \"viaddmax_s32_Gops\": %.1f", ops / t * 1e-6);
    t = time_ms([&] { k_s16x2<<<blocks, 256>>>((unsigned*)out, 1, 2); });
    printf(", \"viaddmax_s16x2_Ginstr\": %.1f, \"viaddmax_s16x2_Gops\": %.1f", ops / t * 1e-6, 2 * ops / t * 1e-6);
    t = time_ms([&] { k_f32<<<blocks, 256>>>((float*)out, 1.f, 2.f); });
    printf(", \"fadd_fmnmx_f32_Gops\": %.1f", ops / t * 1e-6);
    t = time_ms([&] { k_s32_2op<<<blocks, 256>>>((int*)out, 1, 2); });
    printf(", \"iadd_imnmx_s32_Gops\": %.1f", ops / t * 1e-6);
    {
        const int b2 = sms * 2 * 4;
        const double ops2 = (double)b2 * 256 * (ITERS / 32) * 16 * 64;
        t = time_ms([&] { k_smem_tile<<<b2, 256>>>((int*)out, 3); });
        printf(", \"smem_fed_8x8_s32_Gops\": %.1f", ops2 / t * 1e-6);
    }
    printf("}\n");
    return 0;
}
