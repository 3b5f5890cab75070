// Diagnostics: VIADDMNMX.S16x2 throughput of the 8x8 register microtile as a function of warps per SM sub-partition.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dpx_occ dpx_occ.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int PREFETCH>
__global__ void __launch_bounds__(128) k_tile(unsigned* out, const unsigned* __restrict__ src, int stride) {
    __shared__ __align__(16) unsigned sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 128) sm[i] = src[i] & 0x00ff00ffu;
    __syncthreads();
    unsigned acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0xC000C000u;
    const int lane = threadIdx.x & 31;
    const unsigned* pa = sm + (lane & 7) * 4;
    const unsigned* pb = sm + 1024 + (lane >> 3) * 4;
    for (int it = 0; it < ITERS; ++it) {
        const int o = (it * stride) & 511;
        const uint4 a0 = *reinterpret_cast<const uint4*>(pa + o);
        const uint4 a1 = *reinterpret_cast<const uint4*>(pa + o + 32);
        const uint4 b0 = *reinterpret_cast<const uint4*>(pb + o);
        const uint4 b1 = *reinterpret_cast<const uint4*>(pb + o + 32);
        const unsigned a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const unsigned b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = __viaddmax_s16x2(a[i], b[j], acc[i][j]);
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s ^= acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// register-only variant (no shared-memory loads in the loop)
__global__ void __launch_bounds__(128) k_reg(unsigned* out, unsigned a0, unsigned b0) {
    unsigned acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0xC000C000u;
    unsigned a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = a0 + i * threadIdx.x; b[i] = b0 + i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = __viaddmax_s16x2(a[i], b[j], acc[i][j]);
        a[it & 7] += it;
        b[it & 7] ^= it;
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s ^= acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    unsigned *out, *src;
    cudaMalloc(&out, (size_t)sms * 16 * 128 * 4);
    cudaMalloc(&src, 2048 * 4); cudaMemset(src, 1, 2048 * 4);
    for (int B : {1, 2, 3, 4, 6, 8}) {   // blocks of 4 warps per SM = warps per sub-partition
        const double ops = (double)sms * B * 128 * ITERS * 64 * 2;
        float t0 = time_ms([&] { k_reg<<<sms * B, 128>>>(out, 1, 2); });
        float t1 = time_ms([&] { k_tile<0><<<sms * B, 128>>>(out, src, 64); });
        printf("{\"warps_per_smsp\": %d, \"reg_only_Gops\": %.0f, \"smem_fed_Gops\": %.0f}\n", B, ops / t0 * 1e-6, ops / t1 * 1e-6);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
