// Micro-benchmark: does packed fp32x2 (FFMA2/FADD2/FMUL2, sm_100+) raise FP32 throughput or only
// save issue slots?  Prints Gop/s (thread-level scalar flops/2 for fma) per variant.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

__device__ __forceinline__ unsigned long long pack(float a, float b) {
    return ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a);
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__global__ void k_ffma(float* out, float x, float y) {
    float acc[2 * NACC];
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 2 * NACC; ++i) acc[i] = fmaf(acc[i], x, y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd(float* out, float x, float y) {
    float acc[2 * NACC];
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 2 * NACC; ++i) acc[i] = acc[i] + y;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float x, float y) {
    unsigned long long acc[NACC];
    const unsigned long long xx = pack(x, x), yy = pack(y, y);
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = pack(threadIdx.x * 0.001f + i, i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = ffma2(acc[i], xx, yy);
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}
__global__ void k_fadd2(float* out, float x, float y) {
    unsigned long long acc[NACC];
    const unsigned long long yy = pack(y, y);
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = pack(threadIdx.x * 0.001f + i, i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fadd2(acc[i], yy);
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}
// packed FMA with integer (alu-pipe) work interleaved 1:1 in instruction count
__global__ void k_ffma2_mix(float* out, float x, float y, unsigned sel) {
    unsigned long long acc[NACC];
    unsigned z[NACC];
    const unsigned long long xx = pack(x, x), yy = pack(y, y);
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = pack(threadIdx.x * 0.001f + i, i); z[i] = threadIdx.x + i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) { acc[i] = ffma2(acc[i], xx, yy); z[i] = __byte_perm(z[i], 0x47000000u, sel); }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s ^= acc[i] ^ z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s ^ (s >> 32)));
}
__global__ void k_ffma_mix(float* out, float x, float y, unsigned sel) {
    float acc[2 * NACC];
    unsigned z[NACC];
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) acc[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < NACC; ++i) z[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) { acc[2 * i] = fmaf(acc[2 * i], x, y); acc[2 * i + 1] = fmaf(acc[2 * i + 1], x, y); z[i] = __byte_perm(z[i], 0x47000000u, sel); }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    const int blocks = 148 * 8, threads = 256;
    float* out; cudaMalloc(&out, blocks * threads * sizeof(float));
    const double lanes = (double)blocks * threads * ITERS * 2 * NACC;   // scalar fp32 operations per launch
    float ms;
    ms = timeit([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });      printf("FFMA       %8.3f ms  %7.1f Gfma/s\n", ms, lanes / ms / 1e6);
    ms = timeit([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });     printf("FFMA2      %8.3f ms  %7.1f Gfma/s\n", ms, lanes / ms / 1e6);
    ms = timeit([&] { k_fadd<<<blocks, threads>>>(out, 1.0001f, 0.5f); });      printf("FADD       %8.3f ms  %7.1f Gadd/s\n", ms, lanes / ms / 1e6);
    ms = timeit([&] { k_fadd2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });     printf("FADD2      %8.3f ms  %7.1f Gadd/s\n", ms, lanes / ms / 1e6);
    ms = timeit([&] { k_ffma_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f, 0x7604); });  printf("FFMA+PRMT  %8.3f ms  %7.1f Gfma/s (2 FFMA : 1 PRMT)\n", ms, lanes / ms / 1e6);
    ms = timeit([&] { k_ffma2_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f, 0x7604); }); printf("FFMA2+PRMT %8.3f ms  %7.1f Gfma/s (1 FFMA2 : 1 PRMT)\n", ms, lanes / ms / 1e6);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
