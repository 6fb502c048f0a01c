// Micro-benchmark: issue cost of non-FP instructions when mixed with FP32 work on sm_100a.
// For each "other" op X and each FP flavour (scalar FFMA / packed FFMA2) it times a loop body of
// NF fp ops + NX X-ops (independent chains) and prints SM cycles per loop body per SMSP-warp slot.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITERS 2048

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

enum { X_NONE, X_PRMT, X_LOP3, X_IADD, X_IMAD, X_DP4A, X_LDS, X_SHFL, X_HADD2F32, X_I2F, X_FMNMX, X_LDS128 };

template <int X>
__device__ __forceinline__ unsigned xop(unsigned z, unsigned k, const unsigned* sm) {
    if (X == X_PRMT) return __byte_perm(z, 0x47000000u, k);
    if (X == X_LOP3) return (z & k) ^ 0x5a5a5a5au;
    if (X == X_IADD) return z + k;
    if (X == X_IMAD) return z * k + 12345u;
    if (X == X_DP4A) return __dp4a(z, 0x00010001u, k);
    if (X == X_LDS) return sm[(z & 1023)];
    if (X == X_SHFL) return __shfl_xor_sync(0xffffffffu, z, 1);
    if (X == X_HADD2F32) { __half2 hh = *reinterpret_cast<__half2*>(&z); float f = __half2float(hh.x) ; return __float_as_uint(f) | 1u; }
    if (X == X_I2F) return __float_as_uint((float)(z & 0xffff));
    if (X == X_FMNMX) return __float_as_uint(fmaxf(__uint_as_float(z), __uint_as_float(k)));
    return z;
}

template <int X, int NF, int NX, bool PACKED>
__global__ void kern(float* out, float x, float y, unsigned k) {
    __shared__ unsigned sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 7 + k;
    __syncthreads();
    float acc[16];
    unsigned long long pacc[8];
    unsigned z[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { pacc[i] = ((unsigned long long)__float_as_uint(acc[2 * i]) << 32) | __float_as_uint(acc[2 * i + 1]); z[i] = threadIdx.x * 3 + i; }
    const unsigned long long xx = ((unsigned long long)__float_as_uint(x) << 32) | __float_as_uint(x);
    const unsigned long long yy = ((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(y);
    for (int it = 0; it < ITERS; ++it) {
        // NF scalar flop-instructions (or NF/2 packed ones), NX other ops, interleaved
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (PACKED) {
                if (r < NF / 2) pacc[r] = ffma2(pacc[r], xx, yy);
            } else {
                if (2 * r < NF) acc[2 * r] = fmaf(acc[2 * r], x, y);
                if (2 * r + 1 < NF) acc[2 * r + 1] = fmaf(acc[2 * r + 1], x, y);
            }
            if (r < NX) z[r] = xop<X>(z[r], k, sm);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t ^= pacc[i] ^ z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float((unsigned)(t ^ (t >> 32)));
}

static float* g_out;
static double g_clock_ghz = 1.9;

template <int X, int NF, int NX, bool PACKED>
void run(const char* name) {
    const int blocks = 148 * 4, threads = 512;      // 16 warps per SMSP
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    kern<X, NF, NX, PACKED><<<blocks, threads>>>(g_out, 1.0001f, 0.5f, 0x7604);
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) kern<X, NF, NX, PACKED><<<blocks, threads>>>(g_out, 1.0001f, 0.5f, 0x7604);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    // cycles per loop body per warp, per SMSP: 16 warps share one scheduler
    const double cyc = ms * 1e-3 * g_clock_ghz * 1e9 / ITERS / 16.0;
    printf("%-10s NF=%2d NX=%d %-6s  %7.3f ms  %6.2f cyc/body\n", name, NF, NX, PACKED ? "packed" : "scalar", ms, cyc);
}

#define ROW(X, name) \
    run<X, 16, 0, false>(name); run<X, 16, 4, false>(name); run<X, 16, 8, false>(name); \
    run<X, 16, 4, true>(name); run<X, 16, 8, true>(name); run<X, 0, 8, false>(name);

int main() {
    cudaMalloc(&g_out, 148 * 4 * 512 * sizeof(float));
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0); g_clock_ghz = khz * 1e-6;
    printf("clock %.3f GHz (nominal max; cycles assume the GPU runs at it)\n", g_clock_ghz);
    ROW(X_PRMT, "PRMT") ROW(X_LOP3, "LOP3") ROW(X_IADD, "IADD3") ROW(X_IMAD, "IMAD") ROW(X_DP4A, "DP4A")
    ROW(X_LDS, "LDS") ROW(X_SHFL, "SHFL") ROW(X_HADD2F32, "HADD2.F32") ROW(X_I2F, "I2F") ROW(X_FMNMX, "FMNMX")
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
