// Micro-benchmark 4: legacy tensor path (mma.sync.m16n8k16 f16 -> f32, SASS HMMA.16816.F32) on sm_100a:
// issue interval per scheduler, alone and next to packed fp32 (FFMA2) and SHFL, for W warps per scheduler.
// Prints SM cycles per loop body per scheduler, and the dense MAC rate per SM per clock this implies.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

__device__ __forceinline__ void hmma(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void tf32mma(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// NH independent HMMA chains + NF FFMA2 + NS SHFL per loop body; KIND 0 = f16 m16n8k16, 1 = tf32 m16n8k8
template <int KIND, int NH, int NF, int NS>
__global__ void kern(float* out, float x, float y, unsigned k) {
    float d[8][4];
    unsigned a[4], b[2];
    unsigned long long pacc[8];
    unsigned z[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = 0x3c003c00u + threadIdx.x + i * k;
    b[0] = 0x38003800u + k; b[1] = 0x34003400u + k;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) d[i][q] = i + q;
        pacc[i] = ((unsigned long long)__float_as_uint(threadIdx.x * 0.001f + i) << 32) | __float_as_uint(1.f + i);
        z[i] = threadIdx.x * 3 + i;
    }
    const unsigned long long xx = ((unsigned long long)__float_as_uint(x) << 32) | __float_as_uint(x);
    const unsigned long long yy = ((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(y);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (r < NH) { if (KIND == 0) hmma(d[r & 7], a, b); else tf32mma(d[r & 7], a, b); }
            if (r < NF) pacc[r & 7] = ffma2(pacc[r & 7], xx, yy);
            if (r < NS) z[r & 7] = __shfl_xor_sync(0xffffffffu, z[r & 7], 4);
        }
    }
    unsigned long long t = 0;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t ^= pacc[i] ^ z[i]; s += d[i][0] + d[i][1] + d[i][2] + d[i][3]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(t ^ (t >> 32))) + s;
}

static float* g_out;
static double g_clock_ghz = 1.9;

template <int KIND, int NH, int NF, int NS>
double run(int warps_per_sched) {
    const int threads = 128 * warps_per_sched, blocks = 148;   // one CTA per SM
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    kern<KIND, NH, NF, NS><<<blocks, threads>>>(g_out, 1.0001f, 0.5f, 0x7604);
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) kern<KIND, NH, NF, NS><<<blocks, threads>>>(g_out, 1.0001f, 0.5f, 0x7604);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    return ms * 1e-3 * g_clock_ghz * 1e9 / ITERS / warps_per_sched;   // cycles per body per warp-slot of a scheduler
}

int main() {
    cudaMalloc(&g_out, 148 * 1024 * sizeof(float));
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0); g_clock_ghz = khz * 1e-6;
    printf("clock %.3f GHz nominal (cycle figures assume it)\n", g_clock_ghz);
    for (int w = 1; w <= 8; w *= 2) {
        const double c8 = run<0, 8, 0, 0>(w), c16 = run<0, 16, 0, 0>(w);
        printf("f16 HMMA.16816 alone, %d warps/sched: 8/body %.2f cyc (%.2f per HMMA), 16/body %.2f cyc (%.2f per HMMA) -> %.0f MAC/clk/SM\n",
               w, c8, c8 / 8, c16, c16 / 16, 4 * 2048.0 * 16 / c16);
    }
    for (int w = 2; w <= 8; w *= 2) {
        const double c = run<1, 16, 0, 0>(w);
        printf("tf32 HMMA.1688 alone, %d warps/sched: 16/body %.2f cyc (%.2f per MMA) -> %.0f MAC/clk/SM\n", w, c, c / 16, 4 * 1024.0 * 16 / c);
    }
    for (int w = 2; w <= 4; w *= 2) {
        printf("%d warps/sched: 16 FFMA2 alone %.2f | 8 HMMA + 16 FFMA2 %.2f | 16 HMMA + 16 FFMA2 %.2f | 4 HMMA + 16 FFMA2 %.2f\n", w,
               run<0, 0, 16, 0>(w), run<0, 8, 16, 0>(w), run<0, 16, 16, 0>(w), run<0, 4, 16, 0>(w));
        printf("%d warps/sched: 16 SHFL alone %.2f | 8 SHFL + 16 FFMA2 %.2f | 16 SHFL + 16 FFMA2 %.2f | 8 HMMA + 8 SHFL + 16 FFMA2 %.2f\n", w,
               run<0, 0, 0, 16>(w), run<0, 0, 16, 8>(w), run<0, 0, 16, 16>(w), run<0, 8, 16, 8>(w));
    }
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
