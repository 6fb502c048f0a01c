// Micro-benchmark 3: how close to the FMA-pipe rate does the register FFT arithmetic itself run?
// Loops the packed-complex radix-16 butterflies of csrc/fft_cpk.cuh on register data only (no memory,
// no exchange) with W warps per scheduler, and prints cycles per body per scheduler next to the
// ideal (2 cycles per packed instruction).
#include <cstdio>
#include <cuda_runtime.h>

#include "../pyradiotracking_b200/csrc/fft_cpk.cuh"
using namespace rt;

#define ITERS 1024

// MODE 0: cdft16_win only (80 packed / body)
// MODE 1: cdft16_win + 15 twiddle multiplies + cdft16 (190 packed / body)  -- the arithmetic of one segment-thread
// MODE 2: MODE 1 on two independent register sets (380 packed / body)
template <int MODE>
__global__ void kern(float* out, const float* tw, float scale) {
    cpk v[16], u[16];
    float w[16], twr[16];
    unsigned long long twp[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        v[i] = c_make(threadIdx.x * 0.001f + i, 1.f - i * 0.01f);
        u[i] = c_make(threadIdx.x * 0.002f - i, 0.5f + i * 0.01f);
        w[i] = scale + tw[i] * 1e-9f;
        twr[i] = tw[(threadIdx.x * i) & 255];
        twp[i] = cpk_pair(-tw[256 + ((threadIdx.x * i) & 255)], tw[256 + ((threadIdx.x * i) & 255)]);
    }
    for (int it = 0; it < ITERS; ++it) {
        cdft16_win(v, w);
        if (MODE >= 2) cdft16_win(u, w);
        if (MODE >= 1) {
#pragma unroll
            for (int k = 1; k < 16; ++k) v[k] = c_fma_swap_p(v[k], twp[k], c_scale(v[k], twr[k]));
            if (MODE >= 2) {
#pragma unroll
                for (int k = 1; k < 16; ++k) u[k] = c_fma_swap_p(u[k], twp[k], c_scale(u[k], twr[k]));
            }
            cdft16_win(v, w);
            if (MODE >= 2) cdft16_win(u, w);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c_re(v[i]) + c_im(v[i]) + c_re(u[i]) + c_im(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float* g_out;
static float* g_tw;
static double g_clock_ghz = 1.9;

template <int MODE>
void run(const char* name, int packed, int warps_per_sched) {
    const int threads = 128, blocks = 148 * warps_per_sched;     // 128 threads = 1 warp per scheduler per CTA
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    kern<MODE><<<blocks, threads>>>(g_out, g_tw, 0.0625f);
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) kern<MODE><<<blocks, threads>>>(g_out, g_tw, 0.0625f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    const double cyc = ms * 1e-3 * g_clock_ghz * 1e9 / ITERS / warps_per_sched;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern<MODE>);
    printf("%-28s warps/sched %2d regs %3d  %8.1f cycles per body per scheduler   ideal %4d   efficiency %.2f\n", name, warps_per_sched, fa.numRegs,
           cyc, 2 * packed, 2.0 * packed / cyc);
}

int main() {
    cudaMalloc(&g_out, 148 * 16 * 128 * sizeof(float));
    float h[512];
    for (int k = 0; k < 256; ++k) { h[k] = (float)cos(-2 * M_PI * k / 256.0); h[256 + k] = (float)sin(-2 * M_PI * k / 256.0); }
    cudaMalloc(&g_tw, sizeof(h));
    cudaMemcpy(g_tw, h, sizeof(h), cudaMemcpyHostToDevice);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0); g_clock_ghz = khz * 1e-6;
    printf("clock %.3f GHz nominal\n", g_clock_ghz);
    for (int w : {1, 2, 3, 4, 6, 8}) run<0>("dft16 (80 packed)", 80, w);
    for (int w : {1, 2, 3, 4, 6, 8}) run<1>("dft16+tw+dft16 (190)", 190, w);
    for (int w : {1, 2, 3, 4}) run<2>("2 x (dft16+tw+dft16) (380)", 380, w);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
