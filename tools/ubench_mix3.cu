// Micro-benchmark 3 (ubench_mix2.cu with the conversion candidates of spectro_reg256_v8: FHADD, FHFMA, I2F, HADD2, FADD2, FFMA, FMUL): issue/pipe cost of integer/ALU instructions next to packed fp32 (FFMA2) on sm_100a.
// Every "other" op is an `asm volatile` so ptxas cannot fold the dependent chain away (the LOP3/IADD3/FMNMX
// rows of ubench_mix.cu were folded).  Prints SM cycles per loop body per scheduler for
//   NF packed FFMA2 (each = 2 FMA-pipe cycles) + NX X-ops, 16 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

enum { X_LOP3, X_IADD3, X_SHF, X_PRMT, X_ISETP_SEL, X_FSETP_SEL, X_MOV, X_IMAD, X_DP4A, X_FMNMX, X_LEA, X_STS, X_LDS64, X_LDS128, X_FADD, X_FHADD, X_FHFMA, X_I2F, X_HADD2, X_FADD2, X_FFMA, X_FMUL, X_REDUX, X_LDSU16 };

template <int X>
__device__ __forceinline__ unsigned xop(unsigned z, unsigned k, unsigned* sm) {
    unsigned r = z;
    if (X == X_LOP3) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(z), "r"(k), "r"(0x5a5a5a5au));
    if (X == X_IADD3) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(z), "r"(k));
    if (X == X_SHF) asm volatile("shf.l.wrap.b32 %0, %1, %2, 3;" : "=r"(r) : "r"(z), "r"(k));
    if (X == X_PRMT) asm volatile("prmt.b32 %0, %1, %2, 0x7604;" : "=r"(r) : "r"(z), "r"(k));
    if (X == X_ISETP_SEL) asm volatile("{.reg .pred p; setp.lt.u32 p, %1, %2; selp.u32 %0, %1, %2, p;}" : "=r"(r) : "r"(z), "r"(k));
    if (X == X_FSETP_SEL) asm volatile("{.reg .pred p; setp.lt.f32 p, %1, %2; selp.f32 %0, %1, %2, p;}" : "=f"(*(float*)&r) : "f"(__uint_as_float(z)), "f"(__uint_as_float(k)));
    if (X == X_MOV) asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(z));
    if (X == X_IMAD) asm volatile("mad.lo.u32 %0, %1, %2, %1;" : "=r"(r) : "r"(z), "r"(k));
    if (X == X_DP4A) asm volatile("dp4a.u32.u32 %0, %1, %2, %1;" : "=r"(r) : "r"(z), "r"(k));
    if (X == X_FMNMX) asm volatile("max.f32 %0, %1, %2;" : "=f"(*(float*)&r) : "f"(__uint_as_float(z)), "f"(__uint_as_float(k)));
    if (X == X_LEA) asm volatile("mad.lo.u32 %0, %1, 8, %2;" : "=r"(r) : "r"(z), "r"(k));
    if (X == X_STS) { asm volatile("st.shared.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(sm + (threadIdx.x & 1023))), "r"(z) : "memory"); r = z + 1; }
    if (X == X_LDS64) { unsigned hi; asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(r), "=r"(hi) : "r"((unsigned)__cvta_generic_to_shared(sm + 2 * (z & 255))) : "memory"); r ^= hi; }
    if (X == X_LDS128) { unsigned b, c, d; asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r), "=r"(b), "=r"(c), "=r"(d) : "r"((unsigned)__cvta_generic_to_shared(sm + 4 * (z & 127))) : "memory"); r ^= b ^ c ^ d; }
    if (X == X_FADD) asm volatile("add.f32 %0, %1, %2;" : "=f"(*(float*)&r) : "f"(__uint_as_float(z)), "f"(__uint_as_float(k)));
    if (X == X_FHADD) asm volatile("add.rn.f32.f16 %0, %1, %2;" : "=f"(*(float*)&r) : "h"((unsigned short)(z >> 16)), "f"(__uint_as_float(k)));
    if (X == X_FHFMA) asm volatile("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(*(float*)&r) : "h"((unsigned short)(z >> 16)), "h"((unsigned short)k), "f"(__uint_as_float(z)));
    if (X == X_I2F) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(*(float*)&r) : "r"(z));
    if (X == X_HADD2) asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(z), "r"(k));
    if (X == X_FADD2) { unsigned long long d, a = ((unsigned long long)z << 32) | k, b = ((unsigned long long)k << 32) | z; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); r = (unsigned)d ^ (unsigned)(d >> 32); }
    if (X == X_FFMA) asm volatile("fma.rn.f32 %0, %1, %2, %1;" : "=f"(*(float*)&r) : "f"(__uint_as_float(z)), "f"(__uint_as_float(k)));
    if (X == X_FMUL) asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(*(float*)&r) : "f"(__uint_as_float(z)), "f"(__uint_as_float(k)));
    if (X == X_REDUX) asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(z));
    if (X == X_LDSU16) { unsigned short h; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"((unsigned)__cvta_generic_to_shared(sm) + 2 * (z & 1023)) : "memory"); r = h; }
    return r;
}

template <int X, int NF, int NX>
__global__ void kern(float* out, float x, float y, unsigned k) {
    __shared__ unsigned sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 7 + k;
    __syncthreads();
    unsigned long long pacc[8];
    unsigned z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { pacc[i] = ((unsigned long long)__float_as_uint(threadIdx.x * 0.001f + i) << 32) | __float_as_uint(1.f + i); z[i] = threadIdx.x * 3 + i; }
    const unsigned long long xx = ((unsigned long long)__float_as_uint(x) << 32) | __float_as_uint(x);
    const unsigned long long yy = ((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(y);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r < NF) pacc[r] = ffma2(pacc[r], xx, yy);
#pragma unroll
            for (int q = 0; q < (NX + 7) / 8; ++q)
                if (r + 8 * q < NX) z[r] = xop<X>(z[r], k, sm);
        }
    }
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t ^= pacc[i] ^ z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(t ^ (t >> 32)));
}

static float* g_out;
static double g_clock_ghz = 1.9;

template <int X, int NF, int NX>
double run() {
    const int blocks = 148 * 4, threads = 512;      // 16 warps per scheduler
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    kern<X, NF, NX><<<blocks, threads>>>(g_out, 1.0001f, 0.5f, 0x7604);
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) kern<X, NF, NX><<<blocks, threads>>>(g_out, 1.0001f, 0.5f, 0x7604);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    return ms * 1e-3 * g_clock_ghz * 1e9 / ITERS / 16.0;
}

#define ROW(X, name) printf("%-10s  alone x8 %6.2f  x16 %6.2f | 8 FFMA2 + x4 %6.2f  +x8 %6.2f  +x16 %6.2f   (cycles per body per scheduler; 8 FFMA2 alone = 16)\n", name, \
    run<X, 0, 8>(), run<X, 0, 16>(), run<X, 8, 4>(), run<X, 8, 8>(), run<X, 8, 16>());

int main() {
    cudaMalloc(&g_out, 148 * 4 * 512 * sizeof(float));
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0); g_clock_ghz = khz * 1e-6;
    printf("clock %.3f GHz nominal; 8 FFMA2 alone: %.2f cycles\n", g_clock_ghz, run<X_MOV, 8, 0>());
    ROW(X_FADD, "FADD") ROW(X_FHADD, "FHADD") ROW(X_FHFMA, "FHFMA") ROW(X_FFMA, "FFMA") ROW(X_FMUL, "FMUL") ROW(X_FADD2, "FADD2(+2)") ROW(X_HADD2, "HADD2")
    ROW(X_I2F, "I2F.U32") ROW(X_PRMT, "PRMT") ROW(X_IADD3, "IADD3") ROW(X_DP4A, "DP4A") ROW(X_REDUX, "REDUX") ROW(X_LDSU16, "LDS.U16")
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
