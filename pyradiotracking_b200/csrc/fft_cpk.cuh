// Packed-complex in-register FFT building blocks for sm_100a.
//
// A complex fp32 value lives in ONE 64-bit register pair (lo = re, hi = im) and is operated on
// with the packed fp32x2 instructions of sm_100 (PTX add/sub/mul/fma.f32x2 -> SASS FADD2 / FMUL2 /
// FFMA2).  ptxas folds the lane swap (re <-> im) and operand negation into FADD2/FFMA2 operand
// modifiers (.LO_HI, -R), so
//      a +- b                          1 instruction
//      a -+ j b  = a + swap(b)*(+-1,-+1) 1 instruction (FFMA2 with a constant pair)
//      a * (wr + j wi) = a*wr + swap(a)*(-wi, wi)   2 instructions
// i.e. half the instruction count of scalar code with the same number of data registers.
//
// The same source compiles for the HOST (plain float pair) so tests/csrc_host_check.cpp can run the
// identical algebra on the CPU; the build container has no GPU.
#pragma once

#if defined(__CUDACC__)
#define CPK_HD __device__ __forceinline__      // under nvcc everything here is device code
#else
#define CPK_HD inline                          // g++: the CPU emulation of tests/csrc_host_check.cpp
#endif

namespace rt {

#define CPK_C1 0.92387953251128673848f   // cos(pi/8)
#define CPK_S1 0.38268343236508978178f   // sin(pi/8)
#define CPK_R2 0.70710678118654752440f   // sqrt(1/2)

#if defined(__CUDACC__)
// ---------------------------------------------------------------- device: one b64 register pair
struct cpk {
    unsigned long long v;
};
__device__ __forceinline__ unsigned long long cpk_pair(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long cpk_swapped(unsigned long long a) {
    unsigned lo, hi;
    unsigned long long r;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(hi), "r"(lo));
    return r;
}
__device__ __forceinline__ cpk c_make(float re, float im) { return cpk{cpk_pair(re, im)}; }
__device__ __forceinline__ float c_re(cpk a) { return __uint_as_float((unsigned)(a.v & 0xffffffffull)); }
__device__ __forceinline__ float c_im(cpk a) { return __uint_as_float((unsigned)(a.v >> 32)); }
__device__ __forceinline__ cpk c_add(cpk a, cpk b) { cpk d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d; }
__device__ __forceinline__ cpk c_sub(cpk a, cpk b) { cpk d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v)); return d; }
// a * s   (real scalar, both lanes)
__device__ __forceinline__ cpk c_scale(cpk a, float s) { cpk d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(cpk_pair(s, s))); return d; }
// c + swap(a) * (plo, phi)
__device__ __forceinline__ cpk c_fma_swap(cpk a, float plo, float phi, cpk c) {
    cpk d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(cpk_swapped(a.v)), "l"(cpk_pair(plo, phi)), "l"(c.v));
    return d;
}
// c + swap(a) * p   with p a ready-made register pair
__device__ __forceinline__ cpk c_fma_swap_p(cpk a, unsigned long long p, cpk c) {
    cpk d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(cpk_swapped(a.v)), "l"(p), "l"(c.v));
    return d;
}
// c + a * b   (lane-wise)
__device__ __forceinline__ cpk c_fma(cpk a, cpk b, cpk c) { cpk d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return d; }
// c + a * s  and  c - a * s   (real scalar s on both lanes)
__device__ __forceinline__ cpk c_fma_s(cpk a, float s, cpk c) { cpk d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(cpk_pair(s, s)), "l"(c.v)); return d; }
__device__ __forceinline__ cpk c_fnma_s(cpk a, float s, cpk c) { return c_fma_s(a, -s, c); }
#else
// ---------------------------------------------------------------- host: plain float pair, same semantics
struct cpk {
    float re, im;
};
CPK_HD cpk c_make(float re, float im) { return cpk{re, im}; }
CPK_HD float c_re(cpk a) { return a.re; }
CPK_HD float c_im(cpk a) { return a.im; }
CPK_HD cpk c_add(cpk a, cpk b) { return cpk{a.re + b.re, a.im + b.im}; }
CPK_HD cpk c_sub(cpk a, cpk b) { return cpk{a.re - b.re, a.im - b.im}; }
CPK_HD cpk c_scale(cpk a, float s) { return cpk{a.re * s, a.im * s}; }
CPK_HD cpk c_fma_swap(cpk a, float plo, float phi, cpk c) { return cpk{a.im * plo + c.re, a.re * phi + c.im}; }
CPK_HD cpk c_fma(cpk a, cpk b, cpk c) { return cpk{a.re * b.re + c.re, a.im * b.im + c.im}; }
CPK_HD cpk c_fma_s(cpk a, float s, cpk c) { return cpk{a.re * s + c.re, a.im * s + c.im}; }
CPK_HD cpk c_fnma_s(cpk a, float s, cpk c) { return cpk{c.re - a.re * s, c.im - a.im * s}; }
#endif

// a - j b  and  a + j b
CPK_HD cpk c_sub_j(cpk a, cpk b) { return c_fma_swap(b, 1.f, -1.f, a); }
CPK_HD cpk c_add_j(cpk a, cpk b) { return c_fma_swap(b, -1.f, 1.f, a); }
// a * (wr + j wi):  (re wr - im wi, im wr + re wi) = a*wr + swap(a)*(-wi, wi)
CPK_HD cpk c_mul(cpk a, float wr, float wi) { return c_fma_swap(a, -wi, wi, c_scale(a, wr)); }

// forward 4-point DFT (W4 = -j), in place, natural order: 8 packed instructions
CPK_HD void cdft4(cpk& x0, cpk& x1, cpk& x2, cpk& x3) {
    const cpk s02 = c_add(x0, x2), d02 = c_sub(x0, x2);
    const cpk s13 = c_add(x1, x3), d13 = c_sub(x1, x3);
    x0 = c_add(s02, s13);
    x2 = c_sub(s02, s13);
    x1 = c_sub_j(d02, d13);
    x3 = c_add_j(d02, d13);
}
// same, where x2 still has to be multiplied by -j
CPK_HD void cdft4_x2_mj(cpk& x0, cpk& x1, cpk& x2, cpk& x3) {
    const cpk s02 = c_sub_j(x0, x2), d02 = c_add_j(x0, x2);
    const cpk s13 = c_add(x1, x3), d13 = c_sub(x1, x3);
    x0 = c_add(s02, s13);
    x2 = c_sub(s02, s13);
    x1 = c_sub_j(d02, d13);
    x3 = c_add_j(d02, d13);
}

// first radix-4 layer of a 16-point DFT on un-windowed inputs x_i with real weights w_i folded in as FMAs:
// (x0 w0 +- x2 w2) costs 3 packed instructions instead of 4
CPK_HD void cdft4_win(cpk& x0, cpk& x1, cpk& x2, cpk& x3, float w0, float w1, float w2, float w3) {
    const cpk m0 = c_scale(x0, w0), m1 = c_scale(x1, w1);
    const cpk s02 = c_fma_s(x2, w2, m0), d02 = c_fnma_s(x2, w2, m0);
    const cpk s13 = c_fma_s(x3, w3, m1), d13 = c_fnma_s(x3, w3, m1);
    x0 = c_add(s02, s13);
    x2 = c_sub(s02, s13);
    x1 = c_sub_j(d02, d13);
    x3 = c_add_j(d02, d13);
}

// forward 16-point DFT, in place, natural order in and out: 80 packed instructions.
// n = 4a + b, k = c + 4d:  Y[c+4d] = sum_b W16^{bc} W4^{bd} sum_a x[4a+b] W4^{ac}.
CPK_HD void cdft16(cpk (&v)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) cdft4(v[b], v[4 + b], v[8 + b], v[12 + b]);    // t[b][c] left in v[4c+b]
    v[4 + 1] = c_mul(v[4 + 1], CPK_C1, -CPK_S1);      // W16^1
    v[4 + 2] = c_mul(v[4 + 2], CPK_R2, -CPK_R2);      // W16^2
    v[4 + 3] = c_mul(v[4 + 3], CPK_S1, -CPK_C1);      // W16^3
    v[8 + 1] = c_mul(v[8 + 1], CPK_R2, -CPK_R2);      // W16^2
    /* v[8 + 2] * W16^4 = -j: folded into cdft4_x2_mj */
    v[8 + 3] = c_mul(v[8 + 3], -CPK_R2, -CPK_R2);     // W16^6
    v[12 + 1] = c_mul(v[12 + 1], CPK_S1, -CPK_C1);    // W16^3
    v[12 + 2] = c_mul(v[12 + 2], -CPK_R2, -CPK_R2);   // W16^6
    v[12 + 3] = c_mul(v[12 + 3], -CPK_C1, CPK_S1);    // W16^9
    cdft4(v[0], v[1], v[2], v[3]);
    cdft4(v[4], v[5], v[6], v[7]);
    cdft4_x2_mj(v[8], v[9], v[10], v[11]);
    cdft4(v[12], v[13], v[14], v[15]);
    // v[4c+d] holds Y[c+4d]: transpose the 4x4 index grid (register renaming only)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) {
            const cpk t = v[4 * c + d];
            v[4 * c + d] = v[4 * d + c];
            v[4 * d + c] = t;
        }
}

// same, on un-windowed inputs: v[i] <- DFT16(w[i] * v[i])   (72 packed instructions + 0 for the window)
CPK_HD void cdft16_win(cpk (&v)[16], const float (&w)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) cdft4_win(v[b], v[4 + b], v[8 + b], v[12 + b], w[b], w[4 + b], w[8 + b], w[12 + b]);    // t[b][c] left in v[4c+b]
    v[4 + 1] = c_mul(v[4 + 1], CPK_C1, -CPK_S1);      // W16^1
    v[4 + 2] = c_mul(v[4 + 2], CPK_R2, -CPK_R2);      // W16^2
    v[4 + 3] = c_mul(v[4 + 3], CPK_S1, -CPK_C1);      // W16^3
    v[8 + 1] = c_mul(v[8 + 1], CPK_R2, -CPK_R2);      // W16^2
    /* v[8 + 2] * W16^4 = -j: folded into cdft4_x2_mj */
    v[8 + 3] = c_mul(v[8 + 3], -CPK_R2, -CPK_R2);     // W16^6
    v[12 + 1] = c_mul(v[12 + 1], CPK_S1, -CPK_C1);    // W16^3
    v[12 + 2] = c_mul(v[12 + 2], -CPK_R2, -CPK_R2);   // W16^6
    v[12 + 3] = c_mul(v[12 + 3], -CPK_C1, CPK_S1);    // W16^9
    cdft4(v[0], v[1], v[2], v[3]);
    cdft4(v[4], v[5], v[6], v[7]);
    cdft4_x2_mj(v[8], v[9], v[10], v[11]);
    cdft4(v[12], v[13], v[14], v[15]);
    // v[4c+d] holds Y[c+4d]: transpose the 4x4 index grid (register renaming only)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) {
            const cpk t = v[4 * c + d];
            v[4 * c + d] = v[4 * d + c];
            v[4 * d + c] = t;
        }
}

}  // namespace rt
