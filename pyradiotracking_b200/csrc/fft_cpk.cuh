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

#define CPK_T1 0.41421356237309504880f   // tan(pi/8)

// Second radix-4 layer of the 16-point DFT: the groups b = 1, 2, 3 take their inputs times W16^{b c}
// (c = 0..3).  Every W16 power is a real scale (cos(pi/8) or sqrt(1/2)) times a factor of the form
// (1 -+ j t) or a power of j, so the scale is folded into the butterfly's own additions as FMA operands
// and each twiddle costs ONE packed instruction instead of two: 11 + 10 + 11 instead of 14 + 12 + 14.
// b = 1: inputs x (W16^0, W16^1, W16^2, W16^3)
CPK_HD void cdft4_tw1(cpk& y0, cpk& y1, cpk& y2, cpk& y3) {
    const cpk a1 = c_fma_swap(y1, CPK_T1, -CPK_T1, y1);        // y1 (1 - j t):   W16^1 y1 =  c1 a1
    const cpk a3 = c_fma_swap(y3, -CPK_T1, CPK_T1, y3);        // y3 (1 + j t):   W16^3 y3 = -j c1 a3
    const cpk a2 = c_sub_j(y2, y2);                            // y2 (1 - j):     W16^2 y2 =  r2 a2
    const cpk p = c_sub_j(a1, a3), q = c_add_j(a1, a3);        // x1 + x3 = c1 p, x1 - x3 = c1 q
    const cpk s02 = c_fma_s(a2, CPK_R2, y0), d02 = c_fnma_s(a2, CPK_R2, y0);
    y0 = c_fma_s(p, CPK_C1, s02);
    y2 = c_fnma_s(p, CPK_C1, s02);
    y1 = c_fma_swap(q, CPK_C1, -CPK_C1, d02);                  // d02 - j c1 q
    y3 = c_fma_swap(q, -CPK_C1, CPK_C1, d02);                  // d02 + j c1 q
}
// b = 2: inputs x (W16^0, W16^2, W16^4, W16^6)
CPK_HD void cdft4_tw2(cpk& y0, cpk& y1, cpk& y2, cpk& y3) {
    const cpk a1 = c_sub_j(y1, y1);                            // W16^2 y1 =  r2 a1
    const cpk a3 = c_add_j(y3, y3);                            // W16^6 y3 = -r2 a3
    const cpk p = c_sub(a1, a3), q = c_add(a1, a3);            // x1 + x3 = r2 p, x1 - x3 = r2 q
    const cpk s02 = c_sub_j(y0, y2), d02 = c_add_j(y0, y2);    // W16^4 = -j
    y0 = c_fma_s(p, CPK_R2, s02);
    y2 = c_fnma_s(p, CPK_R2, s02);
    y1 = c_fma_swap(q, CPK_R2, -CPK_R2, d02);
    y3 = c_fma_swap(q, -CPK_R2, CPK_R2, d02);
}
// b = 3: inputs x (W16^0, W16^3, W16^6, W16^9)
CPK_HD void cdft4_tw3(cpk& y0, cpk& y1, cpk& y2, cpk& y3) {
    const cpk a1 = c_fma_swap(y1, -CPK_T1, CPK_T1, y1);        // y1 (1 + j t):   W16^3 y1 = -j c1 a1
    const cpk a3 = c_fma_swap(y3, CPK_T1, -CPK_T1, y3);        // y3 (1 - j t):   W16^9 y3 = -c1 a3
    const cpk a2 = c_add_j(y2, y2);                            // y2 (1 + j):     W16^6 y2 = -r2 a2
    const cpk p = c_add_j(a3, a1), q = c_sub_j(a3, a1);        // x1 + x3 = -c1 p, x1 - x3 = c1 q
    const cpk s02 = c_fnma_s(a2, CPK_R2, y0), d02 = c_fma_s(a2, CPK_R2, y0);
    y0 = c_fnma_s(p, CPK_C1, s02);
    y2 = c_fma_s(p, CPK_C1, s02);
    y1 = c_fma_swap(q, CPK_C1, -CPK_C1, d02);
    y3 = c_fma_swap(q, -CPK_C1, CPK_C1, d02);
}

// second layer + output index transpose shared by cdft16 / cdft16_win
CPK_HD void cdft16_tail(cpk (&v)[16]) {
    cdft4(v[0], v[1], v[2], v[3]);
    cdft4_tw1(v[4], v[5], v[6], v[7]);
    cdft4_tw2(v[8], v[9], v[10], v[11]);
    cdft4_tw3(v[12], v[13], v[14], v[15]);
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

// forward 16-point DFT, in place, natural order in and out: 72 packed instructions.
// n = 4a + b, k = c + 4d:  Y[c+4d] = sum_b W16^{bc} W4^{bd} sum_a x[4a+b] W4^{ac}.
CPK_HD void cdft16(cpk (&v)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) cdft4(v[b], v[4 + b], v[8 + b], v[12 + b]);    // t[b][c] left in v[4c+b]
    cdft16_tail(v);
}

// same, on un-windowed inputs: v[i] <- DFT16(w[i] * v[i])   (80 packed instructions including the window)
CPK_HD void cdft16_win(cpk (&v)[16], const float (&w)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) cdft4_win(v[b], v[4 + b], v[8 + b], v[12 + b], w[b], w[4 + b], w[8 + b], w[12 + b]);
    cdft16_tail(v);
}

}  // namespace rt
