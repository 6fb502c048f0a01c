// In-register complex FFT building blocks (fp32), shared by the spectrogram kernels.
// Everything here is __host__ __device__ so tests/csrc_host_check.cpp can run the
// exact same arithmetic on the CPU (the build container has no GPU).
#pragma once

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

namespace rt {

struct cf {
    float re, im;
};

RT_HD cf cadd(cf a, cf b) { return cf{a.re + b.re, a.im + b.im}; }
RT_HD cf csub(cf a, cf b) { return cf{a.re - b.re, a.im - b.im}; }
// a * (wr + j wi)
RT_HD cf cmul(cf a, float wr, float wi) { return cf{a.re * wr - a.im * wi, a.re * wi + a.im * wr}; }
// a * (-j)
RT_HD cf mul_mj(cf a) { return cf{a.im, -a.re}; }

// forward 4-point DFT (W4 = -j), in place, natural order
RT_HD void dft4(cf& x0, cf& x1, cf& x2, cf& x3) {
    cf s02 = cadd(x0, x2), d02 = csub(x0, x2);
    cf s13 = cadd(x1, x3), d13 = csub(x1, x3);
    x0 = cadd(s02, s13);
    x2 = csub(s02, s13);
    x1 = cf{d02.re + d13.im, d02.im - d13.re};   // d02 - j d13
    x3 = cf{d02.re - d13.im, d02.im + d13.re};   // d02 + j d13
}

#define RT_C1 0.92387953251128673848f   // cos(pi/8)
#define RT_S1 0.38268343236508978178f   // sin(pi/8)
#define RT_R2 0.70710678118654752440f   // sqrt(1/2)

// forward 16-point DFT, in place: v[n] (n = 0..15, natural order) -> v[k] (natural order).
// Decomposition n = 4a + b, k = c + 4d:  Y[c+4d] = sum_b W16^{bc} W4^{bd} sum_a x[4a+b] W4^{ac}.
RT_HD void dft16(cf (&v)[16]) {
    // step 1: for each b, 4-point DFT over a of x[4a+b]; result t[b][c] left in v[4c+b]
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
    // step 2: twiddle t[b][c] *= W16^{bc}   (b, c in 1..3)
    // c = 1: exponents b   -> 1, 2, 3
    v[4 + 1] = cmul(v[4 + 1], RT_C1, -RT_S1);
    v[4 + 2] = cf{RT_R2 * (v[4 + 2].re + v[4 + 2].im), RT_R2 * (v[4 + 2].im - v[4 + 2].re)};
    v[4 + 3] = cmul(v[4 + 3], RT_S1, -RT_C1);
    // c = 2: exponents 2b  -> 2, 4, 6
    v[8 + 1] = cf{RT_R2 * (v[8 + 1].re + v[8 + 1].im), RT_R2 * (v[8 + 1].im - v[8 + 1].re)};
    v[8 + 2] = mul_mj(v[8 + 2]);
    v[8 + 3] = cf{RT_R2 * (v[8 + 3].im - v[8 + 3].re), -RT_R2 * (v[8 + 3].re + v[8 + 3].im)};
    // c = 3: exponents 3b  -> 3, 6, 9
    v[12 + 1] = cmul(v[12 + 1], RT_S1, -RT_C1);
    v[12 + 2] = cf{RT_R2 * (v[12 + 2].im - v[12 + 2].re), -RT_R2 * (v[12 + 2].re + v[12 + 2].im)};
    v[12 + 3] = cmul(v[12 + 3], -RT_C1, RT_S1);
    // step 3: for each c, 4-point DFT over b of t[b][c] (held in v[4c+b]); output d lands in v[4c+d]
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4(v[4 * c + 0], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    // v[4c+d] now holds Y[c+4d]: transpose the 4x4 index grid to natural order
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = c + 1; d < 4; ++d) {
            cf t = v[4 * c + d];
            v[4 * c + d] = v[4 * d + c];
            v[4 * d + c] = t;
        }
}

}  // namespace rt
