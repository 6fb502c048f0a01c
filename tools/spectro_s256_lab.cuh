// spectro_s256_lab.cuh -- LAB (tools/build_lab_lib.sh links it into tools/librtb200_lab.so; not part of the product library) --
// uint8 IQ -> power spectrogram cells + per-CTA row sums for nperseg N = 256 R, R = 4 or 16 (1024 / 4096;
// BASELINE.json configs[2], the wideband single stream), built around the in-register 256-point core of spectro256.cuh (sm_100a).
//
// Replaces scipy.signal.spectrogram(..., nperseg=N, noverlap=0, return_onesided=False) as called at
// /root/reference/radiotracking/analyze.py:234-241 (detrend='constant', window, FFT, |X|^2 / (fs * sum w^2)).
//
// Decimation in frequency, N = 256 R, n = n1 + 256 n2 (n2 < R), k = R k1 + k2 (k2 < R):
//     X[R k1 + k2] = sum_{n1 < 256} W_256^{n1 k1} * ( W_N^{n1 k2} * sum_{n2 < R} W_R^{n2 k2} w[n] (x[n] - mean) )
//   stage A   one R-point DFT per n1 across the R quarters / sixteenths of the segment (window folded into its first butterfly
//             layer), times W_N^{n1 k2}: thread n1 of the team (R = 16: 256 threads, one DFT16 each; R = 4: 64 threads, four DFT4
//             each), inputs straight from the TMA-staged bytes (consecutive lanes read consecutive samples), outputs y[k2][n1]
//             into the team's exchange buffer -- the ONLY team-wide exchange, one named barrier per segment;
//   stage B   half-warp k2 of the team runs the 256-point FFT of row y[k2][.] exactly like spectro_reg256_v7 (16 x 16 in
//             registers, one warp-synchronous transpose -- through the memory of its own row, which is dead by then),
//             |X|^2, row sums, four coalesced STG.128 per thread.
// Measured against the product kernel for these sizes (spectro_r16.cuh: three radix-16 Stockham passes, five team barriers per
// segment): faster stand-alone (50 vs 66 us at 1024, 63 vs 70 us at 4096 for a 20 M-sample block) but it needs the whole register
// file (2 x 256 x 128), so the scan kernels of the previous launch cannot run beside it and the STEP is slower (82 vs 69 us, 84 vs
// 85 us); capped at 120 registers with the lean scan beside it: 78 / 85 us.  See profiles/r02_s256_kernel.txt.
//
// S layout PERMR (rt_engine.cu): row t holds, for k2 = 0..R-1, the 256 bins R k1 + k2 in the PERM position order of k1
// (position 256 k2 + 64 (k1 >> 6) + 4 (k1 & 15) + ((k1 >> 4) & 3)): every half-warp store instruction is 256 contiguous bytes.
// Exchange buffers are double-buffered (stage A of segment i + 1 writes while slow half-warps still transpose segment i), the
// segment byte sums for the detrend are taken one segment ahead (no barrier of their own), raw segments arrive through a two-deep
// ring of TMA bulk copies per team.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../pyradiotracking_b200/csrc/spectro256.cuh"

namespace rt {

template <int N>
struct S256Cfg {
    static_assert(N == 1024 || N == 4096, "256-point-core kernel: nperseg 1024 or 4096");
    static constexpr int R = N / 256;
    static constexpr int THREADS = 256;
    static constexpr int BT = 16 * R;                       // threads per team: one half-warp per k2
    static constexpr int TEAMS = THREADS / BT;              // segments in flight per CTA
    static constexpr int TW = BT / 32;                      // warps per team
    static constexpr int ROW = 2304;                        // bytes per row y[k2][.] (256 complex + pad) = the 16 x 144-byte transpose tile
    static constexpr int YBUF = R * ROW;
    static constexpr int OFF_Y = 0;                                      // [TEAMS][2][YBUF]
    static constexpr int OFF_RAW = OFF_Y + TEAMS * 2 * YBUF;             // [TEAMS][2][2N] bytes
    static constexpr int OFF_WIN = OFF_RAW + TEAMS * 2 * 2 * N;          // float[N]: window (with the power scale), natural order
    static constexpr int OFF_SUM = OFF_WIN + 4 * N;                      // [TEAMS][2][TW] uint2 (sum I, sum Q) per warp
    static constexpr int OFF_BAR = OFF_SUM + TEAMS * 2 * TW * 8;         // [TEAMS][2] mbarrier
    static constexpr int SMEM = OFF_BAR + TEAMS * 16;
    static constexpr int CTAS_PER_SM = 2;
    static_assert(OFF_RAW % 16 == 0 && OFF_WIN % 16 == 0 && OFF_SUM % 16 == 0 && OFF_BAR % 8 == 0, "alignment");
    static_assert(2 * (SMEM + 1024) <= 227 * 1024, "two CTAs per SM");
};

#ifndef S256_MAXR
#define S256_MAXR 128       // 2 CTAs x 256 threads x 120 registers leave 4096 of the SM's registers: one lean scan CTA runs beside them
#endif
__device__ __forceinline__ cpk s256_lds(uint32_t addr) {
    cpk v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v.v) : "r"(addr));
    return v;
}

template <int N>
__global__ void __maxnreg__(S256_MAXR) spectro_s256_k(SpectroArgs a) {
    using C = S256Cfg<N>;
    constexpr int R = C::R, BT = C::BT;
    extern __shared__ __align__(16) unsigned char s256_smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int team = tid / BT, b = tid % BT;
    const int wt = b >> 5;                                    // warp of the team
    const int hw = b >> 4, j = b & 15;                        // stage B: half-warp = k2, lane of the 256-point core
    const uint32_t sm0 = smem_u32(s256_smem);
    const uint32_t yb0 = sm0 + C::OFF_Y + team * (2 * C::YBUF);
    const uint32_t rawb = sm0 + C::OFF_RAW + team * (4 * N);
    const uint32_t sumb = sm0 + C::OFF_SUM + team * (2 * C::TW * 8);
    const uint32_t bar0 = sm0 + C::OFF_BAR + team * 16;

    const int s = blockIdx.y;
    const uint8_t* base = a.unit_base(s);
    // segments are dealt round-robin over the (CTA, team) pairs of the unit
    const int step = a.n_chunks * C::TEAMS;
    const int first = blockIdx.x * C::TEAMS + team;
    const int n_it = first < a.T ? (a.T - first + step - 1) / step : 0;

    {
        float* win = reinterpret_cast<float*>(s256_smem + C::OFF_WIN);
        for (int e = tid; e < N; e += C::THREADS) win[e] = a.win[e];
    }
    if (b == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto team_sync = [&]() {
        if (C::TEAMS == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(BT) : "memory");
    };
    auto issue = [&](int it) {                                // one thread of the team: TMA copy of segment `it` into ring slot it & 1
        const uint32_t bar = bar0 + 8 * (it & 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx_a(bar, 2 * N);
        bulk_g2s_a(rawb + (it & 1) * 2 * N, base + (size_t)(first + it * step) * (2 * N), 2 * N, bar);
    };
    if (b == 0) {
        if (n_it > 0) issue(0);
        if (n_it > 1) issue(1);
    }

    // per-thread shared-memory bases (opaque to the compiler so that they stay in registers)
    uint32_t a_raw = rawb + 2 * b;                            // stage A sample (q, n2): + 2 (64 q + 256 n2)      (+ ring slot)
    uint32_t a_sum = rawb + 16 * b;                           // byte sums: + 16 BT                                  (+ ring slot)
    uint32_t a_win = sm0 + C::OFF_WIN + 4 * b;                // window of sample (q, n2): + 4 (64 q + 256 n2)
    uint32_t a_yst = yb0 + 8 * b;                             // stage A store (k2, q): + k2 ROW + 8 * 64 q        (+ buffer)
    uint32_t a_row = yb0 + hw * C::ROW;                       // stage B: this half-warp's row / transpose tile      (+ buffer)
    asm volatile("" : "+r"(a_raw), "+r"(a_sum), "+r"(a_win), "+r"(a_yst), "+r"(a_row));
    const uint32_t a_yld = a_row + 8 * j;                     // row element 16 a + j: + 128 a
    const uint32_t a_tst = a_row + 8 * j;                     // tile store k1: + 144 k1
    const uint32_t a_tld = a_row + 144 * j;                   // tile row j: + 16 c

    // twiddles: stage A  W_N^{n1 k2}  (R = 16: n1 = b, k2 = 1..15;  R = 4: n1 = b + 64 q, k2 = 1..3), stage B  W_256^{j k}
    constexpr int NTA = R == 16 ? 15 : 12;
    float tar[NTA], tai[NTA], tbr[16], tbi[16], acc[16];
#pragma unroll
    for (int i = 0; i < NTA; ++i) {
        const int n1 = R == 16 ? b : b + 64 * (i / 3), k2 = R == 16 ? i + 1 : (i % 3) + 1;
        const float2 t = a.tw[(n1 * k2) & (N - 1)];
        tar[i] = t.x; tai[i] = t.y;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float2 t = a.tw[((j * i) * R) & (N - 1)];
        tbr[i] = t.x; tbi[i] = t.y;
        acc[i] = 0.f;
    }

    // this thread's share of a segment's byte sums -> per-warp slot of the given parity
    auto publish_sums = [&](int slot) {
        const uint4 q0 = lds_128(a_sum + slot * (2 * N)), q1 = lds_128(a_sum + slot * (2 * N) + 16 * BT);
        unsigned sI = __dp4a(q0.x, 0x00010001u, 0u), sQ = __dp4a(q0.x, 0x01000100u, 0u);
        sI = __dp4a(q0.y, 0x00010001u, sI); sQ = __dp4a(q0.y, 0x01000100u, sQ);
        sI = __dp4a(q0.z, 0x00010001u, sI); sQ = __dp4a(q0.z, 0x01000100u, sQ);
        sI = __dp4a(q0.w, 0x00010001u, sI); sQ = __dp4a(q0.w, 0x01000100u, sQ);
        sI = __dp4a(q1.x, 0x00010001u, sI); sQ = __dp4a(q1.x, 0x01000100u, sQ);
        sI = __dp4a(q1.y, 0x00010001u, sI); sQ = __dp4a(q1.y, 0x01000100u, sQ);
        sI = __dp4a(q1.z, 0x00010001u, sI); sQ = __dp4a(q1.z, 0x01000100u, sQ);
        sI = __dp4a(q1.w, 0x00010001u, sI); sQ = __dp4a(q1.w, 0x01000100u, sQ);
        const unsigned tI = __reduce_add_sync(0xffffffffu, sI), tQ = __reduce_add_sync(0xffffffffu, sQ);
        if (lane == 0) asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sumb + slot * (C::TW * 8) + 8 * wt), "r"(tI), "r"(tQ) : "memory");
    };
    if (n_it > 0) {
        mbar_wait(bar0, 0);
        publish_sums(0);
        team_sync();
    }

    for (int it = 0; it < n_it; ++it) {
        const int seg = first + it * step;
        const int st = it & 1;
        const uint32_t rb = a_raw + st * (2 * N);
        // ---- detrend constant of this segment (its byte sums were published one iteration ago, before that iteration's barrier)
        unsigned tI = 0, tQ = 0;
#pragma unroll
        for (int w2 = 0; w2 < C::TW; w2 += 2) {
            const uint4 q = lds_128(sumb + st * (C::TW * 8) + 8 * w2);
            tI += q.x + q.z;
            tQ += q.y + q.w;
        }
        // mean = sum / N is exact in fp32 (sum < 2^24, N a power of two) and so is (float)byte - mean
        const cpk mean = c_make((float)tI * (1.f / N), (float)tQ * (1.f / N));
        const cpk magic = c_make(8388608.f, 8388608.f);

        // ---- stage A: R-point DFTs across the segment's R parts, window folded, times W_N^{n1 k2}
        cpk v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            // R = 16: i = n2, sample b + 256 n2;  R = 4: i = 4 q + n2, sample b + 64 q + 256 n2
            const int off = R == 16 ? 256 * i : 64 * (i >> 2) + 256 * (i & 3);
            const unsigned u = lds_u16(rb + 2 * off);
            // 0x4B0000bb = 2^23 + byte (no I2F); minus 2^23 and minus the mean are both exact
            const cpk f = c_make(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7540)), __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7541)));
            v[i] = c_sub(c_sub(f, magic), mean);
        }
        // the next segment's byte sums (its copy was issued a whole iteration ago)
        if (it + 1 < n_it) {
            mbar_wait(bar0 + 8 * (st ^ 1), ((it + 1) >> 1) & 1);
            publish_sums(st ^ 1);
        }
        const uint32_t ys = a_yst + st * C::YBUF;
        if (R == 16) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {                     // cdft16_win with the window read four values at a time (register pressure)
                const float w0 = lds_f32(a_win + 4 * 256 * g), w1 = lds_f32(a_win + 4 * 256 * (4 + g));
                const float w2 = lds_f32(a_win + 4 * 256 * (8 + g)), w3 = lds_f32(a_win + 4 * 256 * (12 + g));
                cdft4_win(v[g], v[4 + g], v[8 + g], v[12 + g], w0, w1, w2, w3);
            }
            cdft16_tail(v);
            sts_64(ys, v[0].v);
#pragma unroll
            for (int k2 = 1; k2 < 16; ++k2) sts_64(ys + k2 * C::ROW, c_mul(v[k2], tar[k2 - 1], tai[k2 - 1]).v);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float w4[4];
#pragma unroll
                for (int n2 = 0; n2 < 4; ++n2) w4[n2] = lds_f32(a_win + 4 * (64 * q + 256 * n2));
                cdft4_win(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3], w4[0], w4[1], w4[2], w4[3]);
                sts_64(ys + 8 * 64 * q, v[4 * q].v);
#pragma unroll
                for (int k2 = 1; k2 < 4; ++k2)
                    sts_64(ys + k2 * C::ROW + 8 * 64 * q, c_mul(v[4 * q + k2], tar[3 * q + k2 - 1], tai[3 * q + k2 - 1]).v);
            }
        }
        team_sync();
        // every thread of the team has read its bytes of ring slot st (stage A above, byte sums one iteration ago): refill it
        if (b == 0 && it + 2 < n_it) issue(it + 2);

        // ---- stage B: 256-point FFT of row y[k2 = hw][.] (spectro_reg256_v7's two passes), bins R k1 + hw
        const uint32_t yl = a_yld + st * C::YBUF, ts = a_tst + st * C::YBUF, tl = a_tld + st * C::YBUF;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = s256_lds(yl + 128 * i);
        __syncwarp();                                        // the transpose tile below overlays this row
        cdft16(v);                                           // over a (n1 = 16 a + j) -> k1a
        sts_64(ts, v[0].v);
#pragma unroll
        for (int k = 1; k < 16; ++k) sts_64(ts + 144 * k, c_mul(v[k], tbr[k], tbi[k]).v);
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) lds_2x64(tl + 16 * c, v[2 * c].v, v[2 * c + 1].v);
        cdft16(v);                                           // k1 = j + 16 k2'
        const bool valid = seg < a.T;
        float4* dst = reinterpret_cast<float4*>(a.S + (size_t)s * a.S_stream_stride + (size_t)seg * N + 256 * hw + 4 * j);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float p[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const float re = c_re(v[4 * c + d]), im = c_im(v[4 * c + d]);
                p[d] = fmaf(im, im, re * re);
                acc[4 * c + d] += p[d];
            }
            if (valid) dst[16 * c] = make_float4(p[0], p[1], p[2], p[3]);
        }
    }

    // ---- row sums of this CTA: teams added in fixed order, written in FFT bin order (fi = R (j + 16 k2') + hw)
    __syncthreads();
    float* red = reinterpret_cast<float*>(s256_smem);        // [TEAMS][N] floats (the exchange buffers are dead)
    static_assert(C::TEAMS * N * 4 <= C::OFF_RAW, "reduction scratch fits the exchange buffers");
#pragma unroll
    for (int i = 0; i < 16; ++i) red[team * N + R * (j + 16 * i) + hw] = acc[i];
    __syncthreads();
    float* pd = a.part + ((size_t)s * a.n_chunks + blockIdx.x) * N;
    for (int fi = tid; fi < N; fi += C::THREADS) {
        float t = 0.f;
#pragma unroll
        for (int tm = 0; tm < C::TEAMS; ++tm) t += red[tm * N + fi];
        pd[fi] = t;
    }
}

}  // namespace rt
