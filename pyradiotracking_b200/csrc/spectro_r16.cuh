// spectro_r16.cuh -- uint8 IQ -> power spectrogram cells + per-chunk row sums for nperseg 1024 and 4096 (sm_100a):
// BASELINE.json configs[2], the wideband single stream.
//
// Replaces scipy.signal.spectrogram(..., nperseg=N, noverlap=0, return_onesided=False) as called at
// /root/reference/radiotracking/analyze.py:234-241 (detrend='constant', window, FFT, |X|^2 / (fs * sum w^2)).
//
// Self-sorting (Stockham, decimation in frequency) FFT with radix-16 passes: a team of N/16 threads owns one segment,
// every thread runs one packed-complex DFT16 (fft_cpk.cuh, 72 FFMA2-class instructions) per pass and the passes
// exchange through ONE padded shared-memory buffer (read 16 values -> barrier -> write 16 values):
//     N = 4096 = 16 * 16 * 16   three radix-16 passes, team = 256 threads = the CTA
//     N = 1024 = 16 * 16 * 4    two radix-16 passes + one radix-4 pass, team = 64 threads, four segments per CTA
// The first pass takes its inputs straight from the TMA-staged raw bytes (no initial shared-memory write) and the
// last pass leaves thread b with the bins b + (N/R) i in registers (no final write): |X|^2, row sums and the
// coalesced stores to S[stream][t][bin] (LINEAR layout) follow directly.
// Raw segments arrive through a two-deep ring of TMA bulk copies per team (cp.async.bulk + mbarrier).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "spectro256.cuh"

namespace rt {

template <int N>
struct R16Cfg {
    static_assert(N == 1024 || N == 4096, "radix-16 Stockham kernel: nperseg 1024 or 4096");
    static constexpr int THREADS = 256;
    static constexpr int BT = N / 16;                       // threads per team = radix-16 butterflies per pass
    static constexpr int TEAMS = THREADS / BT;              // segments in flight per CTA
    static constexpr int LAST_R = N == 4096 ? 16 : 4;       // radix of the last pass
    static constexpr int PADN = N + N / 16;                 // float2 entries of one exchange buffer (index + index/16)
    static constexpr int OFF_TW = 0;                        // float2[N] twiddles exp(-2 pi i k / N)
    static constexpr int OFF_X = OFF_TW + N * 8;            // [TEAMS][PADN] float2
    static constexpr int OFF_RAW = OFF_X + TEAMS * PADN * 8;   // [TEAMS][2][2N] bytes
    static constexpr int OFF_SUM = OFF_RAW + TEAMS * 2 * 2 * N;    // [TEAMS][2] int
    static constexpr int OFF_BAR = OFF_SUM + TEAMS * 8;     // [TEAMS][2] mbarrier
    static constexpr int SMEM = OFF_BAR + TEAMS * 16;
};

__device__ __forceinline__ int r16_pad(int idx) { return idx + (idx >> 4); }

template <int N>
__global__ void __maxnreg__(104) spectro_r16_k(SpectroArgs a) {   // 2 CTAs/SM and room for the scan kernels of the previous launch
    using C = R16Cfg<N>;
    constexpr int BT = C::BT, M1 = N / 16;
    extern __shared__ __align__(16) unsigned char r16_smem[];
    float2* tw = reinterpret_cast<float2*>(r16_smem + C::OFF_TW);
    const int tid = threadIdx.x, lane = tid & 31;
    const int team = tid / BT, b = tid % BT;                  // b = butterfly index of this thread in every radix-16 pass
    float2* X = reinterpret_cast<float2*>(r16_smem + C::OFF_X) + team * C::PADN;
    unsigned char* raw = r16_smem + C::OFF_RAW + team * (4 * N);
    int* sums = reinterpret_cast<int*>(r16_smem + C::OFF_SUM) + 2 * team;
    const uint32_t bar0 = smem_u32(r16_smem + C::OFF_BAR) + team * 16;

    const int s = blockIdx.y;
    const uint8_t* base = a.unit_base(s);
    // segments are dealt round-robin over the (CTA, team) pairs of the stream: one twiddle-table load per CTA,
    // balanced work for any T; this team's segments: first, first + step, ...
    const int step = a.n_chunks * C::TEAMS;
    const int first = blockIdx.x * C::TEAMS + team;
    const int seg1 = a.T;
    const int n_it = first < seg1 ? (seg1 - first + step - 1) / step : 0;

    for (int i = tid; i < N; i += C::THREADS) tw[i] = a.tw[i];
    if (b == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sums[0] = 0; sums[1] = 0;
    }
    __syncthreads();
    auto team_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(BT) : "memory"); };
    auto issue = [&](int it) {                                // one thread of the team: TMA copy of segment `it` into ring slot it & 1
        const uint32_t bar = bar0 + 8 * (it & 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx_a(bar, 2 * N);
        bulk_g2s_a(smem_u32(raw + (it & 1) * 2 * N), base + (size_t)(first + it * step) * (2 * N), 2 * N, bar);
    };
    if (b == 0 && n_it > 0) issue(0);

    // window of this thread's 16 first-pass samples (index b + M1 i), constant over the segments
    float wj[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) wj[i] = a.win[b + M1 * i];
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;

    for (int it = 0; it < n_it; ++it) {
        const int seg = first + it * step;
        if (b == 0 && it + 1 < n_it) issue(it + 1);            // slot (it+1)&1 was last read two iterations ago
        mbar_wait(bar0 + 8 * (it & 1), (it >> 1) & 1);
        const unsigned char* rb = raw + (it & 1) * 2 * N;

        // ---- gather the 16 first-pass inputs, byte sums for the detrend (exact integers)
        unsigned u[16];
        unsigned packed = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            u[i] = *reinterpret_cast<const unsigned short*>(rb + 2 * (b + M1 * i));
            packed += __byte_perm(u[i], 0, 0x4140);            // I in the low half, Q in the high half (each <= 16 * 255)
        }
        const int sI = __reduce_add_sync(0xffffffffu, (int)(packed & 0xffffu));
        const int sQ = __reduce_add_sync(0xffffffffu, (int)(packed >> 16));
        if (lane == 0) { atomicAdd(&sums[0], sI); atomicAdd(&sums[1], sQ); }
        team_sync();
        // mean = sum / N is exact in fp32 (sum < 2^24, N a power of two) and so is (float)byte - mean
        const cpk mean = c_make((float)sums[0] * (1.f / N), (float)sums[1] * (1.f / N));
        team_sync();
        if (b == 0) { sums[0] = 0; sums[1] = 0; }

        cpk v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            // 0x4B0000bb = 2^23 + byte: no I2F
            const cpk f = c_make(__uint_as_float(__byte_perm(u[i], 0x4B000000u, 0x7540)) - 8388608.f,
                                 __uint_as_float(__byte_perm(u[i], 0x4B000000u, 0x7541)) - 8388608.f);
            v[i] = c_scale(c_sub(f, mean), wj[i]);
        }

        // ---- pass 1: ncur = N, stride 1: p = b, q = 0; y[16 b + i] = W_N^{b i} DFT16(x)[i]
        cdft16(v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            cpk t = v[i];
            if (i > 0) { const float2 w = tw[(b * i) & (N - 1)]; t = c_mul(v[i], w.x, w.y); }
            *reinterpret_cast<unsigned long long*>(&X[r16_pad(16 * b + i)]) = t.v;
        }
        team_sync();
        // ---- pass 2: ncur = N/16, stride 16: p = b / 16, q = b % 16
        {
            const int p = b >> 4, q = b & 15;
            constexpr int M2 = N / 256;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i].v = *reinterpret_cast<const unsigned long long*>(&X[r16_pad(q + 16 * (p + M2 * i))]);
            team_sync();
            cdft16(v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                cpk t = v[i];
                if (i > 0) { const float2 w = tw[(16 * p * i) & (N - 1)]; t = c_mul(v[i], w.x, w.y); }
                *reinterpret_cast<unsigned long long*>(&X[r16_pad(q + 16 * (16 * p + i))]) = t.v;
            }
            team_sync();
        }
        // ---- last pass: stride N / R, p = 0: X[k = q + (N/R) i] = DFT_R(x[q + (N/R) i])[i]; outputs stay in registers
        const bool valid = seg < seg1;
        float* dst = a.S + (size_t)s * a.S_stream_stride + (size_t)seg * N;
        if (N == 4096) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i].v = *reinterpret_cast<const unsigned long long*>(&X[r16_pad(b + 256 * i)]);
            cdft16(v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float re = c_re(v[i]), im = c_im(v[i]);
                const float pw = fmaf(im, im, re * re);
                acc[i] += pw;
                if (valid) dst[b + 256 * i] = pw;
            }
        } else {
            // 256 radix-4 butterflies per segment, 4 per thread: q = b + 64 u
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
                const int q = b + 64 * uu;
                cpk x0, x1, x2, x3;
                x0.v = *reinterpret_cast<const unsigned long long*>(&X[r16_pad(q)]);
                x1.v = *reinterpret_cast<const unsigned long long*>(&X[r16_pad(q + 256)]);
                x2.v = *reinterpret_cast<const unsigned long long*>(&X[r16_pad(q + 512)]);
                x3.v = *reinterpret_cast<const unsigned long long*>(&X[r16_pad(q + 768)]);
                cdft4(x0, x1, x2, x3);
                const cpk o[4] = {x0, x1, x2, x3};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float re = c_re(o[i]), im = c_im(o[i]);
                    const float pw = fmaf(im, im, re * re);
                    acc[4 * uu + i] += pw;
                    if (valid) dst[q + 256 * i] = pw;
                }
            }
        }
        team_sync();                                            // the exchange buffer is rewritten by the next segment's pass 1
    }

    // ---- chunk row sums: teams added in fixed order, written in FFT bin order
    __syncthreads();
    float* red = reinterpret_cast<float*>(r16_smem + C::OFF_X);        // [TEAMS][N]
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int fi = N == 4096 ? (b + 256 * i) : ((b + 64 * (i >> 2)) + 256 * (i & 3));
        red[team * N + fi] = acc[i];
    }
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
