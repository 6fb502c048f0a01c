// spectro256.cuh -- uint8 IQ -> power spectrogram cells + per-chunk row sums for nperseg == 256 (sm_100a).
//
// Replaces scipy.signal.spectrogram(..., nperseg=256, noverlap=0, return_onesided=False) as called at
// /root/reference/radiotracking/analyze.py:234-241 (detrend='constant', window, FFT, |X|^2 / (fs * sum w^2)).
//
// 16 threads (a half-warp) hold one 256-point FFT as a 16x16 Cooley-Tukey in registers; every complex
// value is one packed register pair (fft_cpk.cuh).  A warp works on 2*NSEG consecutive segments per
// round, fed by a per-warp ring of TMA bulk copies.  The kernel is a template so that the engine and
// tools/spectro_lab.cu (variant timing on the GPU box) compile the very same source.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_cpk.cuh"

namespace rt {

struct SpectroArgs {
    const uint8_t* iq;
    size_t stream_stride;
    int n, T, chunk_segs, n_chunks;
    const float* win;      // window * sqrt(1/(fs*sum(w^2))) / 127.5
    const float2* tw;      // exp(-2 pi i k / n)
    float* S;              // generic kernel: [stream][T][n]; register kernel: [stream][T][pos(bin)] (see rt_engine.cu)
    size_t S_stream_stride;  // floats per stream
    float* part;           // [stream][chunk][n]   (FFT bin order; spectro_reg256_v7: PERM position order, like its S rows)
    // probe plane (spectro_reg256_v7 only, optional): a dense copy of the rows the probe loop of extract_signals looks at
    // (analyze.py:364, columns k * probe_stride), [stream][n_probes][256] in the order of the S rows, so that the probe
    // kernel reads 1 KB rows instead of one 32-byte sector per cell
    float* probe = nullptr;
    int probe_stride = 1, n_probes = 0;
};

// compile-time variant selection
template <int STORE_, int NSEG_, int MINB_, int WARPS_, int STAGES_, bool WFOLD_, bool PACKACC_, int SUMS_, bool HINT_>
struct R256Cfg {
    static constexpr int STORE = STORE_;       // 1: write every power cell to S, 0: row sums only (lab)
    static constexpr int NSEG = NSEG_;         // segments per half-warp per round (1 or 2)
    static constexpr int MINB = MINB_;         // __launch_bounds__ min CTAs per SM
    static constexpr int WARPS = WARPS_;       // warps per CTA
    static constexpr int STAGES = STAGES_;     // TMA ring depth per warp (rounds in flight)
    static constexpr bool WFOLD = WFOLD_;      // window folded into the first butterfly layer (FMA)
    static constexpr bool PACKACC = PACKACC_;  // row sums as packed (sum re^2, sum im^2) accumulators
    static constexpr int SUMS = SUMS_;         // byte sums: 0 dp4a (FMA pipe), 1 masked adds (ALU pipe)
    static constexpr bool HINT = HINT_;        // L2 hints: IQ evict-first, S evict-last
    static constexpr int THREADS = WARPS * 32;
    static constexpr int SEGS_PER_WARP = 2 * NSEG;                 // per round
    static constexpr int SEGS_PER_ROUND = WARPS * SEGS_PER_WARP;   // per CTA round
    static constexpr int RAW_STRIDE = 544;                         // 512 B of IQ + 32 B pad
    static constexpr int XROW = 36;                                // floats per exchange row (16 complex + pad)
    static constexpr int XTILE = 16 * XROW;
    static constexpr int RAW_BYTES = WARPS * STAGES * SEGS_PER_WARP * RAW_STRIDE;
    static constexpr int XCH_BYTES = 2 * WARPS * XTILE * 4;        // one tile per half-warp (reused per segment)
    static constexpr int RED_BYTES = 2 * WARPS * 256 * 4;          // final row-sum reduction (aliases raw+xch)
    static constexpr int BAR_OFF = (RAW_BYTES + XCH_BYTES) > RED_BYTES ? (RAW_BYTES + XCH_BYTES) : RED_BYTES;
    static constexpr int SMEM = BAR_OFF + WARPS * STAGES * 8;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void stg128_hint(float4* dst, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// exact byte sums of I and Q over one 256-sample segment held in shared memory (bytes I0 Q0 I1 Q1 ...);
// lane j of the half-warp adds 32 of the 512 bytes, the half-warp total comes back packed I | Q << 16
template <int SUMS>
__device__ __forceinline__ unsigned seg_byte_sums(const unsigned char* rb, int j) {
    const uint4 q0 = *reinterpret_cast<const uint4*>(rb + 16 * j);
    const uint4 q1 = *reinterpret_cast<const uint4*>(rb + 256 + 16 * j);
    unsigned tot;
    if (SUMS == 0) {
        unsigned sI = 0, sQ = 0;
        sI = __dp4a(q0.x, 0x00010001u, sI); sQ = __dp4a(q0.x, 0x01000100u, sQ);
        sI = __dp4a(q0.y, 0x00010001u, sI); sQ = __dp4a(q0.y, 0x01000100u, sQ);
        sI = __dp4a(q0.z, 0x00010001u, sI); sQ = __dp4a(q0.z, 0x01000100u, sQ);
        sI = __dp4a(q0.w, 0x00010001u, sI); sQ = __dp4a(q0.w, 0x01000100u, sQ);
        sI = __dp4a(q1.x, 0x00010001u, sI); sQ = __dp4a(q1.x, 0x01000100u, sQ);
        sI = __dp4a(q1.y, 0x00010001u, sI); sQ = __dp4a(q1.y, 0x01000100u, sQ);
        sI = __dp4a(q1.z, 0x00010001u, sI); sQ = __dp4a(q1.z, 0x01000100u, sQ);
        sI = __dp4a(q1.w, 0x00010001u, sI); sQ = __dp4a(q1.w, 0x01000100u, sQ);
        tot = sI | (sQ << 16);                     // each total <= 255*256 < 2^16
    } else {
        // 16-bit lanes: (I_even | I_odd << 16) and the same for Q; 8 words of <= 255 each stay below 2^16
        const unsigned m = 0x00ff00ffu;
        unsigned aI = (q0.x & m) + (q0.y & m) + (q0.z & m);
        unsigned bI = (q0.w & m) + (q1.x & m) + (q1.y & m);
        unsigned cI = (q1.z & m) + (q1.w & m);
        unsigned aQ = __byte_perm(q0.x, 0, 0x4341) + __byte_perm(q0.y, 0, 0x4341) + __byte_perm(q0.z, 0, 0x4341);
        unsigned bQ = __byte_perm(q0.w, 0, 0x4341) + __byte_perm(q1.x, 0, 0x4341) + __byte_perm(q1.y, 0, 0x4341);
        unsigned cQ = __byte_perm(q1.z, 0, 0x4341) + __byte_perm(q1.w, 0, 0x4341);
        const unsigned sI = aI + bI + cI, sQ = aQ + bQ + cQ;
        // fold the odd-sample lane onto the even one: I total in the low half, Q total in the high half
        tot = ((sI & 0xffffu) + (sI >> 16)) | (((sQ & 0xffffu) + (sQ >> 16)) << 16);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);   // stays inside the half-warp
    return tot;
}

__device__ __forceinline__ cpk detrend_const(unsigned tot) {
    // (32768 + mean_I, 32768 + mean_Q): exact in fp32 (the mean of 256 bytes is a multiple of 2^-8 = ulp(2^15))
    return c_make(32768.f + (float)(tot & 0xffffu) * 0.00390625f, 32768.f + (float)(tot >> 16) * 0.00390625f);
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB) spectro_reg256_k(SpectroArgs a) {
    constexpr int NSEG = C::NSEG, STAGES = C::STAGES, SPW = C::SEGS_PER_WARP, SPR = C::SEGS_PER_ROUND;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    unsigned char* raw = dyn_smem;                                                 // [warp][stage][seg][544]
    float* xch = reinterpret_cast<float*>(dyn_smem + C::RAW_BYTES);                // [half-warp][16][XROW]
    uint64_t* full = reinterpret_cast<uint64_t*>(dyn_smem + C::BAR_OFF);           // [warp][stage]

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31, h = lane >> 4, j = lane & 15;
    const int hw = tid >> 4;
    const int s = blockIdx.y;
    const int seg0 = blockIdx.x * a.chunk_segs;
    const int seg1 = min(a.T, seg0 + a.chunk_segs);
    const uint8_t* base = a.iq + (size_t)s * a.stream_stride;
    // round `it` of this warp covers segments first + SPR*it + [0, SPW): half-warp h takes h*NSEG + [0, NSEG)
    const int first = seg0 + SPW * warp;
    const int n_it = (seg1 - first + SPR - 1) / SPR;
    unsigned char* wraw = raw + warp * (STAGES * SPW * C::RAW_STRIDE);
    uint64_t* wfull = full + warp * STAGES;
    float* xt = xch + hw * C::XTILE;                // [k1][n2] complex

    uint64_t pol_in = 0, pol_out = 0;
    if (C::HINT) { pol_in = policy_evict_first(); pol_out = policy_evict_last(); }

    auto issue = [&](int st, int itx) {             // lane 0: TMA copies of round itx into stage st
        const int sg = first + SPR * itx;
        const int nv = min(SPW, seg1 - sg);
        mbar_expect_tx(&wfull[st], 512 * nv);
#pragma unroll
        for (int q = 0; q < SPW; ++q)
            if (q < nv) {
                if (C::HINT) bulk_g2s_hint(wraw + (st * SPW + q) * C::RAW_STRIDE, base + (size_t)(sg + q) * 512, 512, &wfull[st], pol_in);
                else bulk_g2s(wraw + (st * SPW + q) * C::RAW_STRIDE, base + (size_t)(sg + q) * 512, 512, &wfull[st]);
            }
    };

    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < STAGES; ++st) mbar_init(&wfull[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int st = 0; st < STAGES; ++st)
            if (st < n_it) issue(st, st);
    }

    // per-thread constants: window at samples 16*n1 + j, inter-pass twiddles W256^{j*k1} as (wr, (-wi, wi))
    float wj[16], twr[16];
    unsigned long long twp[16];
    float acc[C::PACKACC ? 1 : 16];
    cpk acc2[C::PACKACC ? 16 : 1];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        wj[i] = a.win[16 * i + j];
        const float2 t = a.tw[(j * i) & 255];
        twr[i] = t.x;
        twp[i] = cpk_pair(-t.y, t.y);
        if (C::PACKACC) acc2[i] = c_make(0.f, 0.f);
        else acc[i] = 0.f;
    }
    __syncwarp();                                   // barriers initialised before anyone polls them

    // detrend constants of the first round (later rounds: computed one round ahead, off the critical path)
    cpk cm[NSEG];
#pragma unroll
    for (int q = 0; q < NSEG; ++q) cm[q] = c_make(0.f, 0.f);
    if (n_it > 0) {
        while (!mbar_try_wait(&wfull[0], 0)) {}
#pragma unroll
        for (int q = 0; q < NSEG; ++q) cm[q] = detrend_const(seg_byte_sums<C::SUMS>(wraw + (h * NSEG + q) * C::RAW_STRIDE, j));
    }

    for (int it = 0; it < n_it; ++it) {
        const int segb = first + SPR * it + h * NSEG;   // this half-warp's first segment of the round
        const int st = it % STAGES;

        // uint8 -> float (0x4700bb00 is 32768 + b, no I2F), detrend (scipy detrend='constant'), window
        cpk v[NSEG][16];
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const unsigned char* rb = wraw + (st * SPW + h * NSEG + q) * C::RAW_STRIDE;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const unsigned u = *reinterpret_cast<const unsigned short*>(rb + 32 * n1 + 2 * j);
                const cpk f = c_make(__uint_as_float(__byte_perm(u, 0x47000000u, 0x7604)),
                                     __uint_as_float(__byte_perm(u, 0x47000000u, 0x7614)));
                v[q][n1] = C::WFOLD ? c_sub(f, cm[q]) : c_scale(c_sub(f, cm[q]), wj[n1]);
            }
        }
        // this stage's bytes are in registers: refill it with the segments STAGES rounds ahead
        __syncwarp();
        if (lane == 0 && it + STAGES < n_it) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(st, it + STAGES);
        }
        // byte sums of the NEXT round: the shuffle chain overlaps the butterflies below
        unsigned tot[NSEG];
#pragma unroll
        for (int q = 0; q < NSEG; ++q) tot[q] = 0;
        if (it + 1 < n_it) {
            const int sn = (it + 1) % STAGES;
            while (!mbar_try_wait(&wfull[sn], ((it + 1) / STAGES) & 1)) {}
#pragma unroll
            for (int q = 0; q < NSEG; ++q) tot[q] = seg_byte_sums<C::SUMS>(wraw + (sn * SPW + h * NSEG + q) * C::RAW_STRIDE, j);
        }

#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            if (C::WFOLD) cdft16_win(v[q], wj);          // over n1 -> k1, for column n2 = j
            else cdft16(v[q]);
        }
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            // inter-pass twiddles, then the 16x16 transpose through shared memory
            if (q > 0) __syncwarp();                    // the tile is reused by the half-warp's next segment
            *reinterpret_cast<unsigned long long*>(&xt[2 * j]) = v[q][0].v;
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) {
                const cpk t = c_fma_swap_p(v[q][k1], twp[k1], c_scale(v[q][k1], twr[k1]));
                *reinterpret_cast<unsigned long long*>(&xt[k1 * C::XROW + 2 * j]) = t.v;
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const ulonglong2 qq = *reinterpret_cast<const ulonglong2*>(&xt[j * C::XROW + 4 * c]);
                v[q][2 * c].v = qq.x;
                v[q][2 * c + 1].v = qq.y;
            }
        }
#pragma unroll
        for (int q = 0; q < NSEG; ++q) cm[q] = detrend_const(tot[q]);
        // (the tile is rewritten only after the next round's __syncwarp)
#pragma unroll
        for (int q = 0; q < NSEG; ++q) cdft16(v[q]);    // over n2 -> k2, for k1 = j: bin = j + 16*k2
#pragma unroll
        for (int q = 0; q < NSEG; ++q) {
            const int seg = segb + q;
            if (seg < seg1) {                           // ragged tail: idle lanes skip the epilogue
                if (C::STORE) {
                    float p[16];
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        const float re = c_re(v[q][k2]), im = c_im(v[q][k2]);
                        p[k2] = re * re + im * im;
                        if (C::PACKACC) acc2[k2] = c_fma(v[q][k2], v[q][k2], acc2[k2]);
                        else acc[k2] += p[k2];
                    }
                    float4* dst = reinterpret_cast<float4*>(a.S + ((size_t)s * a.T + seg) * 256 + 4 * j);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 o = make_float4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
                        if (C::HINT) stg128_hint(dst + 16 * c, o, pol_out);
                        else dst[16 * c] = o;
                    }
                } else {
#pragma unroll
                    for (int k2 = 0; k2 < 16; ++k2) {
                        if (C::PACKACC) acc2[k2] = c_fma(v[q][k2], v[q][k2], acc2[k2]);
                        else {
                            const float re = c_re(v[q][k2]), im = c_im(v[q][k2]);
                            acc[k2] += re * re + im * im;
                        }
                    }
                }
            }
        }
    }

    // chunk row sums: fixed-order reduction over the half-warps, written in FFT bin order (fi = j + 16*k2)
    __syncthreads();
    float* red = reinterpret_cast<float*>(dyn_smem);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2)
        red[hw * 256 + 16 * k2 + j] = C::PACKACC ? (c_re(acc2[k2]) + c_im(acc2[k2])) : acc[k2];
    __syncthreads();
    float* pd = a.part + ((size_t)s * a.n_chunks + blockIdx.x) * 256;
    for (int fi = tid; fi < 256; fi += C::THREADS) {
        float t = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2 * C::WARPS; ++hh) t += red[hh * 256 + fi];
        pd[fi] = t;
    }}

// ---------------------------------------------------------------------------------------------
// v7: same data flow as spectro_reg256_k<NSEG=1>, hand-tightened non-FMA instruction stream:
//   * warp-uniform TMA issue behind elect.sync (no per-lane address loop, no proxy fence: the
//     WAR hazard generic-read -> async-write is ordered by __syncwarp + the mbarrier itself)
//   * running shared/global offsets instead of per-round index arithmetic
//   * byte sums reduced with two REDUX (warp-wide integer add) instead of a 4-deep shuffle chain
//   * detrend constant without I2F: (2^23 + sum) * 2^-8 == 32768 + sum/256 exactly
//   * inter-pass twiddles as (wr, wi) scalars: the lane swap / sign live in FFMA2 operand modifiers
//   * invalid (ragged-tail) lanes masked by a 0/1 factor in the row-sum FMA instead of a branch
// ---------------------------------------------------------------------------------------------
struct R256v7 {
    static constexpr int WARPS = 4, THREADS = 128, STAGES = 4, MINB = 4;
    static constexpr int RAW_STRIDE = 544, XROW = 36, XTILE = 16 * XROW;
    static constexpr int STAGE_BYTES = 2 * RAW_STRIDE;                  // two segments per warp per round
    static constexpr int RAW_BYTES = WARPS * STAGES * STAGE_BYTES;
    static constexpr int XCH_BYTES = 2 * WARPS * XTILE * 4;
    static constexpr int RED_BYTES = 2 * WARPS * 256 * 4;
    static constexpr int BAR_OFF = (RAW_BYTES + XCH_BYTES) > RED_BYTES ? (RAW_BYTES + XCH_BYTES) : RED_BYTES;
    static constexpr int TAB_OFF = BAR_OFF + WARPS * STAGES * 8;         // optional constant tables: tw[16][16] float2, win[256]
    static constexpr int SMEM = TAB_OFF + 2048 + 1024;
    static constexpr int SEGS_PER_ROUND = 2 * WARPS;
};

__device__ __forceinline__ bool elect_one() {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok));
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra W;\n}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src, unsigned bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint_a(uint32_t dst, const void* src, unsigned bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned lds_u16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_2x64(uint32_t addr, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_2f32(uint32_t addr, float& a, float& b) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr));
}
__device__ __forceinline__ void sts_64(uint32_t addr, unsigned long long v) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

// per-lane partial byte sums (I in the low half, Q in the high half) of this lane's 32 bytes of a segment
__device__ __forceinline__ void lane_byte_sums(uint32_t addr, unsigned& sI, unsigned& sQ) {
    const uint4 q0 = lds_128(addr), q1 = lds_128(addr + 256);
    sI = __dp4a(q0.x, 0x00010001u, 0u); sQ = __dp4a(q0.x, 0x01000100u, 0u);
    sI = __dp4a(q0.y, 0x00010001u, sI); sQ = __dp4a(q0.y, 0x01000100u, sQ);
    sI = __dp4a(q0.z, 0x00010001u, sI); sQ = __dp4a(q0.z, 0x01000100u, sQ);
    sI = __dp4a(q0.w, 0x00010001u, sI); sQ = __dp4a(q0.w, 0x01000100u, sQ);
    sI = __dp4a(q1.x, 0x00010001u, sI); sQ = __dp4a(q1.x, 0x01000100u, sQ);
    sI = __dp4a(q1.y, 0x00010001u, sI); sQ = __dp4a(q1.y, 0x01000100u, sQ);
    sI = __dp4a(q1.z, 0x00010001u, sI); sQ = __dp4a(q1.z, 0x01000100u, sQ);
    sI = __dp4a(q1.w, 0x00010001u, sI); sQ = __dp4a(q1.w, 0x01000100u, sQ);
}

// the same sums on the ALU pipe (LOP3 / PRMT field extraction + IADD3) instead of 16 dp4a, which cost two FMA-pipe
// slots each: the kernel is FMA-pipe bound and its ALU pipe is 20 % busy
__device__ __forceinline__ void lane_byte_sums_alu(uint32_t addr, unsigned& sI, unsigned& sQ) {
    const uint4 q0 = lds_128(addr), q1 = lds_128(addr + 256);
    const unsigned m = 0x00ff00ffu;
    // 16-bit fields (even sample | odd sample << 16); eight words of <= 255 each stay below 2^16
    const unsigned aI = (q0.x & m) + (q0.y & m) + (q0.z & m);
    const unsigned bI = (q0.w & m) + (q1.x & m) + (q1.y & m);
    const unsigned cI = (q1.z & m) + (q1.w & m);
    const unsigned aQ = __byte_perm(q0.x, 0, 0x4341) + __byte_perm(q0.y, 0, 0x4341) + __byte_perm(q0.z, 0, 0x4341);
    const unsigned bQ = __byte_perm(q0.w, 0, 0x4341) + __byte_perm(q1.x, 0, 0x4341) + __byte_perm(q1.y, 0, 0x4341);
    const unsigned cQ = __byte_perm(q1.z, 0, 0x4341) + __byte_perm(q1.w, 0, 0x4341);
    const unsigned tI = aI + bI + cI, tQ = aQ + bQ + cQ;
    sI = (tI & 0xffffu) + (tI >> 16);
    sQ = (tQ & 0xffffu) + (tQ >> 16);
}

// T64: time-blocked S layout [t / 64][position / 8][t % 64][position % 8] instead of [t][position]: the 64 time steps of a group
// of 8 row positions are 2 KB of contiguous memory, so the scan kernels' walks along time read consecutive sectors (whole
// 128-byte lines, open DRAM rows) instead of one sector out of every 1 KB row.
template <bool STORE, bool HINT, bool TWS, bool WINS, bool ALUSUM, bool PROBE = false, int T64 = 0, bool PIN = false, bool PACC = false>
__device__ __forceinline__ void spectro_reg256_v7_body(const SpectroArgs& a) {
    using C = R256v7;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, h = lane >> 4, j = lane & 15;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler
    const int s = blockIdx.y;
    const int seg0 = blockIdx.x * a.chunk_segs;
    const int seg1 = min(a.T, seg0 + a.chunk_segs);
    const int first = seg0 + 2 * warp;
    const int n_it = (seg1 - first + C::SEGS_PER_ROUND - 1) / C::SEGS_PER_ROUND;   // warp-uniform

    const uint32_t sm0 = smem_u32(dyn_smem);
    const uint32_t wraw = sm0 + warp * (C::STAGES * C::STAGE_BYTES);
    const uint32_t wbar = sm0 + C::BAR_OFF + warp * (C::STAGES * 8);
    const uint32_t xt = sm0 + C::RAW_BYTES + (tid >> 4) * (C::XTILE * 4);
    const uint32_t my_u16 = wraw + h * C::RAW_STRIDE + 2 * j;        // + stage offset + 32*n1
    const uint32_t my_sum = wraw + h * C::RAW_STRIDE + 16 * j;       // + stage offset (+256)
    uint32_t xt_st = xt + 8 * j;                                      // + k1 * XROW * 4
    uint32_t xt_ld = xt + j * (C::XROW * 4);                          // + 16 * c
    uint32_t my_u16_ = my_u16, my_sum_ = my_sum;
    if (PIN) {
        // opaque to the compiler: without this ptxas re-derives these four addresses from SR_TID.X in every round
        // (S2R -> SHF -> IMAD -> IADD3 in front of the round's first LDS)
        asm volatile("" : "+r"(my_u16_), "+r"(my_sum_), "+r"(xt_st), "+r"(xt_ld));
    }
    int h16 = 16 * h;                                                 // shift of this half-warp's field in the REDUX sums
    if (PIN) asm volatile("" : "+r"(h16));
    // global source of the segment pair of round `it`: base + (first + 8*it) * 512
    const uint8_t* gsrc = a.iq + (size_t)s * a.stream_stride + (size_t)first * 512;
    const int last_seg = seg1 - 1;

    uint64_t pol_in = 0, pol_out = 0;
    if (HINT) { pol_in = policy_evict_first(); pol_out = PIN ? policy_evict_first() : policy_evict_last(); }   // v7h (HINT + PIN): S stores evict-first (streaming)
    auto issue = [&](int itx, int st) {                              // executed by one elected lane
        const int sg = first + C::SEGS_PER_ROUND * itx;
        const uint8_t* p0 = gsrc + (size_t)itx * (C::SEGS_PER_ROUND * 512);
        const uint8_t* p1 = (sg + 1 <= last_seg) ? p0 + 512 : p0;   // ragged tail: copy the same segment twice
        const uint32_t bar = wbar + 8 * st, dst = wraw + st * C::STAGE_BYTES;
        mbar_expect_tx_a(bar, 1024);
        if (HINT) {
            bulk_g2s_hint_a(dst, p0, 512, bar, pol_in);
            bulk_g2s_hint_a(dst + C::RAW_STRIDE, p1, 512, bar, pol_in);
        } else {
            bulk_g2s_a(dst, p0, 512, bar);
            bulk_g2s_a(dst + C::RAW_STRIDE, p1, 512, bar);
        }
    };

    if (elect_one()) {
#pragma unroll
        for (int st = 0; st < C::STAGES; ++st)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(wbar + 8 * st) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int st = 0; st < C::STAGES; ++st)
            if (st < n_it) issue(st, st);
    }

    // per-thread constants: window at samples 16*n1 + j, inter-pass twiddles W256^{j*k1}
    // (in registers, or -- TWS / WINS -- in a shared-memory table to buy a fifth / sixth resident CTA)
    float wj[16], twr[16], twi[16], acc[16];
    cpk acc2[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc2[c] = c_make(0.f, 0.f);
    const uint32_t tab_tw = sm0 + C::TAB_OFF + 8 * j, tab_win = sm0 + C::TAB_OFF + 2048 + 4 * j;
    if (TWS || WINS) {
        float2* ttw = reinterpret_cast<float2*>(dyn_smem + C::TAB_OFF);
        float* twin = reinterpret_cast<float*>(dyn_smem + C::TAB_OFF + 2048);
        for (int i = tid; i < 256; i += C::THREADS) {
            ttw[i] = a.tw[((i & 15) * (i >> 4)) & 255];      // [k1][j]
            twin[i] = a.win[i];                              // [n1][j]
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        wj[i] = a.win[16 * i + j];
        const float2 t = a.tw[(j * i) & 255];
        twr[i] = t.x;
        twi[i] = t.y;
        acc[i] = 0.f;
    }
    __syncwarp();                                   // barriers initialised before anyone polls them

    // segment byte sums -> detrend constant (32768 + mean_I, 32768 + mean_Q), exact in fp32
    auto detrend_of = [&](uint32_t addr) {
        unsigned sI, sQ;
        if (ALUSUM) lane_byte_sums_alu(addr, sI, sQ);
        else lane_byte_sums(addr, sI, sQ);
        // warp-wide REDUX: half-warp 0 in the low 16 bits, half-warp 1 in the high 16 (each total < 2^16)
        const unsigned tI = __reduce_add_sync(0xffffffffu, sI << h16);
        const unsigned tQ = __reduce_add_sync(0xffffffffu, sQ << h16);
        const unsigned mI = (tI >> h16) & 0xffffu, mQ = (tQ >> h16) & 0xffffu;
        return c_scale(c_make(__uint_as_float(0x4B000000u | mI), __uint_as_float(0x4B000000u | mQ)), 0.00390625f);
    };

    cpk cm = c_make(0.f, 0.f);
    if (n_it > 0) {
        mbar_wait(wbar, 0);
        cm = detrend_of(my_sum_);
    }
    float* sdst = a.S + (size_t)s * a.S_stream_stride + (size_t)(first + h) * 256 + 4 * j;   // += 8 * 256 floats per round
    // T64 = positions per block (8 or 32): [t / 64][pos / T64][t % 64][pos % T64]; this thread's positions are 64 c + 4 j ...
    float* const sbase64 = a.S + (size_t)s * a.S_stream_stride + (T64 ? ((4 * j) / (T64 ? T64 : 1)) * (64 * T64) + (4 * j) % (T64 ? T64 : 1) : 0);
    uint32_t st_off = 0, st_bar = wbar;
    unsigned phase = 0;
    const uint8_t* g_next = gsrc + (size_t)C::STAGES * (C::SEGS_PER_ROUND * 512);   // source of the next round to be issued
    int sg_next = first + C::STAGES * C::SEGS_PER_ROUND;
    int seg = first + h;
    // probe plane: seg = pq * probe_stride + prem, kept incrementally (seg advances by SEGS_PER_ROUND per round)
    int pq = 0, prem = 0;
    if (PROBE) { pq = seg / a.probe_stride; prem = seg - pq * a.probe_stride; }

    for (int it = 0; it < n_it; ++it) {
        // uint8 -> float (0x4700bb00 is 32768 + b, no I2F), detrend (scipy detrend='constant'); window folded below
        cpk v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const unsigned u = lds_u16(my_u16_ + st_off + 32 * n1);
            const cpk f = c_make(__uint_as_float(__byte_perm(u, 0x47000000u, 0x7604)),
                                 __uint_as_float(__byte_perm(u, 0x47000000u, 0x7614)));
            v[n1] = c_sub(f, cm);
        }
        // this stage's bytes are in registers: refill it with the segments STAGES rounds ahead
        __syncwarp();
        if (it + C::STAGES < n_it) {
            if (PIN) {
                // running source pointer / ring slot instead of re-deriving both from the round number (uniform datapath)
                if (elect_one()) {
                    const uint8_t* p1 = (sg_next + 1 <= last_seg) ? g_next + 512 : g_next;
                    mbar_expect_tx_a(st_bar, 1024);
                    bulk_g2s_a(wraw + st_off, g_next, 512, st_bar);
                    bulk_g2s_a(wraw + st_off + C::RAW_STRIDE, p1, 512, st_bar);
                }
            } else if (elect_one()) issue(it + C::STAGES, it % C::STAGES);
        }
        g_next += C::SEGS_PER_ROUND * 512;
        sg_next += C::SEGS_PER_ROUND;
        // advance to the next stage; its byte sums overlap the butterflies below
        st_off += C::STAGE_BYTES; st_bar += 8;
        if (st_off == C::STAGES * C::STAGE_BYTES) { st_off = 0; st_bar = wbar; phase ^= 1; }
        cpk cm_next = cm;
        if (it + 1 < n_it) {
            mbar_wait(st_bar, phase);
            cm_next = detrend_of(my_sum_ + st_off);
        }

        if (WINS) {
#pragma unroll
            for (int i = 0; i < 16; ++i) wj[i] = lds_f32(tab_win + 64 * i);
        }
        cdft16_win(v, wj);                          // over n1 -> k1, for column n2 = j
        // inter-pass twiddles, then the 16x16 transpose through shared memory
        sts_64(xt_st, v[0].v);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) {
            float wr = twr[k1], wi = twi[k1];
            if (TWS) lds_2f32(tab_tw + 128 * k1, wr, wi);
            sts_64(xt_st + k1 * (C::XROW * 4), c_mul(v[k1], wr, wi).v);
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) lds_2x64(xt_ld + 16 * c, v[2 * c].v, v[2 * c + 1].v);
        // (the tile is rewritten only after the next round's __syncwarp)
        cdft16(v);                                  // over n2 -> k2, for k1 = j: bin = j + 16*k2
        const bool valid = seg <= last_seg;
        const float m = valid ? 1.f : 0.f;
        float p[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            const float re = c_re(v[k2]), im = c_im(v[k2]);
            p[k2] = fmaf(im, im, re * re);
            if (!PACC) acc[k2] = fmaf(p[k2], m, acc[k2]);
        }
        if (PACC) {                                 // row sums as 8 packed FMAs instead of 16 scalar ones (same lanes, half the issue slots)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc2[c] = c_fma_s(c_make(p[2 * c], p[2 * c + 1]), m, acc2[c]);
        }
        if (STORE) {
            if (valid) {
                float4* dst = T64 ? reinterpret_cast<float4*>(sbase64 + ((size_t)(seg >> 6) << 14) + (seg & 63) * T64) : reinterpret_cast<float4*>(sdst);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 o = make_float4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
                    if (HINT) stg128_hint(dst + 16 * c, o, pol_out);
                    else dst[(T64 ? 1024 : 16) * c] = o;          // 64 positions further: 64 * 64 floats in either blocked layout
                }
            }
            sdst += C::SEGS_PER_ROUND * 256;
            if (PROBE) {
                if (valid && prem == 0) {
                    float4* dst = reinterpret_cast<float4*>(a.probe + ((size_t)s * a.n_probes + pq) * 256 + 4 * j);
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[16 * c] = make_float4(p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
                }
                prem += C::SEGS_PER_ROUND;
                while (prem >= a.probe_stride) { prem -= a.probe_stride; ++pq; }
            }
        }
        seg += C::SEGS_PER_ROUND;
        cm = cm_next;
    }

    // chunk row sums: fixed-order reduction over the half-warps, written in FFT bin order (fi = j + 16*k2)
    __syncthreads();
    float* red = reinterpret_cast<float*>(dyn_smem);
    const int hw = tid >> 4;
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) red[hw * 256 + 16 * k2 + j] = PACC ? ((k2 & 1) ? c_im(acc2[k2 >> 1]) : c_re(acc2[k2 >> 1])) : acc[k2];
    __syncthreads();
    // written in the order of the S rows (PERM position, see rt_engine.cu) so that the probe kernel reads both coalesced
    float* pd = a.part + ((size_t)s * a.n_chunks + blockIdx.x) * 256;
    for (int fi = tid; fi < 256; fi += C::THREADS) {
        float t = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2 * C::WARPS; ++hh) t += red[hh * 256 + fi];
        pd[((fi >> 6) << 6) | ((fi & 15) << 2) | ((fi >> 4) & 3)] = t;
    }
}

template <bool STORE, bool HINT = false, int MINB = 4, bool TWS = false, bool WINS = false, bool ALUSUM = false>
__global__ void __launch_bounds__(R256v7::THREADS, MINB) spectro_reg256_v7(SpectroArgs a) {
    spectro_reg256_v7_body<STORE, HINT, TWS, WINS, ALUSUM>(a);
}

// The engine's variant: registers capped at MAXR instead of "4 CTAs per SM".  At 112 registers four resident CTAs leave 8192
// registers of an SM unused -- room for two 128-thread scan CTAs of 32 registers (rt_engine.cu, lean scan kernels), which then
// run beside the spectrogram of the next launch instead of displacing its CTAs.
template <bool STORE, int MAXR>
__global__ void __maxnreg__(MAXR) spectro_reg256_v7r(SpectroArgs a) {
    spectro_reg256_v7_body<STORE, false, false, false, false>(a);
}

// addresses pinned in registers (see PIN in spectro_reg256_v7_body)
template <bool STORE>
__global__ void __launch_bounds__(R256v7::THREADS, 4) spectro_reg256_v7n(SpectroArgs a) {
    spectro_reg256_v7_body<STORE, false, false, false, false, false, 0, true>(a);
}

// experiment variants of v7n: byte sums on the ALU pipe and / or packed row-sum accumulators
template <bool STORE, bool ALUSUM, bool PACC>
__global__ void __launch_bounds__(R256v7::THREADS, 4) spectro_reg256_v7x(SpectroArgs a) {
    spectro_reg256_v7_body<STORE, false, false, false, ALUSUM, false, 0, true, PACC>(a);
}

// experiment variant of v7n: L2 hints (IQ evict-first, S evict-last) under the overlapped schedule
template <bool STORE>
__global__ void __launch_bounds__(R256v7::THREADS, 4) spectro_reg256_v7h(SpectroArgs a) {
    spectro_reg256_v7_body<STORE, true, false, false, false, false, 0, true, false>(a);
}

// time-blocked S layout (see spectro_reg256_v7_body)
template <bool STORE, int PB>
__global__ void __launch_bounds__(R256v7::THREADS, 4) spectro_reg256_v7t(SpectroArgs a) {
    spectro_reg256_v7_body<STORE, false, false, false, false, false, PB>(a);
}

// variant that also writes the probe plane (SpectroArgs::probe)
template <bool STORE>
__global__ void __launch_bounds__(R256v7::THREADS, 4) spectro_reg256_v7p(SpectroArgs a) {
    spectro_reg256_v7_body<STORE, false, false, false, false, true>(a);
}

}  // namespace rt
