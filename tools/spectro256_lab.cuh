// spectro256_lab.cuh -- the variants of the nperseg-256 spectrogram kernel that were measured and NOT adopted
// (profiles/r01_lab_*.txt): the v1-v6 family (spectro_reg256_k / R256Cfg) and the v7 experiment wrappers
// (launch-bounds only, register caps, ALU byte sums, packed accumulators, L2 hints, time-blocked S, probe plane).
// Lab only: nothing here is compiled into librtb200.so.
#pragma once
#include "../pyradiotracking_b200/csrc/spectro256.cuh"

namespace rt {

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
    const uint8_t* base = a.unit_base(s);
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


// v7n with W warps per CTA and MB resident CTAs per SM (the body is generic in the warp count): how many warps per SM does the
// 118-register body want?  W = 1: 17 x 120 registers fit the register file where 4-warp CTAs stop at 16 warps.
template <bool STORE, int W, int MB>
__global__ void __launch_bounds__(32 * W, MB) spectro_reg256_v7w(SpectroArgs a) {
    spectro_reg256_v7_body<STORE, false, false, false, false, false, 0, true, false, R256v7T<W, MB>>(a);
}

// v8 with a register cap: would a third lean scan CTA fit beside four spectrogram CTAs (104 registers leave 12288 of an SM's 65536)?
template <bool STORE, int MAXR>
__global__ void __maxnreg__(MAXR) spectro_reg256_v8r(SpectroArgs a) {
    spectro_reg256_v8_body<STORE, 2>(a);
}
}  // namespace rt
