// spectro_tc256_lab.cuh -- the pipelined restructuring of the tensor-core spectrogram kernel: measured
// (profiles/r01_tc_pipelined_lab.txt, 246 us) and NOT adopted.  Lab only: nothing here is compiled into librtb200.so.
#pragma once
#include "../pyradiotracking_b200/csrc/spectro_tc256.cuh"

namespace rt {
#ifdef __CUDACC__

// ---------------------------------------------------------------------------------------------------------
// Pipelined variant (lab only, never wired into the engine): the arithmetic of spectro_tc256_k<1> (M = 128 accumulators, batches of
// 128 segments, 16 warps), but without CTA-wide barriers between the phases.  Every warp runs, per batch k,
//     global loads of its 8 rows of batch k+1 (into registers, in flight during the wait)
//     wait d_full(k)            tcgen05.commit of batch k: accumulators ready AND the operand tiles free again
//     convert its 8 rows of batch k+1 -> fp16 operand tiles, arrive a_full(k+1)
//     consume its share of batch k: pass 0, pass 1 (tensor-memory loads, DFT16, |X|^2, row sums, TILE stores);
//     after the loads of pass 1: arrive d_free(k)
// and one elected lane of warp 0 issues the 64 tcgen05.mma of batch k+1 as soon as a_full(k+1) and d_free(k) are
// complete, i.e. while the warps are still busy with the arithmetic of pass 1.  The tensor pipe therefore overlaps the
// second half of the consume phase and the warps drift freely instead of meeting at bar.sync three times per batch.
// ---------------------------------------------------------------------------------------------------------
struct Tcp256 {
    static constexpr int WARPS = 16, THREADS = 512, BATCH = 128;
    static constexpr int A_MAT = BATCH * 64;
    static constexpr int OFF_A = 0;
    static constexpr int A_BYTES = 16 * A_MAT + 128;
    static constexpr int OFF_B = A_BYTES;
    static constexpr int B_BYTES = 16 * 2 * 2048;
    static constexpr int OFF_SUM = OFF_B + B_BYTES;                 // [2][BATCH] u32 segment byte sums (by batch parity)
    static constexpr int OFF_RED = OFF_SUM + 2 * BATCH * 4;         // [16 warps][4][32] floats: row-sum reduction
    static constexpr int OFF_BAR = OFF_RED + WARPS * 4 * 32 * 4;    // a_full, d_full, d_free
    static constexpr int OFF_TMEM = OFF_BAR + 32;
    static constexpr int SMEM = OFF_TMEM + 16;
    static constexpr uint32_t IDESC = Tc256<1>::IDESC;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(Tcp256::THREADS, 1) spectro_tcp256_k(TcArgs a) {
    using C = Tcp256;
    extern __shared__ __align__(128) unsigned char tc_smem[];
    __shared__ unsigned ticket;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t sm0 = smem_u32(tc_smem);
    const uint32_t bar_afull = sm0 + C::OFF_BAR, bar_dfull = bar_afull + 8, bar_dfree = bar_afull + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(tc_smem + C::OFF_TMEM);

    // ---- one-time setup: operand image of the 16 stage-1 matrices, tensor memory, barriers
    {
        uint4* dst = reinterpret_cast<uint4*>(tc_smem + C::OFF_B);
        for (int i = tid; i < C::B_BYTES / 16; i += C::THREADS) dst[i] = a.bmat[i];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm0 + C::OFF_TMEM), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_afull), "r"(C::WARPS) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_dfull) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_dfree), "r"(C::WARPS) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // this CTA's contiguous range of batches ("run")
    const long long G = gridDim.x, Btot = a.total_batches;
    const int run = blockIdx.x;
    const int lo_b = (int)((long long)run * Btot / G), hi_b = (int)((long long)(run + 1) * Btot / G);

    // consume role: lane quarter q (= warp % 4, the tensor-memory access rule), lane half hh, k1 groups j = 2 jsel + ps;
    // two rows (segments) per thread: rowA and rowA + 8
    const int q = warp & 3, hh = (warp >> 2) & 1, jsel = warp >> 3;
    const uint32_t my_tmem = tmem + ((uint32_t)(32 * q + 16 * hh) << 16) + 16 * jsel;
    const int rowA = 32 * q + 16 * hh + (lane >> 2);
    // convert role: rows warp + 16 i (i = 0..7); lane -> (chunk cc of 4 n1, sample pair pp)
    const int cc = lane >> 3, pp = lane & 7;
    const uint32_t a_dst0 = sm0 + C::OFF_A + tc_a_base(2 * pp, C::A_MAT) + cc * 128;       // + (row >> 3) * 512 + (row & 7) * 16
    const uint32_t a_dst1 = sm0 + C::OFF_A + tc_a_base(2 * pp + 1, C::A_MAT) + cc * 128;

    uint32_t w[8][4];
    auto load_rows = [&](int gb) {
        const int s = gb / a.bps, bi = gb - s * a.bps;
        const int nseg = min(C::BATCH, a.T - bi * C::BATCH);
        const uint8_t* src0 = a.iq + (size_t)s * a.stream_stride + (size_t)bi * C::BATCH * 512 + 128 * cc + 4 * pp;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = warp + 16 * i;
            const uint32_t* src = reinterpret_cast<const uint32_t*>(src0 + (size_t)row * 512);
#pragma unroll
            for (int u = 0; u < 4; ++u) w[i][u] = (row < nseg) ? __ldg(src + 8 * u) : 0x80808080u;
        }
    };
    auto prefetch_l2 = [&](int gb) {      // one 128-byte line per thread: the whole batch, pulled into L2 a batch ahead
        const int s = gb / a.bps, bi = gb - s * a.bps;
        const int nseg = min(C::BATCH, a.T - bi * C::BATCH);
        if (tid < nseg * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.iq + (size_t)s * a.stream_stride + (size_t)bi * C::BATCH * 512 + (size_t)tid * 128));
    };
    // uint8 IQ -> fp16 (b - 128, exact) operand tiles + exact byte sums of the segment
    auto convert_rows = [&](uint32_t* segsum) {
        const __half2 off = __halves2half2(__ushort_as_half(0x6480), __ushort_as_half(0x6480));   // 1024 + 128
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = warp + 16 * i;
            uint32_t e0[4], e1[4];
            unsigned sI = 0, sQ = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t x = w[i][u];                     // I(2p) Q(2p) I(2p+1) Q(2p+1) of n1 = 4c + u
                uint32_t h0 = __byte_perm(x, 0x64646464u, 0x4140);   // fp16 pair 1024 + I, 1024 + Q   (sample 2p)
                uint32_t h1 = __byte_perm(x, 0x64646464u, 0x4342);   //                                 (sample 2p+1)
                __half2 v0 = __hsub2(*reinterpret_cast<__half2*>(&h0), off);
                __half2 v1 = __hsub2(*reinterpret_cast<__half2*>(&h1), off);
                e0[u] = *reinterpret_cast<uint32_t*>(&v0);
                e1[u] = *reinterpret_cast<uint32_t*>(&v1);
                sI = __dp4a(x, 0x00010001u, sI);
                sQ = __dp4a(x, 0x01000100u, sQ);
            }
            const uint32_t roff = (uint32_t)(row >> 3) * 512 + (uint32_t)(row & 7) * 16;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_dst0 + roff), "r"(e0[0]), "r"(e0[1]), "r"(e0[2]), "r"(e0[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_dst1 + roff), "r"(e1[0]), "r"(e1[1]), "r"(e1[2]), "r"(e1[3]) : "memory");
            const unsigned tot = __reduce_add_sync(0xffffffffu, sI | (sQ << 16));   // each total <= 65280
            if (lane == 0) segsum[row] = tot;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_afull);
    };
    // stage 1 of one batch on the tensor cores: 16 n2 x 2 K-steps x (hi, lo), M = 128 (one elected lane of warp 0)
    auto issue_mma = [&]() {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 1
            for (int n2 = 0; n2 < 16; ++n2) {
                const uint32_t a_addr = sm0 + C::OFF_A + tc_a_base(n2, C::A_MAT);
                const uint32_t b_addr = sm0 + C::OFF_B + n2 * 4096;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int hl = 0; hl < 2; ++hl)
                        tc_mma(tmem + n2 * 32, tc_desc(a_addr + ks * 256), tc_desc(b_addr + hl * 2048 + ks * 256), C::IDESC, (ks | hl) != 0);
            }
            tc_commit(bar_dfull);
        }
        __syncwarp();
    };

    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    int cur_stream = -1;
    float* red = reinterpret_cast<float*>(tc_smem + C::OFF_RED);
    long long pt[4] = {0, 0, 0, 0}, pc = clock64();
#define RT_PROF(i) { if (a.prof) { const long long now_ = clock64(); pt[i] += now_ - pc; pc = now_; } }

    auto flush = [&](int s) {
        // row sums of this run over stream s: the 8 lanes sharing a k1, then the four lane quarters and two lane halves
        octet_transpose_reduce32(acc, lane);
        __syncthreads();
        // lane holds acc index i = 4 * (lane >> 2) + e (e = 0..3) for k1 offset (lane & 3): red[warp][lane & 3][i]
#pragma unroll
        for (int e = 0; e < 4; ++e) red[(warp * 4 + (lane & 3)) * 32 + 4 * (lane >> 2) + e] = acc[e];
        __syncthreads();
        const int b_first = tc_first_run(s, a.bps, G, Btot);
        float* pd = a.part + ((size_t)s * a.part_slots + (run - b_first)) * 256;
        if (tid < 256) {
            // bin fi = k1 + 16 k2; k1 = 4 (2 jsel + ps) + t3; acc index i = 16 ps + k2; warps 8 js .. 8 js + 7 share jsel = js
            const int fi = tid, k1 = fi & 15, k2 = fi >> 4, j = k1 >> 2, js = j >> 1, ps = j & 1, t3 = k1 & 3, i = 16 * ps + k2;
            float t = 0.f;
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) t += red[((8 * js + qq) * 4 + t3) * 32 + i];
            pd[fi] = t;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        // row means by the last run of the stream (fixed order over the runs)
        if (a.avg != nullptr) {
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                ticket = atomicAdd(&a.ctr[s], 1u);
            }
            __syncthreads();
            const int n_runs = tc_last_run(s, a.bps, G, Btot) - b_first + 1;
            if (ticket == (unsigned)(n_runs - 1)) {
                __threadfence();
                if (tid < 256) {
                    const float* p = a.part + (size_t)s * a.part_slots * 256 + tid;
                    double t = 0.0;
                    for (int c = 0; c < n_runs; ++c) t += (double)__ldcg(p + (size_t)c * 256);
                    a.avg[(size_t)s * 256 + tid] = (float)(t / (double)a.T);
                }
                if (tid == 0) a.ctr[s] = 0;
            }
        }
        __syncthreads();
    };

    // ---- prologue: operands of the first batch, its MMAs
    if (lo_b < hi_b) {
        load_rows(lo_b);
        if (lo_b + 1 < hi_b) prefetch_l2(lo_b + 1);
        convert_rows(reinterpret_cast<uint32_t*>(tc_smem + C::OFF_SUM));
        if (warp == 0) {
            mbar_wait(bar_afull, 0);
            issue_mma();
        }
    }
    RT_PROF(0)

    for (int gb = lo_b; gb < hi_b; ++gb) {
        const int k = gb - lo_b;
        const int s = gb / a.bps, bi = gb - s * a.bps;
        const bool more = gb + 1 < hi_b;
        if (more) load_rows(gb + 1);                             // in flight during the wait below
        if (gb + 2 < hi_b) prefetch_l2(gb + 2);
        mbar_wait(bar_dfull, k & 1);                             // accumulators of batch k ready, operand tiles free
        tc_fence_after();
        RT_PROF(1)
        if (more) convert_rows(reinterpret_cast<uint32_t*>(tc_smem + C::OFF_SUM) + C::BATCH * ((k + 1) & 1));
        RT_PROF(0)
        if (s != cur_stream) {
            if (cur_stream >= 0) flush(cur_stream);
            cur_stream = s;
            RT_PROF(3)
        }
        const int seg0 = bi * C::BATCH;
        const uint32_t* segsum = reinterpret_cast<const uint32_t*>(tc_smem + C::OFF_SUM) + C::BATCH * (k & 1);
        const unsigned totA = segsum[rowA], totB = segsum[rowA + 8];
        // residual mean after the exact -128: (sum - 32768) / 256, exact in fp32
        const cpk mresA = c_make((float)((int)(totA & 0xffffu) - 32768) * 0.00390625f, (float)((int)(totA >> 16) - 32768) * 0.00390625f);
        const cpk mresB = c_make((float)((int)(totB & 0xffffu) - 32768) * 0.00390625f, (float)((int)(totB >> 16) - 32768) * 0.00390625f);
        const int segA = seg0 + rowA, segB = segA + 8;
#pragma unroll
        for (int ps = 0; ps < 2; ++ps) {
            const int j = 2 * jsel + ps, k1 = 4 * j + (lane & 3);
            cpk vA[16], vB[16];
            {
                unsigned x[32], y[32];
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) tc_ld16x256(my_tmem + n2 * 32 + 8 * ps, x[2 * n2], x[2 * n2 + 1], y[2 * n2], y[2 * n2 + 1]);
                tc_wait_ld32(x);
                tc_wait_ld32(y);
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) {
                    asm("mov.b64 %0, {%1, %2};" : "=l"(vA[n2].v) : "r"(x[2 * n2]), "r"(x[2 * n2 + 1]));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(vB[n2].v) : "r"(y[2 * n2]), "r"(y[2 * n2 + 1]));
                }
            }
            if (ps == 1) {
                // every accumulator this warp needs from batch k is in registers
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_dfree);
                if (warp == 0 && more) {
                    mbar_wait(bar_afull, (k + 1) & 1);           // all operand tiles of batch k+1 converted
                    mbar_wait(bar_dfree, k & 1);                 // all warps have read batch k out of tensor memory
                    issue_mma();
                }
            }
            cdft16(vA);                                  // over n2 -> k2: bin = k1 + 16 k2
            cdft16(vB);
            // detrend='constant': only the bins where the window's DFT lives (0, 1, 255) see the segment mean
            if (j == 0) {
                const float wr = k1 == 0 ? a.wc0.x : (k1 == 1 ? a.wc1.x : 0.f), wi = k1 == 0 ? a.wc0.y : (k1 == 1 ? a.wc1.y : 0.f);
                vA[0] = c_sub(vA[0], c_mul(mresA, wr, wi));
                vB[0] = c_sub(vB[0], c_mul(mresB, wr, wi));
            }
            if (j == 3) {
                const float wr = k1 == 15 ? a.wc255.x : 0.f, wi = k1 == 15 ? a.wc255.y : 0.f;
                vA[15] = c_sub(vA[15], c_mul(mresA, wr, wi));
                vB[15] = c_sub(vB[15], c_mul(mresB, wr, wi));
            }
            float pA[16], pB[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
                const float ra = c_re(vA[k2]), ia = c_im(vA[k2]), rb = c_re(vB[k2]), ib = c_im(vB[k2]);
                pA[k2] = fmaf(ia, ia, ra * ra);
                pB[k2] = fmaf(ib, ib, rb * rb);
                acc[16 * ps + k2] += pA[k2] + pB[k2];
            }
            if (a.store) {
                float* sbase = a.S + (size_t)s * a.S_stream_stride + (size_t)k1 * 512;
                if (segA < a.T) {
                    float4* dst = reinterpret_cast<float4*>(sbase + (size_t)(segA >> 5) * 8192 + (size_t)(segA & 31) * 4);
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[32 * c] = make_float4(pA[4 * c], pA[4 * c + 1], pA[4 * c + 2], pA[4 * c + 3]);
                }
                if (segB < a.T) {
                    float4* dst = reinterpret_cast<float4*>(sbase + (size_t)(segB >> 5) * 8192 + (size_t)(segB & 31) * 4);
#pragma unroll
                    for (int c = 0; c < 4; ++c) dst[32 * c] = make_float4(pB[4 * c], pB[4 * c + 1], pB[4 * c + 2], pB[4 * c + 3]);
                }
            }
        }
        RT_PROF(2)
    }
    if (cur_stream >= 0) flush(cur_stream);
    RT_PROF(3)
    if (a.prof && (tid & 31) == 0)
        for (int i = 0; i < 4; ++i) a.prof[(size_t)(run * C::WARPS + warp) * 4 + i] = (unsigned long long)pt[i];
#undef RT_PROF

    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

#endif  // __CUDACC__

}  // namespace rt
