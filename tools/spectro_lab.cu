// spectro_lab.cu -- times variants of the nperseg-256 spectrogram kernel (csrc/spectro256.cuh) on the
// BASELINE configs[1] shape (64 streams x 2.4 M samples) and cross-checks their row sums.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/spectro_lab tools/spectro_lab.cu
// Run on the GPU box:  tools/spectro_lab [streams=64] [reps=10]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <string>
#include <cstring>

#include "spectro256_lab.cuh"

using namespace rt;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static const char* g_filter = nullptr;
static cudaStream_t g_st[4];
static int g_nst = 1;
static int g_pipe = 0;

struct Ctx {
    int streams, T, reps;
    uint8_t* d_iq; size_t stride;
    float *d_win; float2* d_tw; float* d_S; float* d_part; float* d_part_ref;
    size_t part_elems;
    std::vector<float> ref, sref;
};

// group > 0: run the streams in groups of `group`, alternating between two S buffers (L2-resident S experiment)
typedef void (*kern_t)(SpectroArgs);
void run_k(Ctx& c, kern_t kern, int THREADS, int SMEM, const char* name, int chunk, int group = 0) {
    if (g_filter && !strstr(name, g_filter)) return;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, SMEM));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    SpectroArgs a;
    a.iq = c.d_iq; a.stream_stride = c.stride; a.n = 256; a.T = c.T; a.chunk_segs = chunk;
    a.n_chunks = (c.T + chunk - 1) / chunk; a.win = c.d_win; a.tw = c.d_tw; a.S = c.d_S; a.S_stream_stride = (size_t)c.T * 256; a.part = c.d_part;
    if ((size_t)a.n_chunks * c.streams * 256 > c.part_elems) { printf("%s: part buffer too small\n", name); return; }
    CK(cudaMemset(c.d_part, 0, c.part_elems * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto launch = [&]() {
        if (group <= 0) {
            kern<<<dim3(a.n_chunks, c.streams), THREADS, SMEM, g_st[0]>>>(a);
        } else {
            if (g_nst > 1) {   // fork
                cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); cudaEventRecord(ev, g_st[0]);
                for (int q = 1; q < g_nst; ++q) cudaStreamWaitEvent(g_st[q], ev, 0);
                cudaEventDestroy(ev);
            }
            for (int g = 0, k = 0; g < c.streams; g += group, ++k) {
                SpectroArgs b = a;
                b.iq = c.d_iq + (size_t)g * c.stride;
                b.S = c.d_S + (size_t)(k % std::max(2, g_nst)) * group * c.T * 256;      // ring of group buffers
                b.part = c.d_part + (size_t)g * a.n_chunks * 256;
                kern<<<dim3(a.n_chunks, std::min(group, c.streams - g)), THREADS, SMEM, g_st[k % g_nst]>>>(b);
            }
            if (g_nst > 1) {   // join the side streams back into stream 0 (the timing events live there)
                for (int q = 1; q < g_nst; ++q) { cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); cudaEventRecord(ev, g_st[q]); cudaStreamWaitEvent(g_st[0], ev, 0); cudaEventDestroy(ev); }
            }
        }
    };
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    float pipe_us = -1.f;
    if (g_pipe > 0 && group <= 0) {      // steady state: `g_pipe` launches back to back, alternating between two streams (the engine's schedule)
        cudaEvent_t f0, j1; CK(cudaEventCreateWithFlags(&f0, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&j1, cudaEventDisableTiming));
        CK(cudaEventRecord(e0, g_st[0]));
        CK(cudaEventRecord(f0, g_st[0])); CK(cudaStreamWaitEvent(g_st[1], f0, 0));
        for (int i = 0; i < g_pipe; ++i) kern<<<dim3(a.n_chunks, c.streams), THREADS, SMEM, g_st[i & 1]>>>(a);
        CK(cudaEventRecord(j1, g_st[1])); CK(cudaStreamWaitEvent(g_st[0], j1, 0));
        CK(cudaEventRecord(e1, g_st[0]));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        pipe_us = 1e3f * ms / g_pipe;
    }
    float best = 1e9f, tot = 0;
    for (int i = 0; i < c.reps; ++i) {
        CK(cudaEventRecord(e0, g_st[0]));
        launch();
        CK(cudaEventRecord(e1, g_st[0]));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms); tot += ms;
    }
    CK(cudaGetLastError());
    // check: total row sum per (stream, bin) against the first variant run
    std::vector<float> part((size_t)a.n_chunks * c.streams * 256);
    CK(cudaMemcpy(part.data(), c.d_part, part.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<float> rows((size_t)c.streams * 256, 0.f);
    for (int s = 0; s < c.streams; ++s)
        for (int ch = 0; ch < a.n_chunks; ++ch)
            for (int fi = 0; fi < 256; ++fi)      // spectro_reg256_v7 writes its chunk sums in PERM position order
                rows[s * 256 + fi] += part[((size_t)s * a.n_chunks + ch) * 256 + (strncmp(name, "v7", 2) == 0 ? (((fi >> 6) << 6) | ((fi & 15) << 2) | ((fi >> 4) & 3)) : fi)];
    double maxrel = 0;
    if (c.ref.empty()) c.ref = rows;
    else for (size_t i = 0; i < rows.size(); ++i) maxrel = std::max(maxrel, (double)std::fabs(rows[i] - c.ref[i]) / (std::fabs(c.ref[i]) + 1e-30));
    // S check (ungrouped storing variants): last stream, last 16 columns, against the first storing variant
    double smax = -1;
    if (group <= 0 && !strstr(name, "no store") && !strstr(name, "no S store")) {
        std::vector<float> sv(16 * 256);
        CK(cudaMemcpy(sv.data(), c.d_S + ((size_t)(c.streams - 1) * c.T + c.T - 16) * 256, sv.size() * 4, cudaMemcpyDeviceToHost));
        if (c.sref.empty()) c.sref = sv;
        smax = 0;
        for (size_t i = 0; i < sv.size(); ++i) smax = std::max(smax, (double)std::fabs(sv[i] - c.sref[i]) / (std::fabs(c.sref[i]) + 1e-30));
        CK(cudaMemset(c.d_S + ((size_t)(c.streams - 1) * c.T + c.T - 16) * 256, 0, sv.size() * 4));
    }
    const double samples = (double)c.streams * c.T * 256;
    printf("%-34s chunk %4d grp %2d regs %3d occ %2d smem %6d  mean %8.2f us  best %8.2f us  %7.1f GB/s  pipelined %7.2f us  rowsum maxrel %.2e  S maxrel %.2e\n",
           name, chunk, group, fa.numRegs, occ, SMEM, 1e3 * tot / c.reps, 1e3 * best, 2 * samples / (best * 1e-3) / 1e9, pipe_us, maxrel, smax);
    fflush(stdout);
}

template <class C>
void run(Ctx& c, const char* name, int chunk, int group = 0) { run_k(c, spectro_reg256_k<C>, C::THREADS, C::SMEM, name, chunk, group); }

int main(int argc, char** argv) {
    for (auto& st : g_st) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    Ctx c;
    c.streams = argc > 1 ? atoi(argv[1]) : 64;
    c.reps = argc > 2 ? atoi(argv[2]) : 10;
    g_filter = argc > 3 ? argv[3] : nullptr;
    g_pipe = argc > 4 ? atoi(argv[4]) : 0;
    const int block = 2400000;
    c.T = block / 256;
    c.stride = (size_t)2 * block;
    std::vector<uint8_t> h((size_t)c.streams * c.stride);
    unsigned x = 12345;
    for (auto& b : h) {   // ~gaussian-ish bytes around 127.5
        unsigned ssum = 0;
        for (int k = 0; k < 4; ++k) { x = x * 1664525u + 1013904223u; ssum += (x >> 24); }
        b = (uint8_t)std::min(255u, std::max(0u, (ssum + 2) / 4 / 8 + 112));
    }
    CK(cudaMalloc(&c.d_iq, h.size()));
    CK(cudaMemcpy(c.d_iq, h.data(), h.size(), cudaMemcpyHostToDevice));
    std::vector<float> win(256); std::vector<float2> tw(256);
    double sw2 = 0;
    for (int i = 0; i < 256; ++i) { double w = 0.54 - 0.46 * cos(2 * M_PI * i / 256.0); sw2 += w * w; win[i] = (float)w; }
    for (int i = 0; i < 256; ++i) win[i] = (float)(win[i] * sqrt(1.0 / (2.4e6 * sw2)) / 127.5);
    for (int k = 0; k < 256; ++k) tw[k] = make_float2((float)cos(-2 * M_PI * k / 256.0), (float)sin(-2 * M_PI * k / 256.0));
    CK(cudaMalloc(&c.d_win, 1024)); CK(cudaMalloc(&c.d_tw, 2048));
    CK(cudaMemcpy(c.d_win, win.data(), 1024, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.d_tw, tw.data(), 2048, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c.d_S, ((size_t)c.streams * c.T + 64) * 256 * 4));
    c.part_elems = (size_t)c.streams * 1024 * 256;
    CK(cudaMalloc(&c.d_part, c.part_elems * 4));

    //        STORE NSEG MINB WARPS STAGES WFOLD PACKACC SUMS HINT
    using V6   = R256Cfg<1, 1, 4, 4, 4, false, false, 0, false>;
    using V6n  = R256Cfg<0, 1, 4, 4, 4, false, false, 0, false>;
    using Wf   = R256Cfg<1, 1, 4, 4, 4, true,  false, 0, false>;
    using Wfn  = R256Cfg<0, 1, 4, 4, 4, true,  false, 0, false>;
    using Pa   = R256Cfg<1, 1, 4, 4, 4, true,  true,  0, false>;
    using Pan  = R256Cfg<0, 1, 4, 4, 4, true,  true,  0, false>;
    using Al   = R256Cfg<1, 1, 4, 4, 4, true,  false, 1, false>;
    using Aln  = R256Cfg<0, 1, 4, 4, 4, true,  false, 1, false>;
    using M3   = R256Cfg<1, 1, 3, 4, 4, true,  false, 1, false>;
    using M5   = R256Cfg<1, 1, 5, 4, 4, true,  false, 1, false>;
    using W8   = R256Cfg<1, 1, 2, 8, 4, true,  false, 1, false>;
    using W2   = R256Cfg<1, 1, 8, 2, 4, true,  false, 1, false>;
    using N2a  = R256Cfg<1, 2, 3, 4, 3, true,  false, 1, false>;
    using N2b  = R256Cfg<1, 2, 2, 4, 3, true,  false, 1, false>;
    using N2n  = R256Cfg<0, 2, 3, 4, 3, true,  false, 1, false>;
    using N2w8 = R256Cfg<1, 2, 1, 8, 3, true,  false, 1, false>;
    using Hn   = R256Cfg<1, 1, 4, 4, 4, true,  false, 1, true>;
    using S2   = R256Cfg<1, 1, 4, 4, 2, true,  false, 1, false>;
    using S6   = R256Cfg<1, 1, 4, 4, 6, true,  false, 1, false>;

#define RUNW(W, MB, CH) run_k(c, spectro_reg256_v7w<true, W, MB>, 32 * W, R256v7T<W, MB>::SMEM, "v7w W" #W " MB" #MB, CH)
    run_k(c, spectro_reg256_v7n<true>, R256v7::THREADS, R256v7::SMEM, "v7n", 192);
    run_k(c, spectro_reg256_v7n<true, 2>, R256v7::THREADS, R256v7::SMEM, "v7n TG2", 192);
    run_k(c, spectro_reg256_v7n<true, 2, true>, R256v7::THREADS, R256v7T<4, 4, true>::SMEM, "v7n TG2 LMAP", 192);
    run_k(c, spectro_reg256_v8<true, 2>, R256v7::THREADS, R256v7T<4, 4, true>::SMEM, "v8 TG2 (the engine's kernel)", 192);
    run_k(c, spectro_reg256_v8r<true, 104>, R256v7::THREADS, R256v7T<4, 4, true>::SMEM, "v8 TG2 maxnreg 104", 192);
    run_k(c, spectro_reg256_v8r<true, 96>, R256v7::THREADS, R256v7T<4, 4, true>::SMEM, "v8 TG2 maxnreg 96", 192);
    run_k(c, spectro_reg256_v8<false, 2>, R256v7::THREADS, R256v7T<4, 4, true>::SMEM, "v8 TG2 no store", 192);
    run_k(c, spectro_reg256_v8<true, 1>, R256v7::THREADS, R256v7T<4, 4, true>::SMEM, "v8 TG1", 192);
    run_k(c, spectro_reg256_v7n<true, 1, true>, R256v7::THREADS, R256v7T<4, 4, true>::SMEM, "v7n TG1 LMAP", 192);
    run_k(c, spectro_reg256_v7n<true, 4>, R256v7::THREADS, R256v7::SMEM, "v7n TG4", 192);
    run_k(c, spectro_reg256_v7n<true, 8>, R256v7::THREADS, R256v7::SMEM, "v7n TG8", 192);
    RUNW(4, 4, 192);
    RUNW(2, 8, 96);
    RUNW(2, 9, 96);
    RUNW(1, 16, 48);
    RUNW(1, 17, 48);
    RUNW(1, 18, 48);
    RUNW(1, 19, 48);
    RUNW(1, 17, 96);
    RUNW(1, 18, 96);
    RUNW(2, 9, 192);
    run<V6>(c, "v6 (baseline)", 256);
    run<V6n>(c, "v6 no S store", 256);
    run_k(c, spectro_reg256_v7<true>, R256v7::THREADS, R256v7::SMEM, "v7", 256);
    run_k(c, spectro_reg256_v7<false>, R256v7::THREADS, R256v7::SMEM, "v7 no store", 256);
    run_k(c, spectro_reg256_v7<true, false, 4, false, false, true>, R256v7::THREADS, R256v7::SMEM, "v7 alu sums", 256);
    run_k(c, spectro_reg256_v7<false, false, 4, false, false, true>, R256v7::THREADS, R256v7::SMEM, "v7 alu sums no store", 256);
    run_k(c, spectro_reg256_v7<true, false, 4, false, false, true>, R256v7::THREADS, R256v7::SMEM, "v7 alu sums chunk 128", 128);
    run_k(c, spectro_reg256_v7<true>, R256v7::THREADS, R256v7::SMEM, "v7 chunk 128", 128);
    run_k(c, spectro_reg256_v7<true, false, 5, true, false>, R256v7::THREADS, R256v7::SMEM, "v7 tw-smem minb5", 256);
    run_k(c, spectro_reg256_v7<true, false, 6, true, false>, R256v7::THREADS, R256v7::SMEM, "v7 tw-smem minb6", 256);
    run_k(c, spectro_reg256_v7<true, false, 5, false, true>, R256v7::THREADS, R256v7::SMEM, "v7 win-smem minb5", 256);
    run_k(c, spectro_reg256_v7<true, false, 6, true, true>, R256v7::THREADS, R256v7::SMEM, "v7 tw+win-smem minb6", 256);
    run_k(c, spectro_reg256_v7<true, false, 5, true, true>, R256v7::THREADS, R256v7::SMEM, "v7 tw+win-smem minb5", 256);
    run_k(c, spectro_reg256_v7<false, false, 6, true, false>, R256v7::THREADS, R256v7::SMEM, "v7 tw-smem minb6 no store", 256);
    run_k(c, spectro_reg256_v7<true, true>, R256v7::THREADS, R256v7::SMEM, "v7 hints", 256);
    run_k(c, spectro_reg256_v7<true>, R256v7::THREADS, R256v7::SMEM, "v7 grp4 1 stream", 64, 4);
    run_k(c, spectro_reg256_v7<true, true>, R256v7::THREADS, R256v7::SMEM, "v7 grp4 1 stream hints", 64, 4);
    g_nst = 2;
    run_k(c, spectro_reg256_v7<true>, R256v7::THREADS, R256v7::SMEM, "v7 grp4 2 streams", 64, 4);
    run_k(c, spectro_reg256_v7<true, true>, R256v7::THREADS, R256v7::SMEM, "v7 grp4 2 streams hints", 64, 4);
    run_k(c, spectro_reg256_v7<true, true>, R256v7::THREADS, R256v7::SMEM, "v7 grp2 2 streams hints", 32, 2);
    run_k(c, spectro_reg256_v7<true, true>, R256v7::THREADS, R256v7::SMEM, "v7 grp8 2 streams hints", 128, 8);
    g_nst = 4;
    run_k(c, spectro_reg256_v7<true, true>, R256v7::THREADS, R256v7::SMEM, "v7 grp4 4 streams hints", 64, 4);
    g_nst = 1;
    run<Wf>(c, "wfold", 256);
    run<Wfn>(c, "wfold no store", 256);
    run<Pa>(c, "wfold+packacc", 256);
    run<Pan>(c, "wfold+packacc no store", 256);
    run<Al>(c, "wfold+alu sums", 256);
    run<Aln>(c, "wfold+alu sums no store", 256);
    run<M3>(c, "wfold+alu minb3", 256);
    run<M5>(c, "wfold+alu minb5", 256);
    run<W8>(c, "wfold+alu 8 warps x2", 256);
    run<W2>(c, "wfold+alu 2 warps x8", 256);
    run<S2>(c, "wfold+alu stages2", 256);
    run<S6>(c, "wfold+alu stages6", 256);
    run<N2a>(c, "nseg2 minb3", 256);
    run<N2b>(c, "nseg2 minb2", 256);
    run<N2n>(c, "nseg2 minb3 no store", 256);
    run<N2w8>(c, "nseg2 8 warps x1", 256);
    run<Al>(c, "wfold+alu chunk 128", 128);
    run<Al>(c, "wfold+alu chunk 512", 512);
    run<Al>(c, "wfold+alu chunk 64", 64);
    // L2-resident S: groups of streams, ping-pong group buffers
    run<Al>(c, "grouped 4 streams chunk 64", 64, 4);
    run<Hn>(c, "grouped 4 + L2 hints chunk 64", 64, 4);
    run<Al>(c, "grouped 8 streams chunk 128", 128, 8);
    run<Hn>(c, "grouped 8 + L2 hints chunk 128", 128, 8);
    run<Hn>(c, "L2 hints ungrouped", 256);
    CK(cudaDeviceSynchronize());
    return 0;
}
