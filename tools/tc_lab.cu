// tc_lab.cu -- brings up and times the tensor-core spectrogram kernel (csrc/spectro_tc256.cuh) against the
// register kernel spectro_reg256_v7 on the BASELINE configs[1] shape (64 streams x 2.4 M samples).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/tc_lab tools/tc_lab.cu
// Run on the GPU box:  tools/tc_lab [streams=64] [reps=10] [T=9375]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstring>
#include <algorithm>

#include "spectro_tc256_lab.cuh"

using namespace rt;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static int perm_pos(int fi) { return ((fi >> 6) << 6) | ((fi & 15) << 2) | ((fi >> 4) & 3); }

int main(int argc, char** argv) {
    const int streams = argc > 1 ? atoi(argv[1]) : 64;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    const int T = argc > 3 ? atoi(argv[3]) : 9375;
    const int NGsel = argc > 4 ? atoi(argv[4]) : 1;
    const size_t stride = (size_t)512 * T;
    std::vector<uint8_t> h((size_t)streams * stride);
    unsigned x = 12345;
    for (size_t i = 0; i < h.size(); ++i) {
        unsigned ssum = 0;
        for (int k = 0; k < 4; ++k) { x = x * 1664525u + 1013904223u; ssum += (x >> 24); }
        int v = (int)((ssum + 2) / 4 / 8) + 112;
        // a tone in every 5th stream-second quarter so that strong cells and DC offsets are exercised
        const size_t smp = (i % stride) / 2;
        if (((i / stride) % 5) == 1 && (smp / 600000) % 2 == 1) v += (int)lrint(25.0 * ((i & 1) ? sin(0.7 * smp) : cos(0.7 * smp)));
        if (((i / stride) % 7) == 3) v += 9;     // DC offset
        h[i] = (uint8_t)std::min(255, std::max(0, v));
    }
    uint8_t* d_iq; CK(cudaMalloc(&d_iq, h.size()));
    CK(cudaMemcpy(d_iq, h.data(), h.size(), cudaMemcpyHostToDevice));
    std::vector<double> wind(256); std::vector<float> win(256); std::vector<float2> tw(256);
    double sw2 = 0;
    for (int i = 0; i < 256; ++i) { wind[i] = 0.54 - 0.46 * cos(2 * M_PI * i / 256.0); sw2 += wind[i] * wind[i]; }
    const double amp = sqrt(1.0 / (2.4e6 * sw2)) / 127.5;
    for (int i = 0; i < 256; ++i) win[i] = (float)(wind[i] * amp);
    for (int k = 0; k < 256; ++k) tw[k] = make_float2((float)cos(-2 * M_PI * k / 256.0), (float)sin(-2 * M_PI * k / 256.0));
    float* d_win; float2* d_tw;
    CK(cudaMalloc(&d_win, 1024)); CK(cudaMalloc(&d_tw, 2048));
    CK(cudaMemcpy(d_win, win.data(), 1024, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tw, tw.data(), 2048, cudaMemcpyHostToDevice));

    // ---------------- reference: register kernel v7
    const int chunk = 256, n_chunks = (T + chunk - 1) / chunk;
    float *d_S7, *d_part7;
    CK(cudaMalloc(&d_S7, (size_t)streams * T * 256 * 4));
    CK(cudaMalloc(&d_part7, (size_t)streams * n_chunks * 256 * 4));
    SpectroArgs a7;
    a7.iq = d_iq; a7.stream_stride = stride; a7.n = 256; a7.T = T; a7.chunk_segs = chunk; a7.n_chunks = n_chunks;
    a7.win = d_win; a7.tw = d_tw; a7.S = d_S7; a7.S_stream_stride = (size_t)T * 256; a7.part = d_part7;
    CK(cudaFuncSetAttribute(spectro_reg256_v7n<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, R256v7::SMEM));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best7 = 1e9f;
    for (int i = 0; i < 3 + reps; ++i) {
        CK(cudaEventRecord(e0));
        spectro_reg256_v7n<true><<<dim3(n_chunks, streams), R256v7::THREADS, R256v7::SMEM>>>(a7);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (i >= 3) best7 = std::min(best7, ms);
    }
    CK(cudaGetLastError());
    printf("v7 reference: best %.2f us\n", best7 * 1e3); fflush(stdout);

    // ---------------- tensor-core kernel
    TcTables tab = tc_make_tables(wind.data(), amp);
    printf("tables: eligible %d pscale %g wc0 (%g,%g) wc1 (%g,%g) wc255 (%g,%g)\n", (int)tab.eligible, tab.pscale, tab.wc0.x, tab.wc0.y, tab.wc1.x, tab.wc1.y, tab.wc255.x, tab.wc255.y);
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    TcArgs at;
    at.iq = d_iq; at.stream_stride = stride; at.T = T; at.n_streams = streams;
    const bool wspec = NGsel == 3;                  // 3: pipelined kernel spectro_tcp256_k
    const int NGv = NGsel == 2 ? 2 : 1;
    const int rows = 128 / NGv;
    at.bps = (T + rows - 1) / rows; at.total_batches = streams * at.bps;
    const int G = std::max(1, std::min(sms, at.total_batches / NGv));
    const long long GV = (long long)NGv * G;
    void (*kern)(TcArgs) = wspec ? spectro_tcp256_k : (NGsel == 2 ? spectro_tc256_k<2> : spectro_tc256_k<1>);
    const int SMEM = wspec ? Tcp256::SMEM : (NGsel == 2 ? Tc256<2>::SMEM : Tc256<1>::SMEM);
    const int NTHR = 512;
    int slots = 1;
    for (int s = 0; s < streams; ++s) {
        slots = std::max(slots, tc_last_run(s, at.bps, GV, at.total_batches) - tc_first_run(s, at.bps, GV, at.total_batches) + 1);
    }
    at.part_slots = slots;
    uint4* d_bmat; CK(cudaMalloc(&d_bmat, tab.bmat.size() * 2));
    CK(cudaMemcpy(d_bmat, tab.bmat.data(), tab.bmat.size() * 2, cudaMemcpyHostToDevice));
    at.bmat = d_bmat; at.wc0 = tab.wc0; at.wc1 = tab.wc1; at.wc255 = tab.wc255;
    const size_t tiles = (T + 31) / 32;
    at.S_stream_stride = tiles * 8192;
    float *d_St, *d_partt, *d_avg; unsigned* d_ctr;
    CK(cudaMalloc(&d_St, (size_t)streams * at.S_stream_stride * 4));
    CK(cudaMemset(d_St, 0, (size_t)streams * at.S_stream_stride * 4));
    CK(cudaMalloc(&d_partt, (size_t)streams * slots * 256 * 4));
    CK(cudaMemset(d_partt, 0, (size_t)streams * slots * 256 * 4));
    CK(cudaMalloc(&d_avg, (size_t)streams * 256 * 4));
    CK(cudaMalloc(&d_ctr, streams * 4)); CK(cudaMemset(d_ctr, 0, streams * 4));
    at.S = d_St; at.part = d_partt; at.avg = d_avg; at.ctr = d_ctr; at.store = 1; at.prof = nullptr; at.dbg = 0;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    printf("tc kernel NG=%d: regs %d, smem %d, grid %d, part_slots %d, batches/stream %d\n", NGsel, fa.numRegs, SMEM, G, slots, at.bps); fflush(stdout);
    float bestt = 1e9f, tot = 0;
    for (int i = 0; i < 3 + reps; ++i) {
        CK(cudaEventRecord(e0));
        kern<<<G, NTHR, SMEM>>>(at);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (i >= 3) { bestt = std::min(bestt, ms); tot += ms; }
        if (i == 0) { CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); printf("first launch ok: %.2f us\n", ms * 1e3); fflush(stdout); }
    }
    CK(cudaGetLastError());
    const double samples = (double)streams * T * 256;
    printf("tc kernel: mean %.2f us best %.2f us  %.1f GB/s algorithmic\n", 1e3 * tot / reps, 1e3 * bestt, 2 * samples / (bestt * 1e-3) / 1e9);
    at.store = 0;
    float bestn = 1e9f;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(e0));
        kern<<<G, NTHR, SMEM>>>(at);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); bestn = std::min(bestn, ms);
    }
    printf("tc kernel, no S store: best %.2f us\n", 1e3 * bestn);

    if (!wspec) {
        unsigned long long* d_prof; CK(cudaMalloc(&d_prof, G * 32 * NGv));
        at.prof = d_prof;
        for (int st = 1; st >= 0; --st) {
            at.store = st; at.dbg = 0;
            kern<<<G, NTHR, SMEM>>>(at);
            CK(cudaDeviceSynchronize());
            std::vector<unsigned long long> hp(G * 4 * NGv);
            CK(cudaMemcpy(hp.data(), d_prof, G * 32 * NGv, cudaMemcpyDeviceToHost));
            double ph[4] = {0, 0, 0, 0};
            for (int i = 0; i < G * 4 * NGv; ++i) ph[i & 3] += (double)hp[i] / (NGv * G);
            const double nb = (double)at.total_batches / (NGv * G);
            printf("phase cycles per half-batch per warp group (store %d dbg %d): convert %.0f  mma-wait %.0f  consume %.0f  flush(per run) %.0f   [half-batches per run %.1f]\n", at.store, at.dbg, ph[0] / nb, ph[1] / nb, ph[2] / nb, ph[3], nb);
        }
        at.prof = nullptr; at.store = 1; at.dbg = 0;
        kern<<<G, NTHR, SMEM>>>(at);
        CK(cudaDeviceSynchronize());
    }
    if (wspec) {
        const size_t pbytes = (size_t)G * 16 * 4 * 8;
        unsigned long long* d_prof; CK(cudaMalloc(&d_prof, pbytes));
        at.prof = d_prof;
        for (int var = 0; var < 2; ++var) {
            const int st = var == 0 ? 1 : 0;
            at.store = st; at.dbg = 0;
            CK(cudaMemset(d_prof, 0, pbytes));
            kern<<<G, NTHR, SMEM>>>(at);
            CK(cudaDeviceSynchronize());
            std::vector<unsigned long long> hp((size_t)G * 64);
            CK(cudaMemcpy(hp.data(), d_prof, pbytes, cudaMemcpyDeviceToHost));
            const double nb = (double)at.total_batches / G;
            double ph[4] = {0, 0, 0, 0}, w0[4] = {0, 0, 0, 0};
            for (size_t i = 0; i < hp.size(); ++i) { ph[i & 3] += (double)hp[i] / (G * 16.0); if (((i >> 2) & 15) == 0) w0[i & 3] += (double)hp[i] / G; }
            printf("cycles per batch per warp (store %d): convert %.0f  wait d_full %.0f  consume %.0f  flush (per run) %.0f | warp 0 (MMA issuer): convert %.0f  wait %.0f  consume %.0f   [batches per CTA %.1f]\n",
                   st, ph[0] / nb, ph[1] / nb, ph[2] / nb, ph[3], w0[0] / nb, w0[1] / nb, w0[2] / nb, nb);
        }
        at.prof = nullptr; at.store = 1; at.dbg = 0;
        kern<<<G, NTHR, SMEM>>>(at);
        CK(cudaDeviceSynchronize());
    }
    // ---------------- compare
    std::vector<float> p7((size_t)streams * n_chunks * 256), pt((size_t)streams * slots * 256), avg((size_t)streams * 256);
    CK(cudaMemcpy(p7.data(), d_part7, p7.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pt.data(), d_partt, pt.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(avg.data(), d_avg, avg.size() * 4, cudaMemcpyDeviceToHost));
    double rmax = 0, amax = 0; int rworst = -1;
    for (int s = 0; s < streams; ++s)
        for (int fi = 0; fi < 256; ++fi) {
            double r7 = 0, rt_ = 0;
            for (int c = 0; c < n_chunks; ++c) r7 += p7[((size_t)s * n_chunks + c) * 256 + perm_pos(fi)];     // v7 writes its chunk sums in PERM order
            for (int c = 0; c < slots; ++c) rt_ += pt[((size_t)s * slots + c) * 256 + fi];
            rt_ /= tab.pscale;
            const double rel = fabs(rt_ - r7) / (fabs(r7) + 1e-300);
            if (rel > rmax) { rmax = rel; rworst = s * 256 + fi; }
            const double av = (double)avg[(size_t)s * 256 + fi] / tab.pscale, rel2 = fabs(av - r7 / T) / (fabs(r7 / T) + 1e-300);
            amax = std::max(amax, rel2);
        }
    printf("row sums vs v7: max rel %.3e (stream %d bin %d);  row means vs v7: max rel %.3e\n", rmax, rworst / 256, rworst % 256, amax);
    // S cells: a few columns of a few streams
    double smax = 0, smax_big = 0; long long nbad = 0, ncmp = 0;
    const int cols[] = {0, 1, 31, 32, 127, 128, 129, 4000, T - 130, T - 2, T - 1};
    std::vector<float> c7(256), tile(8192);
    for (int s = 0; s < streams; s += std::max(1, streams / 8))
        for (int ci = 0; ci < (int)(sizeof(cols) / sizeof(cols[0])); ++ci) {
            const int t = cols[ci];
            if (t < 0 || t >= T) continue;
            CK(cudaMemcpy(c7.data(), d_S7 + ((size_t)s * T + t) * 256, 1024, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(tile.data(), d_St + (size_t)s * at.S_stream_stride + (size_t)(t >> 5) * 8192, 8192 * 4, cudaMemcpyDeviceToHost));
            float cmaxv = 0;
            for (int fi = 0; fi < 256; ++fi) cmaxv = std::max(cmaxv, c7[perm_pos(fi)]);
            for (int fi = 0; fi < 256; ++fi) {
                const double ref = c7[perm_pos(fi)];
                const double got = (double)tile[tile_cell_off(t & 31, fi)] / tab.pscale;
                const double rel = fabs(got - ref) / (fabs(ref) + 1e-300);
                ++ncmp;
                smax = std::max(smax, rel);
                if (ref > 1e-5 * cmaxv) smax_big = std::max(smax_big, rel);
                if (rel > 1e-3 && nbad < 8) { printf("  cell s %d t %d fi %d: v7 %.6e tc %.6e\n", s, t, fi, ref, got); }
                if (rel > 1e-3) ++nbad;
            }
        }
    printf("S cells vs v7: %lld compared, max rel %.3e (cells within 50 dB of the column max: %.3e), %lld beyond 1e-3\n", ncmp, smax, smax_big, nbad);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
