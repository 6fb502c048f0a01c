// tc_probe.cu -- one-off probe of tcgen05 layouts on sm_100a that the guides do not spell out:
//   (1) where the rows of an M=64 accumulator land in tensor memory (lanes), for D lane offsets 0 and 16
//   (2) which (lane, column) each thread receives from tcgen05.ld.16x256b.x1
// D[r][n] = 64 (r + 1) + n: A[r][0] = r + 1, A[r][1] = 1, B[n][0] = 64, B[n][1] = n, everything else 0.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr) {
    const uint32_t lo = ((addr >> 4) & 0x3fffu) | ((128u >> 4) << 16);
    const uint32_t hi = (512u >> 4) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(128, 1) probe(float* dump32, float* dump16) {
    extern __shared__ __align__(128) unsigned char sm[];
    __half* A = reinterpret_cast<__half*>(sm);                 // 64 rows x 16 halves (one K step): 8 row groups x 512 B (only chunk 0,1 used)
    __half* B = reinterpret_cast<__half*>(sm + 8192);          // 32 rows x 16 halves
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 16400);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8192; i += 128) reinterpret_cast<__half*>(sm)[i] = __float2half(0.f);
    __syncthreads();
    // element (r, k) at (r>>3)*512 + (k>>3)*128 + (r&7)*16 + (k&7)*2 bytes
    if (tid < 64) { A[((tid >> 3) * 512 + (tid & 7) * 16) / 2] = __float2half((float)(tid + 1)); A[((tid >> 3) * 512 + (tid & 7) * 16) / 2 + 1] = __float2half(1.f); }
    if (tid < 32) { B[((tid >> 3) * 512 + (tid & 7) * 16) / 2] = __float2half(64.f); B[((tid >> 3) * 512 + (tid & 7) * 16) / 2 + 1] = __float2half((float)tid); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
    // sentinel -1 in columns 0..63 of every lane
    {
        const uint32_t mine = tmem + ((uint32_t)(32 * warp) << 16);
        const unsigned neg = __float_as_uint(-1.f);
        for (int c = 0; c < 64; ++c) asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(mine + c), "r"(neg) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | ((32u >> 3) << 17) | ((64u >> 4) << 24);     // M=64, N=32
        for (int t = 0; t < 2; ++t) {
            const uint32_t d = tmem + ((uint32_t)(16 * t) << 16) + 32 * t;              // lane offset 0 -> cols 0..31, lane offset 16 -> cols 32..63
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                         ::"r"(d), "l"(desc(smem_u32(A))), "l"(desc(smem_u32(B))), "r"(idesc), "r"(0) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    {
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // (1) dump all 128 lanes x 64 columns with the 32x32b shape (thread = lane)
    {
        const uint32_t mine = tmem + ((uint32_t)(32 * warp) << 16);
        for (int c = 0; c < 64; ++c) {
            unsigned v;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(mine + c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v)::"memory");
            dump32[(32 * warp + lane) * 64 + c] = __uint_as_float(v);
        }
    }
    // (2) 16x256b.x1 at lane offsets 0 and 16 of this warp's quarter, columns 0..7 and 32..39
    for (int t = 0; t < 2; ++t) {
        unsigned r0, r1, r2, r3;
        const uint32_t addr = tmem + ((uint32_t)(32 * warp + 16 * t) << 16) + 32 * t;
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3)::"memory");
        float* o = dump16 + ((t * 4 + warp) * 32 + lane) * 4;
        o[0] = __uint_as_float(r0); o[1] = __uint_as_float(r1); o[2] = __uint_as_float(r2); o[3] = __uint_as_float(r3);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

int main() {
    float *d32, *d16;
    CK(cudaMalloc(&d32, 128 * 64 * 4)); CK(cudaMalloc(&d16, 2 * 4 * 32 * 4 * 4));
    CK(cudaMemset(d32, 0, 128 * 64 * 4)); CK(cudaMemset(d16, 0, 2 * 4 * 32 * 4 * 4));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16640));
    probe<<<1, 128, 16640>>>(d32, d16);
    CK(cudaDeviceSynchronize());
    std::vector<float> h32(128 * 64), h16(2 * 4 * 32 * 4);
    CK(cudaMemcpy(h32.data(), d32, h32.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h16.data(), d16, h16.size() * 4, cudaMemcpyDeviceToHost));
    for (int t = 0; t < 2; ++t) {
        printf("M=64 accumulator written with lane offset %d (columns %d..%d): lane -> row (from column n=0; -1 = untouched)\n", 16 * t, 32 * t, 32 * t + 31);
        for (int l = 0; l < 128; ++l) {
            const float v0 = h32[l * 64 + 32 * t], v1 = h32[l * 64 + 32 * t + 5];
            const int row = v0 > 0 ? (int)v0 / 64 - 1 : -1;
            printf("%s%3d:%3d%s", (l % 16 == 0) ? "  " : " ", l, row, (v0 > 0 && v1 != v0 + 5) ? "!" : "");
            if (l % 16 == 15) printf("\n");
        }
    }
    for (int t = 0; t < 2; ++t) {
        printf("16x256b.x1 at lane offset %d, columns from %d: thread -> (row, col) of r0 r1 r2 r3\n", 16 * t, 32 * t);
        for (int w = 0; w < 4; ++w)
            for (int l = 0; l < 32; ++l) {
                const float* o = &h16[((t * 4 + w) * 32 + l) * 4];
                if (w > 0 && l >= 8) continue;      // first quarter in full, the others abbreviated
                printf("  warp %d lane %2d:", w, l);
                for (int i = 0; i < 4; ++i) {
                    int fr = -1, fc = -1;
                    if (o[i] > 0) {
                        fc = (int)o[i] % 64;
                        const int row = (int)o[i] / 64 - 1;
                        for (int ll = 0; ll < 128; ++ll)
                            if (h32[ll * 64 + 32 * t] == 64.f * (row + 1)) fr = ll;
                    }
                    printf(" %6.0f(lane %3d col %d)", o[i], fr, fc);
                }
                printf("\n");
            }
    }
    return 0;
}
