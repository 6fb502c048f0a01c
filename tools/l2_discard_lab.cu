// l2_discard_lab.cu -- does discard.global.L2 keep dead, dirty lines out of DRAM on the B200?
// Question behind it: power cells are dead once the scan of their stream is done; if the lines can be dropped
// from L2 without a write-back, S never costs DRAM bandwidth no matter how large its footprint is.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/l2_discard_lab tools/l2_discard_lab.cu
// Run:    ncu --cache-control none --clock-control none --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum \
//             --csv --log-file gpurun_out/l2_discard.csv tools/l2_discard_lab
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void write_k(float4* p, size_t n, float v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_float4(v, v, v, v);
}
__global__ void read_k(const float4* p, size_t n, float* out) {     // a "scan" that touches 1/16 of the lines
    float s = 0.f;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128; i < n; i += (size_t)gridDim.x * blockDim.x * 128) s += p[i].x;
    if (s == 12345.f) *out = s;
}
__global__ void discard_k(char* p, size_t bytes) {
    for (size_t off = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128; off < bytes; off += (size_t)gridDim.x * blockDim.x * 128)
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(p + off) : "memory");
}

int main() {
    const size_t group = 38400000;                 // 4 streams of S (9.6 MB each)
    const int n_groups = 16;                       // 614 MB footprint: far beyond L2
    char* buf; CK(cudaMalloc(&buf, group * n_groups));
    float* out; CK(cudaMalloc(&out, 4));
    const size_t n4 = group / 16;
    for (int variant = 0; variant < 3; ++variant) {
        // 0: write every group once (baseline: everything is written back)
        // 1: write, read a little, discard
        // 2: write, read a little, no discard, but rewrite the SAME group buffer (ring of one: absorbed by L2)
        for (int rep = 0; rep < 2; ++rep)
            for (int g = 0; g < n_groups; ++g) {
                char* p = buf + (variant == 2 ? 0 : (size_t)g * group);
                write_k<<<592, 256>>>(reinterpret_cast<float4*>(p), n4, (float)g);
                if (variant >= 1) read_k<<<148, 256>>>(reinterpret_cast<const float4*>(p), n4, out);
                if (variant == 1) discard_k<<<592, 256>>>(p, group);
            }
        CK(cudaDeviceSynchronize());
        printf("variant %d done\n", variant);
    }
    return 0;
}
