// Legacy-path tensor-core rates on B200 (sm_100a): mma.sync int8 (IMMA m16n8k32), bf16 (HMMA m16n8k16),
// tf32 (m16n8k8).  Sizes whether a split-precision (Ozaki-style int8 slices) S(q) contraction on the
// mma.sync path could beat the FP64 DMMA kernel; tcgen05 peaks are the nominal ones in B200_PROFILING.md.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_peaks mma_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NT>
__global__ void __launch_bounds__(256) k_imma(int* out, int iters, uint32_t a0, uint32_t b0) {
    int c[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = i;
    uint32_t a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123456789) out[0] = s;
}
template <int NT>
__global__ void __launch_bounds__(256) k_hmma_bf16(float* out, int iters, uint32_t a0, uint32_t b0) {
    float c[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = (float)i;
    uint32_t a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}
template <int NT>
__global__ void __launch_bounds__(256) k_hmma_tf32(float* out, int iters, uint32_t a0, uint32_t b0) {
    float c[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = (float)i;
    uint32_t a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d\n", p.name, p.major, p.minor, p.multiProcessorCount);
    void* out; CK(cudaMalloc(&out, 1024));
    const int iters = 4096, NT = 8;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = p.multiProcessorCount * bps;
        double warps = (double)grid * 8;
        float ms = timeit([&] { k_imma<NT><<<grid, 256>>>((int*)out, iters, 0x01020304u, 0x05060708u); });
        printf("IMMA  m16n8k32 s8  bps=%d: %.3f ms  %.1f TOP/s\n", bps, ms, warps * iters * NT * 2.0 * 16 * 8 * 32 / ms / 1e9);
        ms = timeit([&] { k_hmma_bf16<NT><<<grid, 256>>>((float*)out, iters, 0x3f803f80u, 0x3f803f80u); });
        printf("HMMA  m16n8k16 bf16 bps=%d: %.3f ms  %.1f TFLOP/s\n", bps, ms, warps * iters * NT * 2.0 * 16 * 8 * 16 / ms / 1e9);
        ms = timeit([&] { k_hmma_tf32<NT><<<grid, 256>>>((float*)out, iters, 0x3f800000u, 0x3f800000u); });
        printf("HMMA  m16n8k8  tf32 bps=%d: %.3f ms  %.1f TFLOP/s\n", bps, ms, warps * iters * NT * 2.0 * 16 * 8 * 8 / ms / 1e9);
    }
    return 0;
}
