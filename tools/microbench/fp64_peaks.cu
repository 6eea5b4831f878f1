// Microbenchmarks that size the S(q) kernel design on B200 (sm_100a):
//   DFMA pipe, DMMA (mma.sync f64) shapes, FFMA pipe, DFMA+DMMA mixed, pinned H2D.
// Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456f) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NT>
__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double c[NT][2];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}
template <int NT, int KK>
__global__ void __launch_bounds__(256) k_dmma16(double* out, int iters, double a0, double b0) {
    double c[NT][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 + i;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            if (KK == 4) dmma1684(c[i], a, b[0]);
            if (KK == 8) dmma1688(c[i], a, b);
            if (KK == 16) dmma16816(c[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// mixed: NT dmma884 tiles + NF dfma chains per iteration
template <int NT, int NF>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
    double c[NT][2], f[NF];
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}

// smem-fed DFMA: 4x4 complex register tile, operands from smem each k (GEMM-like inner loop)
__global__ void __launch_bounds__(256) k_smem_gemm(double* out, int iters) {
    __shared__ double2 sx[32][32];   // [k][m]  32 complex m values
    __shared__ double2 sy[32][32];   // [k][n]
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
        sx[i / 32][i % 32] = make_double2(1e-3 * i, 1.0);
        sy[i / 32][i % 32] = make_double2(1.0, 1e-3 * i);
    }
    __syncthreads();
    int lane = threadIdx.x & 31;
    int tm = (lane >> 2) * 4 % 32, tn = (lane & 3) * 4 + (threadIdx.x >> 5) % 2 * 16;
    double cr[4][4] = {}, ci[4][4] = {};
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            double2 x[4], y[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { x[i] = sx[k][tm + i]; y[i] = sy[k][tn + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    cr[i][j] = fma(x[i].x, y[j].x, cr[i][j]);
                    cr[i][j] = fma(-x[i].y, y[j].y, cr[i][j]);
                    ci[i][j] = fma(x[i].x, y[j].y, ci[i][j]);
                    ci[i][j] = fma(x[i].y, y[j].x, ci[i][j]);
                }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += cr[i][j] + ci[i][j];
    if (s == 123.456) out[0] = s;
}

template <typename F>
float time_ms(F f, int reps = 3) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    double* out; CK(cudaMalloc(&out, 1024));
    int nsm = p.multiProcessorCount;
    for (int bps : {1, 2, 4}) {
        int grid = nsm * bps, iters = 20000;
        {
            float ms = time_ms([&] { k_dfma<16><<<grid, 256>>>(out, iters, 1.000001, 1e-9); });
            double fl = 2.0 * 16 * iters * 256.0 * grid;
            printf("DFMA    bps=%d: %.3f ms  %.2f TFLOP/s (%.1f DFMA/clk/SM @1.965GHz)\n", bps, ms, fl / ms * 1e-9, fl / 2 / (ms * 1e-3) / nsm / 1.965e9);
        }
        {
            float ms = time_ms([&] { k_ffma<16><<<grid, 256>>>((float*)out, iters, 1.000001f, 1e-9f); });
            double fl = 2.0 * 16 * iters * 256.0 * grid;
            printf("FFMA    bps=%d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            int it2 = 4000;
            float ms = time_ms([&] { k_dmma884<8><<<grid, 256>>>(out, it2, 1.000001, 1e-9); });
            double fl = 2.0 * 256 * 8 * it2 * 8.0 * grid;
            printf("DMMA884 bps=%d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            int it2 = 4000;
            float ms = time_ms([&] { k_dmma16<8, 4><<<grid, 256>>>(out, it2, 1.000001, 1e-9); });
            double fl = 2.0 * 16 * 8 * 4 * 8 * it2 * 8.0 * grid;
            printf("DMMA16x8x4 bps=%d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            int it2 = 2000;
            float ms = time_ms([&] { k_dmma16<8, 8><<<grid, 256>>>(out, it2, 1.000001, 1e-9); });
            double fl = 2.0 * 16 * 8 * 8 * 8 * it2 * 8.0 * grid;
            printf("DMMA16x8x8 bps=%d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            int it2 = 1000;
            float ms = time_ms([&] { k_dmma16<8, 16><<<grid, 256>>>(out, it2, 1.000001, 1e-9); });
            double fl = 2.0 * 16 * 8 * 16 * 8 * it2 * 8.0 * grid;
            printf("DMMA16x8x16 bps=%d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            int it2 = 2000;
            float ms = time_ms([&] { k_mixed<4, 32><<<grid, 256>>>(out, it2, 1.000001, 1e-9); });
            double fl = (2.0 * 256 * 4 * 8.0 + 2.0 * 32 * 256) * it2 * grid;
            printf("MIXED(4 dmma884 + 32 dfma per thread-iter) bps=%d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
        {
            int it2 = 200;
            float ms = time_ms([&] { k_smem_gemm<<<grid, 256>>>(out, it2); });
            double fl = 2.0 * 64 * 32 * it2 * 256.0 * grid;
            printf("SMEM-fed 4x4 complex tile DFMA bps=%d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        }
    }
    // pinned H2D / D2H
    size_t nb = 256u << 20;
    void *h, *d; CK(cudaMallocHost(&h, nb)); CK(cudaMalloc(&d, nb));
    float ms = time_ms([&] { CK(cudaMemcpyAsync(d, h, nb, cudaMemcpyHostToDevice)); });
    printf("H2D pinned 256MiB: %.2f GB/s\n", nb / ms * 1e-6);
    ms = time_ms([&] { CK(cudaMemcpyAsync(h, d, nb, cudaMemcpyDeviceToHost)); });
    printf("D2H pinned 256MiB: %.2f GB/s\n", nb / ms * 1e-6);
    return 0;
}
