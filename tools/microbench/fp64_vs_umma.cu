// Do FP64 FMAs and tcgen05 int8 MMAs overlap on a B200 SM?  (ncu reports them as ONE "shared" pipe.)
// One CTA per SM: warp 0 issues `mma_iters` UTCIMMA (M = 128, N = 256, K = 32, operands resident in shared memory),
// warps 1 .. 16 run `fma_iters` rounds of 8 independent DFMA chains each.  Timed: MMAs alone, DFMAs alone, both together.
// If the two kinds of work shared an execution resource the combined time would approach the sum; if they are independent
// it approaches the larger of the two.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_vs_umma.bin fp64_vs_umma.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N>
__global__ void __launch_bounds__(17 * 32) k_mix(int mma_iters, int fma_iters, double a, double b, double* out, int* status) {
    constexpr int M = 128, KSTEP = 32;
    __shared__ __align__(1024) int8_t sA[M * KSTEP];
    __shared__ __align__(1024) int8_t sB[N * KSTEP];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * KSTEP; i += blockDim.x) sA[i] = (int8_t)((i * 37 + 11) & 0xff);
    for (int i = tid; i < N * KSTEP; i += blockDim.x) sB[i] = (int8_t)((i * 53 + 5) & 0xff);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "r"((uint32_t)N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (warp == 0) {
        if (tid == 0 && mma_iters > 0) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
            auto desc = [](uint32_t saddr) { return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46); };
            const uint64_t da = desc(smem_u32(sA)), db = desc(smem_u32(sB));
            for (int it = 0; it < mma_iters; ++it) {
                const uint32_t acc = it > 0;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        }
        if (mma_iters > 0) {
            uint32_t done = 0;
            for (uint32_t spin = 0; !done && spin < (1u << 26); ++spin)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (!done && tid == 0) *status = 1;
        }
    } else {
        double c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = tid * 1e-3 + i;
        for (int it = 0; it < fma_iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i];
        if (s == 123.456) out[0] = s;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((uint32_t)N) : "memory");
}

template <int N>
float run(int n_sm, int mma_iters, int fma_iters, double* out, int* st) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_mix<N><<<n_sm, 17 * 32>>>(mma_iters ? 50 : 0, fma_iters ? 50 : 0, 1.000001, 1e-9, out, st); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); k_mix<N><<<n_sm, 17 * 32>>>(mma_iters, fma_iters, 1.000001, 1e-9, out, st); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double* out; int* st; CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&st, 4)); CK(cudaMemset(st, 0, 4));
    const int n = p.multiProcessorCount;
    printf("device %s SMs=%d: one CTA per SM, warp 0 issues UTCIMMA (M=128, K=32), 16 warps run DFMA chains\n", p.name, n);
    for (int fma_iters : {2000, 4000, 8000}) {
        const int mma_iters = 10000;
        const float t_m256 = run<256>(n, mma_iters, 0, out, st), t_f = run<256>(n, 0, fma_iters, out, st), t_b256 = run<256>(n, mma_iters, fma_iters, out, st);
        const float t_m64 = run<64>(n, mma_iters, 0, out, st), t_b64 = run<64>(n, mma_iters, fma_iters, out, st);
        printf("fma_iters %5d: DFMA alone %.3f ms | N=256: MMA alone %.3f, both %.3f (sum %.3f, max %.3f) | N=64: MMA alone %.3f, both %.3f (sum %.3f, max %.3f)\n",
               fma_iters, t_f, t_m256, t_b256, t_m256 + t_f, t_m256 > t_f ? t_m256 : t_f, t_m64, t_b64, t_m64 + t_f, t_m64 > t_f ? t_m64 : t_f);
    }
    int s = 0; CK(cudaMemcpy(&s, st, 4, cudaMemcpyDeviceToHost));
    printf("timeout flag %d\n", s);
    return 0;
}
