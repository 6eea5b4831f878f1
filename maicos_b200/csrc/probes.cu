// Peak probes of the pipes the roofline fractions are quoted against, run inside the measuring process
// (bench.py prints them next to the clocks of the same run instead of constants from an earlier lease):
//   * tcgen05 kind::i8 dense rate: every SM issues M = 128 x N = 256 x K = 32 int8 MMAs on one shared-memory tile
//     pair (operands resident, accumulator in TMEM) -- the ceiling of k_sf_contract_tc's pipe;
//   * FP64 DMMA rate: mma.sync.m8n8k4.f64 on register operands, 8 independent accumulator tiles per warp -- the
//     ceiling of k_sf_contract's pipe;
//   * scalar FP64 FMA rate -- the generators of k_sf_contract_tc run on it, and an SM cannot run it and the tensor pipe at once;
//   * HBM copy rate: device-to-device copy of 1 GiB (read + write bytes) -- the ceiling of the histogram path.
// Not product compute: nothing here touches user data.
#include "common.cuh"

namespace mb200 {

__device__ __forceinline__ uint32_t pr_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N>
__global__ void __launch_bounds__(128) k_probe_i8(int iters, int *status) {
    constexpr int M = 128, KSTEP = 32;
    __shared__ __align__(1024) int8_t sA[M * KSTEP];
    __shared__ __align__(1024) int8_t sB[N * KSTEP];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * KSTEP; i += 128) sA[i] = (int8_t)((i * 37 + 11) & 0xff);
    for (int i = tid; i < N * KSTEP; i += 128) sB[i] = (int8_t)((i * 53 + 5) & 0xff);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(pr_smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(pr_smem_u32(&tmem_base)), "r"((uint32_t)N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (tid == 0) {
        auto desc = [](uint32_t saddr) {
            return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
        };
        const uint64_t da = desc(pr_smem_u32(sA)), db = desc(pr_smem_u32(sB));
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc = it > 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                         "l"(da), "l"(db), "r"(idesc), "r"(acc)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(pr_smem_u32(&bar)) : "memory");
    }
    {
        const uint32_t a = pr_smem_u32(&bar);
        uint32_t done = 0;
        for (uint32_t spin = 0; !done && spin < (1u << 26); ++spin)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done)
                         : "r"(a), "r"(0u)
                         : "memory");
        if (!done && tid == 0) *status = 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"((uint32_t)N) : "memory");
}

__global__ void __launch_bounds__(256) k_probe_dmma(double *out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) k_probe_dfma(double *out, int iters, double a, double b) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb200_probe_peaks(mb200_ctx *ctx, double out[4]) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    if (!out) return fail(MB200_EINVAL, "null output pointer");
    MB_LOCK(ctx);
    MB_CUDA(cudaSetDevice(ctx->device));
    for (int i = 0; i < 4; ++i) out[i] = 0.0;
    cudaEvent_t e0, e1;
    MB_CUDA(cudaEventCreate(&e0));
    MB_CUDA(cudaEventCreate(&e1));
    DevBuf flag, scratch;
    MB_TRY(flag.ensure(1024));
    MB_CUDA(cudaMemsetAsync(flag.p, 0, 1024, ctx->stream));
    float ms = 0.f;
    // int8 tensor pipe
    {
        const int iters = 20000;
        k_probe_i8<256><<<ctx->n_sm, 128, 0, ctx->stream>>>(200, flag.as<int>());
        double best = 0.0;
        for (int rep = 0; rep < 3; ++rep) {
            MB_CUDA(cudaEventRecord(e0, ctx->stream));
            k_probe_i8<256><<<ctx->n_sm, 128, 0, ctx->stream>>>(iters, flag.as<int>());
            MB_CUDA(cudaEventRecord(e1, ctx->stream));
            MB_CUDA(cudaEventSynchronize(e1));
            MB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            best = std::max(best, 2.0 * 128 * 256 * 32 * (double)iters * ctx->n_sm / (ms * 1e-3) / 1e12);
        }
        int st = 0;
        MB_CUDA(cudaMemcpyAsync(&st, flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        MB_CUDA(cudaStreamSynchronize(ctx->stream));
        out[0] = st ? 0.0 : best;
    }
    // FP64 DMMA pipe
    {
        const int iters = 4000, grid = ctx->n_sm * 2;
        k_probe_dmma<<<grid, 256, 0, ctx->stream>>>(flag.as<double>() + 8, 100, 1.000001, 1e-9);
        double best = 0.0;
        for (int rep = 0; rep < 3; ++rep) {
            MB_CUDA(cudaEventRecord(e0, ctx->stream));
            k_probe_dmma<<<grid, 256, 0, ctx->stream>>>(flag.as<double>() + 8, iters, 1.000001, 1e-9);
            MB_CUDA(cudaEventRecord(e1, ctx->stream));
            MB_CUDA(cudaEventSynchronize(e1));
            MB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            best = std::max(best, 2.0 * 256 * 8 * (double)iters * 8.0 * grid / (ms * 1e-3) / 1e12);
        }
        out[1] = best;
    }
    // scalar FP64 FMA pipe (the generators of k_sf_contract_tc; it does NOT overlap with the tensor pipe on an SM:
    // tools/microbench/fp64_vs_umma.cu)
    {
        const int iters = 10000, grid = ctx->n_sm * 4;
        k_probe_dfma<<<grid, 256, 0, ctx->stream>>>(flag.as<double>() + 16, 100, 1.000001, 1e-9);
        double best = 0.0;
        for (int rep = 0; rep < 3; ++rep) {
            MB_CUDA(cudaEventRecord(e0, ctx->stream));
            k_probe_dfma<<<grid, 256, 0, ctx->stream>>>(flag.as<double>() + 16, iters, 1.000001, 1e-9);
            MB_CUDA(cudaEventRecord(e1, ctx->stream));
            MB_CUDA(cudaEventSynchronize(e1));
            MB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            best = std::max(best, 2.0 * 16 * (double)iters * 256.0 * grid / (ms * 1e-3) / 1e12);
        }
        out[3] = best;
    }
    // HBM copy
    {
        const size_t half = (size_t)512 << 20;
        MB_TRY(scratch.ensure(2 * half));
        MB_CUDA(cudaMemsetAsync(scratch.p, 1, 2 * half, ctx->stream));
        double best = 0.0;
        for (int rep = 0; rep < 4; ++rep) {
            MB_CUDA(cudaEventRecord(e0, ctx->stream));
            MB_CUDA(cudaMemcpyAsync(scratch.as<char>() + half, scratch.p, half, cudaMemcpyDeviceToDevice, ctx->stream));
            MB_CUDA(cudaEventRecord(e1, ctx->stream));
            MB_CUDA(cudaEventSynchronize(e1));
            MB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep) best = std::max(best, 2.0 * (double)half / (ms * 1e-3) / 1e9);
        }
        out[2] = best;
    }
    ctx->launches += 12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return MB200_OK;
}
