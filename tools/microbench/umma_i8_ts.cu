// tcgen05 kind::i8 with the A operand in TENSOR MEMORY (".ts" form): does a thread-written A (tcgen05.st.32x32b,
// row = TMEM lane, 8 columns of four int8 along K) feed the MMA correctly, and at what rate?  Probe for moving the
// generated A operand of the split-precision S(q) contraction out of shared memory.  Not product code.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_ts.bin umma_i8_ts.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int M = 128, KSTEP = 32;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done && spin < (1u << 24); ++spin)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    return done != 0;
}
__device__ __host__ __forceinline__ int canon(int row, int k) { return (row >> 3) * 256 + (k >> 4) * 128 + (row & 7) * 16 + (k & 15); }

// MODE 0: A from shared memory (reference), 1: A from TMEM; NSL = number of A "slices" (column blocks of 8) used round robin
template <int N, int MODE>
__global__ void __launch_bounds__(128) k_umma(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D,
                                              int iters, int* status) {
    __shared__ __align__(1024) int8_t sA[M * KSTEP];
    __shared__ __align__(1024) int8_t sB[N * KSTEP];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < M * KSTEP; i += 128) sA[canon(i / KSTEP, i % KSTEP)] = A[i];
    for (int i = tid; i < N * KSTEP; i += 128) sB[canon(i / KSTEP, i % KSTEP)] = B[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t a_col = 384;  // A operand: columns 384 .. 391
    if (MODE == 1) {
        // thread = row: 32 int8 along K -> eight 32-bit columns, byte b of column j = k = 4 j + b
        uint32_t w[8];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(A + (size_t)tid * KSTEP);
        for (int j = 0; j < 8; ++j) w[j] = src[j];
        const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + a_col;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                     "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (tid == 0) {
        const uint64_t da = make_desc(smem_u32(sA), 128, 256), db = make_desc(smem_u32(sB), 128, 256);
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc = it > 0;
            if (MODE == 0)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n"
                             ::"r"(tmem), "r"(tmem + a_col), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    }
    const bool ok = mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (!ok) { if (tid == 0) *status = 1; }
    if (ok && D) {
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = (int32_t)v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, int MODE>
int run(int n_sm) {
    std::vector<int8_t> hA(M * KSTEP), hB(N * KSTEP);
    srand(1234 + N);
    for (auto& x : hA) x = (int8_t)(rand() % 255 - 127);
    for (auto& x : hB) x = (int8_t)(rand() % 255 - 127);
    int8_t *dA, *dB; int32_t* dD; int* dS;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, sizeof(int32_t) * M * N)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dS, 0, 4)); CK(cudaMemset(dD, 0xff, sizeof(int32_t) * M * N));
    k_umma<N, MODE><<<1, 128>>>(dA, dB, dD, 1, dS);
    CK(cudaDeviceSynchronize());
    int st = 0; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    if (st) { printf("N=%d mode %d: MMA completion never signalled (timeout)\n", N, MODE); return 1; }
    std::vector<int32_t> hD(M * N);
    CK(cudaMemcpy(hD.data(), dD, sizeof(int32_t) * M * N, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            int ref = 0;
            for (int k = 0; k < KSTEP; ++k) ref += (int)hA[i * KSTEP + k] * (int)hB[j * KSTEP + k];
            if (ref != hD[i * N + j]) { if (bad < 4) printf("  mismatch D[%d][%d] = %d, expected %d\n", i, j, hD[i * N + j], ref); ++bad; }
        }
    printf("N=%d A %s: single 128x%dx32 int8 MMA vs host: %s (%d mismatches)\n", N, MODE ? "in TMEM" : "in smem", N, bad ? "FAIL" : "OK", bad);
    const int iters = 20000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_umma<N, MODE><<<n_sm, 128>>>(dA, dB, nullptr, 100, dS); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k_umma<N, MODE><<<n_sm, 128>>>(dA, dB, nullptr, iters, dS); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("N=%d A %s: %d CTAs x %d MMAs: %.3f ms -> %.1f TOP/s int8\n", N, MODE ? "in TMEM" : "in smem", n_sm, iters, ms,
           2.0 * M * N * KSTEP * (double)iters * n_sm / ms / 1e9);
    return bad != 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d\n", p.name, p.major, p.minor, p.multiProcessorCount);
    int rc = 0;
    rc |= run<64, 0>(p.multiProcessorCount);
    rc |= run<64, 1>(p.multiProcessorCount);
    rc |= run<256, 0>(p.multiProcessorCount);
    rc |= run<256, 1>(p.multiProcessorCount);
    return rc;
}
