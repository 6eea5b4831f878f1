// Split-precision (Ozaki-style) complex contraction of ONE tile on tcgen05 kind::i8 -- numerical probe.
//   A[m][l] = sum_j P[m][j] * Z[l][j]   (complex, |P|,|Z| <= 1),  m < 128, l < 32, j < K
// Each real component x is cut into five balanced base-128 digits of X = rint(x * 2^34); the fifteen digit
// products with s + t <= 6 are accumulated exactly in int32 TMEM accumulators grouped by s + t (five 128 x 64
// blocks), issued as six MMAs per 16 atoms (K' = 32 = 16 atoms x {re, im}).  Reports the error against FP64.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ozaki_tile.bin ozaki_tile.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int M = 128, L = 32, NP = 2 * L, KA = 16;  // 16 atoms per MMA K-step (K' = 32)
constexpr int S = 5;
constexpr long long OFFS = 64LL * (268435456LL + 2097152LL + 16384LL + 128LL + 1LL);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; !done && spin < (1u << 24); ++spin)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    return done != 0;
}
__device__ __host__ __forceinline__ int canon(int row, int k) { return (row >> 3) * 256 + (k >> 4) * 128 + (row & 7) * 16 + (k & 15); }

// five balanced digits of X = rint(x * 2^34): d[0] most significant (range [-64, 64]), others [-64, 63]
__device__ __forceinline__ void digits5(double x, int sign, int8_t* d) {
    long long X = __double2ll_rn(x * 17179869184.0);
    unsigned long long U = (unsigned long long)(X + OFFS);
    d[0] = (int8_t)(sign * ((int)(U >> 28) - 64));
#pragma unroll
    for (int s = 1; s < S; ++s) d[s] = (int8_t)(sign * ((int)((U >> (7 * (4 - s))) & 127) - 64));
}

__global__ void __launch_bounds__(160) k_ozaki(const double2* __restrict__ P, const double2* __restrict__ Z, double2* __restrict__ out,
                                               int K, int* status) {
    extern __shared__ __align__(1024) int8_t smem[];
    int8_t* sA = smem;                       // [S][M x 32] canonical
    int8_t* sB = smem + S * M * 32;          // [S][NP x 32] canonical
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base;
    auto idesc = [](int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24); };
    uint32_t phase = 0;
    bool ok = true;
    for (int j0 = 0; j0 < K; j0 += KA) {
        // ---- slices of this K-step (generic-proxy writes) ----
        if (tid < 128) {
            const int m = tid;
            for (int a = 0; a < KA; ++a) {
                const double2 p = P[(size_t)m * K + j0 + a];
                int8_t dr[S], di[S];
                digits5(p.x, 1, dr);
                digits5(p.y, 1, di);
                for (int s = 0; s < S; ++s) {
                    sA[s * M * 32 + canon(m, 2 * a)] = dr[s];
                    sA[s * M * 32 + canon(m, 2 * a + 1)] = di[s];
                }
            }
        } else {
            const int l = tid - 128;
            for (int a = 0; a < KA; ++a) {
                const double2 z = Z[(size_t)l * K + j0 + a];
                int8_t dr[S], di[S], dn[S];
                digits5(z.x, 1, dr);
                digits5(z.y, 1, di);
                digits5(z.y, -1, dn);
                for (int t = 0; t < S; ++t) {
                    int8_t* b = sB + t * NP * 32;
                    b[canon(2 * l, 2 * a)] = dr[t];      b[canon(2 * l, 2 * a + 1)] = dn[t];   // Re row: (Zr, -Zi)
                    b[canon(2 * l + 1, 2 * a)] = di[t];  b[canon(2 * l + 1, 2 * a + 1)] = dr[t];  // Im row: (Zi, Zr)
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t first = j0 == 0 ? 0u : 1u;
            // (slice s of A) x (slices t = 1 .. t_hi of B stacked along N) -> D column blocks g = s + t
            auto mma = [&](int s, int t_lo, int t_hi, uint32_t acc) {
                const uint64_t da = make_desc(smem_u32(sA + (s - 1) * M * 32), 128, 256);
                const uint64_t db = make_desc(smem_u32(sB + (t_lo - 1) * NP * 32), 128, 256);
                const int n = (t_hi - t_lo + 1) * NP;
                const uint32_t d = tmem + (uint32_t)((s + t_lo - 2) * NP);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
                             ::"r"(d), "l"(da), "l"(db), "r"(idesc(n)), "r"(acc) : "memory");
            };
            mma(1, 1, 4, first); mma(1, 5, 5, first);
            mma(2, 1, 4, 1u); mma(3, 1, 3, 1u); mma(4, 1, 2, 1u); mma(5, 1, 1, 1u);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        }
        ok = mbar_wait(&bar, phase) && ok;   // smem may be overwritten once the MMAs are done
        phase ^= 1u;
        if (!ok) break;
    }
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (!ok) { if (tid == 0) *status = 1; }
    if (ok && tid < 128) {
        // row m = tid: A[m][l] = sum_g D_g * 2^(2 - 7g)
        double acc[NP];
        for (int c = 0; c < NP; ++c) acc[c] = 0.0;
        for (int g = 2; g <= 6; ++g) {
            const double scale = ldexp(1.0, 2 - 7 * g);
            for (int c0 = 0; c0 < NP; c0 += 32) {
                uint32_t v[32];
                const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)((g - 2) * NP + c0);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                               "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                               "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                               "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(addr));
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                for (int c = 0; c < 32; ++c) acc[c0 + c] += (double)(int32_t)v[c] * scale;
            }
        }
        for (int l = 0; l < L; ++l) out[(size_t)tid * L + l] = make_double2(acc[2 * l], acc[2 * l + 1]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    const int K = 32768;  // atoms (one flush chunk)
    std::vector<double2> hP((size_t)M * K), hZ((size_t)L * K);
    srand(7);
    auto ph = [] { double t = 6.283185307179586 * rand() / (double)RAND_MAX; return make_double2(cos(t), sin(t)); };
    for (auto& v : hP) v = ph();
    for (auto& v : hZ) v = ph();
    double2 *dP, *dZ, *dO; int* dS;
    CK(cudaMalloc(&dP, hP.size() * 16)); CK(cudaMalloc(&dZ, hZ.size() * 16)); CK(cudaMalloc(&dO, sizeof(double2) * M * L)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dP, hP.data(), hP.size() * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dZ, hZ.data(), hZ.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemset(dS, 0, 4));
    const int smem = S * M * 32 + S * NP * 32;
    CK(cudaFuncSetAttribute(k_ozaki, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_ozaki<<<1, 160, smem>>>(dP, dZ, dO, K, dS);
    CK(cudaDeviceSynchronize());
    int st; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    if (st) { printf("timeout waiting for MMA completion\n"); return 1; }
    std::vector<double2> hO(M * L);
    CK(cudaMemcpy(hO.data(), dO, sizeof(double2) * M * L, cudaMemcpyDeviceToHost));
    double max_abs = 0, max_rel_s = 0, min_amp = 1e300;
    for (int m = 0; m < M; m += 7)
        for (int l = 0; l < L; ++l) {
            long double re = 0, im = 0;
            for (int j = 0; j < K; ++j) {
                const double2 p = hP[(size_t)m * K + j], z = hZ[(size_t)l * K + j];
                re += (long double)p.x * z.x - (long double)p.y * z.y;
                im += (long double)p.x * z.y + (long double)p.y * z.x;
            }
            const double2 o = hO[m * L + l];
            const double e = fmax(fabs((double)(o.x - re)), fabs((double)(o.y - im)));
            const double s_ref = (double)(re * re + im * im), s_gpu = o.x * o.x + o.y * o.y;
            if (e > max_abs) max_abs = e;
            if (fabs(s_gpu / s_ref - 1) > max_rel_s) max_rel_s = fabs(s_gpu / s_ref - 1);
            if (sqrt(s_ref) < min_amp) min_amp = sqrt(s_ref);
        }
    printf("K = %d atoms: max |A - A_fp64| = %.3e, max rel err of S = |A|^2: %.3e (smallest |A| checked: %.3f)\n", K, max_abs, max_rel_s, min_amp);
    return 0;
}
