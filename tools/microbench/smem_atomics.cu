// Shared-memory atomic rate on B200 (sm_100a): the floor of the weighted-histogram path.
// Every thread issues `iters` native 32-bit atomicAdd (SASS ATOMS.ADD) on a 2154-entry table in shared memory, with
//   (a) addresses that are conflict free inside a warp (lane + 32 k),
//   (b) uniformly random bins (what a slab histogram of a liquid sees),
//   (c) all lanes of a warp on ONE address (same-address serialisation),
// and the rate is reported in lanes per clock and SM.  k_hist_sorted needs one such atomic per element, so
// (b) x 12 bytes x 148 SMs x clock is the most the histogram kernel can stream.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics.bin smem_atomics.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NB = 2154;

// OP: 0 atomicAdd(+1), result unused (SASS ATOMS.POPC.INC)   1 atomicAdd(variable), result unused (ATOMS.ADD RZ)
//     2 atomicAdd(+1), result used (returning ATOMS.ADD: what a counting sort needs for the rank)
template <int MODE, int OP>
__global__ void __launch_bounds__(512) k_atoms(unsigned *out, int iters) {
    __shared__ unsigned tab[NB + 32];
    for (int i = threadIdx.x; i < NB + 32; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        unsigned idx;
        if (MODE == 0) idx = lane + 32 * ((x >> 8) % (NB / 32));
        else if (MODE == 1) idx = (x >> 8) % NB;
        else idx = (__shfl_sync(0xffffffffu, x, 0) >> 8) % NB;
        if (OP == 0) atomicAdd(&tab[idx], 1u);
        else if (OP == 1) atomicAdd(&tab[idx], x >> 16);
        else x += atomicAdd(&tab[idx], 1u);
    }
    __syncthreads();
    unsigned s = x == 0xfeedbeefu ? 1u : 0u;
    for (int i = threadIdx.x; i < NB; i += blockDim.x) s += tab[i];
    if (s == 0xdeadbeefu) out[0] = s;
}

template <int MODE, int OP>
void run(const char *name, int n_sm, double clk_ghz) {
    unsigned *out; CK(cudaMalloc(&out, 64));
    const int iters = 20000, ctas = n_sm * 2;
    k_atoms<MODE, OP><<<ctas, 512>>>(out, 100); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); k_atoms<MODE, OP><<<ctas, 512>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double atomics = (double)ctas * 512 * iters;
    const double per_clk_sm = atomics / (best * 1e-3) / n_sm / (clk_ghz * 1e9);
    printf("%-40s %8.3f ms  %7.2f G atomics/s  %5.2f lanes/clk/SM  -> histogram ceiling %6.0f GB/s (12 B per element)\n", name, best,
           atomics / best * 1e-6, per_clk_sm, atomics / (best * 1e-3) * 12 / 1e9);
    CK(cudaFree(out));
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const double clk = p.clockRate * 1e-6;
    printf("device %s SMs=%d clock=%.3f GHz (two CTAs of 512 threads per SM, 2154-entry table)\n", p.name, p.multiProcessorCount, clk);
    const int n = p.multiProcessorCount;
    run<0, 0>("+1, unused     conflict-free lanes", n, clk);
    run<1, 0>("+1, unused     random bins", n, clk);
    run<2, 0>("+1, unused     one address per warp", n, clk);
    run<0, 1>("+v, unused     conflict-free lanes", n, clk);
    run<1, 1>("+v, unused     random bins", n, clk);
    run<2, 1>("+v, unused     one address per warp", n, clk);
    run<0, 2>("+1, returning  conflict-free lanes", n, clk);
    run<1, 2>("+1, returning  random bins", n, clk);
    run<2, 2>("+1, returning  one address per warp", n, clk);
    return 0;
}
