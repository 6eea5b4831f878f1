// Which (TMEM lane, column) does register r of thread i reach with tcgen05.st .16x256b / .16x128b / .16x64b ?
// One warp writes (i << 8 | r) codes with the shape under test and reads the block back with .32x32b (lane = thread,
// column = register).  Probe for the A-in-TMEM generator mapping; not product code.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_st_layout.bin tmem_st_layout.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(uint32_t* out) {
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base;
    for (int shape = 0; shape < 3; ++shape) {
        // clear 32 lanes x 8 columns
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(tmem), "r"(0xffffffffu) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        const uint32_t c = (uint32_t)tid << 8;
        for (int half = 0; half < 2; ++half) {   // lanes 0..15 and 16..31 of the warp's quarter
            const uint32_t addr = tmem + ((uint32_t)(16 * half) << 16);
            const uint32_t h = (uint32_t)half << 16;
            if (shape == 0)
                asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(h | c | 0), "r"(h | c | 1), "r"(h | c | 2), "r"(h | c | 3) : "memory");
            else if (shape == 1)
                asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(h | c | 0), "r"(h | c | 1), "r"(h | c | 2), "r"(h | c | 3) : "memory");
            else
                asm volatile("tcgen05.st.sync.aligned.16x64b.x4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(h | c | 0), "r"(h | c | 1), "r"(h | c | 2), "r"(h | c | 3) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tmem));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int j = 0; j < 8; ++j) out[(shape * 32 + tid) * 8 + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(64u) : "memory");
}
int main() {
    uint32_t* d; cudaMalloc(&d, 3 * 32 * 8 * 4);
    k<<<1, 32>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    uint32_t h[3 * 32 * 8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[3] = {"16x256b.x1", "16x128b.x2", "16x64b.x4"};
    for (int s = 0; s < 3; ++s) {
        printf("%s: TMEM lane -> per column (half:thread.reg)\n", names[s]);
        for (int l = 0; l < 32; ++l) {
            printf(" lane %2d:", l);
            for (int j = 0; j < 8; ++j) { uint32_t v = h[(s * 32 + l) * 8 + j]; if (v == 0xffffffffu) printf("   ----  "); else printf(" %u:%2u.%u ", v >> 16, (v >> 8) & 0xff, v & 0xff); }
            printf("\n");
        }
    }
    return 0;
}
