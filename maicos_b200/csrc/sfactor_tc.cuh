// Split-precision S(q) contraction on the 5th-generation tensor cores (tcgen05 kind::i8, TMEM accumulators).
// Included by sfactor.cu (same translation unit: shares SegDesc / SfPlan / the mbarrier helpers).
//
// Same contraction as k_sf_contract,  A[(h,k), l] = sum_j P_j[(h,k)] * Z_j[l]  (complex), evaluated as a real
// int8 GEMM per 16 atoms (K' = 32 = 16 atoms x {re, im}):
//   * every real component x in [-1, 1] is cut into five balanced base-256 digits (int8, [-128, 127]) of
//     X = rint(x * 2^38): one FP64 add with a magic constant leaves X + bias in the low mantissa bits, whose
//     BYTES xor 0x80 are the digits -- no shifts, no masks;
//   * the fifteen digit products a_s * b_t with s + t <= 6 are accumulated EXACTLY in int32 TMEM accumulators
//     grouped by g = s + t (five 128 x NP column blocks, NP = 2 * L_t for re/im of an l tile), issued as
//     MMAs "slice s of A" x "slices t = 1 .. 6 - s of B stacked along N" (N up to 256);
//   * A = 16 * sum_g D_g * 256^-g in FP64 when an atom chunk (<= 8192 atoms, no int32 overflow) is flushed.
//   Error: input quantisation 2^-39 and the dropped products (s + t >= 7, <= 2^-36 each) give |dA| ~ 1e-9 for
//   1e5 atoms: measured <= 5e-9 relative on EVERY nonzero cell of the raw S cube up to 1M atoms and ~1e-12 on binned
//   S(q) (profiles/r01_tc_notes.md), against a stated tolerance of 1e-6 -- this is the default arithmetic
//   (mb200_ctx_set_precision); the FP64 DMMA kernel stays selectable and takes the plans that cannot be tiled here.
//
// CTA = 16 generator warps + MMA warp + seed loader warp + B loader warp, one CTA per SM (512 TMEM columns):
//   generators : tasks (K-step, 4 row quads x 8 atom pairs) from a shared counter; thread = 2 atoms x the 4 packed (h,k)
//                rows of a quad: P_r = (w Ex[h0] Ey[k0]) Ey[r] from the seed rows the loader warp keeps in a
//                shared-memory ring; digits are written as (re, im) byte pairs either into the canonical K-major
//                no-swizzle UMMA layout in shared memory (cute ((8,n),2):((1,SBO),LBO), SBO = 144, LBO = 2336) or --
//                template parameter ATM, l tiles of at most 32 values -- straight into TENSOR MEMORY
//                (tcgen05.st.16x128b; tasks per TMEM lane quarter);
//   seed loader: 16-byte cp.async gathers of the item's distinct Ex[h0] / Ey[k0] rows and Ey[1], up to 8 K-steps ahead;
//   B loader   : one bulk copy of the pre-sliced Z image of the K-step (written by k_sf_tables) per stage;
//   MMA warp   : waits for A and B of the K-step, one elected lane issues its six MMAs and one tcgen05.commit that
//                frees the stage.  At the end of an item the generator warps read TMEM (tcgen05.ld), rescale and
//                write the partial amplitudes that the usual finalisers consume.
#pragma once
#include "sfactor_digits.cuh"

namespace mb200 {

constexpr int TC_M = 128;          // rows ((h,k) pairs) per tile
constexpr int TC_NSTG = 4;         // A / B stages (power of two: cheap stage / phase arithmetic)
#ifndef MB200_TC_GEN_WARPS
#define MB200_TC_GEN_WARPS 16
#endif
constexpr int TC_GEN_WARPS = MB200_TC_GEN_WARPS;
constexpr int TC_THREADS = (TC_GEN_WARPS + 3) * 32;  // generators + MMA warp + seed loader + B loader
constexpr int TC_W_MMA = TC_GEN_WARPS, TC_W_SEED = TC_GEN_WARPS + 1, TC_W_BLOAD = TC_GEN_WARPS + 2;
constexpr int TC_TASKS = 8;                // generator tasks per K-step: 4 row quads x 8 atom pairs (one warp pass) each
constexpr int TC_SEED_MAX = 8;             // seed ring: at most 8 K-steps in flight ...
constexpr int TC_SEED_BYTES = 40 * 1024;   // ... within this much shared memory (a stage is <= 67 rows x 256 B)
constexpr int TC_MAX_LT = 48;      // l values per tile: 5 * 2 * 48 = 480 TMEM columns
constexpr int TC_CHUNK = 8192;     // atoms per flush: 2 * 5 * 128 * 128 * 8192 < 2^31, the int32 accumulators cannot overflow

struct TcItem {
    int32_t seg, rt, lt, j0, j1, split;
};

struct TcSeg {
    const double2 *ex, *ey;     // FP64 phase tables (blocked layout)
    const int8_t *bimg;         // pre-sliced Z image [n_lt][K-step][5][NP x 32], written by k_sf_tables
    const uint8_t *rows;        // [n_rt * 128][2] packed (h, k) pairs
    double2 *amp;               // [n_split][n_rt * n_lt][128][L_t]
    int32_t rows0, rows1, ld, n_rt, n_lt, lt_size, n_pairs, n_split;
};


// K-major, no swizzle, descriptor version 1; LBO = distance of the two 16-byte K chunks, SBO = distance of 8-row groups
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo = 128, uint32_t sbo = 256) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// A operand (written by the generator warps): the 8 x 16 B core matrices are placed with SBO = 144 and LBO = 2336
// instead of the dense 256 / 128, which shifts neighbouring row groups by one and the second K chunk by two 16-byte
// slots: the 32 lanes of a warp store (2 row groups x 2 quads x 2 K chunks x 4 atom pairs) hit 32 distinct banks.
constexpr int TC_A_SBO = 144, TC_A_LBO = 2336, TC_A_SLICE = 4672;
__device__ __forceinline__ int tc_canon_a(int row, int k) { return (row >> 3) * TC_A_SBO + (k >> 4) * TC_A_LBO + (row & 7) * 16 + (k & 15); }

// ---------------------------------------------------------------------------------------------
// k_sf_contract_tc
// ---------------------------------------------------------------------------------------------
#ifdef MB200_TC_TRACE  // developer build: per-step clock64 stamps of CTA 0's first item (tools/tc_trace.py)
__device__ unsigned long long g_tc_trace[512 * 32];
#define TC_STAMP(slot) do { if (blockIdx.x == 0 && first_item && t < 512) g_tc_trace[t * 32 + (slot)] = clock64(); } while (0)
#else
#define TC_STAMP(slot) do { } while (0)
#endif
__device__ __forceinline__ uint32_t tc_idesc(int n) {  // s8 x s8 -> s32, K-major A and B, M = 128
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
// A operand from tensor memory (row = TMEM lane, four int8 along K per 32-bit column): tools/microbench/umma_i8_ts.cu
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
                 "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
// tcgen05.st.16x128b.x1: thread i of the warp writes r0 to (TMEM lane i / 4, column i % 4) and r1 to (lane i / 4 + 8,
// column i % 4) of the 16-lane block at `addr` (tools/microbench/tmem_st_layout.cu)
__device__ __forceinline__ void tc_st_16x128(uint32_t addr, uint32_t r0, uint32_t r1) {
    asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1, %2};\n" ::"r"(addr), "r"(r0), "r"(r1) : "memory");
}
constexpr int TC_ATM_COL0 = 320;       // A-in-TMEM variant: accumulators in columns 0 .. 5 np - 1 (np <= 64), A stages behind
constexpr int TC_ATM_MAX_LT = 32;

// One lane of a converged warp (the compiler knows a single thread runs the guarded code: operands move to
// uniform registers without a per-value loop).
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ATM = true (l tiles of at most 32 values, i.e. np <= 64 accumulator columns per digit group): the generated A operand
// goes to TENSOR MEMORY instead of shared memory -- tcgen05.st.16x128b from the generator registers, MMAs in the
// "[d], [a], b-desc" form -- which takes the digit stores and the tensor core's A reads (358 of the ~830 wavefronts of a
// K-step) off the shared-memory pipe.  Generator thread (L = lane / 4, pair = lane % 4 + 4 half) of the warp that owns
// TMEM lane quarter `warp % 4` forms the same 4 rows x 2 atoms as before; row 4 o + r of quad o = 8 quarter + L lives in
// TMEM lane 32 quarter + L + 8 r (the fragment layout of the store), and the epilogue undoes that permutation.
template <bool ATM>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_sf_contract_tc(const TcItem *__restrict__ items, int n_items, const TcSeg *__restrict__ segs, int *__restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    // [TC_NSTG] A stages (5 x 128 x 32 B) | [TC_NSTG] B stages (5 x NP_max x 32 B) | barriers | row table
    constexpr int A_STAGE = TC_S * TC_A_SLICE;
    constexpr int B_STAGE = TC_S * 2 * TC_MAX_LT * 32;
    int8_t *sA = reinterpret_cast<int8_t *>(smem);
    int8_t *sB = sA + TC_NSTG * A_STAGE;
    unsigned long long *a_full = reinterpret_cast<unsigned long long *>(sB + TC_NSTG * B_STAGE);  // [NSTG] generators -> MMA
    unsigned long long *b_full = a_full + TC_NSTG;                                               // [NSTG] bulk copy -> MMA
    unsigned long long *s_free = b_full + TC_NSTG;                                               // [NSTG] MMA done -> refill
    unsigned long long *acc_done = s_free + TC_NSTG;                                             // [1]
    unsigned long long *seed_full = acc_done + 1;                                                // [SEED_MAX] loader -> generators
    unsigned long long *seed_free = seed_full + TC_SEED_MAX;                                     // [SEED_MAX] generators -> loader
    uint8_t *s_rows = reinterpret_cast<uint8_t *>(seed_free + TC_SEED_MAX);                                     // [128][2]
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_rows + 2 * TC_M);
    int *s_item = reinterpret_cast<int *>(s_tmem + 1);
    int *s_nseed = s_item + 1;                                                                   // [2] unique h, unique k0
    int *s_task = s_nseed + 2;                                                                   // [4] generator task counters
    uint8_t *s_quad = reinterpret_cast<uint8_t *>(s_task + 4);                                  // [32][2] quad -> seed row (x, y)
    uint16_t *s_seedrow = reinterpret_cast<uint16_t *>(s_quad + 64);                             // [67] table row | axis << 15
    unsigned char *sSeed = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(s_seedrow + 68) + 127) & ~(uintptr_t)127);                  // [depth][n_seed rows][16 atoms] double2

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_gen = warp < TC_GEN_WARPS;

    if (warp == TC_W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *s_tmem;
    bool bars_live = false;
#ifdef MB200_TC_TRACE
    bool first_item = true;
#endif

    for (;;) {
        if (tid == 0) {
            *s_item = atomicAdd(counter, 1);
            s_task[0] = s_task[1] = s_task[2] = s_task[3] = 0;
            for (int s = 0; s < 3 * TC_NSTG + 1 + 2 * TC_SEED_MAX; ++s) {  // re-arm: every item starts at phase 0
                if (bars_live) asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(smem_u32(a_full + s)) : "memory");
                // a_full: one arrive per generator task; b_full / s_free / acc_done: 1; seed_full: one cp.async
                // arrive per loader lane; seed_free: one arrive per generator task
                const unsigned count = s < TC_NSTG ? TC_TASKS : s < 3 * TC_NSTG + 1 ? 1u : s < 3 * TC_NSTG + 1 + TC_SEED_MAX ? 32u :
                                       s < 3 * TC_NSTG + 1 + 2 * TC_SEED_MAX ? TC_TASKS : 1u;
                mbar_init(a_full + s, count);
            }
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        }
        bars_live = true;
        __syncthreads();
        const int item = *s_item;
        if (item >= n_items) break;
        const TcItem it = items[item];
        const TcSeg sg = segs[it.seg];  // by value: the loops below must not re-read the descriptor from global memory
        const int lt_size = sg.lt_size, np = 2 * lt_size;
        const int n_steps = (it.j1 - it.j0) / TC_KA;
        const uint32_t b_bytes = (uint32_t)(TC_S * np * 32);
        for (int i = tid; i < 2 * TC_M; i += TC_THREADS) s_rows[i] = sg.rows[(size_t)it.rt * 2 * TC_M + i];
        __syncthreads();
        // Seed rows of the item: quad o starts at (h0, k0) = rows[4 o]; the distinct h0 (Ex rows), the distinct k0
        // (Ey rows) and Ey[1] are what the generators need per atom -- the loader warp keeps them in a shared-memory
        // ring, TC_SEED_MAX K-steps ahead, so no generator ever waits on a global load.
        if (warp == TC_W_SEED) {
            const int h0 = s_rows[8 * lane], k0 = s_rows[8 * lane + 1];
            const unsigned mh = __match_any_sync(0xffffffffu, h0), mk = __match_any_sync(0xffffffffu, k0);
            const int lh = __ffs(mh) - 1, lk = __ffs(mk) - 1;  // leader lane of my value
            const unsigned leaders_h = __ballot_sync(0xffffffffu, lh == lane), leaders_k = __ballot_sync(0xffffffffu, lk == lane);
            const int ih = __popc(leaders_h & ((1u << lh) - 1u)), ik = __popc(leaders_k & ((1u << lk) - 1u));
            const int n_uh = __popc(leaders_h), n_uk = __popc(leaders_k);
            s_quad[2 * lane] = (uint8_t)ih;
            s_quad[2 * lane + 1] = (uint8_t)(n_uh + ik);
            if (lh == lane) s_seedrow[ih] = (uint16_t)h0;
            if (lk == lane) s_seedrow[n_uh + ik] = (uint16_t)(k0 | 0x8000);
            if (lane == 0) s_seedrow[n_uh + n_uk] = (uint16_t)(min(1, sg.rows1 - 1) | 0x8000);  // Ey[1]
            if (lane == 0) {
                s_nseed[0] = n_uh;
                s_nseed[1] = n_uk;
            }
        }
        __syncthreads();
        const int n_seed = s_nseed[0] + s_nseed[1] + 1;          // rows per seed stage
        const int seed_stage = n_seed * 256;                      // bytes: 16 atoms x double2 per row
        const int seed_shift = seed_stage * 8 <= TC_SEED_BYTES ? 3 : seed_stage * 4 <= TC_SEED_BYTES ? 2 : 1;  // 8, 4 or 2 K-steps
        const int seed_depth = 1 << seed_shift;

        if (is_gen && ATM) {
            // ---- generators, A in tensor memory: tasks (K-step t, column half) per TMEM lane quarter, handed out in
            // order from the quarter's counter to the four warps that can reach it ----
            const int qd = warp & 3, L = lane >> 2;
            const int o = 8 * qd + L;  // quad of the item (rows 4 o .. 4 o + 3)
            const int n_tasks = n_steps * 2;
            const int rx = s_quad[2 * o], ry = s_quad[2 * o + 1], re1 = n_seed - 1;
            for (;;) {
                int task = 0;
                if (lane == 0) task = atomicAdd(s_task + qd, 1);
                task = __shfl_sync(0xffffffffu, task, 0);
                if (task >= n_tasks) break;
                const int t = task >> 1, ch = task & 1;
                const int ap = (lane & 3) + 4 * ch;  // atom pair of the K-step = 32-bit column of the slice
                // seed rows are rotated by four slots on odd rows (see the loader): the two quads of a quarter-warp read
                // different bank halves when their rows differ in parity
                const int sx = (rx * 16 + ((ap + 4 * (rx & 1)) & 7)) * 16, sy = (ry * 16 + ((ap + 4 * (ry & 1)) & 7)) * 16,
                          se = (re1 * 16 + ((ap + 4 * (re1 & 1)) & 7)) * 16;
                const int st = t & (TC_NSTG - 1), ss = t & (seed_depth - 1);
                double2 e[3][2], p[2];
                mbar_wait(seed_full + ss, (unsigned)(t >> seed_shift) & 1u);
                if (lane == 0) TC_STAMP((2 * qd + ch) * 3);
                {
                    const unsigned char *sd = sSeed + ss * seed_stage;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const double2 x = *reinterpret_cast<const double2 *>(sd + sx + 128 * c);
                        const double2 y = *reinterpret_cast<const double2 *>(sd + sy + 128 * c);
                        const double2 e1 = *reinterpret_cast<const double2 *>(sd + se + 128 * c);
                        p[c].x = x.x * y.x - x.y * y.y;
                        p[c].y = x.x * y.y + x.y * y.x;
                        e[0][c] = e1;
                        e[1][c] = make_double2(e1.x * e1.x - e1.y * e1.y, 2.0 * e1.x * e1.y);
                        e[2][c] = make_double2(e[1][c].x * e1.x - e[1][c].y * e1.y, e[1][c].x * e1.y + e[1][c].y * e1.x);
                    }
                }
                if (t >= TC_NSTG) {
                    mbar_wait(s_free + st, (unsigned)((t / TC_NSTG) - 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                }
                if (lane == 0) TC_STAMP((2 * qd + ch) * 3 + 1);
                const uint32_t a_dst = tmem + ((uint32_t)(32 * qd) << 16) + (uint32_t)(TC_ATM_COL0 + st * (TC_S * 8) + 4 * ch);
#pragma unroll
                for (int half = 0; half < 2; ++half) {  // rows r = 0, 1 -> TMEM lanes L, L + 8; r = 2, 3 -> L + 16, L + 24
                    unsigned w[2][TC_S];
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int r = 2 * half + rr;
                        unsigned long long b[4];
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            double pr = p[c].x, pi = p[c].y;
                            if (r > 0) {
                                pr = p[c].x * e[r - 1][c].x - p[c].y * e[r - 1][c].y;
                                pi = p[c].x * e[r - 1][c].y + p[c].y * e[r - 1][c].x;
                            }
                            b[2 * c] = (unsigned long long)__double_as_longlong(pr + TC_MAGIC);
                            b[2 * c + 1] = (unsigned long long)__double_as_longlong(pi + TC_MAGIC);
                        }
                        const unsigned t0 = __byte_perm((unsigned)b[0], (unsigned)b[1], 0x5140), t1 = __byte_perm((unsigned)b[0], (unsigned)b[1], 0x7362),
                                       t2 = __byte_perm((unsigned)b[2], (unsigned)b[3], 0x5140), t3 = __byte_perm((unsigned)b[2], (unsigned)b[3], 0x7362);
                        const unsigned h0 = __byte_perm((unsigned)(b[0] >> 32), (unsigned)(b[1] >> 32), 0x0040),
                                       h1 = __byte_perm((unsigned)(b[2] >> 32), (unsigned)(b[3] >> 32), 0x0040);
                        w[rr][0] = __byte_perm(h0, h1, 0x5410) ^ 0x80808080u;
                        w[rr][1] = __byte_perm(t1, t3, 0x7632) ^ 0x80808080u;
                        w[rr][2] = __byte_perm(t1, t3, 0x5410) ^ 0x80808080u;
                        w[rr][3] = __byte_perm(t0, t2, 0x7632) ^ 0x80808080u;
                        w[rr][4] = __byte_perm(t0, t2, 0x5410) ^ 0x80808080u;
                    }
#pragma unroll
                    for (int s = 0; s < TC_S; ++s) tc_st_16x128(a_dst + ((uint32_t)(16 * half) << 16) + (uint32_t)(8 * s), w[0][s], w[1][s]);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(a_full + st);
                    mbar_arrive(seed_free + ss);
                    TC_STAMP((2 * qd + ch) * 3 + 2);
                }
            }
        } else if (is_gen) {
            // ---- generators: thread (ap, o) -> atoms 2 ap, 2 ap + 1 of the K-step, rows 4 o .. 4 o + 3 ----
            // The plan pads every run of consecutive k to a multiple of four rows, so a quad is always one run
            // (h0, k0 .. k0 + 3):  P_r = (w Ex[h0] Ey[k0]) Ey[r], four independent products per atom (no recurrence
            // chain: eight independent digit pipelines per thread).  Padding rows are computed and never referenced.
            // Tasks (K-step t, pass j = 4 row quads x 8 atom pairs) are handed out in order from a shared counter:
            // the warps of a scheduler that is held up (the MMA warp's UTCIMMA issue stalls the FP64 instructions
            // of its scheduler) simply take fewer of them.
            const int ap = lane & 7;
            const int n_tasks = n_steps * TC_TASKS;
            for (;;) {
                int task = 0;
                if (lane == 0) task = atomicAdd(s_task, 1);
                task = __shfl_sync(0xffffffffu, task, 0);
                if (task >= n_tasks) break;
                const int t = task >> 3, o = 4 * (task & 7) + (lane >> 3);
                const int off0 = tc_canon_a(4 * o, 4 * ap);  // K' = 4 ap .. 4 ap + 3; rows 4 o + r: + 16 r bytes
                // seeds Ex[h0], Ey[k0] and the factors Ey[1..3] of my two atoms, from the seed ring
                const int sx = (s_quad[2 * o] * 16 + ap) * 16, sy = (s_quad[2 * o + 1] * 16 + ap) * 16, se = ((n_seed - 1) * 16 + ap) * 16;
                const int st = t & (TC_NSTG - 1), ss = t & (seed_depth - 1);
                double2 e[3][2], p[2];
                mbar_wait(seed_full + ss, (unsigned)(t >> seed_shift) & 1u);
                if (lane == 0) TC_STAMP((task & 7) * 3);
                {
                    const unsigned char *sd = sSeed + ss * seed_stage;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {  // atom 2 ap + c sits at slot 8 c + ap of its seed row
                        const double2 x = *reinterpret_cast<const double2 *>(sd + sx + 128 * c);
                        const double2 y = *reinterpret_cast<const double2 *>(sd + sy + 128 * c);
                        const double2 e1 = *reinterpret_cast<const double2 *>(sd + se + 128 * c);
                        p[c].x = x.x * y.x - x.y * y.y;
                        p[c].y = x.x * y.y + x.y * y.x;
                        e[0][c] = e1;  // Ey[1], Ey[2] = Ey[1]^2, Ey[3] = Ey[1]^3 (unit-modulus table entries)
                        e[1][c] = make_double2(e1.x * e1.x - e1.y * e1.y, 2.0 * e1.x * e1.y);
                        e[2][c] = make_double2(e[1][c].x * e1.x - e[1][c].y * e1.y, e[1][c].x * e1.y + e[1][c].y * e1.x);
                    }
                }
                if (t >= TC_NSTG) mbar_wait(s_free + st, (unsigned)((t / TC_NSTG) - 1) & 1u);
                if (lane == 0) TC_STAMP((task & 7) * 3 + 1);
                int8_t *dstA = sA + st * A_STAGE + off0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    // biased fixed point of (re0, im0, re1, im1): the low 40 mantissa bits hold the five digits + 128
                    unsigned long long b[4];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        double pr = p[c].x, pi = p[c].y;
                        if (r > 0) {
                            pr = p[c].x * e[r - 1][c].x - p[c].y * e[r - 1][c].y;
                            pi = p[c].x * e[r - 1][c].y + p[c].y * e[r - 1][c].x;
                        }
                        b[2 * c] = (unsigned long long)__double_as_longlong(pr + TC_MAGIC);
                        b[2 * c + 1] = (unsigned long long)__double_as_longlong(pi + TC_MAGIC);
                    }
                    // 4 x 4 byte transpose of the low words: word i = byte i of (re0, im0, re1, im1) = slice 5 - i
                    const unsigned t0 = __byte_perm((unsigned)b[0], (unsigned)b[1], 0x5140), t1 = __byte_perm((unsigned)b[0], (unsigned)b[1], 0x7362),
                                   t2 = __byte_perm((unsigned)b[2], (unsigned)b[3], 0x5140), t3 = __byte_perm((unsigned)b[2], (unsigned)b[3], 0x7362);
                    const unsigned h0 = __byte_perm((unsigned)(b[0] >> 32), (unsigned)(b[1] >> 32), 0x0040),
                                   h1 = __byte_perm((unsigned)(b[2] >> 32), (unsigned)(b[3] >> 32), 0x0040);
                    unsigned w[TC_S];
                    w[0] = __byte_perm(h0, h1, 0x5410);  // slice 1 (most significant digit)
                    w[1] = __byte_perm(t1, t3, 0x7632);  // slice 2 = byte 3
                    w[2] = __byte_perm(t1, t3, 0x5410);  // slice 3 = byte 2
                    w[3] = __byte_perm(t0, t2, 0x7632);  // slice 4 = byte 1
                    w[4] = __byte_perm(t0, t2, 0x5410);  // slice 5 = byte 0
#pragma unroll
                    for (int s = 0; s < TC_S; ++s)       // xor 0x80: biased byte -> balanced int8 digit
                        *reinterpret_cast<uint32_t *>(dstA + s * TC_A_SLICE + 16 * r) = w[s] ^ 0x80808080u;
                }
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(a_full + st);
                    mbar_arrive(seed_free + ss);  // the seeds were consumed long ago: the loader may refill the slot
                    TC_STAMP((task & 7) * 3 + 2);
                }
            }
        } else if (warp == TC_W_SEED) {
            // ---- seed loader: 16-byte cp.async gathers of the item's seed rows, seed_depth K-steps ahead ----
            // lane -> atom a = lane & 15 of the K-step and seed rows (lane >> 4) + 2 i; source pointers of the first
            // eight rows per lane live in registers and advance by a constant per K-step
            const int a = lane & 15, r0 = lane >> 4;
            const char *ex = reinterpret_cast<const char *>(sg.ex), *ey = reinterpret_cast<const char *>(sg.ey);
            const unsigned step_x = (unsigned)sg.rows0 * 256u, step_y = (unsigned)sg.rows1 * 256u;  // two 8-atom blocks per K-step
            auto seed_src = [&](int ri) -> const char * {
                const int code = s_seedrow[ri], row = code & 0x7fff;
                return (code & 0x8000) ? ey + tab_index(it.j0 + a, row, sg.rows1) * 16 : ex + tab_index(it.j0 + a, row, sg.rows0) * 16;
            };
            const char *src[8];
            unsigned stp[8];
            const int n_mine = (n_seed - r0 + 1) >> 1;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ri = min(r0 + 2 * i, n_seed - 1);
                src[i] = seed_src(ri);
                stp[i] = (s_seedrow[ri] & 0x8000) ? step_y : step_x;
            }
            // row layout [atom parity][atom pair]: conflict-free reads; ATM: pairs rotated by four slots on odd rows
            const int dst0 = (r0 * 16 + (a & 1) * 8 + (ATM ? (((a >> 1) + 4 * r0) & 7) : (a >> 1))) * 16;
            for (int t = 0; t < n_steps; ++t) {
                const int ss = t & (seed_depth - 1);
                if (t >= seed_depth) mbar_wait(seed_free + ss, (unsigned)((t >> seed_shift) - 1) & 1u);
                unsigned char *dst = sSeed + ss * seed_stage + dst0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < n_mine) cp_async16(dst + i * 512, src[i]);
                    src[i] += stp[i];
                }
                for (int i = 8; i < n_mine; ++i) {  // more than 16 seed rows: the rest are located on the fly
                    const int ri = r0 + 2 * i;
                    cp_async16(dst + i * 512, seed_src(ri) + (size_t)t * ((s_seedrow[ri] & 0x8000) ? step_y : step_x));
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(seed_full + ss)) : "memory");
            }
        } else if (warp == TC_W_BLOAD) {
            // ---- B loader: one bulk copy of the pre-sliced Z image per K-step, as soon as the stage is free ----
            // (the whole warp walks the loop and waits; one elected lane issues the copy)
            const int8_t *bsrc = sg.bimg + ((size_t)it.lt * (sg.ld / TC_KA) + it.j0 / TC_KA) * (size_t)b_bytes;
            for (int t = 0; t < n_steps; ++t) {
                const int st = t & (TC_NSTG - 1);
                if (t >= TC_NSTG) mbar_wait(s_free + st, (unsigned)((t / TC_NSTG) - 1) & 1u);
                if (tc_elect_one()) {
                    mbar_arrive_expect_tx(b_full + st, b_bytes);
                    bulk_g2s(sB + st * B_STAGE, bsrc + (size_t)t * b_bytes, b_bytes, b_full + st);
                }
                __syncwarp();
            }
        } else if (warp == TC_W_MMA) {
            // ---- MMA issue: the (A slice s) x (B slices t_lo .. t_hi) products of a K-step; everything that does not
            // depend on the stage is tabulated once per item so that the serial issue path is a handful of adds.  The
            // whole warp walks the loop; one elected lane issues (the compiler then keeps the operands in uniform
            // registers).  One commit per K-step and one issuing warp: per-MMA commits, MMAs dealt to four warps and even /
            // odd K-steps on two warps were all measured slower or equal (profiles/r01_tc_notes.md). ----
            const int per = max(1, 256 / np);  // B slices per MMA (N <= 256)
            uint32_t m_a[15], m_b[15], m_d[15], m_i[15];
            int n_mma = 0;
            {
                int s = 1, t_lo = 1;
#pragma unroll
                for (int i = 0; i < 15; ++i) {
                    m_a[i] = m_b[i] = m_d[i] = m_i[i] = 0;
                    if (s <= TC_S) {
                        const int t_hi = min(6 - s, t_lo + per - 1);
                        m_a[i] = ATM ? (uint32_t)((s - 1) * 8) : (uint32_t)((s - 1) * TC_A_SLICE) >> 4;
                        m_b[i] = (uint32_t)((t_lo - 1) * np * 32) >> 4;
                        m_d[i] = (uint32_t)((s + t_lo - 2) * np);          // column block g = s + t
                        m_i[i] = tc_idesc((t_hi - t_lo + 1) * np) | (s > 1 ? 1u : 0u);  // bit 0 (masked off at issue):
                        n_mma = i + 1;                                     // accumulate at t = 0
                        t_lo = t_hi + 1;
                        if (t_lo > 6 - s) { ++s; t_lo = 1; }
                    }
                }
            }
            const uint64_t da0 = tc_desc(smem_u32(sA), TC_A_LBO, TC_A_SBO), db0 = tc_desc(smem_u32(sB));
            for (int t = 0; t < n_steps; ++t) {
                const int st = t & (TC_NSTG - 1);
                const unsigned par = (unsigned)(t / TC_NSTG) & 1u;
                mbar_wait(b_full + st, par);
                mbar_wait(a_full + st, par);
                TC_STAMP(24);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint64_t da = da0 + (uint64_t)(st * (A_STAGE >> 4)), db = db0 + (uint64_t)(st * (B_STAGE >> 4));
                if (tc_elect_one()) {
#pragma unroll
                    for (int i = 0; i < 15; ++i)
                        if (i < n_mma) {
                            if (ATM)
                                tc_mma_ts(tmem + m_d[i], tmem + (uint32_t)(TC_ATM_COL0 + st * (TC_S * 8)) + m_a[i], db + m_b[i], m_i[i] & ~1u,
                                          t > 0 ? 1u : (m_i[i] & 1u));
                            else
                                tc_mma(tmem + m_d[i], da + m_a[i], db + m_b[i], m_i[i] & ~1u, t > 0 ? 1u : (m_i[i] & 1u));
                        }
                    tc_commit(s_free + st);  // frees the A and B stage when these MMAs have read them
                }
                __syncwarp();
                TC_STAMP(26);
            }
            if (tc_elect_one()) tc_commit(acc_done);
            __syncwarp();
        }

        // ---- epilogue: TMEM -> FP64 partial amplitudes (generator warps; warp w reads lanes 32 (w % 4) ..) ----
        if (is_gen) {
            mbar_wait(acc_done, 0u);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            // ATM: TMEM lane 32 q + L + 8 r holds row 32 q + 4 L + r
            const int row = 32 * (warp & 3) + (ATM ? 4 * (lane & 7) + (lane >> 3) : lane);
            const int n_tiles = sg.n_rt * sg.n_lt;
            double2 *out = sg.amp + (((size_t)it.split * n_tiles + (size_t)it.rt * sg.n_lt + it.lt) * TC_M + row) * lt_size;
            // column chunks of 16 (= 8 l values), dealt round robin to the four warps that share a lane quarter
            for (int c0 = 16 * (warp >> 2); c0 < np; c0 += 16 * (TC_GEN_WARPS / 4)) {
                double acc[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[c] = 0.0;
#pragma unroll
                for (int g = 2; g <= 6; ++g) {
                    uint32_t v[16];
                    const uint32_t addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((g - 2) * np + c0);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(addr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                    const double scale = ldexp(1.0, 4 - 8 * g);
#pragma unroll
                    for (int c = 0; c < 16; ++c) acc[c] += (double)(int32_t)v[c] * scale;
                }
#pragma unroll
                for (int c = 0; c < 16; c += 2) out[(c0 + c) >> 1] = make_double2(acc[c], acc[c + 1]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        }
        __syncthreads();  // TMEM and the rings are free for the next item
#ifdef MB200_TC_TRACE
        first_item = false;
#endif
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == TC_W_MMA)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

constexpr int TC_SMEM = 16 + TC_NSTG * (TC_S * TC_A_SLICE + TC_S * 2 * TC_MAX_LT * 32) + (3 * TC_NSTG + 1 + 2 * TC_SEED_MAX) * 8 + 2 * TC_M + 16 + 16 + 64 + 136 +
                        128 + TC_SEED_BYTES + 1024;

}  // namespace mb200
