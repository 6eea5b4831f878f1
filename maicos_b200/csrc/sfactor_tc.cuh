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
//   1e5 atoms, i.e. ~1e-11 relative on binned S(q).  Single vanishing speckle cells of the raw cube can still
//   exceed a relative 1e-6, so this is a separately selected mode (mb200_ctx_set_precision); the FP64 DMMA
//   kernel stays the parity default and the only path of the raw-cube API.
//
// CTA = 8 generator warps + 1 control warp, one CTA per SM (512 TMEM columns):
//   generators : thread (atom pair of the K-step, row quad o) walks 4 packed (h,k) rows for two atoms (two
//                independent dependency chains, 32-bit stores); P = w Ex[h] Ey[k] is
//                seeded from the FP64 tables at the start of a run of consecutive k and advanced by the
//                recurrence P *= Ey[1] inside the run; digits are written as (re, im) byte pairs into the
//                canonical K-major no-swizzle UMMA layout (cute ((8,n),2):((1,SBO),LBO), LBO = 128, SBO = 256);
//   control    : lane 0 bulk-copies the pre-sliced Z image of the K-step (k_sf_bimage) and issues the MMAs;
//                tcgen05.commit frees the stage.  At the end of an item the generator warps read TMEM
//                (tcgen05.ld), rescale and write the partial amplitudes that the usual finalisers consume.
#pragma once

namespace mb200 {

constexpr int TC_M = 128;          // rows ((h,k) pairs) per tile
constexpr int TC_KA = 16;          // atoms per MMA K-step
constexpr int TC_S = 5;            // digits per component
constexpr int TC_NSTG = 4;         // A / B stages (power of two: cheap stage / phase arithmetic)
constexpr int TC_GEN_WARPS = 8;
constexpr int TC_THREADS = (TC_GEN_WARPS + 1) * 32;
constexpr int TC_MAX_LT = 48;      // l values per tile: 5 * 2 * 48 = 480 TMEM columns
constexpr int TC_CHUNK = 8192;     // atoms per flush: 2 * 5 * 128 * 128 * 8192 < 2^31, the int32 accumulators cannot overflow

struct TcItem {
    int32_t seg, rt, lt, j0, j1, split;
};

struct TcSeg {
    const double2 *ex, *ey;     // FP64 phase tables (blocked layout)
    const int8_t *bimg;         // pre-sliced Z image [n_lt][K-step][5][NP x 32]
    const uint8_t *rows;        // [n_rt * 128][2] packed (h, k) pairs
    double2 *amp;               // [n_split][n_rt * n_lt][128][L_t]
    int32_t rows0, rows1, ld, n_rt, n_lt, lt_size, n_pairs, n_split;
};

__device__ __host__ __forceinline__ int tc_canon(int row, int k) { return (row >> 3) * 256 + (k >> 4) * 128 + (row & 7) * 16 + (k & 15); }

// K-major, no swizzle, descriptor version 1; LBO = distance of the two 16-byte K chunks, SBO = distance of 8-row groups
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo = 128, uint32_t sbo = 256) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// A operand (written by the generator warps): the 8 x 16 B core matrices are placed with SBO = 144 and LBO = 2336
// instead of the dense 256 / 128, which shifts neighbouring row groups by one and the second K chunk by two 16-byte
// slots: the 32 lanes of a warp store (2 row groups x 2 quads x 2 K chunks x 4 atom pairs) hit 32 distinct banks.
constexpr int TC_A_SBO = 144, TC_A_LBO = 2336, TC_A_SLICE = 4672;
__device__ __forceinline__ int tc_canon_a(int row, int k) { return (row >> 3) * TC_A_SBO + (k >> 4) * TC_A_LBO + (row & 7) * 16 + (k & 15); }

// Digits: y = x + TC_MAGIC lies in [2^14, 2^15) (ulp 2^-38), so the low 40 mantissa bits of y are
// U = rint(x * 2^38) + 128 * (256^4 + 256^3 + 256^2 + 256 + 1); byte i of U xor 0x80 is the balanced digit of
// weight 256^i.  Returned: lo = digits 5, 4, 3, 2 (bytes 0..3), hi = digit 1 (byte 0), already re-centred.
#define TC_MAGIC (16384.0 + 551911719040.0 / 274877906944.0)
__device__ __forceinline__ void tc_digits(double x, unsigned &lo, unsigned &hi) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x + TC_MAGIC);
    lo = (unsigned)b ^ 0x80808080u;
    hi = ((unsigned)(b >> 32) ^ 0x80u) & 0xffu;
}
// (re, im) digit pair of slice s (1 = most significant) as a 16-bit value, re in the low byte
__device__ __forceinline__ unsigned tc_pair(int s, unsigned rlo, unsigned rhi, unsigned ilo, unsigned ihi) {
    switch (s) {
        case 1: return __byte_perm(rhi, ihi, 0x0040);
        case 2: return __byte_perm(rlo, ilo, 0x0073);
        case 3: return __byte_perm(rlo, ilo, 0x0062);
        case 4: return __byte_perm(rlo, ilo, 0x0051);
        default: return __byte_perm(rlo, ilo, 0x0040);
    }
}

// ---------------------------------------------------------------------------------------------
// k_sf_bimage: Z = Ez rows of an l tile -> int8 digit image in the UMMA smem layout, per K-step
//   grid (K-steps, n_lt, segments); thread = (l, atom) pairs of the K-step
// ---------------------------------------------------------------------------------------------
struct BimgJob {
    const double2 *ez;
    int8_t *bimg;
    int32_t rows2, ld, maxn2, n_lt, lt_size;
};

__global__ void __launch_bounds__(256) k_sf_bimage(const BimgJob *__restrict__ jobs) {
    const BimgJob &jb = jobs[blockIdx.z];
    const int ks = blockIdx.x, lt = blockIdx.y;
    if (ks * TC_KA >= jb.ld || lt >= jb.n_lt) return;
    const int np = 2 * jb.lt_size;
    int8_t *out = jb.bimg + ((size_t)lt * (jb.ld / TC_KA) + ks) * (size_t)(TC_S * np * 32);
    for (int e = threadIdx.x; e < jb.lt_size * TC_KA; e += blockDim.x) {
        const int li = e / TC_KA, a = e % TC_KA;
        const int l = lt * jb.lt_size + li;
        double2 z = make_double2(0.0, 0.0);
        if (l < jb.maxn2) z = jb.ez[tab_index(ks * TC_KA + a, l, jb.rows2)];
        unsigned rlo, rhi, ilo, ihi, nlo, nhi;
        tc_digits(z.x, rlo, rhi);
        tc_digits(z.y, ilo, ihi);
        tc_digits(-z.y, nlo, nhi);
#pragma unroll
        for (int t = 1; t <= TC_S; ++t) {
            int8_t *b = out + (size_t)(t - 1) * np * 32;
            // row 2l (real part of A): (Zr, -Zi); row 2l + 1 (imaginary part): (Zi, Zr)
            *reinterpret_cast<uint16_t *>(b + tc_canon(2 * li, 2 * a)) = (uint16_t)tc_pair(t, rlo, rhi, nlo, nhi);
            *reinterpret_cast<uint16_t *>(b + tc_canon(2 * li + 1, 2 * a)) = (uint16_t)tc_pair(t, ilo, ihi, rlo, rhi);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_sf_contract_tc
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t da, uint64_t db, int n, uint32_t acc) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

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
    uint8_t *s_rows = reinterpret_cast<uint8_t *>(acc_done + 1);                                 // [128][2]
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_rows + 2 * TC_M);
    int *s_item = reinterpret_cast<int *>(s_tmem + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_gen = warp < TC_GEN_WARPS;

    if (warp == TC_GEN_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *s_tmem;
    bool bars_live = false;

    for (;;) {
        if (tid == 0) {
            *s_item = atomicAdd(counter, 1);
            for (int s = 0; s < 3 * TC_NSTG + 1; ++s) {  // re-arm: every item starts at phase 0
                if (bars_live) asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(smem_u32(a_full + s)) : "memory");
                const unsigned count = s < TC_NSTG ? TC_GEN_WARPS : 1u;
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

        if (is_gen) {
            // ---- generators: thread (ap, o) -> atoms 2 ap, 2 ap + 1 of the K-step, rows 4 o .. 4 o + 3 ----
            // The plan pads every run of consecutive k to a multiple of four rows, so a quad is always one run:
            // row 0 is seeded from the FP64 tables, rows 1..3 follow by P *= Ey[1].  Padding rows are computed
            // and never referenced.
            const int ap = tid & 7, o = tid >> 3;
            const int h0 = s_rows[2 * (4 * o)], k0 = s_rows[2 * (4 * o) + 1];
            const int off0 = tc_canon_a(4 * o, 4 * ap);  // K' = 4 ap .. 4 ap + 3; rows 4 o + r: + 16 r bytes
            // seeds of row 0 and the k-step factor are prefetched one K-step ahead; their table addresses advance
            // by a constant stride per K-step (two 8-atom blocks)
            const int row1 = 1 < sg.rows1 ? 1 : 0;
            const double2 *px[2], *py[2], *py1[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int jf = it.j0 + 2 * ap + c;
                px[c] = sg.ex + tab_index(jf, h0, sg.rows0);
                py[c] = sg.ey + tab_index(jf, k0, sg.rows1);
                py1[c] = sg.ey + tab_index(jf, row1, sg.rows1);
            }
            const size_t sx = (size_t)sg.rows0 * 16, sy = (size_t)sg.rows1 * 16;
            double2 nx[2], ny[2], ny1[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) { nx[c] = *px[c]; ny[c] = *py[c]; ny1[c] = *py1[c]; }
            for (int t = 0; t < n_steps; ++t) {
                const int st = t & (TC_NSTG - 1);
                double2 y1[2], p[2];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    y1[c] = ny1[c];
                    p[c].x = nx[c].x * ny[c].x - nx[c].y * ny[c].y;
                    p[c].y = nx[c].x * ny[c].y + nx[c].y * ny[c].x;
                }
                if (t + 1 < n_steps) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        px[c] += sx; py[c] += sy; py1[c] += sy;
                        nx[c] = *px[c]; ny[c] = *py[c]; ny1[c] = *py1[c];
                    }
                }
                if (t >= TC_NSTG) mbar_wait(s_free + st, (unsigned)((t / TC_NSTG) - 1) & 1u);
                int8_t *dstA = sA + st * A_STAGE + off0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (r > 0) {
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const double nr = p[c].x * y1[c].x - p[c].y * y1[c].y, ni = p[c].x * y1[c].y + p[c].y * y1[c].x;
                            p[c].x = nr; p[c].y = ni;
                        }
                    }
                    unsigned lo[2][2], hi[2][2];  // [atom][re / im]
#pragma unroll
                    for (int c = 0; c < 2; ++c) { tc_digits(p[c].x, lo[c][0], hi[c][0]); tc_digits(p[c].y, lo[c][1], hi[c][1]); }
#pragma unroll
                    for (int s = 0; s < TC_S; ++s) {
                        const unsigned w0 = tc_pair(s + 1, lo[0][0], hi[0][0], lo[0][1], hi[0][1]);
                        const unsigned w1 = tc_pair(s + 1, lo[1][0], hi[1][0], lo[1][1], hi[1][1]);
                        *reinterpret_cast<uint32_t *>(dstA + s * TC_A_SLICE + 16 * r) = __byte_perm(w0, w1, 0x5410);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full + st);
            }
        } else if (lane == 0) {
            // ---- control: B image loads + MMA issue ----
            const int8_t *bsrc = sg.bimg + ((size_t)it.lt * (sg.ld / TC_KA) + it.j0 / TC_KA) * (size_t)b_bytes;
            auto load_b = [&](int t) {
                if (t < n_steps) {
                    const int st = t % TC_NSTG;
                    if (t >= TC_NSTG) mbar_wait(s_free + st, (unsigned)((t / TC_NSTG) - 1) & 1u);
                    mbar_arrive_expect_tx(b_full + st, b_bytes);
                    bulk_g2s(sB + st * B_STAGE, bsrc + (size_t)t * b_bytes, b_bytes, b_full + st);
                }
            };
            load_b(0);
            load_b(1);
            for (int t = 0; t < n_steps; ++t) {
                const int st = t % TC_NSTG;
                const unsigned par = (unsigned)(t / TC_NSTG) & 1u;
                mbar_wait(a_full + st, par);
                mbar_wait(b_full + st, par);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t a_base = smem_u32(sA + st * A_STAGE), b_base = smem_u32(sB + st * B_STAGE);
                const int per = max(1, 256 / np);  // B slices per MMA (N <= 256)
#pragma unroll 1
                for (int s = 1; s <= TC_S; ++s) {
                    const uint64_t da = tc_desc(a_base + (uint32_t)((s - 1) * TC_A_SLICE), TC_A_LBO, TC_A_SBO);
                    for (int t_lo = 1; t_lo <= 6 - s; t_lo += per) {
                        const int t_hi = min(6 - s, t_lo + per - 1);
                        const uint64_t db = tc_desc(b_base + (uint32_t)((t_lo - 1) * np * 32));
                        // column blocks g = s + t; on the first K-step the s = 1 MMAs initialise every block
                        tc_mma(tmem + (uint32_t)((s + t_lo - 2) * np), da, db, (t_hi - t_lo + 1) * np, (t > 0 || s > 1) ? 1u : 0u);
                    }
                }
                tc_commit(s_free + st);  // frees the A and B stage when these MMAs have read them
                load_b(t + 2);           // after the issue: waiting for step t-1's stage overlaps with step t's MMAs
            }
            tc_commit(acc_done);
        }

        // ---- epilogue: TMEM -> FP64 partial amplitudes (generator warps; warp w reads lanes 32 (w % 4) ..) ----
        if (is_gen) {
            mbar_wait(acc_done, 0u);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int row = 32 * (warp & 3) + lane;
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
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == TC_GEN_WARPS)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}

constexpr int TC_SMEM = TC_NSTG * (TC_S * TC_A_SLICE + TC_S * 2 * TC_MAX_LT * 32) + (3 * TC_NSTG + 1) * 8 + 2 * TC_M + 16 + 1024;

}  // namespace mb200
