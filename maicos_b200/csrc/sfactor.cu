// S(q) engine of libmaicos_b200 (sm_100a).
//
// Replaces the reference's only native kernel, `_cmath.structure_factor`
// (/root/reference/src/maicos/lib/_cmath.pyx:21-126), and the per-frame bodies built on it
// (modules/saxs.py:175-241, modules/diporderstructurefactor.py:93-155).
//
// Formulation.  With fractional coordinates u = r/L the phase of Miller vector (h,k,l) on atom
// j factorises:  exp(i q.r_j) = Ex_j[h] * Ey_j[k] * Ez_j[l],  E*_j[n] = exp(2 pi i n u*_j).
// The complex amplitudes  A[(h,k), l] = sum_j (w_j Ex_j[h] Ey_j[k]) * Ez_j[l]  are therefore a
// dense complex contraction over atoms (M = (h,k) pairs, N = l, K = atoms).  It is evaluated
// in FP64 on the DMMA pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; measured 37.0 TFLOP/s on
// B200 vs 33 TFLOP/s for scalar DFMA, tools/microbench) with the 3-multiplication complex
// product:  T1 = sum Pr*Zr, T2 = sum Pi*Zi, T3 = sum (Pr+Pi)*(Zr+Zi),
//           Re A = T1 - T2, Im A = T3 - T1 - T2     (3 real MACs per atom*q instead of 4).
//
// Kernels
//   k_sf_tables  : per-atom per-axis phase tables Ex (weighted), Ey, Ez by complex recurrence,
//                  re-seeded with an exact sincospi every 8 steps.  [row][atom] double2.
//   k_sf_contract: one CTA (4 warps) per (segment, atom range, tile group).  cp.async stages
//                  8-atom slices of the needed table rows into shared memory (3-deep), the CTA
//                  forms the P = Ex*Ey rows once per slice (3 planes), every warp owns a
//                  32 (hk) x 16 (l) tile = 4x2 DMMA tiles x 3 planes held in registers.
//                  8x8 DMMA tiles entirely outside the q-sphere are skipped (warp-uniform mask).
//   k_sf_svalues / k_sf_binned: sum the per-split partial amplitudes in fixed order, S = |A|^2,
//                  then either the compact S list (dense cube is scattered on the host) or the
//                  |q| histograms (S, f^2 S, count of S != 0) with one CTA per (segment, bin)
//                  and a fixed-order reduction -> bitwise reproducible run to run.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <map>

#include "common.cuh"
#include <array>
#include "sfactor_digits.cuh"

namespace mb200 {

// ------------------------------------------------------------------------------------------
// tiling constants
// ------------------------------------------------------------------------------------------
constexpr int KT = 8;          // atoms per pipeline stage (two DMMA k-steps of 4)
constexpr int WPC = 4;         // warps per CTA
constexpr int NTHREADS = WPC * 32;
#ifndef MB_NSTAGE
#define MB_NSTAGE 5
#endif
#ifndef MB_LOOKAHEAD
#define MB_LOOKAHEAD (MB_NSTAGE - 2)
#endif
constexpr int NSTAGE = MB_NSTAGE;        // staging ring depth
constexpr int LOOKAHEAD = MB_LOOKAHEAD;  // stages in flight ahead of the one being consumed
constexpr int MAXX = 4;        // distinct (h,k) quads (4h x 8k = 32 rows) a CTA may stage
constexpr int MAXY = 4;        // distinct l-pairs (16 l) a CTA may stage
constexpr int QH = 4, QK = 8;  // quad shape
constexpr int NL = 16;         // l extent of a warp tile
constexpr int RAWROW = 8;      // double2 per staged table row: 8 atoms; odd rows are rotated by one atom (bank spread)
constexpr int XY_ROWS = MAXX * (QH + QK);  // 48
constexpr int Z_ROWS = MAXY * NL;          // 64
constexpr int ST_ROWS = XY_ROWS + Z_ROWS;  // 112 rows per stage
constexpr int SMEM_STAGE = ST_ROWS * RAWROW * 16;       // 16128
constexpr int SMEM_BAR = 2 * NSTAGE * 8;                // full + empty mbarriers
constexpr int SMEM_TOTAL = NSTAGE * SMEM_STAGE + SMEM_BAR + 16;

struct __align__(16) TileGroupDesc {
    int16_t xh0[MAXX];
    int16_t xk0[MAXX];
    int16_t yl0[MAXY];
    uint8_t wx[WPC];
    uint8_t wy[WPC];
    uint8_t wmask[WPC];
    int32_t n_warp, n_x, n_y, wt0;
    int32_t pad[3];
};
static_assert(sizeof(TileGroupDesc) == 64, "TileGroupDesc layout");

struct SegDesc {
    const double2 *ex, *ey, *ez;   // phase tables [row][ld]
    double2 *amp;                  // partial amplitudes [n_split][n_wt][512]
    const TileGroupDesc *tgs;      // plan tile groups (device)
    const int32_t *ent_amp;        // plan entries -> index into a [n_wt][512] amplitude block
    const int32_t *bin_start;      // [n_bins+1] entry ranges per |q| bin (entries sorted by bin)
    const double *ff2;             // [n_valid] squared form factor per entry (may be null)
    int32_t ld, n_split, n_wt, n_valid;
    int32_t rows[3];               // table rows per axis (block stride = rows * 8 double2)
    int32_t pad_;
    long long amp_stride;          // double2 per split in `amp`
    double s_scale;                // S = |A|^2 * s_scale: undoes the power-of-two rescaling of the weights (exact)
    // q-block sharding of the tensor-core path: amplitude tiles (tile_elems double2 each) this rank did not compute read as
    // zero without ever being cleared or folded (null: every tile is owned)
    const uint8_t *tile_owned;
    long long tile_elems;
};

struct ItemDesc {
    int32_t seg, tg, j0, j1, split;
};

struct TableJob {
    const void *pos;       // float32 or float64 coordinates
    const int32_t *index;  // optional gather indices
    const double *weights; // optional weights (null = 1)
    double2 *tab[3];       // outputs (null = skip axis)
    int64_t stride_atom, stride_xyz;
    double box[3];
    float boxf[3];
    int32_t n, ld, rows[3];
    int32_t is_f32, wrap_mode;  // wrap_mode: 0 none, 1 saxs float32 wrap, 2 pack + saxs wrap
    // split-precision mode: the z axis is written as the int8 digit image of the tensor-core kernel instead of an FP64
    // table (sfactor_tc.cuh: [l tile][K-step][5 slices][2 L_t rows x 32 B], UMMA canonical K-major layout)
    int8_t *bimg;
    int32_t img_n_lt, img_lt, img_maxn;
    // split-precision mode: the generators read only the seed rows of the y table (Ey[k0] of the row quads and Ey[1]);
    // bit r of y_mask = row r is written (all rows when use_y_mask == 0).  The table kernel is HBM-write bound.
    int32_t use_y_mask, pad_;
    unsigned long long y_mask[4];
    // weights are multiplied by this exact power of two before they enter the tables (split-precision mode works on
    // entries in [-1, 1] with an ABSOLUTE quantisation of 2^-39: max|w| is brought into (0.5, 1])
    double wscale;
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    // volatile: keeps the issue order written below (dependent DMMAs six instructions apart);
    // without it nvcc pairs the two k-steps of an accumulator back to back and every second
    // DMMA waits for the full latency of the first (ncu: profiles/r01_notes.md).
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Table layout: [atom block of 8][row][8 atoms] double2, atoms of odd rows rotated by one position.
// One pipeline stage of the contraction (8 atoms) is then a handful of contiguous 512 B .. 2 KB
// pieces (4 Ex rows, 8 Ey rows, 16 Ez rows) that land in shared memory bank-conflict free.
__host__ __device__ __forceinline__ size_t tab_index(int atom, int row, int rows) {
    return ((size_t)(atom >> 3) * rows + row) * 8 + (((atom & 7) + (row & 1)) & 7);
}

// ------------------------------------------------------------------------------------------
// k_sf_tables: phase tables
//   grid = (3 ceil(ld/128), 1, n_jobs); thread = one atom on one axis (axis = blockIdx.x % 3).
// ------------------------------------------------------------------------------------------
//   LT: l values per l tile of the digit image known at compile time (32 / 40 / 48), 0: taken from the job.
//
//   Split-precision mode, z axis (tables_z_image): the digits of Z = Ez[l] go straight into the image the tensor-core
//   kernel reads -- rows 2 l (real part of A: (Zr, -Zi)) and 2 l + 1 (imaginary part: (Zi, Zr)) of l tile l / L_t, K position
//   2 (j % 16), five digit slices.  A thread owns an atom and walks l by recurrence; its ten 2-byte digit pairs per l would
//   be ten 2-byte global stores (measured: the image half of the kernel's bytes then moves at 2.8 TB/s against 6.2 TB/s for the
//   16-byte stores of the FP64 rows).  Instead the CTA (128 atoms = 8 K-steps) collects four l values -- one 8-row group of
//   the canonical layout: 256 contiguous bytes per (K-step, slice) -- in shared memory and writes them out with 16-byte
//   stores, 256-byte runs; two buffers, one barrier per group.
template <int LT>
__device__ __forceinline__ void tables_z_image(const TableJob &jb, int j, double u) {
    // shared layout of one 4-l group: [K-step 8][slice 5][K chunk 2][row 8][16 B], padded (K chunk stride 160 B, slice
    // 320 B, K-step 1600 B = 64 mod 128) so that the four 16-byte chunks a warp's 2-byte stores touch (two K-steps x two
    // K chunks) fall into four different bank groups -- unpadded they collide four ways and the L1 pipe saturates (ncu 93 %)
    constexpr int Z_KC = 160, Z_T = 2 * Z_KC, Z_KS = TC_S * Z_T, Z_BUF = 8 * Z_KS;
    __shared__ __align__(16) unsigned char s_img[2][Z_BUF];
    const int ld = jb.ld, n_atoms = jb.n;
    const int img_lt = LT ? LT : jb.img_lt, img_maxn = jb.img_maxn, img_rows = jb.img_n_lt * img_lt;
    const int np = 2 * img_lt;
    const size_t kstep_bytes = (size_t)(TC_S * np * 32);
    const size_t img_tile = (size_t)(ld / TC_KA) * kstep_bytes;
    const int ks = threadIdx.x >> 4, a = threadIdx.x & 15;
    const int kstep0 = ((blockIdx.x / 3) * blockDim.x) / TC_KA, n_ksteps = ld / TC_KA;
    const bool real = j < n_atoms;  // j in [n, ld): zero padding up to the leading dimension
    double s1 = 0.0, c1 = 1.0;
    if (real) {
        const double t = u - rint(u);
        sincospi(2.0 * t, &s1, &c1);
    }
    double cr = 1.0, ci = 0.0;
    unsigned char *mine = &s_img[0][0] + ks * Z_KS + (a >> 3) * Z_KC + (a & 7) * 2;
    for (int l0 = 0; l0 < img_rows; l0 += 4) {
        unsigned char *buf = mine + ((l0 >> 2) & 1) * Z_BUF;
#pragma unroll
        for (int dl = 0; dl < 4; ++dl) {
            const int l = l0 + dl;
            if (real && (l & 7) == 0) {  // exact re-seed: limits recurrence drift to 7 steps
                double t = (double)l * u;
                t -= rint(t);
                sincospi(2.0 * t, &ci, &cr);
            }
            const bool in = real && l < img_maxn;  // l values of the last tile beyond the cube edge are zero rows
            unsigned rlo, rhi, ilo, ihi, nlo, nhi;
            tc_digits(in ? cr : 0.0, rlo, rhi);
            tc_digits(in ? ci : 0.0, ilo, ihi);
            tc_digits(in ? -ci : 0.0, nlo, nhi);
#pragma unroll
            for (int t = 1; t <= TC_S; ++t) {
                *reinterpret_cast<uint16_t *>(buf + (t - 1) * Z_T + (2 * dl) * 16) = (uint16_t)tc_pair(t, rlo, rhi, nlo, nhi);
                *reinterpret_cast<uint16_t *>(buf + (t - 1) * Z_T + (2 * dl + 1) * 16) = (uint16_t)tc_pair(t, ilo, ihi, rlo, rhi);
            }
            const double nr = cr * c1 - ci * s1;
            const double ni = cr * s1 + ci * c1;
            cr = nr;
            ci = ni;
        }
        __syncthreads();
        // 8 K-steps x 5 slices x 256 bytes = 640 chunks of 16 bytes
        const unsigned char *src = &s_img[(l0 >> 2) & 1][0];
        int8_t *dst = jb.bimg + (size_t)(l0 / img_lt) * img_tile + (size_t)((l0 % img_lt) >> 2) * 256;
#pragma unroll
        for (int i = 0; i < TC_S; ++i) {
            const int c = threadIdx.x + 128 * i;
            const int cks = c / (TC_S * 16), rem = c - cks * (TC_S * 16), t = rem >> 4, w = rem & 15;
            if (kstep0 + cks < n_ksteps)
                *reinterpret_cast<uint4 *>(dst + (size_t)(kstep0 + cks) * kstep_bytes + (size_t)t * (np * 32) + w * 16) =
                    *reinterpret_cast<const uint4 *>(src + cks * Z_KS + t * Z_T + (w >> 3) * Z_KC + (w & 7) * 16);
        }
    }
}

template <int LT>
__global__ void __launch_bounds__(128) k_sf_tables(const TableJob *__restrict__ jobs) {
    const TableJob &jb = jobs[blockIdx.z];
    // the axis is the fastest block index: CTAs of the write-heavy x rows, the light y rows and the compute-heavy z digits
    // of the same atoms are resident side by side instead of one phase after the other
    const int axis = blockIdx.x % 3;
    double2 *tab = jb.tab[axis];
    if (tab == nullptr) return;
    const int j = (blockIdx.x / 3) * blockDim.x + threadIdx.x;
    // every field the row loop needs goes to registers up front: the job lives in global memory and the loop's stores
    // would otherwise force the compiler to reload it in every iteration
    const int ld = jb.ld, n_atoms = jb.n;
    const bool to_image = axis == 2 && jb.bimg != nullptr;  // uniform over the CTA
    if (j >= ld && !to_image) return;
    double u = 0.0;
    if (j < n_atoms) {
        const int64_t a = jb.index ? (int64_t)jb.index[j] : (int64_t)j;
        if (jb.is_f32) {
            float p = reinterpret_cast<const float *>(jb.pos)[a * jb.stride_atom + axis * jb.stride_xyz];
            const float L = jb.boxf[axis];
            const int wrap_mode = jb.wrap_mode;
            if (wrap_mode == 2)  // AtomGroup.wrap(compound="atoms"): x -= floor(x/L)*L in float32
                p = __fsub_rn(p, __fmul_rn(floorf(__fdiv_rn(p, L)), L));
            if (wrap_mode >= 1)  // saxs.py:193-195: pos - box*round(pos/box), float32 arithmetic
                p = __fsub_rn(p, __fmul_rn(L, rintf(__fdiv_rn(p, L))));
            u = (double)p / (double)L;
        } else {
            double p = reinterpret_cast<const double *>(jb.pos)[a * jb.stride_atom + axis * jb.stride_xyz];
            u = p / jb.box[axis];
        }
    }
    if (to_image) {  // every thread of the CTA takes part (block barriers), also those beyond ld
        tables_z_image<LT>(jb, j, u);
        return;
    }
    const int rows = jb.rows[axis];
    const bool masked = axis == 1 && jb.use_y_mask != 0;
    const unsigned long long ym0 = masked ? jb.y_mask[0] : ~0ull, ym1 = masked ? jb.y_mask[1] : ~0ull,
                             ym2 = masked ? jb.y_mask[2] : ~0ull, ym3 = masked ? jb.y_mask[3] : ~0ull;
    auto wanted = [&](int r) -> bool {
        const int wd = (r >> 6) & 3;
        const unsigned long long m = wd == 0 ? ym0 : (wd == 1 ? ym1 : (wd == 2 ? ym2 : ym3));
        return (m >> (r & 63)) & 1ull;
    };
    if (j >= n_atoms) {  // zero padding up to the leading dimension
        for (int r = 0; r < rows; ++r)
            if (wanted(r)) tab[tab_index(j, r, rows)] = make_double2(0.0, 0.0);
        return;
    }
    const double w = (axis == 0 && jb.weights) ? jb.weights[j] * jb.wscale : 1.0;
    double s1, c1;
    {
        double t = u - rint(u);
        sincospi(2.0 * t, &s1, &c1);
    }
    double cr = 1.0, ci = 0.0;
    for (int r = 0; r < rows; ++r) {
        if ((r & 7) == 0) {  // exact re-seed: limits recurrence drift to 7 steps
            double t = (double)r * u;
            t -= rint(t);
            sincospi(2.0 * t, &ci, &cr);
        }
        if (wanted(r)) tab[tab_index(j, r, rows)] = make_double2(w * cr, w * ci);
        const double nr = cr * c1 - ci * s1;
        const double ni = cr * s1 + ci * c1;
        cr = nr;
        ci = ni;
    }
}

// ------------------------------------------------------------------------------------------
// k_sf_contract: the DMMA contraction (persistent CTAs, dynamic item queue)
//
//   per stage (8 atoms) a CTA stages the table rows its warps need -- Ex rows h0..h0+3 and Ey rows
//   k0..k0+7 of each of its (h,k) quads, Ez rows l0..l0+15 of each of its l-pairs -- as
//   [row][8 atoms + pad] double2 in an NSTAGE-deep ring.  Thread r owns row r: one 128-byte
//   cp.async.bulk (SASS UBLKCP) per stage, completion counted on the slot's "full" mbarrier; a slot
//   is refilled after every warp has arrived on its "empty" mbarrier.  No block-wide barrier in the
//   atom loop: warps drift up to NSTAGE-1 stages apart.
//   every warp owns a 32 (hk) x 16 (l) tile: 4 x 2 DMMA tiles x 3 planes of accumulators.
//   A fragments P = Ex*Ey (and Pr+Pi) are formed in registers from the staged rows, B fragments
//   are the staged Ez rows (and Zr+Zi); k index of the MMA = atom.  Warp tiles that lie fully
//   inside the q-sphere take an unpredicated instruction stream.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(addr), "r"(parity)
                     : "memory");
        if (spin > (1u << 26)) __trap();  // a lost completion must fail the launch, never hang the GPU
    }
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// MMAs of one 8-atom stage for one warp tile.  FULL: every 8x8 tile valid -> no predication.
// Fragment positions: lane (g, tig) needs atoms 2*tig and 2*tig+1 of its rows; in a row of parity p
// they sit at positions (2*tig + p) & 7 and (2*tig + 1 + p) & 7 (rotation of odd rows).
struct FragOff {
    int ex[2][2];  // [row parity][atom]  (row = h0 + mt, parity = mt & 1; row base added per mt)
    int ey[2];     // Ey row k0 + g
    int ez[2];     // Ez row l0 + g (second n-tile: + 8 rows, same parity)
};

template <bool FULL>
__device__ __forceinline__ void contract_stage(double (&acc)[4][2][3][2], const double2 *__restrict__ st,
                                               const FragOff &fo, unsigned wmask) {
    double br[2][2], bi[2][2], bs[2][2];  // [nt][kstep]
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
        const double2 z0 = st[fo.ez[0] + nt * 8 * RAWROW], z1 = st[fo.ez[1] + nt * 8 * RAWROW];
        br[nt][0] = z0.x; bi[nt][0] = z0.y; bs[nt][0] = z0.x + z0.y;
        br[nt][1] = z1.x; bi[nt][1] = z1.y; bs[nt][1] = z1.x + z1.y;
    }
    const double2 y0 = st[fo.ey[0]], y1 = st[fo.ey[1]];
    double2 xn0 = st[fo.ex[0][0]], xn1 = st[fo.ex[0][1]];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const double2 x0 = xn0, x1 = xn1;
        if (mt < 3) {  // prefetch the next Ex row while this row's DMMAs issue
            xn0 = st[fo.ex[(mt + 1) & 1][0] + (mt + 1) * RAWROW];
            xn1 = st[fo.ex[(mt + 1) & 1][1] + (mt + 1) * RAWROW];
        }
        if (FULL || ((wmask >> (2 * mt)) & 3u)) {
            const double pr0 = x0.x * y0.x - x0.y * y0.y, pi0 = x0.x * y0.y + x0.y * y0.x;
            const double pr1 = x1.x * y1.x - x1.y * y1.y, pi1 = x1.x * y1.y + x1.y * y1.x;
            const double ps0 = pr0 + pi0, ps1 = pr1 + pi1;
            const bool v0 = FULL || ((wmask >> (2 * mt)) & 1u), v1 = FULL || ((wmask >> (2 * mt + 1)) & 1u);
            if (v0) {
                dmma(acc[mt][0][0][0], acc[mt][0][0][1], pr0, br[0][0]);
                dmma(acc[mt][0][1][0], acc[mt][0][1][1], pi0, bi[0][0]);
                dmma(acc[mt][0][2][0], acc[mt][0][2][1], ps0, bs[0][0]);
            }
            if (v1) {
                dmma(acc[mt][1][0][0], acc[mt][1][0][1], pr0, br[1][0]);
                dmma(acc[mt][1][1][0], acc[mt][1][1][1], pi0, bi[1][0]);
                dmma(acc[mt][1][2][0], acc[mt][1][2][1], ps0, bs[1][0]);
            }
            if (v0) {
                dmma(acc[mt][0][0][0], acc[mt][0][0][1], pr1, br[0][1]);
                dmma(acc[mt][0][1][0], acc[mt][0][1][1], pi1, bi[0][1]);
                dmma(acc[mt][0][2][0], acc[mt][0][2][1], ps1, bs[0][1]);
            }
            if (v1) {
                dmma(acc[mt][1][0][0], acc[mt][1][0][1], pr1, br[1][1]);
                dmma(acc[mt][1][1][0], acc[mt][1][1][1], pi1, bi[1][1]);
                dmma(acc[mt][1][2][0], acc[mt][1][2][1], ps1, bs[1][1]);
            }
        }
    }
}

__global__ void __launch_bounds__(NTHREADS, 3)
k_sf_contract(const ItemDesc *__restrict__ items, int n_items, const SegDesc *__restrict__ segs,
              int *__restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2 *raw = reinterpret_cast<double2 *>(smem);  // [NSTAGE][ST_ROWS][RAWROW]
    unsigned long long *full_bar = reinterpret_cast<unsigned long long *>(smem + NSTAGE * SMEM_STAGE);  // [NSTAGE]
    unsigned long long *empty_bar = full_bar + NSTAGE;                                                  // [NSTAGE]
    int *s_item = reinterpret_cast<int *>(empty_bar + NSTAGE);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(full_bar + s, 1);     // one arrive.expect_tx per fill
            mbar_init(empty_bar + s, WPC);  // one arrive per warp per drain
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    unsigned gstage = 0;  // stages issued / consumed by this CTA so far (uniform over threads)

    for (;;) {
        if (tid == 0) *s_item = atomicAdd(counter, 1);
        __syncthreads();  // also: every warp is done with the previous item
        const int item = *s_item;
        if (item >= n_items) break;

        const ItemDesc it = items[item];
        const SegDesc &sg = segs[it.seg];
        const TileGroupDesc *tg = sg.tgs + it.tg;
        const int n_x = tg->n_x, n_y = tg->n_y;

        // staged pieces of one 8-atom block: per (h,k) quad 4 Ex rows (512 B) + 8 Ey rows (1 KB), per l-pair
        // 16 Ez rows (2 KB); thread c < n_copy owns piece c.  Stage layout: [x-slot][4 Ex + 8 Ey][8], [y-slot][16][8].
        const int n_xy = n_x * (QH + QK);
        const int n_rows = n_xy + n_y * NL;
        const int n_copy = 2 * n_x + n_y;
        const double2 *my_src = nullptr;
        unsigned my_bytes = 0, my_dst = 0;
        size_t my_stride = 0;  // double2 per atom block
        if (tid < n_copy) {
            const size_t blk0 = (size_t)(it.j0 >> 3);
            if (tid < 2 * n_x) {
                const int xs = tid >> 1;
                if ((tid & 1) == 0) {
                    my_stride = (size_t)sg.rows[0] * 8;
                    my_src = sg.ex + blk0 * my_stride + (size_t)tg->xh0[xs] * 8;
                    my_bytes = QH * 128u;
                    my_dst = (unsigned)(xs * (QH + QK)) * RAWROW;
                } else {
                    my_stride = (size_t)sg.rows[1] * 8;
                    my_src = sg.ey + blk0 * my_stride + (size_t)tg->xk0[xs] * 8;
                    my_bytes = QK * 128u;
                    my_dst = (unsigned)(xs * (QH + QK) + QH) * RAWROW;
                }
            } else {
                const int ys = tid - 2 * n_x;
                my_stride = (size_t)sg.rows[2] * 8;
                my_src = sg.ez + blk0 * my_stride + (size_t)tg->yl0[ys] * 8;
                my_bytes = NL * 128u;
                my_dst = (unsigned)(n_xy + ys * NL) * RAWROW;
            }
        }
        const bool active = warp < tg->n_warp;
        const int wx = active ? tg->wx[warp] : 0, wy = active ? tg->wy[warp] : 0;
        const unsigned wmask = active ? tg->wmask[warp] : 0u;
        __syncthreads();  // s_item may be rewritten only after everybody has read it

        const int n_stage = (it.j1 - it.j0) / KT;
        const unsigned stage_bytes = (unsigned)n_rows * KT * 16u;

        auto issue = [&](int t) {  // fill ring slot of stage t (thread r: row r)
            if (t < n_stage) {
                const unsigned gs = gstage + (unsigned)t;
                const unsigned slot = gs % NSTAGE, use = gs / NSTAGE;
                if (tid < n_copy) {
                    if (use > 0) mbar_wait(empty_bar + slot, (use - 1) & 1u);  // all warps drained the previous use
                    if (tid == 0) mbar_arrive_expect_tx(full_bar + slot, stage_bytes);
                    bulk_g2s(raw + slot * (ST_ROWS * RAWROW) + my_dst, my_src + (size_t)t * my_stride, my_bytes,
                             full_bar + slot);
                }
            }
        };

        // accumulators: [mt][nt] x planes (T1,T2,T3) x {c0,c1}
        double acc[4][2][3][2];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int p = 0; p < 3; ++p) acc[mt][nt][p][0] = acc[mt][nt][p][1] = 0.0;

#pragma unroll
        for (int s = 0; s < LOOKAHEAD; ++s) issue(s);
        // per-thread fragment offsets inside a stage (double2 units)
        FragOff fo;
        {
            const int pg = g & 1;  // parity of rows k0 + g and l0 (+8) + g; h0 + mt has parity mt & 1
            const int bx = (wx * (QH + QK)) * RAWROW, by = (wx * (QH + QK) + QH + g) * RAWROW,
                      bz = (n_xy + wy * NL + g) * RAWROW;
            fo.ex[0][0] = bx + 2 * tig;               fo.ex[0][1] = bx + 2 * tig + 1;
            fo.ex[1][0] = bx + ((2 * tig + 1) & 7);   fo.ex[1][1] = bx + ((2 * tig + 2) & 7);
            fo.ey[0] = by + ((2 * tig + pg) & 7);     fo.ey[1] = by + ((2 * tig + 1 + pg) & 7);
            fo.ez[0] = bz + ((2 * tig + pg) & 7);     fo.ez[1] = bz + ((2 * tig + 1 + pg) & 7);
        }

        for (int t = 0; t < n_stage; ++t) {
            const unsigned gs = gstage + (unsigned)t;
            const unsigned slot = gs % NSTAGE;
            issue(t + LOOKAHEAD);  // LOOKAHEAD < NSTAGE - 1 leaves slack: warps may drift apart without blocking
            mbar_wait(full_bar + slot, (gs / NSTAGE) & 1u);
            const double2 *st = raw + slot * (ST_ROWS * RAWROW);
            if (wmask == 0xFFu) {
                contract_stage<true>(acc, st, fo, wmask);
            } else if (wmask) {
                contract_stage<false>(acc, st, fo, wmask);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar + slot);
        }
        gstage += (unsigned)n_stage;

        // epilogue: Re = T1 - T2, Im = T3 - T1 - T2 -> amp[split][wt][row][col]
        if (wmask) {
            double2 *out = sg.amp + ((size_t)it.split * sg.n_wt + tg->wt0 + warp) * 512;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    if ((wmask >> (2 * mt + nt)) & 1u) {
                        const int row = mt * 8 + g, col = nt * 8 + 2 * tig;
                        double2 v0, v1;
                        v0.x = acc[mt][nt][0][0] - acc[mt][nt][1][0];
                        v0.y = acc[mt][nt][2][0] - acc[mt][nt][0][0] - acc[mt][nt][1][0];
                        v1.x = acc[mt][nt][0][1] - acc[mt][nt][1][1];
                        v1.y = acc[mt][nt][2][1] - acc[mt][nt][0][1] - acc[mt][nt][1][1];
                        out[row * 16 + col] = v0;
                        out[row * 16 + col + 1] = v1;
                    }
                }
        }
    }
}

// ------------------------------------------------------------------------------------------
// finalisers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double s_of_entry(const SegDesc &sg, int e) {
    const int ai = sg.ent_amp[e];
    if (sg.tile_owned && !sg.tile_owned[ai / (int)sg.tile_elems]) return 0.0;
    double re = 0.0, im = 0.0;
    const size_t stride = (size_t)sg.amp_stride;
    for (int s = 0; s < sg.n_split; ++s) {  // fixed order -> reproducible
        const double2 v = sg.amp[s * stride + ai];
        re += v.x;
        im += v.y;
    }
    return (re * re + im * im) * sg.s_scale;
}

// Split-K partial amplitudes folded into slab 0, amp[0][i] = sum_s amp[s][i], in the fixed order of s_of_entry (the
// result is bitwise the same).  For big frames (config 5: 123 slabs of 16 MB per segment) the finalisers' per-entry gathers
// over all slabs cost more than the whole streaming pass: k_sf_binned 8.0 ms -> fold 0.5 ms + 0.3 ms.
struct FoldDesc {
    double2 *amp;
    long long stride;  // elements per slab
    int32_t n_split, tile_elems;
    const uint8_t *owned;  // per tile of tile_elems elements: folded only when set (null: all)
};
__global__ void __launch_bounds__(256) k_sf_fold_splits(const FoldDesc *__restrict__ descs) {
    const FoldDesc d = descs[blockIdx.y];
    if (d.n_split < 2) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < d.stride; i += (long long)gridDim.x * blockDim.x) {
        if (d.owned && !d.owned[i / d.tile_elems]) continue;
        double re = 0.0, im = 0.0;
        for (int sidx = 0; sidx < d.n_split; ++sidx) {
            const double2 v = d.amp[(size_t)sidx * d.stride + i];
            re += v.x;
            im += v.y;
        }
        d.amp[i] = make_double2(re, im);
    }
}

// compact S list: svals[seg][e]
__global__ void __launch_bounds__(256) k_sf_svalues(const SegDesc *__restrict__ segs, double *__restrict__ svals,
                                                   int sval_stride) {
    const SegDesc &sg = segs[blockIdx.y];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < sg.n_valid) svals[(size_t)blockIdx.y * sval_stride + e] = s_of_entry(sg, e);
}

// |q| histograms: out[seg][{S, f2S}][n_bins], cnt[seg][n_bins].  The entries of a (segment, bin) are cut into
// gridDim.z equal parts; a CTA reduces its part in a fixed order (thread-strided sums, shuffle tree, four warp
// partials) and writes the part's result to `part` [seg][bin][z][3]; k_sf_binned_final adds the parts in order.
// The result depends on n_parts (a function of the plan size alone), never on scheduling: bitwise reproducible.
__global__ void __launch_bounds__(128) k_sf_binned(const SegDesc *__restrict__ segs, double *__restrict__ part, int n_bins) {
    const SegDesc &sg = segs[blockIdx.y];
    const int b = blockIdx.x;
    const int n_parts = gridDim.z, z = blockIdx.z;
    const int e_lo = sg.bin_start[b], e_hi = sg.bin_start[b + 1];
    const int per = (e_hi - e_lo + n_parts - 1) / n_parts;
    const int e0 = e_lo + z * per, e1 = min(e_hi, e0 + per);
    double s = 0.0, si = 0.0;
    long long c = 0;
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const double v = s_of_entry(sg, e);
        s += v;
        if (sg.ff2) si += sg.ff2[e] * v;
        c += (v != 0.0);
    }
    __shared__ double sh_s[4], sh_i[4];
    __shared__ long long sh_c[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        si += __shfl_down_sync(0xffffffffu, si, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh_s[threadIdx.x >> 5] = s;
        sh_i[threadIdx.x >> 5] = si;
        sh_c[threadIdx.x >> 5] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double *o = part + (((size_t)blockIdx.y * n_bins + b) * n_parts + z) * 3;
        o[0] = (sh_s[0] + sh_s[1]) + (sh_s[2] + sh_s[3]);
        o[1] = (sh_i[0] + sh_i[1]) + (sh_i[2] + sh_i[3]);
        o[2] = __longlong_as_double(sh_c[0] + sh_c[1] + sh_c[2] + sh_c[3]);
    }
}

__global__ void __launch_bounds__(128) k_sf_binned_final(const double *__restrict__ part, double *__restrict__ out_s,
                                                        double *__restrict__ out_i, long long *__restrict__ out_c,
                                                        int n_parts, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    double s = 0.0, si = 0.0;
    long long c = 0;
    for (int z = 0; z < n_parts; ++z) {
        const double *p = part + ((size_t)i * n_parts + z) * 3;
        s += p[0];
        si += p[1];
        c += __double_as_longlong(p[2]);
    }
    out_s[i] = s;
    if (out_i) out_i[i] = si;
    out_c[i] = c;
}

}  // namespace mb200
#include "sfactor_tc.cuh"
namespace mb200 {

// ------------------------------------------------------------------------------------------
// host: q-plan
// ------------------------------------------------------------------------------------------
struct SfPlan {
    // key
    double dims[3], qmin, qmax, thmin, thmax;
    int32_t n_bins;  // 0: entries in C order of the cube; >0: entries sorted by |q| bin
    // geometry
    int32_t maxn[3];
    int32_t rows[3];  // padded table rows
    // entries (valid q vectors)
    std::vector<int32_t> cell;     // h*maxn1*maxn2 + k*maxn2 + l
    std::vector<double> qrr;       // |q|
    std::vector<int32_t> ent_amp;  // index into [n_wt][512]
    std::vector<int32_t> bin_start;
    // tiles
    std::vector<TileGroupDesc> tgs;
    int32_t n_wt = 0;
    int64_t n_mma_tiles = 0;  // valid 8x8 DMMA tiles
    // tensor-core (split precision) tiling: packed (h,k) rows in tiles of 128, l tiles of <= 48
    std::vector<uint8_t> tc_rows;       // [tc_n_rt * 128][2]
    unsigned long long tc_ymask[4] = {0, 0, 0, 0};  // y-table rows the tensor-core generators read (k0 of every quad, row 1)
    std::vector<int32_t> ent_amp_tc;    // index into [tc_n_rt * tc_n_lt][128][tc_lt]
    std::vector<uint8_t> tc_tile_used;  // [tc_n_rt * tc_n_lt] 1: the (row tile, l tile) block holds at least one valid q
    int32_t tc_n_rt = 0, tc_n_lt = 0, tc_lt = 0, tc_n_pairs = 0, tc_n_used = 0;
    bool tc_ok = false;
    // device copies
    DevBuf d_tgs, d_ent_amp, d_bin_start, d_ff2, d_tc_rows, d_ent_amp_tc;
    std::vector<double> ff2_key;  // cm params the cached ff2 was built for
    int32_t ff2_groups = 0;
    std::vector<double> ff2_host;      // squared form factors computed ahead (prepare_plans), uploaded by plan_ff2
    std::vector<double> ff2_host_key;

    // a retired plan becomes the next one: vectors and device buffers keep their capacity (no cudaMalloc / cudaFree --
    // which synchronises the device -- per frame of an NPT trajectory)
    void reset() {
        cell.clear(); qrr.clear(); ent_amp.clear(); bin_start.clear(); tgs.clear(); tc_rows.clear(); ent_amp_tc.clear();
        tc_tile_used.clear(); ff2_key.clear(); ff2_host.clear(); ff2_host_key.clear();
        n_wt = 0; n_mma_tiles = 0; tc_n_rt = tc_n_lt = tc_lt = tc_n_pairs = tc_n_used = 0; tc_ok = false; ff2_groups = 0;
        for (auto &m : tc_ymask) m = 0ull;
    }
};

struct PlanPool {
    std::mutex m;
    std::vector<SfPlan *> free;
    SfPlan *take() {
        std::lock_guard<std::mutex> lock(m);
        if (free.empty()) return nullptr;
        SfPlan *p = free.back();
        free.pop_back();
        return p;
    }
    void give(SfPlan *p) {
        {
            std::lock_guard<std::mutex> lock(m);
            if (free.size() < 320) { free.push_back(p); return; }
        }
        delete p;
    }
    ~PlanPool() { for (SfPlan *p : free) delete p; }
};

static void sfactor_maxn_host(const double *dims, double qmax, int32_t *maxn) {
    for (int i = 0; i < 3; ++i) {
        const double q_factor = 2 * M_PI / dims[i];
        // _cmath.pyx:96: the divisor is cast to float32 before the (double) division
        maxn[i] = (int32_t)std::ceil(qmax / (double)(float)q_factor);
    }
}

// numpy.histogram uniform-bin index for x in [first, last] (numpy/lib/_histograms_impl.py:840-873)
static inline int32_t numpy_bin(double x, double first, double last, int32_t n, const std::vector<double> &edges) {
    const double denom = last - first;
    double f = ((x - first) / denom) * (double)n;
    int64_t idx = (int64_t)f;
    if (idx == n) idx -= 1;
    if (x < edges[idx]) idx -= 1;
    if (x >= edges[idx + 1] && idx != n - 1) idx += 1;
    return (int32_t)idx;
}

static int build_plan(const double *dims, double qmin, double qmax, double thmin, double thmax, int32_t n_bins,
                      std::shared_ptr<SfPlan> &out, const std::shared_ptr<PlanPool> &pool = nullptr) {
    std::shared_ptr<SfPlan> pl;
    if (pool) {  // the plan goes back to the pool when its last holder lets go of it
        SfPlan *raw = pool->take();
        if (raw) raw->reset(); else raw = new SfPlan();
        pl = std::shared_ptr<SfPlan>(raw, [pool](SfPlan *p) { pool->give(p); });
    } else {
        pl = std::make_shared<SfPlan>();
    }
    for (int i = 0; i < 3; ++i) pl->dims[i] = dims[i];
    pl->qmin = qmin; pl->qmax = qmax; pl->thmin = thmin; pl->thmax = thmax; pl->n_bins = n_bins;
    sfactor_maxn_host(dims, qmax, pl->maxn);
    const int m0 = std::max(pl->maxn[0], 0), m1 = std::max(pl->maxn[1], 0), m2 = std::max(pl->maxn[2], 0);
    pl->maxn[0] = m0; pl->maxn[1] = m1; pl->maxn[2] = m2;
    if (m0 > 30000 || m1 > 30000 || m2 > 30000 || (double)m0 * m1 * m2 > 2.0e9)
        return fail(MB200_EINVAL, "q-range too large for this box: maxn = (%d, %d, %d)", m0, m1, m2);
    pl->rows[0] = (m0 + QH - 1) / QH * QH;
    pl->rows[1] = (m1 + QK - 1) / QK * QK;
    pl->rows[2] = (m2 + NL - 1) / NL * NL;
    double qf[3];
    for (int i = 0; i < 3; ++i) qf[i] = 2 * M_PI / dims[i];

    static const bool trace_plan = std::getenv("MAICOS_B200_TRACE_PLAN") != nullptr;  // developer aid: phase times to stderr
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (!trace_plan) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[build_plan] %-18s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    const size_t cells = (size_t)m0 * m1 * m2;
    std::vector<double> qcube(cells, 0.0);  // 0 = rejected
    const bool need_theta = !(thmin <= 0.0 && thmax >= 1.5707963267948968);
    const size_t par_h = cells >= 200000 ? 2 : (size_t)m0 + 1;  // threads only for big cubes
    // the acceptance test is the reference's (sqrt, then the comparisons with qmin / qmax); cells whose q^2 lies clearly
    // outside [qmin^2, qmax^2] skip the square root -- the margin is far wider than the rounding of q^2 and of the root
    const double q2_hi = qmax > 0.0 ? qmax * qmax * (1.0 + 1e-12) : 0.0, q2_lo = qmin > 0.0 ? qmin * qmin * (1.0 - 1e-12) : -1.0;
    std::vector<double> qz_of(m2), qz2_of(m2);
    for (int l = 0; l < m2; ++l) { qz_of[l] = l * qf[2]; qz2_of[l] = qz_of[l] * qz_of[l]; }
    parallel_chunks((size_t)m0, par_h, [&](size_t h_lo, size_t h_hi) {
        for (int h = (int)h_lo; h < (int)h_hi; ++h) {
            const double qx = h * qf[0];
            for (int k = 0; k < m1; ++k) {
                const double qy = k * qf[1];
                const double qxy2 = qx * qx + qy * qy;  // (qx qx + qy qy) + qz qz: the reference's order of additions
                if (qxy2 > q2_hi) continue;
                double *row = qcube.data() + ((size_t)h * m1 + k) * m2;
                for (int l = (h + k == 0) ? 1 : 0; l < m2; ++l) {
                    const double q2 = qxy2 + qz2_of[l];
                    if (q2 > q2_hi) break;  // q grows with l
                    if (q2 < q2_lo) continue;
                    const double qrr = std::sqrt(q2);
                    if (!(qrr >= qmin && qrr <= qmax)) continue;
                    if (need_theta) {
                        const double theta = std::acos(qz_of[l] / qrr);
                        if (!(theta >= thmin && theta <= thmax)) continue;
                    }
                    row[l] = qrr;
                }
            }
        }
    });
    mark("q cube");
    // warp tiles: (quad, l-pair) with at least one valid q; DMMA-tile validity mask
    const int nqh = pl->rows[0] / QH, nqk = pl->rows[1] / QK, nlp = pl->rows[2] / NL;
    std::vector<int32_t> wt_of((size_t)nqh * nqk * nlp, -1);
    struct WT { int16_t h0, k0, l0; uint8_t mask; };
    std::vector<WT> wts;
    std::vector<std::vector<WT>> wts_a(nqh);  // per quad row a, in (b, c) order; concatenated in a order below
    parallel_chunks((size_t)nqh, cells >= 200000 ? 2 : (size_t)nqh + 1, [&](size_t a_lo, size_t a_hi) {
    for (int a = (int)a_lo; a < (int)a_hi; ++a)
        for (int b = 0; b < nqk; ++b)
            for (int c = 0; c < nlp; ++c) {
                unsigned mask = 0;
                for (int mt = 0; mt < 4; ++mt) {
                    const int h = a * QH + mt;
                    if (h >= m0) continue;
                    for (int nt = 0; nt < 2; ++nt) {
                        bool any = false;
                        for (int kk = 0; kk < QK && !any; ++kk) {
                            const int k = b * QK + kk;
                            if (k >= m1) break;
                            for (int ll = 0; ll < 8; ++ll) {
                                const int l = c * NL + nt * 8 + ll;
                                if (l >= m2) break;
                                if (qcube[((size_t)h * m1 + k) * m2 + l] != 0.0) { any = true; break; }
                            }
                        }
                        if (any) mask |= 1u << (2 * mt + nt);
                    }
                }
                if (mask) wts_a[a].push_back({(int16_t)(a * QH), (int16_t)(b * QK), (int16_t)(c * NL), (uint8_t)mask});
            }
    });
    for (int a = 0; a < nqh; ++a)
        for (const WT &t : wts_a[a]) {
            wt_of[((size_t)a * nqk + t.k0 / QK) * nlp + t.l0 / NL] = (int32_t)wts.size();
            wts.push_back(t);
            pl->n_mma_tiles += __builtin_popcount(t.mask);
        }
    pl->n_wt = (int32_t)wts.size();
    // pack consecutive warp tiles into CTAs
    for (size_t i = 0; i < wts.size(); i += WPC) {
        TileGroupDesc tg;
        std::memset(&tg, 0, sizeof(tg));
        tg.wt0 = (int32_t)i;
        tg.n_warp = (int32_t)std::min<size_t>(WPC, wts.size() - i);
        for (int w = 0; w < tg.n_warp; ++w) {
            const WT &t = wts[i + w];
            int xs = -1, ys = -1;
            for (int s = 0; s < tg.n_x; ++s)
                if (tg.xh0[s] == t.h0 && tg.xk0[s] == t.k0) xs = s;
            if (xs < 0) { xs = tg.n_x++; tg.xh0[xs] = t.h0; tg.xk0[xs] = t.k0; }
            for (int s = 0; s < tg.n_y; ++s)
                if (tg.yl0[s] == t.l0) ys = s;
            if (ys < 0) { ys = tg.n_y++; tg.yl0[ys] = t.l0; }
            tg.wx[w] = (uint8_t)xs; tg.wy[w] = (uint8_t)ys; tg.wmask[w] = t.mask;
        }
        pl->tgs.push_back(tg);
    }
    mark("dmma tiles");
    // entries
    std::vector<int32_t> cell; std::vector<double> qv; std::vector<int32_t> amp;
    std::vector<uint32_t> hkl, hkl_sorted;  // (h << 20 | k << 10 | l) of every entry when it fits: spares the tilings' divisions
    const bool pack_hkl = m0 <= 1023 && m1 <= 1023 && m2 <= 1023;
    {
        std::vector<size_t> h_start((size_t)m0 + 1, 0);  // entries are listed in C order of the cube: offsets per h slab
        parallel_chunks((size_t)m0, par_h, [&](size_t h_lo, size_t h_hi) {
            for (size_t h = h_lo; h < h_hi; ++h) {
                size_t c = 0;
                const double *slab = qcube.data() + h * (size_t)m1 * m2;
                for (size_t i = 0; i < (size_t)m1 * m2; ++i) c += slab[i] != 0.0;
                h_start[h + 1] = c;
            }
        });
        for (int h = 0; h < m0; ++h) h_start[h + 1] += h_start[h];
        const size_t total = h_start[m0];
        cell.resize(total); qv.resize(total); amp.resize(total);
        if (pack_hkl) hkl.resize(total);
        parallel_chunks((size_t)m0, par_h, [&](size_t h_lo, size_t h_hi) {
            for (int h = (int)h_lo; h < (int)h_hi; ++h) {
                size_t o = h_start[h];
                for (int k = 0; k < m1; ++k) {
                    const size_t ci0 = ((size_t)h * m1 + k) * m2;
                    const size_t wt0 = ((size_t)(h / QH) * nqk + (k / QK)) * nlp;
                    const int32_t amp_hk = (h % QH) * 8 + (k % QK);
                    const uint32_t hk_bits = ((uint32_t)h << 20) | ((uint32_t)k << 10);
                    for (int l = 0; l < m2; ++l) {
                        const size_t ci = ci0 + l;
                        if (qcube[ci] == 0.0) continue;
                        const int32_t wt = wt_of[wt0 + (l / NL)];
                        cell[o] = (int32_t)ci;
                        qv[o] = qcube[ci];
                        amp[o] = (wt * 32 + amp_hk) * 16 + (l % NL);
                        if (pack_hkl) hkl[o] = hk_bits | (uint32_t)l;
                        ++o;
                    }
                }
            }
        });
    }
    const size_t nv = cell.size();
    mark("entries");
    if (n_bins > 0) {
        // np.histogram(a=|q|, bins=n_bins, range=(qmin, qmax)) -- saxs.py:222-233
        double first = qmin, last = qmax;
        if (first == last) { first -= 0.5; last += 0.5; }
        std::vector<double> edges(n_bins + 1);
        const double step = (last - first) / (double)n_bins;  // np.linspace
        for (int i = 0; i <= n_bins; ++i) edges[i] = (double)i * step + first;
        edges[n_bins] = last;
        std::vector<int32_t> bin(nv);
        std::vector<int32_t> count(n_bins + 1, 0);
        parallel_chunks(nv, 100000, [&](size_t e_lo, size_t e_hi) {
            for (size_t e = e_lo; e < e_hi; ++e) bin[e] = numpy_bin(qv[e], first, last, n_bins, edges);
        });
        for (size_t e = 0; e < nv; ++e) count[bin[e] + 1]++;
        for (int b = 0; b < n_bins; ++b) count[b + 1] += count[b];
        pl->bin_start = count;
        std::vector<int32_t> pos(count.begin(), count.end() - 1);
        pl->cell.resize(nv); pl->qrr.resize(nv); pl->ent_amp.resize(nv);
        if (pack_hkl) hkl_sorted.resize(nv);
        for (size_t e = 0; e < nv; ++e) {  // stable counting sort by bin
            const int32_t d = pos[bin[e]]++;
            pl->cell[d] = cell[e]; pl->qrr[d] = qv[e]; pl->ent_amp[d] = amp[e];
            if (pack_hkl) hkl_sorted[d] = hkl[e];
        }
        hkl.swap(hkl_sorted);
    } else {
        pl->cell.swap(cell); pl->qrr.swap(qv); pl->ent_amp.swap(amp);
    }
    mark("bins");
    // tensor-core tiling: (h,k) pairs with at least one valid l, packed in quads (h, k0 .. k0 + 3) -- the generators form the
    // four rows of a quad from one seed product -- 32 quads = 128 rows per row tile, l tiles of <= 48 values
    if (m0 <= 255 && m1 <= 255 && m2 >= 1 && nv > 0) {
        pl->tc_n_lt = (m2 + TC_MAX_LT - 1) / TC_MAX_LT;
        pl->tc_lt = (((m2 + pl->tc_n_lt - 1) / pl->tc_n_lt) + 7) / 8 * 8;
        const int n_lt = pl->tc_n_lt, lt_size = pl->tc_lt;
        struct Quad { uint8_t h, k0; int8_t top; };
        std::vector<Quad> quads;
        std::vector<int32_t> quad_of((size_t)m0 * m1, -1);  // (h,k) -> quad (h-major order)
        std::vector<int8_t> top_hk((size_t)m0 * m1, -1);  // highest l tile with a valid cell of every (h,k) pair
        parallel_chunks((size_t)m0, par_h, [&](size_t h_lo, size_t h_hi) {
            for (size_t h = h_lo; h < h_hi; ++h)
                for (int k = 0; k < m1; ++k)
                    for (int l = m2 - 1; l >= 0; --l)
                        if (qcube[(h * m1 + k) * m2 + l] != 0.0) { top_hk[h * m1 + k] = (int8_t)(l / lt_size); break; }
        });
        for (int h = 0; h < m0; ++h) {
            int prev_k = -2, run_k0 = 0;
            for (int k = 0; k < m1; ++k) {
                const int top = top_hk[(size_t)h * m1 + k];
                if (top < 0) continue;
                if (k != prev_k + 1) run_k0 = k;                     // a new run of consecutive k starts its own quads
                if ((k - run_k0) % 4 == 0) quads.push_back({(uint8_t)h, (uint8_t)k, (int8_t)-1});
                Quad &qd = quads.back();
                qd.top = (int8_t)std::max<int>(qd.top, top);
                quad_of[(size_t)h * m1 + k] = (int32_t)quads.size() - 1;
                prev_k = k;
            }
        }
        // Cubes with several l tiles: quads whose highest populated l tile is the same go into the same row tiles (stable
        // sort, h-major inside a class), so that most row tiles of the outer (h,k) region have no valid cell in the upper
        // l tiles -- those (row tile, l tile) items are never launched (run_segments_tc) -- instead of every row tile
        // touching every l tile: 204 -> 154 items at maxn = 103, 360 -> 287 at maxn = 138 (the generated operand is
        // produced once per item, so items are what the contraction's cost counts).
        std::vector<int32_t> order(quads.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
        if (n_lt > 1)
            std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return quads[a].top > quads[b].top; });
        std::vector<int32_t> slot_of(quads.size());
        pl->tc_rows.reserve(quads.size() * 8 + 2 * TC_M);
        for (size_t sidx = 0; sidx < order.size(); ++sidx) {
            const Quad &qd = quads[order[sidx]];
            slot_of[order[sidx]] = (int32_t)sidx;
            for (int r = 0; r < 4; ++r) {
                pl->tc_rows.push_back(qd.h);
                pl->tc_rows.push_back((uint8_t)std::min(255, (int)qd.k0 + r));
            }
            pl->tc_ymask[qd.k0 >> 6] |= 1ull << (qd.k0 & 63);
        }
        pl->tc_n_pairs = (int32_t)(pl->tc_rows.size() / 2);
        pl->tc_n_rt = (pl->tc_n_pairs + TC_M - 1) / TC_M;
        pl->tc_rows.resize((size_t)pl->tc_n_rt * TC_M * 2, 0);
        pl->tc_ymask[0] |= 1ull;  // the padding quads of the last row tile start at k0 = 0
        {
            const int r1 = std::min(1, std::max(0, (int)pl->rows[1] - 1));
            pl->tc_ymask[r1 >> 6] |= 1ull << (r1 & 63);
        }
        pl->ent_amp_tc.resize(nv);
        pl->tc_tile_used.assign((size_t)pl->tc_n_rt * n_lt, 0);
        // row of every (h,k) pair in the packed tiling, l tile and position inside it of every l: per entry only look-ups
        std::vector<int32_t> row_of((size_t)m0 * m1, -1);
        for (size_t hk = 0; hk < row_of.size(); ++hk) {
            const int32_t qi = quad_of[hk];
            if (qi >= 0) row_of[hk] = slot_of[qi] * 4 + ((int)(hk % (size_t)m1) - (int)quads[qi].k0);
        }
        std::vector<int32_t> lt_of(m2), lin_of(m2);
        for (int l = 0; l < m2; ++l) { lt_of[l] = l / lt_size; lin_of[l] = l % lt_size; }
        parallel_chunks(nv, 100000, [&](size_t e_lo, size_t e_hi) {
            for (size_t e = e_lo; e < e_hi; ++e) {
                int h, k, l;
                if (pack_hkl) {
                    const uint32_t v = hkl[e];
                    h = (int)(v >> 20); k = (int)((v >> 10) & 1023u); l = (int)(v & 1023u);
                } else {
                    const int32_t ci = pl->cell[e];
                    l = ci % m2; k = (ci / m2) % m1; h = ci / (m2 * m1);
                }
                const int32_t pp = row_of[(size_t)h * m1 + k];
                const int32_t tile = (pp / TC_M) * n_lt + lt_of[l];
                pl->ent_amp_tc[e] = (tile * TC_M + pp % TC_M) * lt_size + lin_of[l];
            }
        });
        for (size_t e = 0; e < nv; ++e) pl->tc_tile_used[pl->ent_amp_tc[e] / (TC_M * lt_size)] = 1;
        pl->tc_n_used = 0;
        for (uint8_t u : pl->tc_tile_used) pl->tc_n_used += u;
        pl->tc_ok = true;
    }
    mark("tc tiling");
    out = pl;
    return MB200_OK;
}

static int upload_plan(mb200_ctx *ctx, SfPlan &pl) {
    if (!pl.tgs.empty()) {
        MB_TRY(pl.d_tgs.ensure(pl.tgs.size() * sizeof(TileGroupDesc)));
        MB_CUDA(cudaMemcpyAsync(pl.d_tgs.p, pl.tgs.data(), pl.tgs.size() * sizeof(TileGroupDesc),
                                cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!pl.ent_amp.empty()) {
        MB_TRY(pl.d_ent_amp.ensure(pl.ent_amp.size() * 4));
        MB_CUDA(cudaMemcpyAsync(pl.d_ent_amp.p, pl.ent_amp.data(), pl.ent_amp.size() * 4, cudaMemcpyHostToDevice,
                                ctx->stream));
    }
    if (!pl.bin_start.empty()) {
        MB_TRY(pl.d_bin_start.ensure(pl.bin_start.size() * 4));
        MB_CUDA(cudaMemcpyAsync(pl.d_bin_start.p, pl.bin_start.data(), pl.bin_start.size() * 4,
                                cudaMemcpyHostToDevice, ctx->stream));
    }
    if (pl.tc_ok) {
        MB_TRY(pl.d_tc_rows.ensure(pl.tc_rows.size()));
        MB_CUDA(cudaMemcpyAsync(pl.d_tc_rows.p, pl.tc_rows.data(), pl.tc_rows.size(), cudaMemcpyHostToDevice, ctx->stream));
        MB_TRY(pl.d_ent_amp_tc.ensure(pl.ent_amp_tc.size() * 4));
        MB_CUDA(cudaMemcpyAsync(pl.d_ent_amp_tc.p, pl.ent_amp_tc.data(), pl.ent_amp_tc.size() * 4, cudaMemcpyHostToDevice,
                                ctx->stream));
    }
    // no synchronisation: the sources are pageable, cudaMemcpyAsync returns once it has staged them
    return MB200_OK;
}

static inline bool plan_is(const SfPlan &p, const double *d, double qmin, double qmax, double thmin, double thmax, int32_t n_bins) {
    return p.dims[0] == d[0] && p.dims[1] == d[1] && p.dims[2] == d[2] && p.qmin == qmin && p.qmax == qmax && p.thmin == thmin &&
           p.thmax == thmax && p.n_bins == n_bins;
}

static std::shared_ptr<PlanPool> plan_pool_of(mb200_ctx *ctx) {
    std::lock_guard<std::mutex> lock(ctx->plan_mutex);
    if (!ctx->plan_pool) ctx->plan_pool = std::make_shared<PlanPool>();
    return ctx->plan_pool;
}

// (called under the context's call mutex; plan_mutex only guards the lists against mb200_plans_ahead on another thread)
static int get_plan(mb200_ctx *ctx, const double *dims, double qmin, double qmax, double thmin, double thmax,
                    int32_t n_bins, std::shared_ptr<SfPlan> &out) {
    {
        std::lock_guard<std::mutex> lock(ctx->plan_mutex);
        for (auto it = ctx->plans.begin(); it != ctx->plans.end(); ++it)
            if (plan_is(**it, dims, qmin, qmax, thmin, thmax, n_bins)) {
                out = *it;
                ctx->plans.splice(ctx->plans.begin(), ctx->plans, it);
                return MB200_OK;
            }
        for (auto it = ctx->plans_ahead.begin(); it != ctx->plans_ahead.end(); ++it)
            if (plan_is(**it, dims, qmin, qmax, thmin, thmax, n_bins)) {  // built ahead, not uploaded yet
                out = *it;
                ctx->plans_ahead.erase(it);
                break;
            }
    }
    if (!out) MB_TRY(build_plan(dims, qmin, qmax, thmin, thmax, n_bins, out, plan_pool_of(ctx)));
    MB_TRY(upload_plan(ctx, *out));
    std::vector<std::shared_ptr<SfPlan>> retired;  // released outside the lock
    {
        std::lock_guard<std::mutex> lock(ctx->plan_mutex);
        ctx->plans.push_front(out);
        while (ctx->plans.size() > ctx->plan_cap) { retired.push_back(ctx->plans.back()); ctx->plans.pop_back(); }
    }
    return MB200_OK;
}

static void compute_ff2(const SfPlan &pl, const double *cm, int n_groups, std::vector<double> &ff2, bool threads);

// The q-plans of all frames of a call, built ahead: an NPT trajectory has a new cell -- a new plan -- in every frame, and
// one plan costs about a millisecond of host time (maxn = 32).  Boxes that are not cached are built on the host threads
// side by side (with their squared form factors when `cm` is given), then uploaded one after the other; the calls'
// get_plan lookups then hit the cache.  dims_of(f, d[3]) yields the cell of frame f.
// Boxes of `n_frames` frames that have no plan yet -- neither cached nor built ahead -- each listed once.
template <typename DimsOf>
static std::vector<std::array<double, 3>> missing_plans(mb200_ctx *ctx, size_t n_frames, DimsOf dims_of, double qmin, double qmax,
                                                        double thmin, double thmax, int32_t n_bins) {
    std::vector<std::array<double, 3>> miss;
    std::lock_guard<std::mutex> lock(ctx->plan_mutex);
    double last[3] = {-1.0, -1.0, -1.0};
    for (size_t f = 0; f < n_frames; ++f) {
        double d[3];
        dims_of(f, d);
        if (d[0] == last[0] && d[1] == last[1] && d[2] == last[2]) continue;  // NVT: one look-up per call
        last[0] = d[0]; last[1] = d[1]; last[2] = d[2];
        bool known = false;
        for (const auto &sp : ctx->plans)
            if (plan_is(*sp, d, qmin, qmax, thmin, thmax, n_bins)) { known = true; break; }
        for (size_t i = 0; i < ctx->plans_ahead.size() && !known; ++i)
            known = plan_is(*ctx->plans_ahead[i], d, qmin, qmax, thmin, thmax, n_bins);
        for (size_t i = 0; i < miss.size() && !known; ++i) known = miss[i][0] == d[0] && miss[i][1] == d[1] && miss[i][2] == d[2];
        if (!known) miss.push_back({d[0], d[1], d[2]});
    }
    return miss;
}

// Host side of the plans of `miss`, side by side on the host threads (with their squared form factors when `cm` is given).
static int build_plans(mb200_ctx *ctx, const std::vector<std::array<double, 3>> &miss, double qmin, double qmax, double thmin,
                       double thmax, int32_t n_bins, const double *cm, int n_groups, std::vector<std::shared_ptr<SfPlan>> &built) {
    const std::shared_ptr<PlanPool> pool = plan_pool_of(ctx);
    built.assign(miss.size(), nullptr);
    std::vector<int> rc(miss.size(), MB200_OK);
    parallel_chunks(miss.size(), 2, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            rc[i] = build_plan(miss[i].data(), qmin, qmax, thmin, thmax, n_bins, built[i], pool);
            if (rc[i] == MB200_OK && cm && n_groups > 0) {
                compute_ff2(*built[i], cm, n_groups, built[i]->ff2_host, false);
                built[i]->ff2_host_key.assign(cm, cm + (size_t)n_groups * 20);
            }
        }
    });
    for (int r : rc)
        if (r != MB200_OK) return r;
    return MB200_OK;
}

// The q-plans of all frames of a call: an NPT trajectory has a new cell -- a new plan -- in every frame, and one plan
// costs about a millisecond of host time (maxn = 32).  Plans that mb200_plans_ahead did not build already (on the
// staging thread, under the previous call) are built here side by side; the calls' get_plan lookups then find them.
// dims_of(f, d[3]) yields the cell of frame f.
template <typename DimsOf>
static int prepare_plans(mb200_ctx *ctx, size_t n_frames, DimsOf dims_of, double qmin, double qmax, double thmin, double thmax,
                         int32_t n_bins, const double *cm, int n_groups) {
    {
        std::lock_guard<std::mutex> lock(ctx->plan_mutex);
        ctx->plan_cap = std::max<size_t>(ctx->plan_cap, std::min<size_t>(n_frames + 16, 288));
    }
    const auto miss = missing_plans(ctx, n_frames, dims_of, qmin, qmax, thmin, thmax, n_bins);
    if (miss.size() < 2) return MB200_OK;  // a single new cell: get_plan builds it (with its own threads if it is large)
    std::vector<std::shared_ptr<SfPlan>> built;
    MB_TRY(build_plans(ctx, miss, qmin, qmax, thmin, thmax, n_bins, cm, n_groups, built));
    std::lock_guard<std::mutex> lock(ctx->plan_mutex);
    for (auto &pl : built) ctx->plans_ahead.push_back(pl);  // uploaded by get_plan when their frame comes up
    return MB200_OK;
}

extern "C" int mb200_plans_ahead(mb200_ctx *ctx, const double *boxes, int64_t n_frames, double qmin, double qmax, double thetamin,
                                 double thetamax, int32_t n_bins, const double *cm_params, int32_t n_groups) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    if (!boxes || n_frames < 0 || n_bins < 0 || (n_groups > 0 && !cm_params)) return fail(MB200_EINVAL, "invalid argument to mb200_plans_ahead");
    // host work only, and NOT under the call mutex: it runs beside the compute call of the previous batch
    const auto miss = missing_plans(ctx, (size_t)n_frames, [&](size_t f, double *d) { for (int i = 0; i < 3; ++i) d[i] = boxes[f * 3 + i]; },
                                    qmin, qmax, thetamin, thetamax, n_bins);
    if (miss.empty()) return MB200_OK;
    std::vector<std::shared_ptr<SfPlan>> built;
    MB_TRY(build_plans(ctx, miss, qmin, qmax, thetamin, thetamax, n_bins, cm_params, n_groups, built));
    std::lock_guard<std::mutex> lock(ctx->plan_mutex);
    for (auto &pl : built) ctx->plans_ahead.push_back(pl);
    while (ctx->plans_ahead.size() > 512) ctx->plans_ahead.erase(ctx->plans_ahead.begin());  // announced, never computed
    return MB200_OK;
}

// Squared form factor per entry and group (lib/math.py:692-700): f = sum_i a_i exp(-b_i (q/4pi)^2) + c,
// up to two element sets per group for united atoms (math.py:671-684).
static void compute_ff2(const SfPlan &pl, const double *cm, int n_groups, std::vector<double> &ff2, bool threads) {
    const size_t nv = pl.qrr.size();
    ff2.resize((size_t)n_groups * nv);
    for (int gidx = 0; gidx < n_groups; ++gidx)
        parallel_chunks(nv, threads ? (size_t)50000 : nv + 1, [&](size_t e_lo, size_t e_hi) {
        for (size_t e = e_lo; e < e_hi; ++e) {
            const double q = pl.qrr[e] / (4 * M_PI);
            const double q2 = q * q;
            double f = 0.0;
            for (int s = 0; s < 2; ++s) {
                const double *p = cm + ((size_t)gidx * 2 + s) * 10;
                if (p[0] == 0.0) continue;
                double fs = 0.0;
                for (int i = 0; i < 4; ++i) fs += p[1 + i] * std::exp(-p[5 + i] * q2);
                f += p[0] * (fs + p[9]);
            }
            ff2[(size_t)gidx * nv + e] = f * f;
        }
        });
}

static int plan_ff2(mb200_ctx *ctx, SfPlan &pl, const double *cm, int n_groups) {
    std::vector<double> key(cm, cm + (size_t)n_groups * 20);
    if (pl.ff2_groups == n_groups && pl.ff2_key == key) return MB200_OK;
    if (pl.ff2_host_key != key) {  // not computed ahead by prepare_plans
        compute_ff2(pl, cm, n_groups, pl.ff2_host, true);
        pl.ff2_host_key = key;
    }
    if (!pl.ff2_host.empty()) {
        MB_TRY(pl.d_ff2.ensure(pl.ff2_host.size() * 8));
        // (pageable source: the call returns once it has staged the data)
        MB_CUDA(cudaMemcpyAsync(pl.d_ff2.p, pl.ff2_host.data(), pl.ff2_host.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    pl.ff2_key.swap(key);
    pl.ff2_groups = n_groups;
    return MB200_OK;
}

// launches k_sf_tables with the l-tile size of the jobs' digit images as a compile-time constant when they agree on one
// (the positions of a batch are read by this launch only: a prefetch slot they came from is released right behind it, so
// that the next batch's host -> device copy runs under the contraction -- also when the batch after is already queued)
static void launch_tables(mb200_ctx *ctx, dim3 grid, const TableJob *d_jobs, const std::vector<int> &img_lts) {
    int lt = -1;
    for (int v : img_lts) {
        if (v <= 0) continue;  // no digit image in this job
        lt = (lt < 0 || lt == v) ? v : 0;
        if (lt == 0) break;
    }
    switch (lt) {
        case 32: k_sf_tables<32><<<grid, 128, 0, ctx->stream>>>(d_jobs); break;
        case 40: k_sf_tables<40><<<grid, 128, 0, ctx->stream>>>(d_jobs); break;
        case 48: k_sf_tables<48><<<grid, 128, 0, ctx->stream>>>(d_jobs); break;
        default: k_sf_tables<0><<<grid, 128, 0, ctx->stream>>>(d_jobs); break;
    }
    if (ctx->release_after_tables) {
        ctx->release_after_tables->release();
        ctx->release_after_tables = nullptr;
    }
}

// ------------------------------------------------------------------------------------------
// host: engine
// ------------------------------------------------------------------------------------------
struct SegSpec {       // one contraction problem (one reference structure_factor call)
    SfPlan *plan;
    TableJob job;       // table inputs (tab pointers filled by the engine)
    int share_yz_with;  // >= 0: reuse Ey/Ez tables of that earlier segment (same positions)
    int ff2_group;      // index into plan ff2, -1 none
    double s_scale = 1.0;  // see SegDesc::s_scale
};

// max|w| of weight sets on the device, as the bit pattern of |w| (unsigned order: finite < inf < NaN)
__global__ void __launch_bounds__(256) k_sf_wmax(const double *__restrict__ w, long long n, unsigned long long *__restrict__ out) {
    const double *ws = w + (size_t)blockIdx.y * (size_t)n;
    unsigned long long m = 0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(ws[i]) & 0x7fffffffffffffffull;
        m = b > m ? b : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_down_sync(0xffffffffu, m, o);
        m = v > m ? v : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out + blockIdx.y, m);
}

// Power-of-two rescaling of a weight set with max|w| = wmax for the split-precision mode: w * wscale has its maximum in
// (0.5, 1], S is multiplied back by s_scale = wscale^-2 (both exact).  false: the set cannot take that mode (non-finite
// weights, or a scale whose square leaves the double range) and the call uses the FP64 kernel.
static bool weight_scale(double wmax, double &wscale, double &s_scale) {
    wscale = 1.0; s_scale = 1.0;
    if (!(wmax <= 1.7e308)) return false;
    if (wmax == 0.0) return true;
    int ex = 0;
    const double f = std::frexp(wmax, &ex);  // wmax = f * 2^ex, 0.5 <= f < 1
    if (f == 0.5) --ex;                      // an exact power of two stays the maximum 1
    if (ex == 0) return true;
    if (ex > 480 || ex < -480) return false;
    wscale = std::ldexp(1.0, -ex);
    s_scale = std::ldexp(1.0, 2 * ex);
    return true;
}




// Split-precision variant of run_segments: FP64 x / y tables + int8 Z image (one kernel) -> tcgen05 contraction.
static int run_segments_tc(mb200_ctx *ctx, std::vector<SegSpec> &specs, int32_t block, int32_t n_blocks,
                           std::vector<SegDesc> &segs_host) {
    const int ns = (int)specs.size();
    // ---- workspaces: FP64 tables (ld padded to the 16-atom K-step), Z images, partial amplitudes
    size_t tab_elems = 0, img_bytes = 0, amp_elems = 0;
    std::vector<size_t> off_x(ns), off_y(ns), off_z(ns), off_img(ns), off_amp(ns);
    std::vector<int> n_split(ns, 1);
    std::vector<TcItem> items;
    int64_t atomq = 0, n_mma = 0;
    int64_t total_tiles = 0;
    for (int s = 0; s < ns; ++s) total_tiles += ((int64_t)specs[s].plan->tc_n_used + n_blocks - 1) / n_blocks;
    // Items are handed out in list order from one counter and all take about the same time, so the launch ends with up to
    // one item of idle time per CTA (3 % on the bench workload): the segments at the END of the list are cut into twice as
    // many atom chunks until the last 2 x n_sm items are half-size.
    std::vector<int> split_factor(ns, 1);
    {
        int64_t tail_items = 0;
        for (int s = ns - 1; s >= 0 && tail_items < 2 * (int64_t)ctx->n_sm; --s) {
            const SfPlan &pl = *specs[s].plan;
            const int ld = (specs[s].job.n + TC_KA - 1) / TC_KA * TC_KA;
            const int64_t tiles = ((int64_t)pl.tc_n_used + n_blocks - 1) / n_blocks;
            const int base = std::max<int>(1, (ld + TC_CHUNK - 1) / TC_CHUNK);
            if (ld / (2 * base) >= 64 * TC_KA) {
                split_factor[s] = 2;
                tail_items += 2 * base * tiles;
            } else {
                tail_items += base * tiles;
            }
        }
    }
    for (int s = 0; s < ns; ++s) {
        SegSpec &sp = specs[s];
        const SfPlan &pl = *sp.plan;
        const int ld = (sp.job.n + TC_KA - 1) / TC_KA * TC_KA;
        sp.job.ld = ld;
        for (int a = 0; a < 3; ++a) sp.job.rows[a] = pl.rows[a];
        off_x[s] = tab_elems; tab_elems += (size_t)pl.rows[0] * ld;
        if (sp.share_yz_with >= 0) {
            off_y[s] = off_y[sp.share_yz_with]; off_z[s] = off_z[sp.share_yz_with]; off_img[s] = off_img[sp.share_yz_with];
        } else {
            off_y[s] = tab_elems; tab_elems += (size_t)pl.rows[1] * ld;
            off_z[s] = off_y[s];  // no FP64 z table in this mode: k_sf_tables writes the digit image instead
            off_img[s] = img_bytes;
            img_bytes += (size_t)pl.tc_n_lt * (ld / TC_KA) * (size_t)(TC_S * 2 * pl.tc_lt * 32);
        }
        // atom chunks: at most TC_CHUNK atoms, enough items to fill the SMs a few times over
        const int64_t tiles = (int64_t)pl.tc_n_rt * pl.tc_n_lt;
        int ns_s = std::max<int>(1, (ld + TC_CHUNK - 1) / TC_CHUNK) * split_factor[s];
        const int64_t want = ((int64_t)ctx->n_sm * 3 + total_tiles - 1) / std::max<int64_t>(total_tiles, 1);
        ns_s = (int)std::max<int64_t>(ns_s, std::min<int64_t>(want, ld / (64 * TC_KA)));
        ns_s = std::max(ns_s, 1);
        const int steps_total = ld / TC_KA;
        const int steps_per = (steps_total + ns_s - 1) / ns_s;
        ns_s = steps_per ? (steps_total + steps_per - 1) / steps_per : 1;
        n_split[s] = std::max(ns_s, 1);
        off_amp[s] = amp_elems;
        amp_elems += (size_t)n_split[s] * tiles * TC_M * pl.tc_lt;
        if (ld == 0) continue;
        for (int sp_i = 0; sp_i < n_split[s]; ++sp_i) {
            const int j0 = sp_i * steps_per * TC_KA, j1 = std::min(ld, (sp_i + 1) * steps_per * TC_KA);
            if (j1 <= j0) continue;
            int used_idx = 0;  // q-blocks deal the populated (row tile, l tile) blocks round robin
            for (int rt = 0; rt < pl.tc_n_rt; ++rt)
                for (int lt = 0; lt < pl.tc_n_lt; ++lt) {
                    if (!pl.tc_tile_used[(size_t)rt * pl.tc_n_lt + lt]) continue;  // no valid q in this block: never launched
                    if (used_idx++ % n_blocks != block) continue;
                    items.push_back({s, rt, lt, j0, j1, sp_i});
                    n_mma += (int64_t)((j1 - j0) / TC_KA) * 15;
                }
        }
        atomq += (int64_t)sp.job.n * (int64_t)pl.cell.size();
    }
    MB_TRY(ctx->d_tables.ensure(std::max<size_t>(tab_elems, 1) * 16));
    MB_TRY(ctx->d_bimg.ensure(std::max<size_t>(img_bytes, 16)));
    MB_TRY(ctx->d_amp.ensure(std::max<size_t>(amp_elems, 1) * 16));
    double2 *tab = ctx->d_tables.as<double2>();
    // q-blocks: per distinct plan, which (row tile, l tile) blocks this rank computes -- the finalisers and the fold pass
    // skip the others, so nothing outside the rank's share of the amplitude workspace is ever touched
    std::vector<uint8_t> owned_host;
    std::vector<size_t> owned_off(ns, 0);
    if (n_blocks > 1) {
        std::map<const SfPlan *, size_t> seen;
        for (int s = 0; s < ns; ++s) {
            const SfPlan &pl = *specs[s].plan;
            auto it = seen.find(&pl);
            if (it != seen.end()) { owned_off[s] = it->second; continue; }
            owned_off[s] = seen[&pl] = owned_host.size();
            int used_idx = 0;
            for (size_t t = 0; t < pl.tc_tile_used.size(); ++t)
                owned_host.push_back(pl.tc_tile_used[t] && (used_idx++ % n_blocks) == block ? 1 : 0);
        }
        MB_TRY(ctx->d_owned.ensure(std::max<size_t>(owned_host.size(), 16)));
    }
    // empty segments (no atoms) are never written: they must read as zero
    if (n_blocks == 1 && items.empty())
        MB_CUDA(cudaMemsetAsync(ctx->d_amp.p, 0, std::max<size_t>(amp_elems, 1) * 16, ctx->stream));
    std::vector<TcSeg> tsegs(ns);
    std::vector<FoldDesc> folds(ns);
    bool any_fold = false;
    for (int s = 0; s < ns; ++s) {
        SegSpec &sp = specs[s];
        const SfPlan &pl = *sp.plan;
        SegDesc &sd = segs_host[s];
        sd.ld = sp.job.ld; sd.n_split = n_split[s]; sd.n_wt = pl.n_wt; sd.n_valid = (int32_t)pl.cell.size();
        sd.amp_stride = (long long)pl.tc_n_rt * pl.tc_n_lt * TC_M * pl.tc_lt;
        sd.s_scale = sp.s_scale;
        // many large slabs: fold them into slab 0 with one streaming pass (k_sf_fold_splits) before the finalisers gather
        folds[s].stride = sd.amp_stride;
        folds[s].n_split = (n_split[s] >= 2 && (size_t)n_split[s] * (size_t)sd.amp_stride * 16 >= ((size_t)96 << 20)) ? n_split[s] : 1;
        folds[s].tile_elems = TC_M * pl.tc_lt;
        folds[s].owned = n_blocks > 1 ? ctx->d_owned.as<uint8_t>() + owned_off[s] : nullptr;
        sd.tile_owned = folds[s].owned;
        sd.tile_elems = TC_M * pl.tc_lt;
        if (folds[s].n_split > 1) { sd.n_split = 1; any_fold = true; }
        for (int a = 0; a < 3; ++a) sd.rows[a] = pl.rows[a];
        sd.tgs = nullptr;
        sd.ent_amp = pl.d_ent_amp_tc.as<int32_t>();
        sd.bin_start = pl.d_bin_start.as<int32_t>();
        sd.ff2 = (sp.ff2_group >= 0 && pl.d_ff2.p) ? pl.d_ff2.as<double>() + (size_t)sp.ff2_group * sd.n_valid : nullptr;
        sd.ex = tab + off_x[s]; sd.ey = tab + off_y[s]; sd.ez = tab + off_z[s];
        sd.amp = ctx->d_amp.as<double2>() + off_amp[s];
        folds[s].amp = sd.amp;
        sp.job.tab[0] = tab + off_x[s];
        sp.job.tab[1] = sp.share_yz_with >= 0 ? nullptr : tab + off_y[s];
        sp.job.tab[2] = sp.share_yz_with >= 0 ? nullptr : tab + off_z[s];
        if (sp.job.ld == 0 && amp_elems)
            MB_CUDA(cudaMemsetAsync(sd.amp, 0, (size_t)n_split[s] * sd.amp_stride * 16, ctx->stream));
        TcSeg &ts = tsegs[s];
        ts.ex = sd.ex; ts.ey = sd.ey;
        ts.bimg = ctx->d_bimg.as<int8_t>() + off_img[s];
        ts.rows = pl.d_tc_rows.as<uint8_t>();
        ts.amp = sd.amp;
        ts.rows0 = pl.rows[0]; ts.rows1 = pl.rows[1]; ts.ld = sp.job.ld;
        ts.n_rt = pl.tc_n_rt; ts.n_lt = pl.tc_n_lt; ts.lt_size = pl.tc_lt; ts.n_pairs = pl.tc_n_pairs; ts.n_split = n_split[s];
        sp.job.use_y_mask = 1;
        for (int i = 0; i < 4; ++i) sp.job.y_mask[i] = pl.tc_ymask[i];
        if (sp.share_yz_with < 0) {  // the z axis goes straight into the digit image (k_sf_tables)
            sp.job.bimg = ctx->d_bimg.as<int8_t>() + off_img[s];
            sp.job.img_n_lt = pl.tc_n_lt; sp.job.img_lt = pl.tc_lt; sp.job.img_maxn = pl.maxn[2];
        }
    }
    // ---- descriptors to the device through the pinned staging buffer
    const size_t b_jobs = (size_t)ns * sizeof(TableJob), b_segs = (size_t)ns * sizeof(SegDesc),
                 b_tsegs = ((size_t)ns * sizeof(TcSeg) + 15) / 16 * 16 + (size_t)ns * sizeof(FoldDesc),  // TcSeg[ns] | FoldDesc[ns]
                 b_fold_off = ((size_t)ns * sizeof(TcSeg) + 15) / 16 * 16, b_items = items.size() * sizeof(TcItem);
    const size_t b_owned = (owned_host.size() + 15) / 16 * 16;
    // two staging buffers alternate, so that this call can be prepared while the previous one still runs (its copies out
    // of the other buffer may not have executed yet); a buffer is rewritten once the copies that read it are done
    const int par = ctx->tc_parity;
    ctx->tc_parity ^= 1;
    if (ctx->tc_stage_used[par]) MB_CUDA(cudaEventSynchronize(ctx->tc_stage_ev[par]));
    MB_TRY(ctx->h_tc_stage[par].ensure(b_jobs + b_segs + b_tsegs + b_items + b_owned + 64));
    char *hs = ctx->h_tc_stage[par].as<char>();
    for (int s = 0; s < ns; ++s) std::memcpy(hs + (size_t)s * sizeof(TableJob), &specs[s].job, sizeof(TableJob));
    std::memcpy(hs + b_jobs, segs_host.data(), b_segs);
    std::memcpy(hs + b_jobs + b_segs, tsegs.data(), (size_t)ns * sizeof(TcSeg));
    std::memcpy(hs + b_jobs + b_segs + b_fold_off, folds.data(), (size_t)ns * sizeof(FoldDesc));
    if (b_items) std::memcpy(hs + b_jobs + b_segs + b_tsegs, items.data(), b_items);
    if (!owned_host.empty()) {
        std::memcpy(hs + b_jobs + b_segs + b_tsegs + b_items, owned_host.data(), owned_host.size());
        MB_CUDA(cudaMemcpyAsync(ctx->d_owned.p, hs + b_jobs + b_segs + b_tsegs + b_items, owned_host.size(), cudaMemcpyHostToDevice,
                                ctx->stream));
    }
    MB_TRY(ctx->d_jobs.ensure(b_jobs));
    MB_TRY(ctx->d_segs.ensure(b_segs));
    MB_TRY(ctx->d_tcdesc.ensure(b_tsegs + 64));
    MB_TRY(ctx->d_items.ensure(std::max<size_t>(b_items, 16)));
    MB_TRY(ctx->d_counter.ensure(16));
    MB_CUDA(cudaMemcpyAsync(ctx->d_jobs.p, hs, b_jobs, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaMemcpyAsync(ctx->d_segs.p, hs + b_jobs, b_segs, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaMemcpyAsync(ctx->d_tcdesc.p, hs + b_jobs + b_segs, b_tsegs, cudaMemcpyHostToDevice, ctx->stream));
    if (b_items)
        MB_CUDA(cudaMemcpyAsync(ctx->d_items.p, hs + b_jobs + b_segs + b_tsegs, b_items, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaEventRecord(ctx->tc_stage_ev[par], ctx->stream));
    ctx->tc_stage_used[par] = true;
    // ---- launches
    int max_ld = 0;
    for (int s = 0; s < ns; ++s) max_ld = std::max(max_ld, specs[s].job.ld);
    int64_t launched = 0;
    if (max_ld > 0) {
        std::vector<int> lts(ns);
        for (int s = 0; s < ns; ++s) lts[s] = specs[s].job.bimg ? specs[s].job.img_lt : 0;
        launch_tables(ctx, dim3(3 * ((max_ld + 127) / 128), 1, ns), ctx->d_jobs.as<TableJob>(), lts);
        MB_CUDA(cudaGetLastError());
        ++launched;
    }
    if (!items.empty()) {
        if (!ctx->tc_attr_set) {
            MB_CUDA(cudaFuncSetAttribute(k_sf_contract_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
            MB_CUDA(cudaFuncSetAttribute(k_sf_contract_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
            ctx->tc_attr_set = true;
        }
        // l tiles of at most 32 values leave room in tensor memory for the generated A operand (sfactor_tc.cuh)
        bool a_in_tmem = ctx->tc_a_tmem;
        for (int s = 0; s < ns; ++s) a_in_tmem = a_in_tmem && specs[s].plan->tc_lt <= TC_ATM_MAX_LT;
        if (ctx->timing) MB_CUDA(cudaEventRecord(ctx->tc_ev0[par], ctx->stream));
        MB_CUDA(cudaMemsetAsync(ctx->d_counter.p, 0, 4, ctx->stream));
        const unsigned n_cta = (unsigned)std::min<size_t>(items.size(), (size_t)ctx->n_sm);
        if (a_in_tmem)
            k_sf_contract_tc<true><<<n_cta, TC_THREADS, TC_SMEM, ctx->stream>>>(ctx->d_items.as<TcItem>(), (int)items.size(),
                                                                               ctx->d_tcdesc.as<TcSeg>(), ctx->d_counter.as<int>());
        else
            k_sf_contract_tc<false><<<n_cta, TC_THREADS, TC_SMEM, ctx->stream>>>(ctx->d_items.as<TcItem>(), (int)items.size(),
                                                                                ctx->d_tcdesc.as<TcSeg>(), ctx->d_counter.as<int>());
        MB_CUDA(cudaGetLastError());
        if (ctx->timing) MB_CUDA(cudaEventRecord(ctx->tc_ev1[par], ctx->stream));
        ++launched;
    }
    ctx->t_ev0 = ctx->tc_ev0[par]; ctx->t_ev1 = ctx->tc_ev1[par]; ctx->t_parity = par;
    if (any_fold) {
        long long max_stride = 0;
        for (int s = 0; s < ns; ++s)
            if (folds[s].n_split > 1) max_stride = std::max(max_stride, folds[s].stride);
        const unsigned gx = (unsigned)std::min<long long>((max_stride + 255) / 256, (long long)ctx->n_sm * 8);
        k_sf_fold_splits<<<dim3(gx, (unsigned)ns), 256, 0, ctx->stream>>>(
            reinterpret_cast<const FoldDesc *>(ctx->d_tcdesc.as<char>() + b_fold_off));
        MB_CUDA(cudaGetLastError());
        ++launched;
    }
    ctx->launches += launched;
    ctx->stats[0] = launched;
    ctx->stats[1] = ns ? (int64_t)specs[0].plan->cell.size() : 0;
    ctx->stats[2] = (int64_t)items.size();
    ctx->stats[3] = n_mma;
    ctx->stats[4] = atomq;
    ctx->stats[6] = ns ? n_split[0] : 0;
    ctx->stats[7] = ns ? specs[0].plan->tc_n_rt * specs[0].plan->tc_n_lt : 0;
    return MB200_OK;
}

// Runs tables + contraction for all segments; leaves SegDesc array on the device (ctx->d_segs).
static int run_segments(mb200_ctx *ctx, std::vector<SegSpec> &specs, int32_t block, int32_t n_blocks,
                        std::vector<SegDesc> &segs_host) {
    const int ns = (int)specs.size();
    segs_host.assign(ns, SegDesc());
    if (ns == 0) return MB200_OK;
    if (n_blocks < 1 || block < 0 || block >= n_blocks) return fail(MB200_EINVAL, "invalid q-block %d of %d", block, n_blocks);
    if (ctx->precision == 1 && !ctx->force_fp64) {  // split-precision tensor-core mode when every plan can be tiled for it
        bool ok = true;
        for (int s = 0; s < ns; ++s) ok = ok && specs[s].plan->tc_ok;
        if (ok) return run_segments_tc(ctx, specs, block, n_blocks, segs_host);
    }

    // ---- table workspace layout
    size_t tab_elems = 0;
    std::vector<size_t> off_x(ns), off_y(ns), off_z(ns);
    for (int s = 0; s < ns; ++s) {
        SegSpec &sp = specs[s];
        const int ld = (sp.job.n + KT - 1) / KT * KT;
        sp.job.ld = ld;
        for (int a = 0; a < 3; ++a) sp.job.rows[a] = sp.plan->rows[a];
        sp.job.use_y_mask = 0;  // the FP64 kernel reads every row
        sp.job.bimg = nullptr;
        off_x[s] = tab_elems; tab_elems += (size_t)sp.plan->rows[0] * ld;
        if (sp.share_yz_with >= 0) {
            off_y[s] = off_y[sp.share_yz_with]; off_z[s] = off_z[sp.share_yz_with];
        } else {
            off_y[s] = tab_elems; tab_elems += (size_t)sp.plan->rows[1] * ld;
            off_z[s] = tab_elems; tab_elems += (size_t)sp.plan->rows[2] * ld;
        }
    }
    MB_TRY(ctx->d_tables.ensure(std::max<size_t>(tab_elems, 1) * 16));
    double2 *tab = ctx->d_tables.as<double2>();

    // ---- split / item plan
    int64_t total_tg = 0;
    for (int s = 0; s < ns; ++s) {
        const int ntg = (int)specs[s].plan->tgs.size();
        for (int g = 0; g < ntg; ++g) total_tg += (g % n_blocks == block) && specs[s].job.ld > 0;
    }
    const int64_t target_ctas = (int64_t)ctx->n_sm * 3 * 6;  // ~6 items per resident CTA (dynamic queue)
    std::vector<ItemDesc> items;
    size_t amp_elems = 0;
    std::vector<size_t> off_amp(ns);
    int64_t mma_steps = 0, atomq = 0;
    int max_split = 1;
    for (int s = 0; s < ns; ++s) {
        SegSpec &sp = specs[s];
        const int ld = sp.job.ld;
        const int n_stage_total = ld / KT;
        int n_split = 1;
        if (total_tg > 0 && total_tg < target_ctas) n_split = (int)((target_ctas + total_tg - 1) / total_tg);
        n_split = std::max(1, std::min(n_split, n_stage_total / 32));  // >= 256 atoms per split
        const int stages_per = n_split ? (n_stage_total + n_split - 1) / n_split : 0;
        if (stages_per > 0) n_split = (n_stage_total + stages_per - 1) / stages_per;
        max_split = std::max(max_split, n_split);
        off_amp[s] = amp_elems;
        amp_elems += (size_t)n_split * sp.plan->n_wt * 512;
        SegDesc &sd = segs_host[s];
        sd.ld = ld; sd.n_split = n_split; sd.n_wt = sp.plan->n_wt; sd.n_valid = (int32_t)sp.plan->cell.size();
        sd.amp_stride = (long long)sp.plan->n_wt * 512;
        sd.s_scale = sp.s_scale;
        for (int a = 0; a < 3; ++a) sd.rows[a] = sp.plan->rows[a];
        sd.tgs = sp.plan->d_tgs.as<TileGroupDesc>();
        sd.ent_amp = sp.plan->d_ent_amp.as<int32_t>();
        sd.bin_start = sp.plan->d_bin_start.as<int32_t>();
        sd.ff2 = (sp.ff2_group >= 0 && sp.plan->d_ff2.p) ? sp.plan->d_ff2.as<double>() + (size_t)sp.ff2_group * sd.n_valid : nullptr;
        sd.ex = tab + off_x[s]; sd.ey = tab + off_y[s]; sd.ez = tab + off_z[s];
        sp.job.tab[0] = tab + off_x[s];
        sp.job.tab[1] = sp.share_yz_with >= 0 ? nullptr : tab + off_y[s];
        sp.job.tab[2] = sp.share_yz_with >= 0 ? nullptr : tab + off_z[s];
        if (ld == 0) continue;
        const int ntg = (int)sp.plan->tgs.size();
        for (int sp_i = 0; sp_i < n_split; ++sp_i) {
            const int j0 = sp_i * stages_per * KT, j1 = std::min(ld, (sp_i + 1) * stages_per * KT);
            if (j1 <= j0) continue;
            for (int g = 0; g < ntg; ++g) {
                if (g % n_blocks != block) continue;
                items.push_back({s, g, j0, j1, sp_i});
                int tiles = 0;
                for (int w = 0; w < sp.plan->tgs[g].n_warp; ++w) tiles += __builtin_popcount(sp.plan->tgs[g].wmask[w]);
                mma_steps += (int64_t)tiles * 3 * ((j1 - j0) / 4);
            }
        }
        atomq += (int64_t)sp.job.n * sd.n_valid;
    }
    MB_TRY(ctx->d_amp.ensure(std::max<size_t>(amp_elems, 1) * 16));
    for (int s = 0; s < ns; ++s) segs_host[s].amp = ctx->d_amp.as<double2>() + off_amp[s];
    if (n_blocks > 1 && amp_elems) {
        MB_CUDA(cudaMemsetAsync(ctx->d_amp.p, 0, amp_elems * 16, ctx->stream));
    } else {
        // splits beyond a segment's populated range never exist; unowned tiles never read. Zero-length
        // segments (n = 0) are never written, so clear their amplitude blocks.
        for (int s = 0; s < ns; ++s)
            if (specs[s].job.ld == 0 && specs[s].plan->n_wt)
                MB_CUDA(cudaMemsetAsync(segs_host[s].amp, 0, (size_t)segs_host[s].n_split * specs[s].plan->n_wt * 512 * 16, ctx->stream));
    }

    // ---- upload descriptors (pinned staging keeps the copies asynchronous and ordered)
    const size_t b_jobs = (size_t)ns * sizeof(TableJob), b_segs = (size_t)ns * sizeof(SegDesc),
                 b_items = items.size() * sizeof(ItemDesc);
    MB_TRY(ctx->h_stage.ensure(b_jobs + b_segs + b_items + 64));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));  // previous call may still read the staging buffer
    char *hs = ctx->h_stage.as<char>();
    for (int s = 0; s < ns; ++s) std::memcpy(hs + (size_t)s * sizeof(TableJob), &specs[s].job, sizeof(TableJob));
    std::memcpy(hs + b_jobs, segs_host.data(), b_segs);
    if (b_items) std::memcpy(hs + b_jobs + b_segs, items.data(), b_items);
    MB_TRY(ctx->d_jobs.ensure(b_jobs));
    MB_TRY(ctx->d_segs.ensure(b_segs));
    MB_TRY(ctx->d_items.ensure(std::max<size_t>(b_items, 16)));
    MB_TRY(ctx->d_counter.ensure(16));
    MB_CUDA(cudaMemcpyAsync(ctx->d_jobs.p, hs, b_jobs, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaMemcpyAsync(ctx->d_segs.p, hs + b_jobs, b_segs, cudaMemcpyHostToDevice, ctx->stream));
    if (b_items)
        MB_CUDA(cudaMemcpyAsync(ctx->d_items.p, hs + b_jobs + b_segs, b_items, cudaMemcpyHostToDevice, ctx->stream));

    // ---- launch
    int max_ld = 0;
    for (int s = 0; s < ns; ++s) max_ld = std::max(max_ld, specs[s].job.ld);
    int64_t launched = 0;
    if (max_ld > 0) {
        dim3 grid(3 * ((max_ld + 127) / 128), 1, ns);
        launch_tables(ctx, grid, ctx->d_jobs.as<TableJob>(), std::vector<int>());  // FP64 mode: no digit images
        MB_CUDA(cudaGetLastError());
        ++launched;
    }
    if (!items.empty()) {
        if (!ctx->sf_attr_set) {  // per device: a process may own several contexts
            MB_CUDA(cudaFuncSetAttribute(k_sf_contract, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
            ctx->sf_attr_set = true;
        }
        const int tpar = ctx->tc_parity;  // timing events alternate like the split-precision path's (deferred batches)
        ctx->tc_parity ^= 1;
        if (ctx->timing) MB_CUDA(cudaEventRecord(ctx->tc_ev0[tpar], ctx->stream));
        MB_CUDA(cudaMemsetAsync(ctx->d_counter.p, 0, 4, ctx->stream));
        const unsigned n_cta = (unsigned)std::min<size_t>(items.size(), (size_t)ctx->n_sm * 3);
        k_sf_contract<<<n_cta, NTHREADS, SMEM_TOTAL, ctx->stream>>>(ctx->d_items.as<ItemDesc>(), (int)items.size(),
                                                                   ctx->d_segs.as<SegDesc>(), ctx->d_counter.as<int>());
        MB_CUDA(cudaGetLastError());
        if (ctx->timing) MB_CUDA(cudaEventRecord(ctx->tc_ev1[tpar], ctx->stream));
        ctx->t_ev0 = ctx->tc_ev0[tpar]; ctx->t_ev1 = ctx->tc_ev1[tpar]; ctx->t_parity = tpar;
        ++launched;
    }
    ctx->launches += launched;
    ctx->stats[0] = launched;
    ctx->stats[1] = ns ? (int64_t)specs[0].plan->cell.size() : 0;
    ctx->stats[2] = (int64_t)items.size();
    ctx->stats[3] = mma_steps;
    ctx->stats[4] = atomq;
    ctx->stats[6] = max_split;
    ctx->stats[7] = ns ? specs[0].plan->n_wt : 0;
    return MB200_OK;
}

// |q| histograms of all segments of a call (k_sf_binned + k_sf_binned_final).  The number of parts per (segment, bin)
// depends on the plan alone: about 2048 entries per part.
static int launch_binned(mb200_ctx *ctx, const std::vector<SegSpec> &specs, double *d_s, double *d_i, long long *d_c,
                         int32_t n_bins) {
    const size_t ns = specs.size();
    if (ns == 0) return MB200_OK;
    size_t max_valid = 0;
    for (const SegSpec &sp : specs) max_valid = std::max(max_valid, sp.plan->cell.size());
    // a function of the plan alone (not of the batch size): the same frame gives the same bits in any batch
    const size_t per_bin = max_valid / (size_t)std::max(n_bins, 1) + 1;
    const int n_parts = (int)std::max<size_t>(1, std::min<size_t>({per_bin / 2048 + 1, (size_t)64, (size_t)8192 / (size_t)n_bins + 1}));
    MB_TRY(ctx->d_sval.ensure(ns * (size_t)n_bins * n_parts * 24));
    k_sf_binned<<<dim3((unsigned)n_bins, (unsigned)ns, (unsigned)n_parts), 128, 0, ctx->stream>>>(
        ctx->d_segs.as<SegDesc>(), ctx->d_sval.as<double>(), n_bins);
    MB_CUDA(cudaGetLastError());
    const long long total = (long long)ns * n_bins;
    k_sf_binned_final<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_sval.as<double>(), d_s, d_i, d_c, n_parts, total);
    MB_CUDA(cudaGetLastError());
    ctx->launches += 2; ctx->stats[0] += 2;
    return MB200_OK;
}

static int finish_timing(mb200_ctx *ctx) {
    ctx->stats[5] = 0;
    if (ctx->timing && ctx->stats[2] > 0 && ctx->t_ev0 && ctx->t_ev1) {
        float ms = 0.f;
        MB_CUDA(cudaEventSynchronize(ctx->t_ev1));
        MB_CUDA(cudaEventElapsedTime(&ms, ctx->t_ev0, ctx->t_ev1));
        ctx->stats[5] = (int64_t)(ms * 1.0e6);
    }
    return MB200_OK;
}

}  // namespace mb200

using namespace mb200;

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int mb200_ctx_set_precision(mb200_ctx *ctx, int mode) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context");
    if (mode < 0 || mode > 2) return fail(MB200_EINVAL, "precision mode must be 0 (fp64), 1 (i8x5 split precision) or 2 (i8x5, A operand in shared memory)");
    ctx->precision = mode ? 1 : 0;
    ctx->tc_a_tmem = mode != 2;
    return MB200_OK;
}

// Diagnostic (host only, no device needed): build the q-plan of a box / q window and describe it.
extern "C" int mb200_plan_describe(const double dimensions[3], double qmin, double qmax, double thetamin, double thetamax,
                                   int32_t n_bins, int64_t out[16]) {
    if (!dimensions || !out) return fail(MB200_EINVAL, "null argument");
    std::shared_ptr<SfPlan> pl;
    MB_TRY(build_plan(dimensions, qmin, qmax, thetamin, thetamax, n_bins, pl));
    auto fnv = [](const void *p, size_t bytes) {
        unsigned long long h = 1469598103934665603ull;
        const unsigned char *c = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < bytes; ++i) h = (h ^ c[i]) * 1099511628211ull;
        return (int64_t)(h >> 1);
    };
    for (int i = 0; i < 16; ++i) out[i] = 0;
    out[0] = (int64_t)pl->cell.size();
    out[1] = pl->maxn[0]; out[2] = pl->maxn[1]; out[3] = pl->maxn[2];
    out[4] = pl->tc_ok; out[5] = pl->tc_n_rt; out[6] = pl->tc_n_lt; out[7] = pl->tc_lt; out[8] = pl->tc_n_used;
    out[9] = pl->n_wt;
    out[10] = fnv(pl->cell.data(), pl->cell.size() * 4);
    out[11] = fnv(pl->ent_amp.data(), pl->ent_amp.size() * 4);
    out[12] = fnv(pl->ent_amp_tc.data(), pl->ent_amp_tc.size() * 4);
    out[13] = fnv(pl->tc_rows.data(), pl->tc_rows.size());
    out[14] = fnv(pl->bin_start.data(), pl->bin_start.size() * 4);
    out[15] = fnv(pl->tgs.data(), pl->tgs.size() * sizeof(TileGroupDesc));
    return MB200_OK;
}

extern "C" int mb200_sfactor_maxn(const double dimensions[3], double qmax, int32_t maxn[3]) {
    if (!dimensions || !maxn) return fail(MB200_EINVAL, "null argument");
    sfactor_maxn_host(dimensions, qmax, maxn);
    return MB200_OK;
}

extern "C" int mb200_structure_factor_block(mb200_ctx *ctx, const double *positions, int64_t stride_atom,
                                            int64_t stride_xyz, int64_t n_atoms, const double dimensions[3],
                                            double qmin, double qmax, double thetamin, double thetamax,
                                            const double *weights, int32_t block, int32_t n_blocks,
                                            double *scattering_vectors, double *structure_factors) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (!dimensions || !scattering_vectors || !structure_factors || n_atoms < 0 || (n_atoms && (!positions || !weights)))
        return fail(MB200_EINVAL, "invalid argument to mb200_structure_factor");
    if (n_atoms > 2000000000LL) return fail(MB200_EINVAL, "too many atoms");
    MB_CUDA(cudaSetDevice(ctx->device));
    std::shared_ptr<SfPlan> plan;
    MB_TRY(get_plan(ctx, dimensions, qmin, qmax, thetamin, thetamax, 0, plan));
    const size_t cells = (size_t)plan->maxn[0] * plan->maxn[1] * plan->maxn[2];
    std::memset(scattering_vectors, 0, cells * 8);
    std::memset(structure_factors, 0, cells * 8);
    const size_t nv = plan->cell.size();
    for (size_t e = 0; e < nv; ++e) scattering_vectors[plan->cell[e]] = plan->qrr[e];
    if (nv == 0) return MB200_OK;

    // stage inputs: positions compacted to [n][3] f64, weights [n]
    const size_t n = (size_t)n_atoms;
    MB_TRY(ctx->h_out.ensure(std::max<size_t>(n * 32, nv * 8) + 64));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    double *hp = ctx->h_out.as<double>();
    for (size_t j = 0; j < n; ++j)
        for (int d = 0; d < 3; ++d) hp[j * 3 + d] = positions[(int64_t)j * stride_atom + d * stride_xyz];
    // Split-precision mode works on table entries in [-1, 1] with an absolute quantisation: the weights are multiplied by
    // the power of two that brings max|w| into (0.5, 1] (exact, applied by k_sf_tables) and S is multiplied back by its
    // inverse square (exact, applied by the finalisers).  Non-finite weights take the FP64 kernel.
    double wscale = 1.0, s_scale = 1.0;
    bool tc_weights = ctx->precision == 1;
    if (tc_weights) {
        double wmax = 0.0;
        for (size_t j = 0; j < n; ++j) {
            const double a = std::fabs(weights[j]);
            if (!(a <= 1.7e308)) { wmax = a; break; }
            wmax = std::max(wmax, a);
        }
        tc_weights = weight_scale(wmax, wscale, s_scale);
        if (!tc_weights) { wscale = 1.0; s_scale = 1.0; }
    }
    if (n) std::memcpy(hp + n * 3, weights, n * 8);
    MB_TRY(ctx->d_pos.ensure(std::max<size_t>(n, 1) * 32));
    if (n) MB_CUDA(cudaMemcpyAsync(ctx->d_pos.p, hp, n * 32, cudaMemcpyHostToDevice, ctx->stream));

    std::vector<SegSpec> specs(1);
    SegSpec &sp = specs[0];
    sp.plan = plan.get(); sp.share_yz_with = -1; sp.ff2_group = -1;
    std::memset(&sp.job, 0, sizeof(TableJob));
                sp.job.wscale = 1.0;
    sp.job.pos = ctx->d_pos.p; sp.job.index = nullptr;
    sp.job.weights = ctx->d_pos.as<double>() + n * 3;
    sp.job.stride_atom = 3; sp.job.stride_xyz = 1;
    for (int d = 0; d < 3; ++d) { sp.job.box[d] = dimensions[d]; sp.job.boxf[d] = (float)dimensions[d]; }
    sp.job.n = (int32_t)n; sp.job.is_f32 = 0; sp.job.wrap_mode = 0;
    sp.job.wscale = wscale; sp.s_scale = s_scale;
    std::vector<SegDesc> segs;
    ctx->force_fp64 = !tc_weights;
    const int rc_seg = run_segments(ctx, specs, block, n_blocks, segs);
    ctx->force_fp64 = false;
    MB_TRY(rc_seg);

    MB_TRY(ctx->d_sval.ensure(nv * 8));
    dim3 grid((unsigned)((nv + 255) / 256), 1);
    k_sf_svalues<<<grid, 256, 0, ctx->stream>>>(ctx->d_segs.as<SegDesc>(), ctx->d_sval.as<double>(), (int)nv);
    MB_CUDA(cudaGetLastError());
    ctx->launches += 1; ctx->stats[0] += 1;
    MB_CUDA(cudaMemcpyAsync(hp, ctx->d_sval.p, nv * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    MB_TRY(finish_timing(ctx));
    for (size_t e = 0; e < nv; ++e) structure_factors[plan->cell[e]] = hp[e];
    return MB200_OK;
}

extern "C" int mb200_structure_factor(mb200_ctx *ctx, const double *positions, int64_t stride_atom, int64_t stride_xyz,
                                      int64_t n_atoms, const double dimensions[3], double qmin, double qmax,
                                      double thetamin, double thetamax, const double *weights,
                                      double *scattering_vectors, double *structure_factors) {
    return mb200_structure_factor_block(ctx, positions, stride_atom, stride_xyz, n_atoms, dimensions, qmin, qmax,
                                        thetamin, thetamax, weights, 0, 1, scattering_vectors, structure_factors);
}

// Element groups of a Saxs analysis, resident on the device for the whole run (mb200_groups_create).
static int saxs_frames_impl(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                            int64_t n_atoms_total, const float *boxes, const int32_t *d_group_index,
                            const int64_t *group_ptr, int32_t n_groups, const double *cm_params, double qmin,
                            double qmax, double thetamin, double thetamax, int32_t n_bins, int pack_atoms,
                            int32_t block, int32_t n_blocks, double *s_binned, double *i_binned, int64_t *bincount,
                            int32_t *ticket = nullptr);

extern "C" int mb200_groups_create(mb200_ctx *ctx, const int32_t *group_index, const int64_t *group_ptr, int32_t n_groups,
                                   int64_t n_atoms_total, mb200_groups **out) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (!out) return fail(MB200_EINVAL, "null output pointer");
    *out = nullptr;
    if (n_groups < 0 || !group_ptr || n_atoms_total < 0) return fail(MB200_EINVAL, "invalid argument to mb200_groups_create");
    for (int gi = 0; gi < n_groups; ++gi)
        if (group_ptr[gi + 1] < group_ptr[gi]) return fail(MB200_EINVAL, "group_ptr must be non-decreasing");
    const size_t n_idx = (size_t)group_ptr[n_groups];
    if (n_idx && !group_index) return fail(MB200_EINVAL, "null group_index");
    for (size_t i = 0; i < n_idx; ++i)
        if (group_index[i] < 0 || group_index[i] >= n_atoms_total) return fail(MB200_EINVAL, "group index out of range");
    MB_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<mb200_groups> g(new mb200_groups());
    MB_TRY(g->idx.ensure(std::max<size_t>(n_idx, 1) * 4));
    if (n_idx) {
        MB_CUDA(cudaMemcpyAsync(g->idx.p, group_index, n_idx * 4, cudaMemcpyHostToDevice, ctx->stream));
        MB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    g->ptr.assign(group_ptr, group_ptr + n_groups + 1);
    g->n_atoms_total = n_atoms_total;
    g->device = ctx->device;
    *out = g.release();
    return MB200_OK;
}

extern "C" void mb200_groups_destroy(mb200_ctx *ctx, mb200_groups *groups) {
    if (!groups) return;
    if (ctx) {
        MB_LOCK(ctx);
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        delete groups;
    } else {
        delete groups;
    }
}

extern "C" int mb200_saxs_frames_groups(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                                        const float *boxes, const mb200_groups *groups, const double *cm_params, double qmin,
                                        double qmax, double thetamin, double thetamax, int32_t n_bins, int pack_atoms,
                                        int32_t block, int32_t n_blocks, double *s_binned, double *i_binned,
                                        int64_t *bincount) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (!groups) return fail(MB200_EINVAL, "null groups handle");
    if (groups->device != ctx->device) return fail(MB200_EINVAL, "groups belong to another device");
    return saxs_frames_impl(ctx, positions, positions_on_device, n_frames, groups->n_atoms_total, boxes,
                            groups->idx.as<int32_t>(), groups->ptr.data(), (int32_t)groups->ptr.size() - 1, cm_params, qmin, qmax,
                            thetamin, thetamax, n_bins, pack_atoms, block, n_blocks, s_binned, i_binned, bincount);
}

// Deferred form of mb200_saxs_frames_groups: the batch is prepared and queued, the call returns while its kernels run (or
// wait behind the previous batch), so that the host side of batch k + 1 -- plan look-up, descriptors, the launches
// themselves -- overlaps with the device work of batch k.  At most two batches wait at a time.
extern "C" int mb200_saxs_frames_submit(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                                        const float *boxes, const mb200_groups *groups, const double *cm_params, double qmin,
                                        double qmax, double thetamin, double thetamax, int32_t n_bins, int pack_atoms,
                                        int32_t block, int32_t n_blocks, int32_t *ticket) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (!groups || !ticket) return fail(MB200_EINVAL, "null groups handle or ticket");
    if (groups->device != ctx->device) return fail(MB200_EINVAL, "groups belong to another device");
    return saxs_frames_impl(ctx, positions, positions_on_device, n_frames, groups->n_atoms_total, boxes,
                            groups->idx.as<int32_t>(), groups->ptr.data(), (int32_t)groups->ptr.size() - 1, cm_params, qmin, qmax,
                            thetamin, thetamax, n_bins, pack_atoms, block, n_blocks, nullptr, nullptr, nullptr, ticket);
}

extern "C" int mb200_saxs_frames_collect(mb200_ctx *ctx, int32_t ticket, double *s_binned, double *i_binned, int64_t *bincount) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    if (ticket < 0 || ticket >= mb200_ctx::N_PENDING || !s_binned || !i_binned || !bincount)
        return fail(MB200_EINVAL, "invalid argument to mb200_saxs_frames_collect");
    mb200_ctx::Pending &pd = ctx->pending[ticket];
    if (!pd.open) return fail(MB200_EINVAL, "no deferred batch waits under ticket %d", (int)ticket);
    // the wait happens outside the call mutex: the next batch may be submitted meanwhile
    MB_CUDA(cudaSetDevice(ctx->device));
    const cudaError_t waited = cudaEventSynchronize(pd.done);
    MB_LOCK(ctx);
    if (waited != cudaSuccess) {  // the batch failed on the device: the ticket is spent all the same
        pd.open = false;
        return fail(MB200_ECUDA, "deferred batch %d failed: %s", (int)ticket, cudaGetErrorString(waited));
    }
    const double *h_s = ctx->h_res[ticket].as<double>();
    std::memcpy(s_binned, h_s, pd.n_out * 8);
    std::memcpy(i_binned, h_s + pd.n_out, pd.n_out * 8);
    std::memcpy(bincount, h_s + 2 * pd.n_out, pd.n_out * 8);
    for (int i = 0; i < 8; ++i) ctx->stats[i] = pd.stats[i];
    ctx->stats[5] = 0;
    if (ctx->timing && pd.parity >= 0 && pd.stats[2] > 0) {
        float ms = 0.f;
        MB_CUDA(cudaEventElapsedTime(&ms, ctx->tc_ev0[pd.parity], ctx->tc_ev1[pd.parity]));
        ctx->stats[5] = (int64_t)(ms * 1.0e6);
    }
    pd.open = false;
    return MB200_OK;
}

extern "C" int mb200_saxs_frames(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                                 int64_t n_atoms_total, const float *boxes, const int32_t *group_index,
                                 const int64_t *group_ptr, int32_t n_groups, const double *cm_params, double qmin,
                                 double qmax, double thetamin, double thetamax, int32_t n_bins, int pack_atoms,
                                 int32_t block, int32_t n_blocks, double *s_binned, double *i_binned,
                                 int64_t *bincount) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (n_groups < 0 || !group_ptr) return fail(MB200_EINVAL, "invalid argument to mb200_saxs_frames");
    if (n_frames == 0 || n_groups == 0) return MB200_OK;
    MB_CUDA(cudaSetDevice(ctx->device));
    const size_t n_idx = (size_t)group_ptr[n_groups];
    for (int gi = 0; gi < n_groups; ++gi)
        if (group_ptr[gi + 1] < group_ptr[gi]) return fail(MB200_EINVAL, "group_ptr must be non-decreasing");
    if (n_idx && !group_index) return fail(MB200_EINVAL, "null group_index");
    // group indices -> device (kept from the previous call when the array has the same content: a Saxs run sends the same
    // selection with every batch; size + plain and position-weighted sums as the fingerprint).  Callers that keep a
    // selection for a whole run use mb200_groups_create + mb200_saxs_frames_groups and pay nothing per call.
    {
        unsigned long long s1 = 0, s2 = 0;
        for (size_t i = 0; i < n_idx; ++i) {
            const unsigned long long x = (unsigned)group_index[i];
            s1 += x;
            s2 += x * (unsigned long long)(2 * i + 1);
        }
        const unsigned long long fp = (s1 * 0x9e3779b97f4a7c15ull) ^ s2 ^ ((unsigned long long)n_idx << 40);
        if (!(ctx->idx_valid && ctx->idx_n == n_idx && ctx->idx_fp == fp && ctx->idx_natoms == n_atoms_total)) {
            for (size_t i = 0; i < n_idx; ++i)
                if (group_index[i] < 0 || group_index[i] >= n_atoms_total) return fail(MB200_EINVAL, "group index out of range");
            ctx->idx_valid = false;
            MB_TRY(ctx->d_idx.ensure(std::max<size_t>(n_idx, 1) * 4));
            if (n_idx) MB_CUDA(cudaMemcpyAsync(ctx->d_idx.p, group_index, n_idx * 4, cudaMemcpyHostToDevice, ctx->stream));
            ctx->idx_n = n_idx; ctx->idx_fp = fp; ctx->idx_natoms = n_atoms_total; ctx->idx_valid = true;
        }
    }
    return saxs_frames_impl(ctx, positions, positions_on_device, n_frames, n_atoms_total, boxes, ctx->d_idx.as<int32_t>(),
                            group_ptr, n_groups, cm_params, qmin, qmax, thetamin, thetamax, n_bins, pack_atoms, block, n_blocks,
                            s_binned, i_binned, bincount);
}

static int saxs_frames_impl(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                            int64_t n_atoms_total, const float *boxes, const int32_t *d_group_index,
                            const int64_t *group_ptr, int32_t n_groups, const double *cm_params, double qmin,
                            double qmax, double thetamin, double thetamax, int32_t n_bins, int pack_atoms,
                            int32_t block, int32_t n_blocks, double *s_binned, double *i_binned, int64_t *bincount,
                            int32_t *ticket) {
    // ticket != null: deferred -- the results of the (last chunk of the) batch stay in page-locked memory and the call returns
    // with its kernels still queued; mb200_saxs_frames_collect(ticket) waits for them and hands the results out
    const bool deferred = ticket != nullptr;
    if (n_frames < 0 || n_groups < 0 || n_bins <= 0 || !boxes || !group_ptr || !cm_params ||
        (!deferred && (!s_binned || !i_binned || !bincount)) || (n_atoms_total && !positions))
        return fail(MB200_EINVAL, "invalid argument to mb200_saxs_frames");
    MB_CUDA(cudaSetDevice(ctx->device));
    int slot = -1;
    double *h_s = nullptr, *h_i = nullptr;
    int64_t *h_c = nullptr;
    if (deferred) {
        for (int i = 0; i < mb200_ctx::N_PENDING && slot < 0; ++i)
            if (!ctx->pending[i].open) slot = i;
        if (slot < 0) return fail(MB200_EBUSY, "every deferred-batch slot is taken: collect one first (or use the synchronous call)");
        const size_t n_out = (size_t)n_frames * (size_t)n_groups * (size_t)n_bins;
        MB_TRY(ctx->h_res[slot].ensure(std::max<size_t>(n_out, 1) * 24));
        h_s = ctx->h_res[slot].as<double>(); h_i = h_s + n_out; h_c = reinterpret_cast<int64_t *>(h_i + n_out);
        s_binned = h_s; i_binned = h_i; bincount = h_c;
        ctx->pending[slot].n_out = n_out;
        *ticket = slot;
    }
    if (n_frames == 0 || n_groups == 0) {
        if (deferred) {
            ctx->pending[slot].open = true;
            ctx->pending[slot].parity = -1;
            MB_CUDA(cudaEventRecord(ctx->pending[slot].done, ctx->stream));
        }
        return MB200_OK;
    }
    // developer aid: MAICOS_B200_TRACE_HOST=1 prints the host-side time of every phase of the call to stderr
    static const bool trace_host = std::getenv("MAICOS_B200_TRACE_HOST") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (!trace_host) return;
        const auto now = std::chrono::steady_clock::now();
        static const auto t_zero = now;  // absolute time of the first traced call: shows the gaps between calls as well
        std::fprintf(stderr, "[mb200_saxs_frames] %-28s %8.3f ms   @ %10.3f ms\n", what,
                     std::chrono::duration<double, std::milli>(now - t_last).count(),
                     std::chrono::duration<double, std::milli>(now - t_zero).count());
        t_last = now;
    };
    mark("entry");
    // frames are processed in chunks bounded by the table workspace
    const size_t frame_elems = (size_t)n_atoms_total * 3;
    size_t done = 0;
    const float *pref = nullptr;
    bool pref_checked = false;
    PrefetchRead pref_read(ctx);
    MB_TRY(prepare_plans(ctx, (size_t)n_frames,
                         [&](size_t f, double *d) { for (int i = 0; i < 3; ++i) d[i] = (double)boxes[f * 3 + i]; },
                         qmin, qmax, thetamin, thetamax, n_bins, cm_params, n_groups));
    mark("plans ahead");
    while (done < (size_t)n_frames) {
        // chunk size: keep tables under ~12 GB
        size_t chunk = 0, tab_bytes = 0;
        std::vector<std::shared_ptr<SfPlan>> plans;
        while (done + chunk < (size_t)n_frames) {
            const float *bx = boxes + (done + chunk) * 3;
            double dims[3] = {(double)bx[0], (double)bx[1], (double)bx[2]};
            std::shared_ptr<SfPlan> pl;
            MB_TRY(get_plan(ctx, dims, qmin, qmax, thetamin, thetamax, n_bins, pl));
            size_t fb = 0;
            for (int gi = 0; gi < n_groups; ++gi) {
                const size_t ld = ((size_t)(group_ptr[gi + 1] - group_ptr[gi]) + KT - 1) / KT * KT;
                fb += ld * (size_t)(pl->rows[0] + pl->rows[1] + pl->rows[2]) * 16;
            }
            if (chunk > 0 && tab_bytes + fb > ((size_t)12 << 30)) break;
            tab_bytes += fb;
            plans.push_back(pl);
            ++chunk;
            if (chunk >= 256) break;
        }
        mark("plans");
        for (auto &pl : plans) MB_TRY(plan_ff2(ctx, *pl, cm_params, n_groups));
        mark("form factors");

        const float *dpos;
        if (positions_on_device) {
            dpos = positions + done * frame_elems;
        } else if (!pref_checked && (pref = reinterpret_cast<const float *>(
                                         pref_read.take(positions, (size_t)n_frames * frame_elems * 4))) != nullptr) {
            pref_checked = true;
            dpos = pref + done * frame_elems;  // the whole call was copied ahead by mb200_prefetch_frames
        } else if (pref != nullptr) {
            dpos = pref + done * frame_elems;
        } else {
            pref_checked = true;
            MB_TRY(ctx->d_pos.ensure(std::max<size_t>(chunk * frame_elems, 1) * 4));
            if (frame_elems)
                MB_CUDA(cudaMemcpyAsync(ctx->d_pos.p, positions + done * frame_elems, chunk * frame_elems * 4,
                                        cudaMemcpyHostToDevice, ctx->stream));
            dpos = ctx->d_pos.as<float>();
        }
        std::vector<SegSpec> specs(chunk * n_groups);
        for (size_t f = 0; f < chunk; ++f)
            for (int gi = 0; gi < n_groups; ++gi) {
                SegSpec &sp = specs[f * n_groups + gi];
                sp.plan = plans[f].get(); sp.share_yz_with = -1; sp.ff2_group = gi;
                std::memset(&sp.job, 0, sizeof(TableJob));
                sp.job.wscale = 1.0;
                sp.job.pos = dpos + f * frame_elems;
                sp.job.index = d_group_index + group_ptr[gi];
                sp.job.weights = nullptr;
                sp.job.stride_atom = 3; sp.job.stride_xyz = 1;
                const float *bx = boxes + (done + f) * 3;
                for (int d = 0; d < 3; ++d) { sp.job.boxf[d] = bx[d]; sp.job.box[d] = (double)bx[d]; }
                sp.job.n = (int32_t)(group_ptr[gi + 1] - group_ptr[gi]);
                sp.job.is_f32 = 1; sp.job.wrap_mode = pack_atoms ? 2 : 1;
            }
        std::vector<SegDesc> segs;
        mark("positions + specs");
        if (done + chunk == (size_t)n_frames) ctx->release_after_tables = &pref_read;  // last reader of the prefetched frames
        const int rc_seg = run_segments(ctx, specs, block, n_blocks, segs);
        ctx->release_after_tables = nullptr;
        MB_TRY(rc_seg);
        mark("run_segments (host side)");
        const size_t ns = specs.size();
        const size_t ob = ns * (size_t)n_bins;
        MB_TRY(ctx->d_out.ensure(ob * 24));
        double *d_s = ctx->d_out.as<double>(), *d_i = d_s + ob;
        long long *d_c = reinterpret_cast<long long *>(d_i + ob);
        MB_TRY(launch_binned(ctx, specs, d_s, d_i, d_c, n_bins));
        const size_t oo = done * (size_t)n_groups * n_bins;
        MB_CUDA(cudaMemcpyAsync(s_binned + oo, d_s, ob * 8, cudaMemcpyDeviceToHost, ctx->stream));
        MB_CUDA(cudaMemcpyAsync(i_binned + oo, d_i, ob * 8, cudaMemcpyDeviceToHost, ctx->stream));
        MB_CUDA(cudaMemcpyAsync(bincount + oo, d_c, ob * 8, cudaMemcpyDeviceToHost, ctx->stream));
        done += chunk;
        if (deferred && done == (size_t)n_frames) break;  // the last chunk is waited for by the collect call
        MB_CUDA(cudaStreamSynchronize(ctx->stream));
        mark("device (wait)");
        MB_TRY(finish_timing(ctx));
    }
    pref_read.release();  // the prefetch slot this call read may be refilled from here on
    if (deferred) {
        mb200_ctx::Pending &pd = ctx->pending[slot];
        MB_CUDA(cudaEventRecord(pd.done, ctx->stream));
        pd.open = true;
        pd.parity = ctx->t_parity;
        for (int i = 0; i < 8; ++i) pd.stats[i] = ctx->stats[i];
        mark("queued (deferred)");
    }
    return MB200_OK;
}

// Saxs with bin_spectrum = False (saxs.py:188-189, 238-239): the per-cell structure factors of every frame and element
// group as compact lists over the accepted q vectors of the (constant) cell, in C order of the reciprocal cube.  One
// tables launch and one contraction launch per batch instead of one mb200_structure_factor call per frame and group.
extern "C" int mb200_saxs_frames_cells(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                                       const float box[3], const mb200_groups *groups, double qmin, double qmax,
                                       double thetamin, double thetamax, int pack_atoms, int64_t n_valid,
                                       int32_t *cells, double *scattering_vectors, double *s_values) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (!groups || !box || n_frames < 0 || n_valid < 0 || (n_frames && n_valid && !s_values) ||
        (groups->n_atoms_total && n_frames && !positions))
        return fail(MB200_EINVAL, "invalid argument to mb200_saxs_frames_cells");
    if (groups->device != ctx->device) return fail(MB200_EINVAL, "groups belong to another device");
    MB_CUDA(cudaSetDevice(ctx->device));
    const double dims[3] = {(double)box[0], (double)box[1], (double)box[2]};
    std::shared_ptr<SfPlan> plan;
    MB_TRY(get_plan(ctx, dims, qmin, qmax, thetamin, thetamax, 0, plan));
    const size_t nv = plan->cell.size();
    if ((int64_t)nv != n_valid)
        return fail(MB200_EINVAL, "n_valid = %lld, the cell has %zu accepted q vectors (mb200_plan_describe out[0])", (long long)n_valid, nv);
    for (size_t e = 0; e < nv; ++e) {
        if (cells) cells[e] = plan->cell[e];
        if (scattering_vectors) scattering_vectors[e] = plan->qrr[e];
    }
    const int n_groups = (int)groups->ptr.size() - 1;
    if (n_frames == 0 || n_groups <= 0 || nv == 0) return MB200_OK;
    const size_t frame_elems = (size_t)groups->n_atoms_total * 3;
    size_t done = 0;
    while (done < (size_t)n_frames) {
        size_t fb = 0;  // table bytes of one frame
        for (int gi = 0; gi < n_groups; ++gi) {
            const size_t ld = ((size_t)(groups->ptr[gi + 1] - groups->ptr[gi]) + KT - 1) / KT * KT;
            fb += ld * (size_t)(plan->rows[0] + plan->rows[1] + plan->rows[2]) * 16;
        }
        const size_t chunk = std::max<size_t>(1, std::min<size_t>({(size_t)n_frames - done, ((size_t)12 << 30) / std::max<size_t>(fb, 1), (size_t)256}));
        const float *dpos;
        if (positions_on_device) {
            dpos = positions + done * frame_elems;
        } else {
            MB_TRY(ctx->d_pos.ensure(std::max<size_t>(chunk * frame_elems, 1) * 4));
            if (frame_elems)
                MB_CUDA(cudaMemcpyAsync(ctx->d_pos.p, positions + done * frame_elems, chunk * frame_elems * 4, cudaMemcpyHostToDevice,
                                        ctx->stream));
            dpos = ctx->d_pos.as<float>();
        }
        std::vector<SegSpec> specs(chunk * n_groups);
        for (size_t f = 0; f < chunk; ++f)
            for (int gi = 0; gi < n_groups; ++gi) {
                SegSpec &sp = specs[f * n_groups + gi];
                sp.plan = plan.get(); sp.share_yz_with = -1; sp.ff2_group = -1;
                std::memset(&sp.job, 0, sizeof(TableJob));
                sp.job.wscale = 1.0;
                sp.job.pos = dpos + f * frame_elems;
                sp.job.index = groups->idx.as<int32_t>() + groups->ptr[gi];
                sp.job.weights = nullptr;
                sp.job.stride_atom = 3; sp.job.stride_xyz = 1;
                for (int d = 0; d < 3; ++d) { sp.job.boxf[d] = box[d]; sp.job.box[d] = dims[d]; }
                sp.job.n = (int32_t)(groups->ptr[gi + 1] - groups->ptr[gi]);
                sp.job.is_f32 = 1; sp.job.wrap_mode = pack_atoms ? 2 : 1;
            }
        std::vector<SegDesc> segs;
        MB_TRY(run_segments(ctx, specs, 0, 1, segs));
        const size_t ns = specs.size();
        MB_TRY(ctx->d_sval.ensure(ns * nv * 8));
        k_sf_svalues<<<dim3((unsigned)((nv + 255) / 256), (unsigned)ns), 256, 0, ctx->stream>>>(ctx->d_segs.as<SegDesc>(),
                                                                                              ctx->d_sval.as<double>(), (int)nv);
        MB_CUDA(cudaGetLastError());
        ctx->launches += 1; ctx->stats[0] += 1;
        MB_CUDA(cudaMemcpyAsync(s_values + done * (size_t)n_groups * nv, ctx->d_sval.p, ns * nv * 8, cudaMemcpyDeviceToHost, ctx->stream));
        MB_CUDA(cudaStreamSynchronize(ctx->stream));
        MB_TRY(finish_timing(ctx));
        done += chunk;
    }
    return MB200_OK;
}

extern "C" int mb200_dipsf_frames(mb200_ctx *ctx, const double *positions, const double *weights, int inputs_on_device,
                                  int64_t n_frames, int64_t n, const double *boxes, double qmin, double qmax, int32_t n_bins,
                                  int32_t block, int32_t n_blocks, double *s_binned, int64_t *bincount) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (n_frames < 0 || n < 0 || n_bins <= 0 || !boxes || !s_binned || !bincount || (n && (!positions || !weights)))
        return fail(MB200_EINVAL, "invalid argument to mb200_dipsf_frames");
    if (n_frames == 0) return MB200_OK;
    MB_CUDA(cudaSetDevice(ctx->device));
    const size_t nn = (size_t)n;
    size_t done = 0;
    MB_TRY(prepare_plans(ctx, (size_t)n_frames,
                         [&](size_t f, double *d) { for (int i = 0; i < 3; ++i) d[i] = boxes[f * 3 + i]; },
                         qmin, qmax, 0.0, M_PI, n_bins, nullptr, 0));
    while (done < (size_t)n_frames) {
        size_t chunk = 0, tab_bytes = 0;
        std::vector<std::shared_ptr<SfPlan>> plans;
        while (done + chunk < (size_t)n_frames) {
            std::shared_ptr<SfPlan> pl;
            MB_TRY(get_plan(ctx, boxes + (done + chunk) * 3, qmin, qmax, 0.0, M_PI, n_bins, pl));
            const size_t ld = (nn + KT - 1) / KT * KT;
            const size_t fb = ld * (size_t)(3 * pl->rows[0] + pl->rows[1] + pl->rows[2]) * 16;
            if (chunk > 0 && tab_bytes + fb > ((size_t)12 << 30)) break;
            tab_bytes += fb;
            plans.push_back(pl);
            if (++chunk >= 64) break;
        }
        // inputs: positions [chunk][n][3], weights [chunk][3][n]
        double *dp, *dw;
        if (inputs_on_device) {
            dp = const_cast<double *>(positions) + done * nn * 3;
            dw = const_cast<double *>(weights) + done * nn * 3;
        } else {
            MB_TRY(ctx->d_pos.ensure(std::max<size_t>(chunk * nn, 1) * 48));
            dp = ctx->d_pos.as<double>();
            dw = dp + chunk * nn * 3;
        }
        if (nn && !inputs_on_device) {
            MB_CUDA(cudaMemcpyAsync(dp, positions + done * nn * 3, chunk * nn * 24, cudaMemcpyHostToDevice, ctx->stream));
            MB_CUDA(cudaMemcpyAsync(dw, weights + done * nn * 3, chunk * nn * 24, cudaMemcpyHostToDevice, ctx->stream));
        }
        // per (frame, component) max|w| -> power-of-two rescaling for the split-precision mode; a non-finite weight
        // anywhere in the chunk (a zero dipole: mu / |mu| = NaN, the reference then returns NaN) sends it to FP64
        std::vector<double> wsc(chunk * 3, 1.0), ssc(chunk * 3, 1.0);
        bool tc_weights = ctx->precision == 1;
        if (tc_weights && nn) {
            MB_TRY(ctx->d_w.ensure(chunk * 3 * 8));
            MB_TRY(ctx->h_out.ensure(chunk * 3 * 8 + 64));
            MB_CUDA(cudaMemsetAsync(ctx->d_w.p, 0, chunk * 3 * 8, ctx->stream));
            const unsigned bx = (unsigned)std::max<size_t>(1, std::min<size_t>((nn + 2047) / 2048, (size_t)ctx->n_sm * 8 / (chunk * 3) + 1));
            k_sf_wmax<<<dim3(bx, (unsigned)(chunk * 3)), 256, 0, ctx->stream>>>(dw, (long long)nn, ctx->d_w.as<unsigned long long>());
            MB_CUDA(cudaGetLastError());
            ctx->launches += 1;
            MB_CUDA(cudaMemcpyAsync(ctx->h_out.p, ctx->d_w.p, chunk * 3 * 8, cudaMemcpyDeviceToHost, ctx->stream));
            MB_CUDA(cudaStreamSynchronize(ctx->stream));
            const double *hm = ctx->h_out.as<double>();
            for (size_t i = 0; i < chunk * 3 && tc_weights; ++i) tc_weights = weight_scale(hm[i], wsc[i], ssc[i]);
        }
        std::vector<SegSpec> specs(chunk * 3);
        for (size_t f = 0; f < chunk; ++f)
            for (int c = 0; c < 3; ++c) {
                SegSpec &sp = specs[f * 3 + c];
                sp.plan = plans[f].get(); sp.share_yz_with = c == 0 ? -1 : (int)(f * 3); sp.ff2_group = -1;
                std::memset(&sp.job, 0, sizeof(TableJob));
                sp.job.wscale = tc_weights ? wsc[f * 3 + c] : 1.0;
                sp.s_scale = tc_weights ? ssc[f * 3 + c] : 1.0;
                sp.job.pos = dp + f * nn * 3; sp.job.index = nullptr;
                sp.job.weights = dw + (f * 3 + c) * nn;
                sp.job.stride_atom = 3; sp.job.stride_xyz = 1;
                for (int d = 0; d < 3; ++d) { sp.job.box[d] = boxes[(done + f) * 3 + d]; sp.job.boxf[d] = (float)sp.job.box[d]; }
                sp.job.n = (int32_t)nn; sp.job.is_f32 = 0; sp.job.wrap_mode = 0;
            }
        std::vector<SegDesc> segs;
        ctx->force_fp64 = !tc_weights;
        const int rc_seg = run_segments(ctx, specs, block, n_blocks, segs);
        ctx->force_fp64 = false;
        MB_TRY(rc_seg);
        const size_t ns = specs.size(), ob = ns * (size_t)n_bins;
        MB_TRY(ctx->d_out.ensure(ob * 16));
        double *d_s = ctx->d_out.as<double>();
        long long *d_c = reinterpret_cast<long long *>(d_s + ob);
        MB_TRY(launch_binned(ctx, specs, d_s, nullptr, d_c, n_bins));
        const size_t oo = done * 3 * (size_t)n_bins;
        MB_CUDA(cudaMemcpyAsync(s_binned + oo, d_s, ob * 8, cudaMemcpyDeviceToHost, ctx->stream));
        MB_CUDA(cudaMemcpyAsync(bincount + oo, d_c, ob * 8, cudaMemcpyDeviceToHost, ctx->stream));
        MB_CUDA(cudaStreamSynchronize(ctx->stream));
        MB_TRY(finish_timing(ctx));
        done += chunk;
    }
    return MB200_OK;
}

#ifdef MB200_TC_TRACE
// Developer build only (make TRACE=1): clock64 stamps written by CTA 0 of the last k_sf_contract_tc launch.
extern "C" int mb200_debug_tc_trace(unsigned long long *out, int n) {
    return cudaMemcpyFromSymbol(out, mb200::g_tc_trace, (size_t)std::min(n, 512 * 32) * 8) == cudaSuccess ? 0 : -1;
}
#endif
