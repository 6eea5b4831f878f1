// Weighted slab / shell histograms of libmaicos_b200 (sm_100a).
//
// Replaces the two numpy passes of ProfileBase._single_frame
// (/root/reference/src/maicos/core/base.py:815-818) and the three geometry overrides
//   planar   core/planar.py:235-243    np.histogram(pos[:, dim], n_bins, (zmin, zmax), weights)
//   cylinder core/cylinder.py:229-243  transform_cylinder (lib/math.py:705-744) + np.histogram2d
//   sphere   core/sphere.py:215-223    transform_sphere (lib/math.py:747-778) + np.histogram
// with ONE pass that produces the weighted sums (float64) and the counts (int64) together,
// batched over frames.  numpy bin semantics (numpy/lib/_histograms_impl.py:816-873 and
// 1037-1069) are reproduced exactly: x promoted to float64, [first, last] inclusive, bin i
// holds edges[i] <= x < edges[i+1]; the arithmetic index guess is corrected against the
// np.linspace edges the host passes in, so counts are bit-exact.
//
// HBM-bound: 12 B (float32 xyz) + 8 B (per-frame weight) per element.  Shared memory has native
// atomics for 32-bit integers only (u64 / f32 / f64 adds compile to ATOMS.CAST.SPIN loops), so the
// per-CTA privatised histogram keeps, per bin, a 32-bit count and the weighted sum as a 96-bit
// FIXED-POINT integer in three 32-bit limbs with carries:
//     fixed(w) = rint((w + B) / Q),  B = 2^k >= max|w| of the frame,  Q = 2^(k-52)
// i.e. the largest weights keep their full 53-bit mantissa and every weight is quantised with an
// absolute error <= Q/2 = 2^-53 * B -- the resolution float64 accumulation has once the running sum
// reaches max|w|.  Integer sums are exact and order independent, so profiles are bitwise
// reproducible run to run.  CTAs merge into 128-bit global accumulators (two native u64 atomics with
// carry) and k_hist_merge removes the bias (count * 2^52) exactly in __int128 and rescales.
// Non-finite weights and histograms with more than 4096 bins fall back to float64 global atomics.
//
// Bound in practice: native ATOMS retire about one LANE per cycle per SM, so the three shared
// atomics per element (count, low limb, middle limb) cap the kernel near 20 % of HBM bandwidth.
// Tried and rejected (profiles/r01_hist_notes.md): float64 CAS loops (slower), and an atomic-free
// warp-private variant using __match_any_sync + plain read-modify-write (2x slower: MATCH latency,
// 12.5 % occupancy at 2154 bins).
#include "common.cuh"
#include <chrono>
#include <cstdlib>

namespace mb200 {

constexpr int H_THREADS = 256;
constexpr int H_MAX_SMEM_BINS = 4096;

struct HistParams {
    const void *pos;       // [F][n][3]
    const double *weights; // [n] or [F][n] or null
    const double *edges;   // [F][n_bins+1]
    const double *center;  // [F][3] or null
    const double *zrange;  // [F][2] or null
    const unsigned long long *wmax_bits;  // [F] or [1]: bit pattern of max |w| (null without weights)
    unsigned long long *acc_lo;           // [F][n_bins] low 64 bits of the fixed-point sums
    unsigned long long *acc_hi;           // [F][n_bins] high 64 bits
    double *acc_f;                        // [F][n_bins] float64 fallback sums
    unsigned long long *acc_c;            // [F][n_bins] counts
    long long n;
    long long fstride;     // elements of `pos` between two frames (n * 3, n for scalars; more when the group is a window of whole frames)
    int n_bins, dim, geometry, is_f32, w_per_frame, digitize;
};

// exponent k with 2^k >= wmax (wmax finite, > 0), clamped so that Q = 2^(k-52) stays normal
__device__ __forceinline__ int bias_exponent(double wmax) {
    int e;
    const double m = frexp(wmax, &e);  // wmax = m * 2^e, m in [0.5, 1)
    int k = (m == 0.5) ? e - 1 : e;
    return max(k, -960);
}

__device__ __forceinline__ int bin_of(double x, const double *__restrict__ edges, int n_bins, double first, double last,
                                      double inv_width) {
    if (!(x >= first && x <= last)) return -1;
    // arithmetic guess, then exact correction against the edges (numpy: edges[i] <= x < edges[i+1],
    // right edge inclusive)
    int idx = (int)((x - first) * inv_width);
    idx = min(max(idx, 0), n_bins - 1);
    while (idx > 0 && x < edges[idx]) --idx;
    while (idx < n_bins - 1 && x >= edges[idx + 1]) ++idx;
    return idx;
}

// np.digitize(x, edges[1:-1]) (dielectric modules: dielectricplanar.py:170-172, dielectriccylinder.py:134,
// dielectricsphere.py:113): the bin is the number of interior edges <= x, i.e. values left of the range go to bin 0,
// values right of it (and NaN, which searchsorted orders last) to bin n_bins - 1.  Nothing is dropped.
__device__ __forceinline__ int bin_digitize(double x, const double *__restrict__ edges, int n_bins, double first,
                                            double inv_width) {
    if (x != x) return n_bins - 1;
    const double gd = (x - first) * inv_width;
    int idx = gd <= 0.0 ? 0 : (gd >= (double)(n_bins - 1) ? n_bins - 1 : (int)gd);
    while (idx > 0 && x < edges[idx]) --idx;
    while (idx < n_bins - 1 && x >= edges[idx + 1]) ++idx;
    return idx;
}

// max |w| per frame as an integer max over the (sign-stripped) bit patterns; NaN -> +inf
__global__ void __launch_bounds__(256) k_hist_wmax(const double *__restrict__ w, long long n, unsigned long long *out) {
    const double *wf = w + (size_t)blockIdx.y * n;
    unsigned long long m = 0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const double v = fabs(wf[j]);
        const unsigned long long b = (v != v) ? 0x7ff0000000000000ull : (unsigned long long)__double_as_longlong(v);
        m = b > m ? b : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
        m = t > m ? t : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out + blockIdx.y, m);
}

// Dense float32 copy [F][n][3] of a group of atoms of whole frames, in front of the histogram:
//   index != null : the group is index[0 .. n) (any selection), else atoms 0 .. n of `src` (a window);
//   boxes != null : AtomGroup.wrap(compound="atoms"), x -= floor(x / L) * L with every operation rounded to float32 as
//                   numpy does it (core/base.py:446-448 of the reference calls it on the host for every frame).
__global__ void __launch_bounds__(256) k_pack_atoms(const float *__restrict__ src, long long src_fstride, float *__restrict__ dst,
                                                    const float *__restrict__ boxes, const int32_t *__restrict__ index,
                                                    long long n3) {
    const long long f = blockIdx.y;
    const float *s = src + (size_t)f * src_fstride;
    float *d = dst + (size_t)f * n3;
    float L[3] = {1.f, 1.f, 1.f};
    if (boxes) { L[0] = boxes[f * 3]; L[1] = boxes[f * 3 + 1]; L[2] = boxes[f * 3 + 2]; }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n3; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / 3;
        const int c = (int)(e - i * 3);
        float x = index ? s[(size_t)index[i] * 3 + c] : s[e];
        if (boxes) {
            const float l = L[c];
            x = __fsub_rn(x, __fmul_rn(floorf(__fdiv_rn(x, l)), l));
        }
        d[e] = x;
    }
}

struct FrameGeom {
    double cx, cy, cz, zlo, zhi, first, last, inv_width;
    int d0, d1, d2, geometry, nb, digitize;
};

__device__ __forceinline__ double pick(double x, double y, double z, int d) { return d == 0 ? x : (d == 1 ? y : z); }

__device__ __forceinline__ int atom_bin(const FrameGeom &g, double x, double y, double z, const double *edges) {
    double v;
    if (g.geometry == MB200_GEOM_PLANAR) {
        v = pick(x, y, z, g.d0);
    } else if (g.geometry == MB200_GEOM_CYLINDER) {
        const double ax = pick(x, y, z, g.d0);
        if (!g.digitize && !(ax >= g.zlo && ax <= g.zhi)) return -1;  // histogram2d drops outliers of either dimension
        const double a = __dsub_rn(pick(x, y, z, g.d1), pick(g.cx, g.cy, g.cz, g.d1));
        const double b = __dsub_rn(pick(x, y, z, g.d2), pick(g.cx, g.cy, g.cz, g.d2));
        v = sqrt(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
    } else {
        const double a = __dsub_rn(x, g.cx), b = __dsub_rn(y, g.cy), c = __dsub_rn(z, g.cz);
        v = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
    }
    return g.digitize ? bin_digitize(v, edges, g.nb, g.first, g.inv_width) : bin_of(v, edges, g.nb, g.first, g.last, g.inv_width);
}
__device__ __forceinline__ int scalar_bin(const FrameGeom &g, double v, const double *edges) {
    return g.digitize ? bin_digitize(v, edges, g.nb, g.first, g.inv_width) : bin_of(v, edges, g.nb, g.first, g.last, g.inv_width);
}

// MODE 0: shared-memory privatised, fixed-point weights   1: shared counts only (no weights)
// MODE 2: global float64 atomics (non-finite weights or > 4096 bins)
template <int MODE>
__global__ void __launch_bounds__(H_THREADS) k_hist(HistParams p) {
    extern __shared__ __align__(16) unsigned char hsm[];
    const int f = blockIdx.y;
    const int nb = p.n_bins;
    double *sh_edges = reinterpret_cast<double *>(hsm);                       // [nb + 1]
    unsigned int *sh_c = reinterpret_cast<unsigned int *>(sh_edges + nb + 1); // [nb]
    unsigned int *sh_lo = sh_c + nb, *sh_mid = sh_lo + nb, *sh_hi = sh_mid + nb;
    const double *g_edges = p.edges + (size_t)f * (nb + 1);
    const double *edges = g_edges;
    if (MODE != 2) {
        for (int i = threadIdx.x; i <= nb; i += H_THREADS) sh_edges[i] = g_edges[i];
        for (int i = threadIdx.x; i < nb * (MODE == 0 ? 4 : 1); i += H_THREADS) sh_c[i] = 0u;
        edges = sh_edges;
        __syncthreads();
    }
    FrameGeom g;
    g.geometry = p.geometry; g.nb = nb; g.digitize = p.digitize;
    g.d0 = p.dim; g.d1 = (p.dim + 1) % 3; g.d2 = (p.dim + 2) % 3;  // odims = roll(arange(3), -dim)[1:]
    g.cx = g.cy = g.cz = g.zlo = g.zhi = 0;
    if (p.geometry == MB200_GEOM_CYLINDER || p.geometry == MB200_GEOM_SPHERE) { g.cx = p.center[f * 3]; g.cy = p.center[f * 3 + 1]; g.cz = p.center[f * 3 + 2]; }
    if (p.geometry == MB200_GEOM_CYLINDER && p.zrange) { g.zlo = p.zrange[f * 2]; g.zhi = p.zrange[f * 2 + 1]; }
    g.first = g_edges[0]; g.last = g_edges[nb];
    g.inv_width = (double)nb / (g.last - g.first);
    const double *w = p.weights ? p.weights + (p.w_per_frame ? (size_t)f * p.n : 0) : nullptr;

    double bias = 0.0, inv_q = 0.0;
    bool fx = false;  // fixed-point accumulation possible (all weights of the frame finite)
    if (MODE == 0) {
        const unsigned long long wb = p.wmax_bits[p.w_per_frame ? f : 0];
        fx = wb < 0x7ff0000000000000ull;
        const double wmax = __longlong_as_double((long long)wb);
        const int k = bias_exponent(fx && wmax > 0.0 ? wmax : 1.0);
        bias = ldexp(1.0, k);
        inv_q = ldexp(1.0, 52 - k);
    }
    double *gf = p.acc_f + (size_t)f * nb;
    unsigned long long *gc = p.acc_c + (size_t)f * nb;

    auto deposit = [&](int idx, long long j) {
        if (idx < 0) return;
        if (MODE == 2) {
            atomicAdd(&gc[idx], 1ull);
            if (w) atomicAdd(&gf[idx], w[j]);
            return;
        }
        atomicAdd(&sh_c[idx], 1u);
        if (MODE == 0 && !fx) {
            atomicAdd(&gf[idx], w[j]);  // inf / nan weights: float64 atomics keep numpy's result
        } else if (MODE == 0) {
            const unsigned long long v = (unsigned long long)__double2ll_rn(__dmul_rn(__dadd_rn(w[j], bias), inv_q));
            const unsigned int lo = (unsigned int)v, mid = (unsigned int)(v >> 32);  // v <= 2^53
            const unsigned int old = atomicAdd(&sh_lo[idx], lo);
            const unsigned int m2 = mid + ((old + lo) < old ? 1u : 0u);  // mid < 2^22: cannot wrap here
            if (m2) {
                const unsigned int oldm = atomicAdd(&sh_mid[idx], m2);
                if ((oldm + m2) < oldm) atomicAdd(&sh_hi[idx], 1u);
            }
        }
    };

    const long long n = p.n;
    if (p.geometry == MB200_GEOM_SCALAR) {  // the binned coordinate itself, [F][n]
        const long long gtid = (long long)blockIdx.x * H_THREADS + threadIdx.x, gstride = (long long)gridDim.x * H_THREADS;
        if (p.is_f32) {
            const float *base = reinterpret_cast<const float *>(p.pos) + (size_t)f * p.fstride;
            for (long long j = gtid; j < n; j += gstride) deposit(scalar_bin(g, (double)base[j], edges), j);
        } else {
            const double *base = reinterpret_cast<const double *>(p.pos) + (size_t)f * p.fstride;
            for (long long j = gtid; j < n; j += gstride) deposit(scalar_bin(g, base[j], edges), j);
        }
    } else if (p.is_f32) {
        const float *base = reinterpret_cast<const float *>(p.pos) + (size_t)f * p.fstride;
        // scalar head up to 16-byte alignment, float4 body (4 atoms = 3 x float4), scalar tail
        const unsigned long long addr = (unsigned long long)(uintptr_t)base;
        long long a0 = (addr & 3ull) ? n : 0;  // not even float aligned: all scalar
        while (a0 < n && ((addr + (unsigned long long)a0 * 12ull) & 15ull)) ++a0;
        const long long n_quads = (n - a0) / 4;
        const long long tail0 = a0 + n_quads * 4;
        const long long gtid = (long long)blockIdx.x * H_THREADS + threadIdx.x;
        const long long gstride = (long long)gridDim.x * H_THREADS;
        for (long long j = gtid; j < a0; j += gstride)
            deposit(atom_bin(g, (double)base[j * 3], (double)base[j * 3 + 1], (double)base[j * 3 + 2], edges), j);
        const float4 *vb = reinterpret_cast<const float4 *>(base + a0 * 3);
        for (long long qd = gtid; qd < n_quads; qd += gstride) {
            const float4 v0 = __ldg(vb + qd * 3), v1 = __ldg(vb + qd * 3 + 1), v2 = __ldg(vb + qd * 3 + 2);
            const long long j = a0 + qd * 4;
            const int b0 = atom_bin(g, (double)v0.x, (double)v0.y, (double)v0.z, edges);
            const int b1 = atom_bin(g, (double)v0.w, (double)v1.x, (double)v1.y, edges);
            const int b2 = atom_bin(g, (double)v1.z, (double)v1.w, (double)v2.x, edges);
            const int b3 = atom_bin(g, (double)v2.y, (double)v2.z, (double)v2.w, edges);
            deposit(b0, j); deposit(b1, j + 1); deposit(b2, j + 2); deposit(b3, j + 3);
        }
        for (long long j = tail0 + gtid; j < n; j += gstride)
            deposit(atom_bin(g, (double)base[j * 3], (double)base[j * 3 + 1], (double)base[j * 3 + 2], edges), j);
    } else {
        const double *base = reinterpret_cast<const double *>(p.pos) + (size_t)f * p.fstride;
        for (long long j = (long long)blockIdx.x * H_THREADS + threadIdx.x; j < n; j += (long long)gridDim.x * H_THREADS)
            deposit(atom_bin(g, base[j * 3], base[j * 3 + 1], base[j * 3 + 2], edges), j);
    }

    if (MODE != 2) {
        __syncthreads();
        unsigned long long *glo = p.acc_lo + (size_t)f * nb, *ghi = p.acc_hi + (size_t)f * nb;
        for (int i = threadIdx.x; i < nb; i += H_THREADS) {
            const unsigned int c = sh_c[i];
            if (!c) continue;
            atomicAdd(&gc[i], (unsigned long long)c);
            if (MODE == 0 && fx) {
                const unsigned long long lo64 = ((unsigned long long)sh_mid[i] << 32) | sh_lo[i];
                const unsigned long long old = atomicAdd(&glo[i], lo64);
                const unsigned long long hi = (unsigned long long)sh_hi[i] + ((old + lo64) < old ? 1ull : 0ull);
                if (hi) atomicAdd(&ghi[i], hi);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// k_hist_sorted: the weighted mode with ONE shared-memory atomic per element (default for finite weights).
//
// k_hist<0> spends three native ATOMS per element (count, low limb, middle limb) and the shared-memory atomic unit
// retires about one lane per cycle per SM, which caps it near 20 % of the HBM rate.  Here a CTA of 512 threads works on
// tiles of 4096 elements and counting-sorts each tile by bin:
//   1. bin of the element and rank = atomicAdd(&cnt[bin], 1) -- the only atomic, it is also the count.  The bin is
//      idx = trunc((x - first) * n / (last - first)) whenever the fractional part of that product is further than
//      `eps` from 0 and 1, eps being a bound on everything that separates the product from the position of x among
//      the float64 edges (three roundings of the product, two of each np.linspace edge); otherwise -- values sitting
//      on an edge, about one element in 1e11 for continuous data -- the edges decide exactly as in k_hist<0>, either
//      evaluated the way np.linspace does (fl(fl(i * step) + first), after the CTA has verified that the table it was
//      given is exactly that) or read from the table;
//   2. block scan of the tile counts -> segment offsets (order: thread, round; thread t owns bins t + 512 r);
//   3. scatter of the 54-bit fixed-point weights (same fixed(w) as k_hist<0>) to offset[bin] + rank;
//   4. the owner of a bin sums its segment with plain 64-bit adds and folds it into the CTA's private
//      {sum, carry, count} record of the bin (plain read-modify-write, the owner is the only writer).
// Integer sums are exact and order independent, so the result is BITWISE the one of k_hist<0>; the flush into the
// 128-bit global accumulators and k_hist_merge are shared.  Frames with non-finite weights take float64 global atomics.
constexpr int HS_THREADS = 512;
constexpr int HS_PER = 8;
constexpr int HS_TILE = HS_THREADS * HS_PER;

// shared memory through 32-bit window addresses (keeps the address arithmetic out of the generic space)
__device__ __forceinline__ unsigned int lds32(unsigned int a) { unsigned int v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(unsigned int a, unsigned int v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long lds64(unsigned int a) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(unsigned int a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ double ldsf64(unsigned int a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned int atoms_inc(unsigned int a) { unsigned int v; asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(a) : "memory"); return v; }
// per-bin record of a CTA: 96-bit sum (three limbs) + 32-bit count
__device__ __forceinline__ void lds_acc(unsigned int a, unsigned int &s0, unsigned int &s1, unsigned int &s2, unsigned int &count) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(count) : "r"(a));
}
__device__ __forceinline__ void sts_acc(unsigned int a, unsigned int s0, unsigned int s1, unsigned int s2, unsigned int count) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(s0), "r"(s1), "r"(s2), "r"(count) : "memory");
}
// {s2, s1, s0} += {carry, sum}
__device__ __forceinline__ void add96(unsigned int &s0, unsigned int &s1, unsigned int &s2, unsigned long long sum, unsigned int carry) {
    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, %5;"
        : "+r"(s0), "+r"(s1), "+r"(s2)
        : "r"((unsigned int)sum), "r"((unsigned int)(sum >> 32)), "r"(carry));
}

struct EdgeRef {
    unsigned int tab;  // shared-memory address of the edge table
    double first, last, step, inv_width, eps;
    int nb, arith;
    __device__ __forceinline__ double at(int i) const {
        if (arith) return i == nb ? last : __dadd_rn(__dmul_rn((double)i, step), first);
        return ldsf64(tab + 8u * (unsigned int)i);
    }
};

// same decisions as bin_of / bin_digitize above, edges through EdgeRef (rare path: kept out of line, scalar arguments
// so that the caller's EdgeRef stays in registers)
__device__ __noinline__ int find_bin_exact(unsigned int tab, double first, double last, double step, double inv_width,
                                           int nb, int arith, double x, int digitize) {
    EdgeRef e;
    e.tab = tab; e.first = first; e.last = last; e.step = step; e.inv_width = inv_width; e.nb = nb; e.arith = arith;
    int idx;
    if (digitize) {
        if (x != x) return nb - 1;
        const double gd = (x - first) * inv_width;
        idx = gd <= 0.0 ? 0 : (gd >= (double)(nb - 1) ? nb - 1 : (int)gd);
    } else {
        if (!(x >= first && x <= last)) return -1;
        idx = (int)((x - first) * inv_width);
        idx = min(max(idx, 0), nb - 1);
    }
    while (idx > 0 && x < e.at(idx)) --idx;
    while (idx < nb - 1 && x >= e.at(idx + 1)) ++idx;
    return idx;
}

__device__ __forceinline__ int find_bin(const EdgeRef &e, double x, int digitize) {
    const double gd = __dmul_rn(__dsub_rn(x, e.first), e.inv_width);
    const int idx = (int)gd;
    const double fr = gd - (double)idx;
    // strictly inside bin idx whatever the roundings (this also excludes NaN, out-of-range values and x == last)
    if (fr > e.eps && fr < 1.0 - e.eps && gd > 0.0 && idx < e.nb) return idx;
    return find_bin_exact(e.tab, e.first, e.last, e.step, e.inv_width, e.nb, e.arith, x, digitize);
}

// binned coordinate of an atom given as float32; false when histogram2d's axial window drops it
__device__ __forceinline__ bool atom_value_f32(const FrameGeom &g, float x, float y, float z, double &v) {
    if (g.geometry == MB200_GEOM_PLANAR) {
        v = (double)(g.d0 == 0 ? x : (g.d0 == 1 ? y : z));
        return true;
    }
    const double dx = (double)x, dy = (double)y, dz = (double)z;
    if (g.geometry == MB200_GEOM_CYLINDER) {
        const double ax = pick(dx, dy, dz, g.d0);
        if (!g.digitize && !(ax >= g.zlo && ax <= g.zhi)) return false;
        const double a = __dsub_rn(pick(dx, dy, dz, g.d1), pick(g.cx, g.cy, g.cz, g.d1));
        const double b = __dsub_rn(pick(dx, dy, dz, g.d2), pick(g.cx, g.cy, g.cz, g.d2));
        v = sqrt(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
    } else {
        const double a = __dsub_rn(dx, g.cx), b = __dsub_rn(dy, g.cy), c = __dsub_rn(dz, g.cz);
        v = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
    }
    return true;
}
__device__ __forceinline__ bool atom_value(const FrameGeom &g, double x, double y, double z, double &v) {
    if (g.geometry == MB200_GEOM_PLANAR) {
        v = pick(x, y, z, g.d0);
    } else if (g.geometry == MB200_GEOM_CYLINDER) {
        const double ax = pick(x, y, z, g.d0);
        if (!g.digitize && !(ax >= g.zlo && ax <= g.zhi)) return false;
        const double a = __dsub_rn(pick(x, y, z, g.d1), pick(g.cx, g.cy, g.cz, g.d1));
        const double b = __dsub_rn(pick(x, y, z, g.d2), pick(g.cx, g.cy, g.cz, g.d2));
        v = sqrt(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
    } else {
        const double a = __dsub_rn(x, g.cx), b = __dsub_rn(y, g.cy), c = __dsub_rn(z, g.cz);
        v = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
    }
    return true;
}

template <int BPT>
__global__ void __launch_bounds__(HS_THREADS, 2) k_hist_sorted(HistParams p) {
    extern __shared__ __align__(16) unsigned char hsm[];
    const int f = blockIdx.y, nb = p.n_bins, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // shared layout: sorted [HS_TILE] x 8 B | warp totals [16] x 4 B | acc [nb] x 16 B | edges [nb + 1] x 8 B |
    // cnt [nb + 1] x 4 B (the last counter collects the dropped elements).  The window addresses are made opaque to
    // the compiler so that they stay in registers instead of being rebuilt at every use.
    unsigned int s_sorted = (unsigned int)__cvta_generic_to_shared(hsm);
    unsigned int s_wt = s_sorted + 8u * HS_TILE;
    unsigned int s_acc = s_wt + 64u;
    unsigned int s_edges = s_acc + 16u * (unsigned int)nb;
    unsigned int s_cnt = s_edges + 8u * (unsigned int)(nb + 1);
    asm volatile("" : "+r"(s_sorted), "+r"(s_wt), "+r"(s_acc), "+r"(s_cnt));

    const double *g_edges = p.edges + (size_t)f * (nb + 1);
    EdgeRef E;
    E.tab = s_edges; E.nb = nb;
    E.first = g_edges[0]; E.last = g_edges[nb];
    const double width = __dsub_rn(E.last, E.first);
    E.step = __ddiv_rn(width, (double)nb);
    E.inv_width = (double)nb / width;
    {
        // |gd - t(x)| <= 3 nb u and |t(edge_i) - i| <= nb u (1 + 6 A / width), u = 2^-53, A = max(|first|, |last|); x2 margin
        const double A = fmax(fabs(E.first), fabs(E.last));
        const double e = 8.0 * (double)nb * 1.1102230246251565e-16 * (1.0 + 2.0 * A / fabs(width));
        E.eps = (width > 0.0 && e < 0.25) ? e : 2.0;  // 2: the quick test never passes
    }
    int lin = 1;
    for (int i = t; i <= nb; i += HS_THREADS) {
        const double e = g_edges[i];
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(s_edges + 8u * (unsigned int)i), "d"(e) : "memory");
        lin &= (e == (i == nb ? E.last : __dadd_rn(__dmul_rn((double)i, E.step), E.first))) ? 1 : 0;
    }
    for (int i = t; i <= nb; i += HS_THREADS) {
        sts32(s_cnt + 4u * (unsigned int)i, 0u);
        if (i < nb) sts_acc(s_acc + 16u * (unsigned int)i, 0u, 0u, 0u, 0u);
    }
    E.arith = __syncthreads_and(lin);
    if (!E.arith) E.eps = 2.0;  // hand-made edges: always decided by the table

    FrameGeom g;
    g.geometry = p.geometry; g.nb = nb; g.digitize = p.digitize;
    g.d0 = p.dim; g.d1 = (p.dim + 1) % 3; g.d2 = (p.dim + 2) % 3;
    g.cx = g.cy = g.cz = g.zlo = g.zhi = 0;
    if (p.geometry == MB200_GEOM_CYLINDER || p.geometry == MB200_GEOM_SPHERE) { g.cx = p.center[f * 3]; g.cy = p.center[f * 3 + 1]; g.cz = p.center[f * 3 + 2]; }
    if (p.geometry == MB200_GEOM_CYLINDER && p.zrange) { g.zlo = p.zrange[f * 2]; g.zhi = p.zrange[f * 2 + 1]; }
    const int dig = p.digitize;
    if (dig) E.eps = 2.0;  // digitize clamps instead of dropping: exact path
    const long long n = p.n;
    const double *w = p.weights + (p.w_per_frame ? (size_t)f * n : 0);
    double *gf = p.acc_f + (size_t)f * nb;
    unsigned long long *gc = p.acc_c + (size_t)f * nb;
    unsigned long long *glo = p.acc_lo + (size_t)f * nb, *ghi = p.acc_hi + (size_t)f * nb;

    const unsigned long long wb = p.wmax_bits[p.w_per_frame ? f : 0];
    const bool fx = wb < 0x7ff0000000000000ull;
    const double wmax = __longlong_as_double((long long)wb);
    const int kexp = bias_exponent(fx && wmax > 0.0 ? wmax : 1.0);
    const double bias = ldexp(1.0, kexp), inv_q = ldexp(1.0, 52 - kexp);

    // element j of the frame -> bin (or -1); generic fetch for every layout
    auto bin_generic = [&](long long j) -> int {
        double v;
        if (p.geometry == MB200_GEOM_SCALAR) {
            v = p.is_f32 ? (double)(reinterpret_cast<const float *>(p.pos) + (size_t)f * p.fstride)[j]
                         : (reinterpret_cast<const double *>(p.pos) + (size_t)f * p.fstride)[j];
        } else {
            double x, y, z;
            if (p.is_f32) {
                const float *b = reinterpret_cast<const float *>(p.pos) + (size_t)f * p.fstride + (size_t)j * 3;
                x = (double)b[0]; y = (double)b[1]; z = (double)b[2];
            } else {
                const double *b = reinterpret_cast<const double *>(p.pos) + (size_t)f * p.fstride + (size_t)j * 3;
                x = b[0]; y = b[1]; z = b[2];
            }
            if (!atom_value(g, x, y, z, v)) return -1;
        }
        return find_bin(E, v, dig);
    };
    auto fixed = [&](double wj) -> unsigned long long {
        return (unsigned long long)__double2ll_rn(__dmul_rn(__dadd_rn(wj, bias), inv_q));
    };

    if (!fx) {  // inf / nan weights in this frame: float64 atomics keep numpy's result
        for (long long j = (long long)blockIdx.x * HS_THREADS + t; j < n; j += (long long)gridDim.x * HS_THREADS) {
            const int b = bin_generic(j);
            if (b < 0) continue;
            atomicAdd(&gc[b], 1ull);
            atomicAdd(&gf[b], w[j]);
        }
        return;
    }

    // float32 [n][3] frames: 16-byte aligned body of 4-atom groups (3 x float4), at most 3 + 3 atoms around it
    long long a0 = 0, n_quads = 0;
    const bool fast = p.geometry != MB200_GEOM_SCALAR && p.is_f32 &&
                      (((unsigned long long)(uintptr_t)p.pos) & 3ull) == 0;
    const float *fbase = reinterpret_cast<const float *>(p.pos) + (size_t)f * p.fstride;
    if (fast) {
        const unsigned long long addr = (unsigned long long)(uintptr_t)fbase;
        while (a0 < n && ((addr + (unsigned long long)a0 * 12ull) & 15ull)) ++a0;
        n_quads = (n - a0) / 4;
    }
    const long long n_tiles = fast ? (n_quads + HS_TILE / 4 - 1) / (HS_TILE / 4) : (n + HS_TILE - 1) / HS_TILE;

    if (fast && blockIdx.x == 0) {  // the unaligned head and the tail of the frame: straight into the global accumulators
        const long long tail0 = a0 + n_quads * 4;
        const long long n_extra = a0 + (n - tail0);
        if (t < n_extra) {
            const long long j = t < a0 ? t : tail0 + (t - a0);
            const int b = bin_generic(j);
            if (b >= 0) {
                const unsigned long long v = fixed(w[j]);
                atomicAdd(&gc[b], 1ull);
                const unsigned long long old = atomicAdd(&glo[b], v);
                if ((old + v) < old) atomicAdd(&ghi[b], 1ull);
            }
        }
    }

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        unsigned int pk[HS_PER];
        unsigned long long val[HS_PER];
        int bins[HS_PER];
        if (fast) {
            const float4 *vb = reinterpret_cast<const float4 *>(fbase + a0 * 3);
            if (t == 0 && tile + gridDim.x < n_tiles) {  // the CTA's next tile on its way into L2 while this one is sorted
                const long long q0 = (tile + gridDim.x) * (HS_TILE / 4);
                const long long nq = min((long long)(HS_TILE / 4), n_quads - q0);
                const unsigned long long wa = (unsigned long long)(uintptr_t)(w + a0 + q0 * 4) & ~15ull;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(vb + q0 * 3), "r"((unsigned int)(nq * 48)) : "memory");
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(wa), "r"((unsigned int)(nq * 32)) : "memory");
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long qd = tile * (HS_TILE / 4) + h * HS_THREADS + t;
                if (qd < n_quads) {
                    const float4 v0 = __ldg(vb + qd * 3), v1 = __ldg(vb + qd * 3 + 1), v2 = __ldg(vb + qd * 3 + 2);
                    const double *wq = w + a0 + qd * 4;
                    const double w0 = __ldg(wq), w1 = __ldg(wq + 1), w2 = __ldg(wq + 2), w3 = __ldg(wq + 3);
                    double c;
                    bins[h * 4 + 0] = atom_value_f32(g, v0.x, v0.y, v0.z, c) ? find_bin(E, c, dig) : -1;
                    bins[h * 4 + 1] = atom_value_f32(g, v0.w, v1.x, v1.y, c) ? find_bin(E, c, dig) : -1;
                    bins[h * 4 + 2] = atom_value_f32(g, v1.z, v1.w, v2.x, c) ? find_bin(E, c, dig) : -1;
                    bins[h * 4 + 3] = atom_value_f32(g, v2.y, v2.z, v2.w, c) ? find_bin(E, c, dig) : -1;
                    val[h * 4 + 0] = fixed(w0); val[h * 4 + 1] = fixed(w1); val[h * 4 + 2] = fixed(w2); val[h * 4 + 3] = fixed(w3);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { bins[h * 4 + i] = -1; val[h * 4 + i] = 0ull; }
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < HS_PER; ++s) {
                const long long j = tile * HS_TILE + s * HS_THREADS + t;
                bins[s] = -1; val[s] = 0ull;
                if (j < n) {
                    bins[s] = bin_generic(j);
                    val[s] = fixed(w[j]);
                }
            }
        }
        __syncthreads();  // E: the previous tile's owners are done with cnt[] and sorted[]
        // dropped elements (and the empty slots of a partial tile) are counted in cnt[nb] and sorted behind everything
#pragma unroll
        for (int s = 0; s < HS_PER; ++s) {
            const unsigned int b = bins[s] >= 0 ? (unsigned int)bins[s] : (unsigned int)nb;
            pk[s] = (b << 13) | atoms_inc(s_cnt + 4u * b);
        }
        __syncthreads();  // A

        // segment offsets, order (thread, round): thread t owns bins t + 512 r
        unsigned int c[BPT];
        unsigned int mine = 0;
#pragma unroll
        for (int r = 0; r < BPT; ++r) {
            const int b = t + r * HS_THREADS;
            c[r] = b < nb ? lds32(s_cnt + 4u * (unsigned int)b) : 0u;
            mine += c[r];
        }
        unsigned int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) sts32(s_wt + 4u * (unsigned int)warp, incl);
        __syncthreads();  // B
        unsigned int base;
        {
            const unsigned int wv = lane < HS_THREADS / 32 ? lds32(s_wt + 4u * (unsigned int)lane) : 0u;
            unsigned int wi = wv;
#pragma unroll
            for (int o = 1; o < HS_THREADS / 32; o <<= 1) {
                const unsigned int up = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += up;
            }
            base = __shfl_sync(0xffffffffu, wi - wv, warp) + incl - mine;
            unsigned int o = base;
#pragma unroll
            for (int r = 0; r < BPT; ++r) {
                const int b = t + r * HS_THREADS;
                if (b < nb) sts32(s_cnt + 4u * (unsigned int)b, o);
                o += c[r];
            }
            if (t == HS_THREADS - 1) sts32(s_cnt + 4u * (unsigned int)nb, o);  // the dropped elements go last
        }
        __syncthreads();  // C
#pragma unroll
        for (int s = 0; s < HS_PER; ++s)
            sts64(s_sorted + 8u * (lds32(s_cnt + ((pk[s] >> 11) & 0x1ffffcu)) + (pk[s] & 8191u)), val[s]);
        __syncthreads();  // D
        unsigned int src = s_sorted + 8u * base;
        if (t == HS_THREADS - 1) sts32(s_cnt + 4u * (unsigned int)nb, 0u);
        const bool wide = mine > 2048u;  // 2048 values below 2^53 cannot wrap a 64-bit sum
#pragma unroll
        for (int r = 0; r < BPT; ++r) {
            const int b = t + r * HS_THREADS;
            if (c[r]) {
                unsigned long long sum = 0ull;
                unsigned int carry = 0u;
                const unsigned int end = src + 8u * c[r];
                if (!wide) {
#pragma unroll 1
                    for (; src != end; src += 8u) sum += lds64(src);
                } else {
#pragma unroll 1
                    for (; src != end; src += 8u) {
                        const unsigned long long x = lds64(src);
                        sum += x;
                        carry += sum < x ? 1u : 0u;
                    }
                }
                unsigned int s0, s1, s2, cn;
                const unsigned int aa = s_acc + 16u * (unsigned int)b;
                lds_acc(aa, s0, s1, s2, cn);
                add96(s0, s1, s2, sum, carry);
                sts_acc(aa, s0, s1, s2, cn + c[r]);
            }
            if (b < nb) sts32(s_cnt + 4u * (unsigned int)b, 0u);  // it held the segment offset
        }
    }
    __syncthreads();

    // owners flush their bins into the 128-bit global accumulators
#pragma unroll
    for (int r = 0; r < BPT; ++r) {
        const int b = t + r * HS_THREADS;
        if (b >= nb) continue;
        unsigned int s0, s1, s2, cn;
        lds_acc(s_acc + 16u * (unsigned int)b, s0, s1, s2, cn);
        if (!cn) continue;
        atomicAdd(&gc[b], (unsigned long long)cn);
        const unsigned long long lo64 = ((unsigned long long)s1 << 32) | s0;
        const unsigned long long old = atomicAdd(&glo[b], lo64);
        const unsigned long long hi = (unsigned long long)s2 + ((old + lo64) < old ? 1ull : 0ull);
        if (hi) atomicAdd(&ghi[b], hi);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_hist_chunk: the weighted mode with FOUR non-returning shared-memory atomics per element and nothing else
// (default for finite weights since round 2).
//
// Round 1 assumed that the shared-memory atomic unit retires one lane per clock and SM and built the counting sort
// above around that.  tools/microbench/smem_atomics.cu (profiles/r02_smem_atomics.log) measures 8.5 - 10 lanes per clock
// and SM for 32-bit adds on random bins, returning or not: what bounded k_hist<1> / k_hist_sorted was everything AROUND
// the atomic (data-dependent edge loads, the sort's scan, scatter and segment sums on the shared-memory pipe).  So:
//   * the bin comes from arithmetic alone (find_bin: the edge table is only consulted for values within eps of an edge);
//   * the 54-bit fixed-point weight fixed(w) (same as k_hist<0>) is cut into three 18-bit chunks and each chunk is added
//     to its own 32-bit counter of the bin with a plain `red.shared.add.u32`; the count is a fourth one.  16384 chunks
//     below 2^18 cannot wrap a 32-bit counter, so after every four tiles of 4096 elements the owner of a bin (thread
//     t owns bins t + 512 r) folds the four counters into the CTA's private {96-bit sum, count} record of the bin
//     with plain loads and stores and clears them: no sort, no scan, two block barriers per 16384 elements;
//   * the flush into the 128-bit global accumulators and k_hist_merge are the ones of the other kernels: integer sums
//     are exact and order independent, the result is BITWISE theirs.
// 5 x 4 lanes of atomics per element against ~9.7 lanes per clock: 2.4 elements per clock and SM = 8.4 TB/s of 12-byte
// elements -- above what HBM delivers, so the kernel streams at the memory rate.
constexpr int HC_THREADS = 512;
constexpr int HC_PER = 8;
constexpr int HC_TILE = HC_THREADS * HC_PER;
constexpr int HC_FOLD_TILES = 4;  // 4 x 4096 = 16384 elements between folds: 16384 x (2^18 - 1) < 2^32

__device__ __forceinline__ void red_add32(unsigned int a, unsigned int v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

__global__ void __launch_bounds__(HC_THREADS, 2) k_hist_chunk(HistParams p) {
    extern __shared__ __align__(16) unsigned char hsm[];
    const int f = blockIdx.y, nb = p.n_bins, t = threadIdx.x;
    // shared layout: acc [nb] x 16 B | edges [nb + 1] x 8 B | cnt, c0, c1, c2 [nb] x 4 B each
    unsigned int s_acc = (unsigned int)__cvta_generic_to_shared(hsm);
    unsigned int s_edges = s_acc + 16u * (unsigned int)nb;
    unsigned int s_cnt = s_edges + 8u * (unsigned int)(nb + 1);
    const unsigned int plane = 4u * (unsigned int)nb;  // byte distance of the four counter planes
    asm volatile("" : "+r"(s_acc), "+r"(s_cnt));

    const double *g_edges = p.edges + (size_t)f * (nb + 1);
    EdgeRef E;
    E.tab = s_edges; E.nb = nb;
    E.first = g_edges[0]; E.last = g_edges[nb];
    const double width = __dsub_rn(E.last, E.first);
    E.step = __ddiv_rn(width, (double)nb);
    E.inv_width = (double)nb / width;
    {
        const double A = fmax(fabs(E.first), fabs(E.last));
        const double e = 8.0 * (double)nb * 1.1102230246251565e-16 * (1.0 + 2.0 * A / fabs(width));
        E.eps = (width > 0.0 && e < 0.25) ? e : 2.0;
    }
    int lin = 1;
    for (int i = t; i <= nb; i += HC_THREADS) {
        const double e = g_edges[i];
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(s_edges + 8u * (unsigned int)i), "d"(e) : "memory");
        lin &= (e == (i == nb ? E.last : __dadd_rn(__dmul_rn((double)i, E.step), E.first))) ? 1 : 0;
    }
    for (int i = t; i < nb; i += HC_THREADS) {
        sts_acc(s_acc + 16u * (unsigned int)i, 0u, 0u, 0u, 0u);
#pragma unroll
        for (unsigned int k = 0; k < 4u; ++k) sts32(s_cnt + k * plane + 4u * (unsigned int)i, 0u);
    }
    E.arith = __syncthreads_and(lin);
    if (!E.arith) E.eps = 2.0;  // hand-made edges: always decided by the table

    FrameGeom g;
    g.geometry = p.geometry; g.nb = nb; g.digitize = p.digitize;
    g.d0 = p.dim; g.d1 = (p.dim + 1) % 3; g.d2 = (p.dim + 2) % 3;
    g.cx = g.cy = g.cz = g.zlo = g.zhi = 0;
    if (p.geometry == MB200_GEOM_CYLINDER || p.geometry == MB200_GEOM_SPHERE) { g.cx = p.center[f * 3]; g.cy = p.center[f * 3 + 1]; g.cz = p.center[f * 3 + 2]; }
    if (p.geometry == MB200_GEOM_CYLINDER && p.zrange) { g.zlo = p.zrange[f * 2]; g.zhi = p.zrange[f * 2 + 1]; }
    const int dig = p.digitize;
    if (dig) E.eps = 2.0;
    const long long n = p.n;
    const double *w = p.weights + (p.w_per_frame ? (size_t)f * n : 0);
    double *gf = p.acc_f + (size_t)f * nb;
    unsigned long long *gc = p.acc_c + (size_t)f * nb;
    unsigned long long *glo = p.acc_lo + (size_t)f * nb, *ghi = p.acc_hi + (size_t)f * nb;

    const unsigned long long wb = p.wmax_bits[p.w_per_frame ? f : 0];
    const bool fx = wb < 0x7ff0000000000000ull;
    const double wmax = __longlong_as_double((long long)wb);
    const int kexp = bias_exponent(fx && wmax > 0.0 ? wmax : 1.0);
    const double bias = ldexp(1.0, kexp), inv_q = ldexp(1.0, 52 - kexp);

    auto bin_generic = [&](long long j) -> int {
        double v;
        if (p.geometry == MB200_GEOM_SCALAR) {
            v = p.is_f32 ? (double)(reinterpret_cast<const float *>(p.pos) + (size_t)f * p.fstride)[j]
                         : (reinterpret_cast<const double *>(p.pos) + (size_t)f * p.fstride)[j];
        } else {
            double x, y, z;
            if (p.is_f32) {
                const float *b = reinterpret_cast<const float *>(p.pos) + (size_t)f * p.fstride + (size_t)j * 3;
                x = (double)b[0]; y = (double)b[1]; z = (double)b[2];
            } else {
                const double *b = reinterpret_cast<const double *>(p.pos) + (size_t)f * p.fstride + (size_t)j * 3;
                x = b[0]; y = b[1]; z = b[2];
            }
            if (!atom_value(g, x, y, z, v)) return -1;
        }
        return find_bin(E, v, dig);
    };
    auto fixed = [&](double wj) -> unsigned long long {
        return (unsigned long long)__double2ll_rn(__dmul_rn(__dadd_rn(wj, bias), inv_q));
    };
    auto deposit = [&](int b, unsigned long long v) {  // v <= 2^53: chunks of 18 bits
        if (b < 0) return;
        const unsigned int a = s_cnt + 4u * (unsigned int)b;
        red_add32(a, 1u);
        red_add32(a + plane, (unsigned int)v & 0x3ffffu);
        red_add32(a + 2u * plane, (unsigned int)(v >> 18) & 0x3ffffu);
        red_add32(a + 3u * plane, (unsigned int)(v >> 36));
    };
    auto fold = [&]() {  // between two block barriers: owners move the counters into the wide records
        for (int b = t; b < nb; b += HC_THREADS) {
            const unsigned int a = s_cnt + 4u * (unsigned int)b;
            const unsigned int cn = lds32(a);
            if (!cn) continue;
            const unsigned long long c0 = lds32(a + plane), c1 = lds32(a + 2u * plane), c2 = lds32(a + 3u * plane);
            // c0 + c1 2^18 + c2 2^36 < 2^69: low 64 bits with carry, then the top limb
            const unsigned long long lo_a = c0 + (c1 << 18);                 // < 2^51
            const unsigned long long lo_b = c2 << 36;                        // low 64 bits of c2 2^36
            const unsigned long long lo = lo_a + lo_b;
            const unsigned int top = (unsigned int)(c2 >> 28) + (lo < lo_a ? 1u : 0u);
            unsigned int s0, s1, s2, count;
            lds_acc(s_acc + 16u * (unsigned int)b, s0, s1, s2, count);
            add96(s0, s1, s2, lo, top);
            sts_acc(s_acc + 16u * (unsigned int)b, s0, s1, s2, count + cn);
            sts32(a, 0u); sts32(a + plane, 0u); sts32(a + 2u * plane, 0u); sts32(a + 3u * plane, 0u);
        }
    };

    if (!fx) {  // inf / nan weights in this frame: float64 atomics keep numpy's result
        for (long long j = (long long)blockIdx.x * HC_THREADS + t; j < n; j += (long long)gridDim.x * HC_THREADS) {
            const int b = bin_generic(j);
            if (b < 0) continue;
            atomicAdd(&gc[b], 1ull);
            atomicAdd(&gf[b], w[j]);
        }
        return;
    }

    long long a0 = 0, n_quads = 0;
    // float32 [n][3] frames.  Slab profiles need ONE coordinate per atom: consecutive lanes read it from consecutive atoms
    // (stride 12 B: a warp touches three cache lines per load, every fetched sector is used by the neighbouring lanes'
    // other loads) -- the three float4 per four atoms of the radial geometries cost three times the L1 wavefronts
    // (ncu on the first version of this kernel: L1TEX pipe 83 % busy, 64 % excessive sectors).
    const bool planar32 = p.geometry == MB200_GEOM_PLANAR && p.is_f32 && !dig;
    const bool fast = !planar32 && p.geometry != MB200_GEOM_SCALAR && p.is_f32 && (((unsigned long long)(uintptr_t)p.pos) & 3ull) == 0;
    const float *fbase = reinterpret_cast<const float *>(p.pos) + (size_t)f * p.fstride;
    if (fast) {
        const unsigned long long addr = (unsigned long long)(uintptr_t)fbase;
        while (a0 < n && ((addr + (unsigned long long)a0 * 12ull) & 15ull)) ++a0;
        n_quads = (n - a0) / 4;
    }
    const long long n_tiles = fast ? (n_quads + HC_TILE / 4 - 1) / (HC_TILE / 4) : (n + HC_TILE - 1) / HC_TILE;

    if (fast && blockIdx.x == 0) {  // the unaligned head and the tail of the frame: straight into the global accumulators
        const long long tail0 = a0 + n_quads * 4;
        const long long n_extra = a0 + (n - tail0);
        if (t < n_extra) {
            const long long j = t < a0 ? t : tail0 + (t - a0);
            const int b = bin_generic(j);
            if (b >= 0) {
                const unsigned long long v = fixed(w[j]);
                atomicAdd(&gc[b], 1ull);
                const unsigned long long old = atomicAdd(&glo[b], v);
                if ((old + v) < old) atomicAdd(&ghi[b], 1ull);
            }
        }
    }

    int since_fold = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (planar32) {
            float zv[HC_PER];
            double wv[HC_PER];
            const float *zb = fbase + g.d0;
#pragma unroll
            for (int s = 0; s < HC_PER; ++s) {  // all loads of the tile first: eight independent pairs in flight per thread
                const long long j = tile * HC_TILE + s * HC_THREADS + t;
                zv[s] = j < n ? __ldg(zb + 3 * j) : 0.f;
                wv[s] = j < n ? __ldg(w + j) : 0.0;
            }
#pragma unroll
            for (int s = 0; s < HC_PER; ++s) {
                const long long j = tile * HC_TILE + s * HC_THREADS + t;
                if (j < n) deposit(find_bin(E, (double)zv[s], 0), fixed(wv[s]));
            }
        } else if (fast) {
            const float4 *vb = reinterpret_cast<const float4 *>(fbase + a0 * 3);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long qd = tile * (HC_TILE / 4) + h * HC_THREADS + t;
                if (qd < n_quads) {
                    const float4 v0 = __ldg(vb + qd * 3), v1 = __ldg(vb + qd * 3 + 1), v2 = __ldg(vb + qd * 3 + 2);
                    const double *wq = w + a0 + qd * 4;
                    const double w0 = __ldg(wq), w1 = __ldg(wq + 1), w2 = __ldg(wq + 2), w3 = __ldg(wq + 3);
                    double c;
                    const int b0 = atom_value_f32(g, v0.x, v0.y, v0.z, c) ? find_bin(E, c, dig) : -1;
                    const int b1 = atom_value_f32(g, v0.w, v1.x, v1.y, c) ? find_bin(E, c, dig) : -1;
                    const int b2 = atom_value_f32(g, v1.z, v1.w, v2.x, c) ? find_bin(E, c, dig) : -1;
                    const int b3 = atom_value_f32(g, v2.y, v2.z, v2.w, c) ? find_bin(E, c, dig) : -1;
                    deposit(b0, fixed(w0)); deposit(b1, fixed(w1)); deposit(b2, fixed(w2)); deposit(b3, fixed(w3));
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < HC_PER; ++s) {
                const long long j = tile * HC_TILE + s * HC_THREADS + t;
                if (j < n) deposit(bin_generic(j), fixed(w[j]));
            }
        }
        if (++since_fold == HC_FOLD_TILES) {
            __syncthreads();
            fold();
            __syncthreads();
            since_fold = 0;
        }
    }
    __syncthreads();
    fold();
    __syncthreads();

    // owners flush their bins into the 128-bit global accumulators
    for (int b = t; b < nb; b += HC_THREADS) {
        unsigned int s0, s1, s2, cn;
        lds_acc(s_acc + 16u * (unsigned int)b, s0, s1, s2, cn);
        if (!cn) continue;
        atomicAdd(&gc[b], (unsigned long long)cn);
        const unsigned long long lo64 = ((unsigned long long)s1 << 32) | s0;
        const unsigned long long old = atomicAdd(&glo[b], lo64);
        const unsigned long long hi = (unsigned long long)s2 + ((old + lo64) < old ? 1ull : 0ull);
        if (hi) atomicAdd(&ghi[b], hi);
    }
}

// out_w = (acc - count * 2^52) * Q  (exact in __int128), or the float64 fallback sums
__global__ void k_hist_merge(HistParams p, int fixed_point, double *__restrict__ out_w, long long *__restrict__ out_c,
                             long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long f = i / p.n_bins;
    const unsigned long long c = p.acc_c[i];
    out_c[i] = (long long)c;
    if (!out_w) return;
    if (!fixed_point) {
        out_w[i] = p.acc_f[i];
        return;
    }
    const unsigned long long wb = p.wmax_bits[p.w_per_frame ? f : 0];
    if (wb >= 0x7ff0000000000000ull) {
        out_w[i] = p.acc_f[i];
        return;
    }
    const double wmax = __longlong_as_double((long long)wb);
    const int k = bias_exponent(wmax > 0.0 ? wmax : 1.0);
    const unsigned __int128 acc = ((unsigned __int128)p.acc_hi[i] << 64) | p.acc_lo[i];
    const __int128 v = (__int128)acc - ((__int128)c << 52);
    // |v| < 2^(54 + log2 n): split into two exactly representable halves before scaling
    const bool neg = v < 0;
    const unsigned __int128 a = neg ? (unsigned __int128)(-v) : (unsigned __int128)v;
    const double hi = (double)(unsigned long long)(a >> 32), lo = (double)(unsigned long long)(a & 0xffffffffull);
    const double mag = ldexp(hi, 32) + lo;  // one rounding (hi * 2^32 is exact when a < 2^85)
    out_w[i] = ldexp(neg ? -mag : mag, k - 52);
}


}  // namespace mb200

using namespace mb200;


extern "C" int mb200_profile_hist(mb200_ctx *ctx, int geometry, const void *positions, int pos_is_f32,
                                  int positions_on_device, int64_t n_frames, int64_t n, int dim, const double *weights,
                                  int weights_per_frame, int weights_on_device, const double *edges, int32_t n_bins,
                                  const double *center, const double *zrange, double *hist_w, int64_t *hist_c) {
    return mb200_profile_hist_window(ctx, geometry, positions, pos_is_f32, positions_on_device, n_frames, n, 0, n, nullptr,
                                     nullptr, dim, weights, weights_per_frame, weights_on_device, edges, n_bins, center, zrange,
                                     hist_w, hist_c);
}

extern "C" int mb200_profile_hist_window(mb200_ctx *ctx, int geometry, const void *positions, int pos_is_f32,
                                         int positions_on_device, int64_t n_frames, int64_t n_total, int64_t first_atom,
                                         int64_t n, const mb200_groups *selection, const float *pack_boxes, int dim,
                                         const double *weights, int weights_per_frame,
                                         int weights_on_device, const double *edges, int32_t n_bins, const double *center,
                                         const double *zrange, double *hist_w, int64_t *hist_c) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    const int digitize = (geometry & MB200_BIN_DIGITIZE) ? 1 : 0;
    geometry &= ~MB200_BIN_DIGITIZE;
    if (geometry < 0 || geometry > 3 || n_frames < 0 || n < 0 || n_bins <= 0 || dim < 0 || dim > 2 || !edges || !hist_c ||
        (n && !positions) || (weights && !hist_w) || first_atom < 0 || first_atom + n > n_total)
        return fail(MB200_EINVAL, "invalid argument to mb200_profile_hist");
    if ((pack_boxes || selection) && (!pos_is_f32 || geometry == MB200_GEOM_SCALAR))
        return fail(MB200_EINVAL, "pack_boxes / selection apply to float32 atom positions");
    if (selection && (selection->device != ctx->device || selection->ptr.size() != 2 || selection->ptr[1] != n ||
                      selection->n_atoms_total != n_total || first_atom != 0))
        return fail(MB200_EINVAL, "selection must be one group of n atoms of frames of n_total atoms on this device");
    if ((geometry == MB200_GEOM_CYLINDER || geometry == MB200_GEOM_SPHERE) && !center)
        return fail(MB200_EINVAL, "center required for radial geometries");
    if (geometry == MB200_GEOM_CYLINDER && !zrange && !digitize) return fail(MB200_EINVAL, "zrange required for cylinder geometry");
    if (n_frames == 0) return MB200_OK;
    if (n_frames > 65535) return fail(MB200_EINVAL, "at most 65535 frames per call");
    MB_CUDA(cudaSetDevice(ctx->device));
    const size_t F = (size_t)n_frames, N = (size_t)n, NB = (size_t)n_bins;
    const size_t esz = pos_is_f32 ? 4 : 8;
    const size_t ncomp = geometry == MB200_GEOM_SCALAR ? 1 : 3;  // values per element in `positions`

    // developer aid: MAICOS_B200_TRACE_HOST=1 prints the host-side time of every phase of the call to stderr
    static const bool trace_host = std::getenv("MAICOS_B200_TRACE_HOST") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (!trace_host) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[mb200_profile_hist] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    mark("entry");
    HistParams p;
    std::memset(&p, 0, sizeof(p));
    PrefetchRead pref_read(ctx);
    // the group is atoms [first_atom, first_atom + n) of whole frames of n_total atoms (a window; n_total = n: the frames
    // hold the group only)
    const size_t NT = (size_t)n_total, win = (size_t)first_atom * ncomp * esz;
    p.fstride = (long long)(NT * ncomp);
    if (positions_on_device || N == 0) {
        p.pos = reinterpret_cast<const char *>(positions) + win;
    } else if (const void *pref = pref_read.take(positions, F * NT * ncomp * esz)) {
        // announced with mb200_prefetch_frames: copied on the copy stream under the previous call
        p.pos = reinterpret_cast<const char *>(pref) + win;
    } else {
        MB_TRY(ctx->d_hpos.ensure(F * (selection ? NT : N) * ncomp * esz));
        if (NT == N || selection) {  // (a selection is gathered on the device: its whole frames travel)
            MB_CUDA(cudaMemcpyAsync(ctx->d_hpos.p, positions, F * NT * ncomp * esz, cudaMemcpyHostToDevice, ctx->stream));
        } else {  // only the window travels
            MB_CUDA(cudaMemcpy2DAsync(ctx->d_hpos.p, N * ncomp * esz, reinterpret_cast<const char *>(positions) + win,
                                      NT * ncomp * esz, N * ncomp * esz, F, cudaMemcpyHostToDevice, ctx->stream));
        }
        p.pos = ctx->d_hpos.p;
        p.fstride = (long long)((selection ? NT : N) * ncomp);
    }
    const bool weighted = weights && N;
    const size_t Fw = weights_per_frame ? F : 1;
    if (weighted) {
        if (weights_on_device) {
            p.weights = weights;
        } else {
            MB_TRY(ctx->d_hw.ensure(Fw * N * 8));
            MB_CUDA(cudaMemcpyAsync(ctx->d_hw.p, weights, Fw * N * 8, cudaMemcpyHostToDevice, ctx->stream));
            p.weights = ctx->d_hw.as<double>();
        }
    }
    // small per-frame parameters: edges | center | zrange | wmax bits -- gathered in the pinned staging block, ONE copy
    const size_t n_box = pack_boxes ? (F * 3 + 1) / 2 : 0;  // float32 [F][3] in units of doubles
    const size_t n_par = F * (NB + 1) + F * 3 + F * 2 + n_box + F;
    MB_TRY(ctx->d_hpar.ensure(n_par * 8));
    MB_TRY(ctx->h_stage.ensure(n_par * 8));  // the previous call synchronised the stream before it returned
    double *dpar = ctx->d_hpar.as<double>();
    double *hpar = ctx->h_stage.as<double>();
    std::memcpy(hpar, edges, F * (NB + 1) * 8);
    p.edges = dpar;
    if (center) {
        std::memcpy(hpar + F * (NB + 1), center, F * 24);
        p.center = dpar + F * (NB + 1);
    }
    if (zrange) {
        std::memcpy(hpar + F * (NB + 1) + F * 3, zrange, F * 16);
        p.zrange = dpar + F * (NB + 1) + F * 3;
    }
    if (pack_boxes) std::memcpy(hpar + F * (NB + 1) + F * 5, pack_boxes, F * 12);
    MB_CUDA(cudaMemcpyAsync(dpar, hpar, (F * (NB + 1) + F * 5 + n_box) * 8, cudaMemcpyHostToDevice, ctx->stream));
    unsigned long long *d_wmax = reinterpret_cast<unsigned long long *>(dpar + F * (NB + 1) + F * 5 + n_box);
    if ((pack_boxes || selection) && N) {  // gather / wrap into the primary cell on the device: dense float32 copy of the group
        MB_TRY(ctx->d_hpack.ensure(F * N * 12));
        const unsigned gx = (unsigned)std::max<size_t>(1, std::min<size_t>((N * 3 + 256 * 8 - 1) / (256 * 8), (size_t)ctx->n_sm * 8 / F + 1));
        k_pack_atoms<<<dim3(gx, (unsigned)F), 256, 0, ctx->stream>>>(reinterpret_cast<const float *>(p.pos), p.fstride,
                                                                     ctx->d_hpack.as<float>(),
                                                                     pack_boxes ? reinterpret_cast<const float *>(dpar + F * (NB + 1) + F * 5) : nullptr,
                                                                     selection ? selection->idx.as<int32_t>() : nullptr,
                                                                     (long long)(N * 3));
        MB_CUDA(cudaGetLastError());
        ctx->launches += 1;
        p.pos = ctx->d_hpack.p;
        p.fstride = (long long)(N * 3);
    }
    p.wmax_bits = d_wmax;
    // accumulators [F][NB]: lo, hi, f64, count + outputs [F][NB] (w, c)
    const size_t acc = F * NB;
    MB_TRY(ctx->d_hout.ensure(acc * 8 * 6));
    p.acc_lo = ctx->d_hout.as<unsigned long long>();
    p.acc_hi = p.acc_lo + acc;
    p.acc_f = reinterpret_cast<double *>(p.acc_hi + acc);
    p.acc_c = reinterpret_cast<unsigned long long *>(p.acc_f + acc);
    double *d_ow = reinterpret_cast<double *>(p.acc_c + acc);
    long long *d_oc = reinterpret_cast<long long *>(d_ow + acc);
    MB_CUDA(cudaMemsetAsync(p.acc_lo, 0, acc * 8 * 4, ctx->stream));
    p.n = n; p.n_bins = n_bins; p.dim = dim; p.geometry = geometry; p.is_f32 = pos_is_f32; p.w_per_frame = weights_per_frame; p.digitize = digitize;

    const long long total = (long long)(F * NB);
    {
    // mode: 0 fixed-point smem atomics, 1 counts-only smem atomics, 2 float64 global atomics
    int mode = weighted ? 0 : 1;
    if (n_bins > H_MAX_SMEM_BINS) mode = 2;
    if (weighted && mode == 0) {
        MB_CUDA(cudaMemsetAsync(d_wmax, 0, Fw * 8, ctx->stream));
        const unsigned bx = (unsigned)std::max<size_t>(1, std::min<size_t>((N + 256 * 8 - 1) / (256 * 8), (size_t)ctx->n_sm * 4 / Fw + 1));
        k_hist_wmax<<<dim3(bx, (unsigned)Fw), 256, 0, ctx->stream>>>(p.weights, n, d_wmax);
        MB_CUDA(cudaGetLastError());
        ctx->launches += 1;
        // frames with non-finite weights are detected on the device and use float64 atomics
    }
    if (N) {
        // enough CTAs to fill the GPU across frames, each streaming >= 16 elements per thread
        const size_t smem = mode == 2 ? 0 : (NB + 1) * 8 + NB * 4 * (mode == 0 ? 4 : 1);
        if (!ctx->hist_attr_set) {  // per device: a process may own several contexts
            MB_CUDA(cudaFuncSetAttribute(k_hist<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            MB_CUDA(cudaFuncSetAttribute(k_hist<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            ctx->hist_attr_set = true;
        }
        long long per_frame = (long long)((N + (size_t)H_THREADS * 16 - 1) / ((size_t)H_THREADS * 16));
        long long want = ((long long)ctx->n_sm * 4 + (long long)F - 1) / (long long)F;
        unsigned bx = (unsigned)std::max<long long>(1, std::min(per_frame, std::max<long long>(want, 1)));
        dim3 grid(bx, (unsigned)F);
        if (ctx->timing) MB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        if (mode == 0 && ctx->hist_mode == 0) {
            // chunked non-returning atomics: one or two resident CTAs of 512 threads per SM, the grid fills them once
            const size_t csm = NB * 16 + (NB + 1) * 8 + NB * 16 + 16;
            if (!ctx->hc_attr_set) {
                MB_CUDA(cudaFuncSetAttribute(k_hist_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                ctx->hc_attr_set = true;
            }
            const long long slots = (long long)ctx->n_sm * (csm <= 112 * 1024 ? 2 : 1);
            const long long n_tiles = (long long)((N + HC_TILE - 1) / HC_TILE);
            const unsigned cbx = (unsigned)std::max<long long>(1, std::min(n_tiles, std::max<long long>(1, slots / (long long)F)));
            k_hist_chunk<<<dim3(cbx, (unsigned)F), HC_THREADS, csm, ctx->stream>>>(p);
        } else if (mode == 0 && ctx->hist_mode == 2) {
            // counting-sort kernel: one or two resident CTAs of 512 threads per SM, the grid fills them once
            const size_t ssm = NB * 16 + (size_t)HS_TILE * 8 + (NB + 1) * 8 + (NB + 1) * 4 + 64 + 16;
            const long long slots = (long long)ctx->n_sm * (ssm <= 112 * 1024 ? 2 : 1);
            const long long n_tiles = (long long)((N + HS_TILE - 1) / HS_TILE);
            const unsigned sbx = (unsigned)std::max<long long>(1, std::min(n_tiles, slots / (long long)F));
            const dim3 sgrid(sbx, (unsigned)F);
            const int bpt = (int)((NB + HS_THREADS - 1) / HS_THREADS);
#define MB_HS_CASE(B)                                                                                              \
    case B:                                                                                                        \
        if (!(ctx->hs_attr_set & (1u << B))) {                                                                     \
            MB_CUDA(cudaFuncSetAttribute(k_hist_sorted<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); \
            ctx->hs_attr_set |= 1u << B;                                                                           \
        }                                                                                                          \
        k_hist_sorted<B><<<sgrid, HS_THREADS, ssm, ctx->stream>>>(p);                                              \
        break;
            switch (bpt) {
                MB_HS_CASE(1) MB_HS_CASE(2) MB_HS_CASE(3) MB_HS_CASE(4) MB_HS_CASE(5) MB_HS_CASE(6) MB_HS_CASE(7) MB_HS_CASE(8)
                default: return fail(MB200_EINVAL, "internal: bins per thread out of range");
            }
#undef MB_HS_CASE
        } else if (mode == 0) k_hist<0><<<grid, H_THREADS, smem, ctx->stream>>>(p);
        else if (mode == 1) k_hist<1><<<grid, H_THREADS, smem, ctx->stream>>>(p);
        else k_hist<2><<<grid, H_THREADS, 0, ctx->stream>>>(p);
        MB_CUDA(cudaGetLastError());
        if (ctx->timing) MB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        ctx->launches += 1;
    }
    k_hist_merge<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(p, mode == 0, hist_w ? d_ow : nullptr, d_oc, total);
    MB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    }
    mark(pref_read.slot >= 0 ? "queued (prefetched)" : "queued (own copy)");
    pref_read.release();  // the prefetch slot this call read may be refilled from here on
    // results: ONE device -> pinned copy of (weighted | counts), then into the caller's (pageable) arrays
    MB_TRY(ctx->h_out.ensure(F * NB * 16));
    char *hres = ctx->h_out.as<char>();
    if (hist_w) {
        MB_CUDA(cudaMemcpyAsync(hres, d_ow, F * NB * 16, cudaMemcpyDeviceToHost, ctx->stream));  // d_ow | d_oc are adjacent
    } else {
        MB_CUDA(cudaMemcpyAsync(hres + F * NB * 8, d_oc, F * NB * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    mark("device (wait)");
    if (hist_w) std::memcpy(hist_w, hres, F * NB * 8);
    std::memcpy(hist_c, hres + F * NB * 8, F * NB * 8);
    ctx->stats[5] = 0;
    if (ctx->timing && N) {  // time of the histogram kernel proper (mb200_ctx_stats, out[5])
        float ms = 0.f;
        MB_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->stats[5] = (int64_t)(ms * 1.0e6);
    }
    return MB200_OK;
}

extern "C" int mb200_ctx_set_hist_mode(mb200_ctx *ctx, int mode) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context");
    if (mode < 0 || mode > 2)
        return fail(MB200_EINVAL, "histogram mode must be 0 (chunked atomics), 1 (three returning atomics) or 2 (counting sort)");
    ctx->hist_mode = mode;
    return MB200_OK;
}
