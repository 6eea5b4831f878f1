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
            const float *base = reinterpret_cast<const float *>(p.pos) + (size_t)f * n;
            for (long long j = gtid; j < n; j += gstride) deposit(scalar_bin(g, (double)base[j], edges), j);
        } else {
            const double *base = reinterpret_cast<const double *>(p.pos) + (size_t)f * n;
            for (long long j = gtid; j < n; j += gstride) deposit(scalar_bin(g, base[j], edges), j);
        }
    } else if (p.is_f32) {
        const float *base = reinterpret_cast<const float *>(p.pos) + (size_t)f * n * 3;
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
        const double *base = reinterpret_cast<const double *>(p.pos) + (size_t)f * n * 3;
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
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    const int digitize = (geometry & MB200_BIN_DIGITIZE) ? 1 : 0;
    geometry &= ~MB200_BIN_DIGITIZE;
    if (geometry < 0 || geometry > 3 || n_frames < 0 || n < 0 || n_bins <= 0 || dim < 0 || dim > 2 || !edges || !hist_c ||
        (n && !positions) || (weights && !hist_w))
        return fail(MB200_EINVAL, "invalid argument to mb200_profile_hist");
    if ((geometry == MB200_GEOM_CYLINDER || geometry == MB200_GEOM_SPHERE) && !center)
        return fail(MB200_EINVAL, "center required for radial geometries");
    if (geometry == MB200_GEOM_CYLINDER && !zrange && !digitize) return fail(MB200_EINVAL, "zrange required for cylinder geometry");
    if (n_frames == 0) return MB200_OK;
    if (n_frames > 65535) return fail(MB200_EINVAL, "at most 65535 frames per call");
    MB_CUDA(cudaSetDevice(ctx->device));
    const size_t F = (size_t)n_frames, N = (size_t)n, NB = (size_t)n_bins;
    const size_t esz = pos_is_f32 ? 4 : 8;
    const size_t ncomp = geometry == MB200_GEOM_SCALAR ? 1 : 3;  // values per element in `positions`

    HistParams p;
    std::memset(&p, 0, sizeof(p));
    if (positions_on_device || N == 0) {
        p.pos = positions;
    } else {
        MB_TRY(ctx->d_hpos.ensure(F * N * ncomp * esz));
        MB_CUDA(cudaMemcpyAsync(ctx->d_hpos.p, positions, F * N * ncomp * esz, cudaMemcpyHostToDevice, ctx->stream));
        p.pos = ctx->d_hpos.p;
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
    // small per-frame parameters: edges | center | zrange | wmax bits
    const size_t n_par = F * (NB + 1) + F * 3 + F * 2 + F;
    MB_TRY(ctx->d_hpar.ensure(n_par * 8));
    double *dpar = ctx->d_hpar.as<double>();
    MB_CUDA(cudaMemcpyAsync(dpar, edges, F * (NB + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    p.edges = dpar;
    if (center) {
        MB_CUDA(cudaMemcpyAsync(dpar + F * (NB + 1), center, F * 24, cudaMemcpyHostToDevice, ctx->stream));
        p.center = dpar + F * (NB + 1);
    }
    if (zrange) {
        MB_CUDA(cudaMemcpyAsync(dpar + F * (NB + 1) + F * 3, zrange, F * 16, cudaMemcpyHostToDevice, ctx->stream));
        p.zrange = dpar + F * (NB + 1) + F * 3;
    }
    unsigned long long *d_wmax = reinterpret_cast<unsigned long long *>(dpar + F * (NB + 1) + F * 5);
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
        if (mode == 0) k_hist<0><<<grid, H_THREADS, smem, ctx->stream>>>(p);
        else if (mode == 1) k_hist<1><<<grid, H_THREADS, smem, ctx->stream>>>(p);
        else k_hist<2><<<grid, H_THREADS, 0, ctx->stream>>>(p);
        MB_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    k_hist_merge<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(p, mode == 0, hist_w ? d_ow : nullptr, d_oc, total);
    MB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    }
    if (hist_w) MB_CUDA(cudaMemcpyAsync(hist_w, d_ow, F * NB * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaMemcpyAsync(hist_c, d_oc, F * NB * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MB200_OK;
}
