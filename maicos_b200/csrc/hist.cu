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
// HBM-bound: 12 B (float32 xyz) + 8 B (per-frame weight) per element.  Each CTA privatises the
// histogram in shared memory (32-bit counts: native ATOMS; float64 sums: ATOMS.CAS loop) and
// merges with native REDG.ADD.F64 / REDG.ADD.64 into one of NCOPY global copies; a second
// kernel sums the copies in fixed order.
#include "common.cuh"

namespace mb200 {

constexpr int H_THREADS = 256;
constexpr int H_NCOPY = 8;
constexpr int H_MAX_SMEM_BINS = 4096;  // 4096 * (8 + 4) B = 48 KB static-equivalent

struct HistParams {
    const void *pos;       // [F][n][3]
    const double *weights; // [n] or [F][n] or null
    const double *edges;   // [F][n_bins+1]
    const double *center;  // [F][3] or null
    const double *zrange;  // [F][2] or null
    double *acc_w;         // [F][NCOPY][n_bins]
    unsigned long long *acc_c;
    long long n;
    int n_bins, dim, geometry, is_f32, w_per_frame;
};

__device__ __forceinline__ int bin_of(double x, const double *__restrict__ edges, int n_bins) {
    const double first = edges[0], last = edges[n_bins];
    if (!(x >= first && x <= last)) return -1;
    // arithmetic guess (numpy: ((x - first) / (last - first)) * n), then exact correction
    int idx = (int)(__ddiv_rn(__dsub_rn(x, first), __dsub_rn(last, first)) * (double)n_bins);
    idx = min(max(idx, 0), n_bins - 1);
    while (idx > 0 && x < edges[idx]) --idx;
    while (idx < n_bins - 1 && x >= edges[idx + 1]) ++idx;
    return idx;
}

template <bool SMEM>
__global__ void __launch_bounds__(H_THREADS) k_hist(HistParams p) {
    extern __shared__ __align__(16) unsigned char hsm[];
    double *sh_w = reinterpret_cast<double *>(hsm);
    unsigned int *sh_c = reinterpret_cast<unsigned int *>(hsm + (size_t)p.n_bins * 8);
    const int f = blockIdx.y;
    const int nb = p.n_bins;
    if (SMEM) {
        for (int i = threadIdx.x; i < nb; i += H_THREADS) { sh_w[i] = 0.0; sh_c[i] = 0u; }
        __syncthreads();
    }
    const double *edges = p.edges + (size_t)f * (nb + 1);
    const double *w = p.weights ? p.weights + (p.w_per_frame ? (size_t)f * p.n : 0) : nullptr;
    double cx = 0, cy = 0, cz = 0, zlo = 0, zhi = 0;
    if (p.geometry != MB200_GEOM_PLANAR) { cx = p.center[f * 3]; cy = p.center[f * 3 + 1]; cz = p.center[f * 3 + 2]; }
    if (p.geometry == MB200_GEOM_CYLINDER) { zlo = p.zrange[f * 2]; zhi = p.zrange[f * 2 + 1]; }
    const int copy = blockIdx.x % H_NCOPY;
    double *gw = p.acc_w + ((size_t)f * H_NCOPY + copy) * nb;
    unsigned long long *gc = p.acc_c + ((size_t)f * H_NCOPY + copy) * nb;
    const int d0 = p.dim, d1 = (p.dim + 1) % 3, d2 = (p.dim + 2) % 3;  // odims = roll(arange(3), -dim)[1:]

    for (long long j = (long long)blockIdx.x * H_THREADS + threadIdx.x; j < p.n; j += (long long)gridDim.x * H_THREADS) {
        double x, y, z;
        if (p.is_f32) {
            const float *q = reinterpret_cast<const float *>(p.pos) + ((size_t)f * p.n + j) * 3;
            x = (double)q[0]; y = (double)q[1]; z = (double)q[2];
        } else {
            const double *q = reinterpret_cast<const double *>(p.pos) + ((size_t)f * p.n + j) * 3;
            x = q[0]; y = q[1]; z = q[2];
        }
        const double c3[3] = {x, y, z};
        double v;
        bool keep = true;
        if (p.geometry == MB200_GEOM_PLANAR) {
            v = c3[d0];
        } else if (p.geometry == MB200_GEOM_CYLINDER) {
            const double ctr[3] = {cx, cy, cz};
            const double a = __dsub_rn(c3[d1], ctr[d1]), b = __dsub_rn(c3[d2], ctr[d2]);
            v = sqrt(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
            const double ax = c3[d0];
            keep = (ax >= zlo && ax <= zhi);  // histogram2d drops outliers of either dimension
        } else {
            const double a = __dsub_rn(x, cx), b = __dsub_rn(y, cy), c = __dsub_rn(z, cz);
            v = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
        }
        const int idx = keep ? bin_of(v, edges, nb) : -1;
        if (idx >= 0) {
            if (SMEM) {
                atomicAdd(&sh_c[idx], 1u);
                if (w) atomicAdd(&sh_w[idx], w[j]);
            } else {
                atomicAdd(&gc[idx], 1ull);
                if (w) atomicAdd(&gw[idx], w[j]);
            }
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < nb; i += H_THREADS) {
            const unsigned int c = sh_c[i];
            if (c) {
                atomicAdd(&gc[i], (unsigned long long)c);
                if (w) atomicAdd(&gw[i], sh_w[i]);
            }
        }
    }
}

__global__ void k_hist_merge(const double *__restrict__ acc_w, const unsigned long long *__restrict__ acc_c,
                             double *__restrict__ out_w, long long *__restrict__ out_c, int n_bins, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long f = i / n_bins, b = i % n_bins;
    double s = 0.0;
    unsigned long long c = 0;
    for (int k = 0; k < H_NCOPY; ++k) {
        s += acc_w[(f * H_NCOPY + k) * n_bins + b];
        c += acc_c[(f * H_NCOPY + k) * n_bins + b];
    }
    out_w[i] = s;
    out_c[i] = (long long)c;
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb200_profile_hist(mb200_ctx *ctx, int geometry, const void *positions, int pos_is_f32,
                                  int positions_on_device, int64_t n_frames, int64_t n, int dim, const double *weights,
                                  int weights_per_frame, int weights_on_device, const double *edges, int32_t n_bins,
                                  const double *center, const double *zrange, double *hist_w, int64_t *hist_c) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    if (geometry < 0 || geometry > 2 || n_frames < 0 || n < 0 || n_bins <= 0 || dim < 0 || dim > 2 || !edges || !hist_c ||
        (n && !positions) || (weights && !hist_w))
        return fail(MB200_EINVAL, "invalid argument to mb200_profile_hist");
    if (geometry != MB200_GEOM_PLANAR && !center) return fail(MB200_EINVAL, "center required for radial geometries");
    if (geometry == MB200_GEOM_CYLINDER && !zrange) return fail(MB200_EINVAL, "zrange required for cylinder geometry");
    if (n_frames == 0) return MB200_OK;
    if (n_frames > 65535) return fail(MB200_EINVAL, "at most 65535 frames per call");
    MB_CUDA(cudaSetDevice(ctx->device));
    const size_t F = (size_t)n_frames, N = (size_t)n, NB = (size_t)n_bins;
    const size_t esz = pos_is_f32 ? 4 : 8;

    HistParams p;
    std::memset(&p, 0, sizeof(p));
    if (positions_on_device || N == 0) {
        p.pos = positions;
    } else {
        MB_TRY(ctx->d_hpos.ensure(F * N * 3 * esz));
        MB_CUDA(cudaMemcpyAsync(ctx->d_hpos.p, positions, F * N * 3 * esz, cudaMemcpyHostToDevice, ctx->stream));
        p.pos = ctx->d_hpos.p;
    }
    if (weights && N) {
        if (weights_on_device) {
            p.weights = weights;
        } else {
            const size_t wn = weights_per_frame ? F * N : N;
            MB_TRY(ctx->d_hw.ensure(wn * 8));
            MB_CUDA(cudaMemcpyAsync(ctx->d_hw.p, weights, wn * 8, cudaMemcpyHostToDevice, ctx->stream));
            p.weights = ctx->d_hw.as<double>();
        }
    }
    // small per-frame parameters: edges | center | zrange
    const size_t n_par = F * (NB + 1) + F * 3 + F * 2;
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
    // accumulators [F][NCOPY][NB] (w, c) + outputs [F][NB] (w, c)
    const size_t acc = F * H_NCOPY * NB;
    MB_TRY(ctx->d_hout.ensure(acc * 16 + F * NB * 16));
    p.acc_w = ctx->d_hout.as<double>();
    p.acc_c = reinterpret_cast<unsigned long long *>(p.acc_w + acc);
    double *d_ow = reinterpret_cast<double *>(p.acc_c + acc);
    long long *d_oc = reinterpret_cast<long long *>(d_ow + F * NB);
    MB_CUDA(cudaMemsetAsync(p.acc_w, 0, acc * 16, ctx->stream));
    p.n = n; p.n_bins = n_bins; p.dim = dim; p.geometry = geometry; p.is_f32 = pos_is_f32; p.w_per_frame = weights_per_frame;

    if (N) {
        // enough CTAs to fill the GPU across frames, each streaming >= 8 elements per thread
        long long per_frame = (long long)((N + (size_t)H_THREADS * 8 - 1) / ((size_t)H_THREADS * 8));
        long long want = ((long long)ctx->n_sm * 8 + (long long)F - 1) / (long long)F;
        unsigned bx = (unsigned)std::max<long long>(1, std::min(per_frame, std::max<long long>(want, 1)));
        dim3 grid(bx, (unsigned)F);
        if (n_bins <= H_MAX_SMEM_BINS) {
            k_hist<true><<<grid, H_THREADS, NB * 12, ctx->stream>>>(p);
        } else {
            k_hist<false><<<grid, H_THREADS, 0, ctx->stream>>>(p);
        }
        MB_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    const long long total = (long long)(F * NB);
    k_hist_merge<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(p.acc_w, p.acc_c, d_ow, d_oc, n_bins, total);
    MB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    if (hist_w) MB_CUDA(cudaMemcpyAsync(hist_w, d_ow, F * NB * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaMemcpyAsync(hist_c, d_oc, F * NB * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MB200_OK;
}
