// Per-frame compound prologue of libmaicos_b200 (sm_100a): centres and dipole-derived weights of
// residues / molecules straight from the float32 atom coordinates of a batch of frames.
//
// Replaces, for compounds that are contiguous runs of atoms, the host calls the reference makes every
// frame before the two hot kernels:
//   atoms.unwrap(compound) / atoms.wrap(compound)   core/base.py:438-448   (MDAnalysis)
//   get_center(ag, bin_method, compound)            lib/util.py:642-680    -> ag.center(weights, compound)
//   diporder_weights(ag, compound, order_parameter) lib/weights.py:131-178 -> ag.dipole_vector(compound)
// The arithmetic restates maicos_b200/mdshim/universe.py operation by operation (float32 minimum
// image about the first atom of a compound, float64 sequential sums in atom order, centre-of-mass
// wrap), so the device prologue and the host prologue agree bit for bit; one thread owns one compound.
#include "common.cuh"
#include <chrono>
#include <cstdlib>

namespace mb200 {

struct CompoundParams {
    const float *pos;        // [F][n_atoms][3]
    const float *boxes;      // [F][3]
    const long long *ptr;    // [n_c + 1]
    const double *cw;        // [n_atoms] centre weights (masses, |charges|) or null (geometry)
    const double *mass;      // [n_atoms] masses for the wrap / dipole reference point, or null
    const double *charge;    // [n_atoms] or null
    double *centers;         // [F][n_c][3]
    double *weights;         // [F][n_c] or [F][3][n_c] or null
    long long n_atoms, n_c;
    long long fstride;       // floats between two frames of `pos` (n_atoms * 3, more for a window of whole frames)
    int make_whole, wrap_center, weight_mode, pdim;
    // unit vector the dipole is projected on: 0 Cartesian e_pdim (lib/util.py:684-706), 1 radial direction of a cylinder
    // with axis uv_dim about the box centre (util.py:710-756), 2 radial direction of a sphere about the box centre (760-786)
    int uv_mode, uv_dim;
};

constexpr int MAX_COMPOUND_ATOMS = 64;  // larger compounds take the host path

__global__ void __launch_bounds__(128) k_compound(CompoundParams p) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long f = blockIdx.y;
    if (c >= p.n_c) return;
    const long long a0 = p.ptr[c], a1 = p.ptr[c + 1];
    const int m = (int)(a1 - a0);
    const float *q = p.pos + (size_t)f * p.fstride + (size_t)a0 * 3;
    const float bx = p.boxes[f * 3], by = p.boxes[f * 3 + 1], bz = p.boxes[f * 3 + 2];

    // pass 1: make whole (float32, about the first atom) and the mass-weighted centre for wrapping
    float ax = q[0], ay = q[1], az = q[2];
    auto whole = [&](int i, float &x, float &y, float &z) {
        x = q[i * 3]; y = q[i * 3 + 1]; z = q[i * 3 + 2];
        if (p.make_whole) {
            float dx = __fsub_rn(x, ax), dy = __fsub_rn(y, ay), dz = __fsub_rn(z, az);
            dx = __fsub_rn(dx, __fmul_rn(bx, rintf(__fdiv_rn(dx, bx))));
            dy = __fsub_rn(dy, __fmul_rn(by, rintf(__fdiv_rn(dy, by))));
            dz = __fsub_rn(dz, __fmul_rn(bz, rintf(__fdiv_rn(dz, bz))));
            x = __fadd_rn(ax, dx); y = __fadd_rn(ay, dy); z = __fadd_rn(az, dz);
        }
    };
    float sx = 0.f, sy = 0.f, sz = 0.f;  // float32 shift of the whole compound into the cell
    if (p.wrap_center) {
        double nx = 0, ny = 0, nz = 0, ws = 0;
        for (int i = 0; i < m; ++i) {
            float x, y, z;
            whole(i, x, y, z);
            const double w = p.mass ? p.mass[a0 + i] : 1.0;
            nx = __dadd_rn(nx, __dmul_rn(w, (double)x));
            ny = __dadd_rn(ny, __dmul_rn(w, (double)y));
            nz = __dadd_rn(nz, __dmul_rn(w, (double)z));
            ws = __dadd_rn(ws, w);
        }
        const double cx = __ddiv_rn(nx, ws), cy = __ddiv_rn(ny, ws), cz = __ddiv_rn(nz, ws);
        sx = (float)__dmul_rn(-floor(__ddiv_rn(cx, (double)bx)), (double)bx);
        sy = (float)__dmul_rn(-floor(__ddiv_rn(cy, (double)by)), (double)by);
        sz = (float)__dmul_rn(-floor(__ddiv_rn(cz, (double)bz)), (double)bz);
    }
    // pass 2: requested centre (weights cw) and the centre of mass (dipole reference), float64 sums in atom order
    double nx = 0, ny = 0, nz = 0, ws = 0, mx = 0, my = 0, mz = 0, ms = 0;
    for (int i = 0; i < m; ++i) {
        float x, y, z;
        whole(i, x, y, z);
        x = __fadd_rn(x, sx); y = __fadd_rn(y, sy); z = __fadd_rn(z, sz);
        const double w = p.cw ? p.cw[a0 + i] : 1.0;
        nx = __dadd_rn(nx, __dmul_rn(w, (double)x));
        ny = __dadd_rn(ny, __dmul_rn(w, (double)y));
        nz = __dadd_rn(nz, __dmul_rn(w, (double)z));
        ws = __dadd_rn(ws, w);
        if (p.weight_mode) {
            const double mw = p.mass[a0 + i];
            mx = __dadd_rn(mx, __dmul_rn(mw, (double)x));
            my = __dadd_rn(my, __dmul_rn(mw, (double)y));
            mz = __dadd_rn(mz, __dmul_rn(mw, (double)z));
            ms = __dadd_rn(ms, mw);
        }
    }
    double *out = p.centers + ((size_t)f * p.n_c + c) * 3;
    out[0] = __ddiv_rn(nx, ws); out[1] = __ddiv_rn(ny, ws); out[2] = __ddiv_rn(nz, ws);
    if (!p.weight_mode) return;

    // dipole about the centre of mass: sum_i q_i (r_i - R_com)
    const double rx = __ddiv_rn(mx, ms), ry = __ddiv_rn(my, ms), rz = __ddiv_rn(mz, ms);
    double dx = 0, dy = 0, dz = 0;
    for (int i = 0; i < m; ++i) {
        float x, y, z;
        whole(i, x, y, z);
        x = __fadd_rn(x, sx); y = __fadd_rn(y, sy); z = __fadd_rn(z, sz);
        const double qi = p.charge[a0 + i];
        dx = __dadd_rn(dx, __dmul_rn(qi, __dsub_rn((double)x, rx)));
        dy = __dadd_rn(dy, __dmul_rn(qi, __dsub_rn((double)y, ry)));
        dz = __dadd_rn(dz, __dmul_rn(qi, __dsub_rn((double)z, rz)));
    }
    const double mu[3] = {dx, dy, dz};
    if (p.uv_mode != 0 && p.weight_mode != 4) {
        // e = (centre - box / 2), axial component dropped for the cylinder, normalised (numpy order: ((x^2 + y^2) + z^2));
        // the centre is the binned one (bin_method), the box centre is dimensions[:3] / 2 in float32 (util.py:745-755, 779-785)
        double e[3] = {__dsub_rn(out[0], (double)(bx * 0.5f)), __dsub_rn(out[1], (double)(by * 0.5f)),
                       __dsub_rn(out[2], (double)(bz * 0.5f))};
        if (p.uv_mode == 1) e[p.uv_dim] = 0.0;
        const double en = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(e[0], e[0]), __dmul_rn(e[1], e[1])), __dmul_rn(e[2], e[2])));
        for (int d = 0; d < 3; ++d) e[d] = __ddiv_rn(e[d], en);
        double m3[3] = {mu[0], mu[1], mu[2]};
        if (p.weight_mode != 1) {  // cos_theta: (mu / |mu|) . e  (weights.py:163-167)
            const double norm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            for (int d = 0; d < 3; ++d) m3[d] = __ddiv_rn(m3[d], norm);
        }
        const double w = __dadd_rn(__dadd_rn(__dmul_rn(m3[0], e[0]), __dmul_rn(m3[1], e[1])), __dmul_rn(m3[2], e[2]));
        p.weights[(size_t)f * p.n_c + c] = p.weight_mode == 3 ? __dmul_rn(w, w) : w;
        return;
    }
    if (p.weight_mode == 1) {  // P0: mu . e
        p.weights[(size_t)f * p.n_c + c] = mu[p.pdim];
        return;
    }
    const double norm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    if (p.weight_mode == 4) {  // the three components of mu / |mu|: [F][3][n_c]
        for (int d = 0; d < 3; ++d) p.weights[((size_t)f * 3 + d) * p.n_c + c] = __ddiv_rn(mu[d], norm);
        return;
    }
    const double cs = __ddiv_rn(mu[p.pdim], norm);
    p.weights[(size_t)f * p.n_c + c] = p.weight_mode == 3 ? __dmul_rn(cs, cs) : cs;
}

// Velocity-derived weights of the Velocity* / Temperature* profile modules from the raw float32 velocities of a batch of
// frames (lib/weights.py:101-127, 195-225); one thread per compound:
//   mode 0  v[:, vdim] per atom (grouping "atoms": compounds are single atoms)
//   mode 1  sum_i m_i v_i[vdim] / sum_i m_i per compound, float64 sums in atom order
//   mode 2  kinetic temperature per atom, ((vx^2 + vy^2) + vz^2 in float32) * m / 2 * prefac  (weights.py:125-127)
struct VelocityParams {
    const float *vel;       // [F][n_atoms][3]
    const long long *ptr;   // [n_c + 1] or null (every atom its own compound)
    const double *mass;     // [n_atoms] or null (mode 0)
    double *weights;        // [F][n_c]
    long long n_atoms, n_c;
    double prefac;
    int mode, vdim;
};

__global__ void __launch_bounds__(256) k_velocity_weights(VelocityParams p) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long f = blockIdx.y;
    if (c >= p.n_c) return;
    const long long a0 = p.ptr ? p.ptr[c] : c, a1 = p.ptr ? p.ptr[c + 1] : c + 1;
    const float *v = p.vel + ((size_t)f * p.n_atoms) * 3;
    double w;
    if (p.mode == 0) {
        w = (double)v[a0 * 3 + p.vdim];
    } else if (p.mode == 1) {
        double num = 0.0, den = 0.0;
        for (long long a = a0; a < a1; ++a) {
            const double m = p.mass[a];
            num = __dadd_rn(num, __dmul_rn((double)v[a * 3 + p.vdim], m));
            den = __dadd_rn(den, m);
        }
        w = __ddiv_rn(num, den);
    } else {
        const float vx = v[a0 * 3], vy = v[a0 * 3 + 1], vz = v[a0 * 3 + 2];
        const float s = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
        w = __dmul_rn(__ddiv_rn(__dmul_rn((double)s, p.mass[a0]), 2.0), p.prefac);
    }
    p.weights[(size_t)f * p.n_c + c] = w;
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb200_dev_alloc(mb200_ctx *ctx, size_t bytes, void **out) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context");
    MB_LOCK(ctx);
    if (!out) return fail(MB200_EINVAL, "null output pointer");
    MB_CUDA(cudaSetDevice(ctx->device));
    *out = nullptr;
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        *out = nullptr;
        return fail(MB200_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    return MB200_OK;
}

extern "C" int mb200_dev_free(mb200_ctx *ctx, void *p) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context");
    MB_LOCK(ctx);
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (p) MB_CUDA(cudaFree(p));
    return MB200_OK;
}

extern "C" int mb200_dev_read(mb200_ctx *ctx, const void *dev, void *host, size_t bytes) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context");
    MB_LOCK(ctx);
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MB200_OK;
}

extern "C" int mb200_dev_write(mb200_ctx *ctx, void *dev, const void *host, size_t bytes) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context");
    MB_LOCK(ctx);
    if (bytes && (!dev || !host)) return fail(MB200_EINVAL, "null pointer in mb200_dev_write");
    MB_CUDA(cudaSetDevice(ctx->device));
    MB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));  // `host` may be reused as soon as the call returns
    return MB200_OK;
}

// Topology of the compounds, uploaded once per analysis (compound_ptr | centre weights | masses | charges): the per-call
// variant below used to re-send these 8 + 24 bytes per atom with every batch of frames.
struct mb200_topology {
    mb200::DevBuf buf;
    const long long *ptr = nullptr;
    const double *cw = nullptr, *mass = nullptr, *charge = nullptr;
    int64_t n_atoms = 0, n_c = 0;
    int device = 0;
};

extern "C" int mb200_topology_create(mb200_ctx *ctx, const int64_t *compound_ptr, int64_t n_c, int64_t n_atoms,
                                     const double *center_weights, const double *masses, const double *charges,
                                     mb200_topology **out) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (!out) return fail(MB200_EINVAL, "null output pointer");
    *out = nullptr;
    if (n_atoms < 0 || n_c <= 0) return fail(MB200_EINVAL, "invalid argument to mb200_topology_create");
    if (!compound_ptr && n_c != n_atoms) return fail(MB200_EINVAL, "without compound_ptr every atom is its own compound");
    if (compound_ptr) {
        if (compound_ptr[0] != 0 || compound_ptr[n_c] != n_atoms) return fail(MB200_EINVAL, "compound_ptr must cover all atoms");
        for (int64_t c = 0; c < n_c; ++c) {
            const int64_t m = compound_ptr[c + 1] - compound_ptr[c];
            if (m <= 0 || m > MAX_COMPOUND_ATOMS)
                return fail(MB200_EINVAL, "compound %lld has %lld atoms (supported: 1..%d)", (long long)c, (long long)m,
                            MAX_COMPOUND_ATOMS);
        }
    }
    MB_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<mb200_topology> t(new mb200_topology());
    const size_t NC = (size_t)n_c, N = (size_t)n_atoms;
    const size_t b_ptr = compound_ptr ? (NC + 1) * 8 : 0, b_at = N * 8;
    MB_TRY(t->buf.ensure(b_ptr + 3 * b_at + 64));
    char *base = t->buf.as<char>();
    if (compound_ptr) {
        MB_CUDA(cudaMemcpyAsync(base, compound_ptr, b_ptr, cudaMemcpyHostToDevice, ctx->stream));
        t->ptr = reinterpret_cast<const long long *>(base);
    }
    size_t off = b_ptr;
    const double *src[3] = {center_weights, masses, charges};
    const double **dst[3] = {&t->cw, &t->mass, &t->charge};
    for (int i = 0; i < 3; ++i, off += b_at)
        if (src[i] && b_at) {
            MB_CUDA(cudaMemcpyAsync(base + off, src[i], b_at, cudaMemcpyHostToDevice, ctx->stream));
            *dst[i] = reinterpret_cast<const double *>(base + off);
        }
    MB_CUDA(cudaStreamSynchronize(ctx->stream));  // the host arrays may be reused by the caller
    t->n_atoms = n_atoms; t->n_c = n_c; t->device = ctx->device;
    *out = t.release();
    return MB200_OK;
}

extern "C" void mb200_topology_destroy(mb200_ctx *ctx, mb200_topology *topo) {
    if (!topo) return;
    if (ctx) {
        MB_LOCK(ctx);
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        delete topo;
    } else {
        delete topo;
    }
}

extern "C" int mb200_compound_prologue_topo(mb200_ctx *ctx, const float *positions, int positions_on_device,
                                            int64_t n_frames, const float *boxes, const mb200_topology *topo,
                                            int make_whole, int wrap_center, int weight_mode, int pdim,
                                            int uv_mode, int uv_dim, double *centers_dev, double *weights_dev) {
    return mb200_compound_prologue_window(ctx, positions, positions_on_device, n_frames, topo ? topo->n_atoms : 0, 0, boxes, topo,
                                          make_whole, wrap_center, weight_mode, pdim, uv_mode, uv_dim, centers_dev, weights_dev);
}

extern "C" int mb200_compound_prologue_window(mb200_ctx *ctx, const float *positions, int positions_on_device,
                                              int64_t n_frames, int64_t n_total, int64_t first_atom, const float *boxes,
                                              const mb200_topology *topo, int make_whole, int wrap_center, int weight_mode,
                                              int pdim, int uv_mode, int uv_dim, double *centers_dev, double *weights_dev) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (topo && (first_atom < 0 || first_atom + topo->n_atoms > n_total))
        return fail(MB200_EINVAL, "the window [first_atom, first_atom + n_atoms) does not fit the frames");
    if (!topo || n_frames < 0 || !boxes || !centers_dev || (topo->n_atoms && !positions) || weight_mode < 0 || weight_mode > 4 ||
        pdim < 0 || pdim > 2 || uv_mode < 0 || uv_mode > 2 || uv_dim < 0 || uv_dim > 2 || !topo->ptr)
        return fail(MB200_EINVAL, "invalid argument to mb200_compound_prologue");
    if (topo->device != ctx->device) return fail(MB200_EINVAL, "topology belongs to another device");
    if (weight_mode && (!topo->mass || !topo->charge || !weights_dev))
        return fail(MB200_EINVAL, "dipole weights need masses, charges and an output buffer");
    if (n_frames == 0) return MB200_OK;
    if (n_frames > 65535) return fail(MB200_EINVAL, "at most 65535 frames per call");
    MB_CUDA(cudaSetDevice(ctx->device));
    const size_t F = (size_t)n_frames, N = (size_t)topo->n_atoms, NC = (size_t)topo->n_c;
    // developer aid: MAICOS_B200_TRACE_HOST=1 prints the host-side time of every phase of the call to stderr
    static const bool trace_host = std::getenv("MAICOS_B200_TRACE_HOST") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (!trace_host) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[mb200_compound_prologue] %-22s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    mark("entry");
    CompoundParams p;
    std::memset(&p, 0, sizeof(p));
    PrefetchRead pref_read(ctx);
    // the topology's atoms are [first_atom, first_atom + N) of whole frames of n_total atoms
    const size_t NT = (size_t)n_total;
    p.fstride = (long long)(NT * 3);
    if (positions_on_device) {
        p.pos = positions + (size_t)first_atom * 3;
    } else if (const void *pref = pref_read.take(positions, F * NT * 12)) {
        p.pos = reinterpret_cast<const float *>(pref) + (size_t)first_atom * 3;  // copied on the copy stream under the previous call
    } else {
        MB_TRY(ctx->d_hpos.ensure(F * N * 12));
        if (NT == N) {
            MB_CUDA(cudaMemcpyAsync(ctx->d_hpos.p, positions, F * N * 12, cudaMemcpyHostToDevice, ctx->stream));
        } else {  // only the window travels
            MB_CUDA(cudaMemcpy2DAsync(ctx->d_hpos.p, N * 12, positions + (size_t)first_atom * 3, NT * 12, N * 12, F,
                                      cudaMemcpyHostToDevice, ctx->stream));
        }
        p.pos = ctx->d_hpos.as<float>();
        p.fstride = (long long)(N * 3);
    }
    MB_TRY(ctx->d_ctopo.ensure(F * 12 + 64));
    MB_CUDA(cudaMemcpyAsync(ctx->d_ctopo.p, boxes, F * 12, cudaMemcpyHostToDevice, ctx->stream));
    p.boxes = ctx->d_ctopo.as<float>();
    p.ptr = topo->ptr; p.cw = topo->cw; p.mass = topo->mass; p.charge = topo->charge;
    p.centers = centers_dev; p.weights = weights_dev;
    p.n_atoms = topo->n_atoms; p.n_c = topo->n_c;
    p.make_whole = make_whole; p.wrap_center = wrap_center; p.weight_mode = weight_mode; p.pdim = pdim;
    p.uv_mode = uv_mode; p.uv_dim = uv_dim;
    dim3 grid((unsigned)((NC + 127) / 128), (unsigned)F);
    k_compound<<<grid, 128, 0, ctx->stream>>>(p);
    MB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    mark(pref_read.slot >= 0 ? "queued (prefetched)" : "queued (own copy)");
    pref_read.release();
    MB_CUDA(cudaStreamSynchronize(ctx->stream));  // `positions` / `boxes` may be reused by the caller
    mark("device (wait)");
    return MB200_OK;
}

extern "C" int mb200_velocity_weights(mb200_ctx *ctx, const float *velocities, int velocities_on_device, int64_t n_frames,
                                      const mb200_topology *topo, int mode, int vdim, double *weights_dev) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (!topo || n_frames < 0 || !weights_dev || (topo->n_atoms && !velocities) || mode < 0 || mode > 2 || vdim < 0 || vdim > 2)
        return fail(MB200_EINVAL, "invalid argument to mb200_velocity_weights");
    if (topo->device != ctx->device) return fail(MB200_EINVAL, "topology belongs to another device");
    if (mode != 0 && !topo->mass) return fail(MB200_EINVAL, "this mode needs masses in the topology");
    if (mode != 1 && topo->n_c != topo->n_atoms) return fail(MB200_EINVAL, "per-atom modes need one compound per atom");
    if (n_frames == 0 || topo->n_c == 0) return MB200_OK;
    if (n_frames > 65535) return fail(MB200_EINVAL, "at most 65535 frames per call");
    MB_CUDA(cudaSetDevice(ctx->device));
    const size_t F = (size_t)n_frames, N = (size_t)topo->n_atoms;
    VelocityParams p;
    std::memset(&p, 0, sizeof(p));
    if (velocities_on_device) {
        p.vel = velocities;
    } else {
        MB_TRY(ctx->d_hw2.ensure(F * N * 12));
        MB_CUDA(cudaMemcpyAsync(ctx->d_hw2.p, velocities, F * N * 12, cudaMemcpyHostToDevice, ctx->stream));
        p.vel = ctx->d_hw2.as<float>();
    }
    p.ptr = topo->n_c == topo->n_atoms ? nullptr : topo->ptr;
    p.mass = topo->mass; p.weights = weights_dev;
    p.n_atoms = topo->n_atoms; p.n_c = topo->n_c;
    // ((1 u * A^2) / ps^2) / k_B with scipy.constants' CODATA values (weights.py:125-126)
    p.prefac = 1.66053906660e-27 * 1e4 / 1.380649e-23;
    p.mode = mode; p.vdim = vdim;
    dim3 grid((unsigned)((topo->n_c + 255) / 256), (unsigned)F);
    k_velocity_weights<<<grid, 256, 0, ctx->stream>>>(p);
    MB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    MB_CUDA(cudaStreamSynchronize(ctx->stream));  // `velocities` may be reused by the caller
    return MB200_OK;
}

extern "C" int mb200_compound_prologue(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                                       int64_t n_atoms, const float *boxes, const int64_t *compound_ptr, int64_t n_c,
                                       const double *center_weights, const double *masses, const double *charges,
                                       int make_whole, int wrap_center, int weight_mode, int pdim,
                                       double *centers_dev, double *weights_dev) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (n_frames < 0 || n_atoms < 0 || n_c < 0 || !boxes || !compound_ptr || !centers_dev || (n_atoms && !positions) ||
        weight_mode < 0 || weight_mode > 4 || pdim < 0 || pdim > 2)
        return fail(MB200_EINVAL, "invalid argument to mb200_compound_prologue");
    if (weight_mode && (!masses || !charges || !weights_dev))
        return fail(MB200_EINVAL, "dipole weights need masses, charges and an output buffer");
    if (n_frames == 0 || n_c == 0) return MB200_OK;
    mb200_topology *topo = nullptr;  // one-shot form: the topology travels with the call
    MB_TRY(mb200_topology_create(ctx, compound_ptr, n_c, n_atoms, center_weights, masses, charges, &topo));
    const int rc = mb200_compound_prologue_topo(ctx, positions, positions_on_device, n_frames, boxes, topo, make_whole,
                                                wrap_center, weight_mode, pdim, 0, 0, centers_dev, weights_dev);
    mb200_topology_destroy(ctx, topo);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// Dielectric profile modules: per-atom quantities of the "virtual cut" estimator.
//
// Reference, per frame (modules/dielectricplanar.py:180-213, modules/dielectriccylinder.py:163-190):
//   testpos = centre of |charge| of the atom's compound along the profile coordinate (planar: x[dim];
//             cylinder: the radial coordinate r itself is averaged, sum(r |q|) / sum(|q|)),
//   cbin    = np.digitize(x[cut direction], arange(n_cut)[1:] * (L / n_cut)),
//   curQ    = bincount(zbin(testpos) + n_bins * cbin, charges).reshape(n_cut, n_bins),
//   m       = -cumsum(curQ / A, axis=0).mean(axis=0).
// Since  mean_c cumsum_c curQ[c, z] = (1 / n_cut) sum_c (n_cut - c) curQ[c, z],  the two-dimensional table is a
// one-dimensional weighted histogram over testpos with weights  q_i (n_cut - cbin_i):  this kernel writes
// testpos and those weights per atom, mb200_profile_hist(MB200_GEOM_SCALAR | MB200_BIN_DIGITIZE) bins them
// (exact fixed-point accumulation), the host applies -1 / (A n_cut).  One thread owns one compound (contiguous
// atoms, float64 sequential sums in atom order like AtomGroup.center / accumulate of the shim).
// ---------------------------------------------------------------------------------------------------------
namespace mb200 {

struct DielParams {
    const float *pos;        // [F][n_atoms][3]
    const long long *ptr;    // [n_c + 1]
    const double *charge;    // [n_atoms]
    const double *center;    // [F][3] (cylinder) or null
    const int *cut_bins;     // [F][n_cut]
    const double *cut_step;  // [F][n_cut]
    double *coord;           // [F][n_atoms]
    double *weights;         // [F][n_cut][n_atoms]
    long long n_atoms, n_c;
    int geometry, dim, n_cut, cut_dims[2];
};

__global__ void __launch_bounds__(128) k_dielectric_prepare(DielParams p) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long f = blockIdx.y;
    if (c >= p.n_c) return;
    const long long a0 = p.ptr[c], a1 = p.ptr[c + 1];
    const float *q = p.pos + ((size_t)f * p.n_atoms + a0) * 3;
    const int d1 = (p.dim + 1) % 3, d2 = (p.dim + 2) % 3;  // odims = roll(arange(3), -dim)[1:]
    double cx1 = 0.0, cx2 = 0.0;
    if (p.geometry == 1) { cx1 = p.center[f * 3 + d1]; cx2 = p.center[f * 3 + d2]; }
    auto profile_coord = [&](long long i) -> double {
        if (p.geometry == 0) return (double)q[i * 3 + p.dim];
        const double a = __dsub_rn((double)q[i * 3 + d1], cx1), b = __dsub_rn((double)q[i * 3 + d2], cx2);
        return sqrt(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));  // transform_cylinder: norm of the odims components
    };
    double num = 0.0, den = 0.0;
    for (long long i = 0; i < a1 - a0; ++i) {
        const double w = fabs(p.charge[a0 + i]);
        num = __dadd_rn(num, __dmul_rn(w, profile_coord(i)));
        den = __dadd_rn(den, w);
    }
    const double testpos = __ddiv_rn(num, den);
    for (long long i = 0; i < a1 - a0; ++i) {
        p.coord[(size_t)f * p.n_atoms + a0 + i] = testpos;
        const double qi = p.charge[a0 + i];
        for (int j = 0; j < p.n_cut; ++j) {
            const int nb = p.cut_bins[f * p.n_cut + j];
            const double step = p.cut_step[f * p.n_cut + j];
            const double x = (double)q[i * 3 + p.cut_dims[j]];
            // np.digitize(x, arange(nb)[1:] * step): number of edges k * step, k = 1 .. nb - 1, that are <= x
            int b;
            if (x != x) {
                b = nb - 1;
            } else {
                const double g = x / step;
                b = g <= 0.0 ? 0 : (g >= (double)(nb - 1) ? nb - 1 : (int)g);
                while (b > 0 && x < __dmul_rn((double)b, step)) --b;
                while (b < nb - 1 && x >= __dmul_rn((double)(b + 1), step)) ++b;
            }
            p.weights[((size_t)f * p.n_cut + j) * p.n_atoms + a0 + i] = __dmul_rn(qi, (double)(nb - b));
        }
    }
}

}  // namespace mb200

extern "C" int mb200_dielectric_prepare(mb200_ctx *ctx, const float *positions, int positions_on_device, int64_t n_frames,
                                        int64_t n_atoms, const int64_t *compound_ptr, int64_t n_c, const double *charges,
                                        int geometry, int dim, const double *center, int n_cut, const int32_t *cut_dims,
                                        const int32_t *cut_bins, const double *cut_step, double *coord_dev,
                                        double *weights_dev) {
    using namespace mb200;
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context (there is no CPU fallback)");
    MB_LOCK(ctx);
    if (n_frames < 0 || n_atoms < 0 || n_c < 0 || !compound_ptr || !charges || !coord_dev || (n_atoms && !positions) ||
        geometry < 0 || geometry > 1 || dim < 0 || dim > 2 || n_cut < 0 || n_cut > 2 ||
        (n_cut && (!cut_dims || !cut_bins || !cut_step || !weights_dev)) || (geometry == 1 && !center))
        return fail(MB200_EINVAL, "invalid argument to mb200_dielectric_prepare");
    if (n_frames == 0 || n_c == 0) return MB200_OK;
    if (n_frames > 65535) return fail(MB200_EINVAL, "at most 65535 frames per call");
    if (compound_ptr[0] != 0 || compound_ptr[n_c] != n_atoms) return fail(MB200_EINVAL, "compound_ptr must cover all atoms");
    for (int64_t c = 0; c < n_c; ++c)
        if (compound_ptr[c + 1] <= compound_ptr[c]) return fail(MB200_EINVAL, "empty compound %lld", (long long)c);
    for (int j = 0; j < n_cut; ++j)
        if (cut_dims[j] < 0 || cut_dims[j] > 2) return fail(MB200_EINVAL, "cut direction out of range");
    for (int64_t i = 0; i < n_frames * n_cut; ++i)
        if (cut_bins[i] < 1 || !(cut_step[i] > 0.0)) return fail(MB200_EINVAL, "virtual cuts need n >= 1 and a positive spacing");
    MB_CUDA(cudaSetDevice(ctx->device));
    const size_t F = (size_t)n_frames, N = (size_t)n_atoms, NC = (size_t)n_c;
    DielParams p;
    std::memset(&p, 0, sizeof(p));
    if (positions_on_device) {
        p.pos = positions;
    } else {
        MB_TRY(ctx->d_pos.ensure(F * N * 12));
        MB_CUDA(cudaMemcpyAsync(ctx->d_pos.p, positions, F * N * 12, cudaMemcpyHostToDevice, ctx->stream));
        p.pos = ctx->d_pos.as<float>();
    }
    // ptr | charge | center | cut_bins | cut_step
    const size_t b_ptr = (NC + 1) * 8, b_q = N * 8, b_c = F * 24, b_nb = ((F * 2 * 4 + 7) / 8) * 8, b_st = F * 2 * 8;
    MB_TRY(ctx->d_ctopo.ensure(b_ptr + b_q + b_c + b_nb + b_st + 64));
    char *base = ctx->d_ctopo.as<char>();
    MB_CUDA(cudaMemcpyAsync(base, compound_ptr, b_ptr, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaMemcpyAsync(base + b_ptr, charges, b_q, cudaMemcpyHostToDevice, ctx->stream));
    if (center) MB_CUDA(cudaMemcpyAsync(base + b_ptr + b_q, center, b_c, cudaMemcpyHostToDevice, ctx->stream));
    if (n_cut) {
        MB_CUDA(cudaMemcpyAsync(base + b_ptr + b_q + b_c, cut_bins, F * n_cut * 4, cudaMemcpyHostToDevice, ctx->stream));
        MB_CUDA(cudaMemcpyAsync(base + b_ptr + b_q + b_c + b_nb, cut_step, F * n_cut * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    p.ptr = reinterpret_cast<const long long *>(base);
    p.charge = reinterpret_cast<const double *>(base + b_ptr);
    p.center = center ? reinterpret_cast<const double *>(base + b_ptr + b_q) : nullptr;
    p.cut_bins = reinterpret_cast<const int *>(base + b_ptr + b_q + b_c);
    p.cut_step = reinterpret_cast<const double *>(base + b_ptr + b_q + b_c + b_nb);
    p.coord = coord_dev; p.weights = weights_dev;
    p.n_atoms = n_atoms; p.n_c = n_c; p.geometry = geometry; p.dim = dim; p.n_cut = n_cut;
    for (int j = 0; j < n_cut; ++j) p.cut_dims[j] = cut_dims[j];
    dim3 grid((unsigned)((NC + 127) / 128), (unsigned)F);
    k_dielectric_prepare<<<grid, 128, 0, ctx->stream>>>(p);
    MB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    MB_CUDA(cudaStreamSynchronize(ctx->stream));  // host arrays may be reused by the caller
    return MB200_OK;
}
