// Context / library-level entry points of the C ABI (include/maicos_b200.h).
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <thread>

namespace mb200 {
thread_local std::string g_last_error;
}
using namespace mb200;

mb200_ctx::~mb200_ctx() {
    plans.clear();
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (int i = 0; i < 2; ++i) {
        if (tc_stage_ev[i]) cudaEventDestroy(tc_stage_ev[i]);
        if (tc_ev0[i]) cudaEventDestroy(tc_ev0[i]);
        if (tc_ev1[i]) cudaEventDestroy(tc_ev1[i]);
    }
    for (int i = 0; i < N_PENDING; ++i)
        if (pending[i].done) cudaEventDestroy(pending[i].done);
    for (int i = 0; i < N_PREF; ++i)
        if (pref_ev[i]) cudaEventDestroy(pref_ev[i]);
    for (int i = 0; i < N_PREF; ++i)
        if (pref_done_ev[i]) cudaEventDestroy(pref_done_ev[i]);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (stream) cudaStreamDestroy(stream);
}

extern "C" int mb200_abi_version(void) { return 2; }

extern "C" const char *mb200_last_error(void) { return g_last_error.c_str(); }

extern "C" int mb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int mb200_ctx_create(int device, mb200_ctx **out) {
    if (!out) return fail(MB200_EINVAL, "null output pointer");
    *out = nullptr;
    int n = mb200_device_count();
    if (n <= 0) return fail(MB200_ENODEVICE, "no CUDA device visible: maicos_b200 has no CPU fallback");
    if (device < 0 || device >= n) return fail(MB200_EINVAL, "device %d out of range (have %d)", device, n);
    MB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(MB200_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    mb200_ctx *ctx = new mb200_ctx();
    ctx->device = device;
    ctx->n_sm = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&ctx->tc_stage_ev[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->tc_ev0[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->tc_ev1[i]);
    }
    for (int i = 0; i < mb200_ctx::N_PENDING && e == cudaSuccess; ++i)
        e = cudaEventCreateWithFlags(&ctx->pending[i].done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < mb200_ctx::N_PREF && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->pref_ev[i], cudaEventDisableTiming);
    for (int i = 0; i < mb200_ctx::N_PREF && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->pref_done_ev[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete ctx;
        return fail(MB200_ECUDA, "context creation failed: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return MB200_OK;
}

extern "C" void mb200_ctx_destroy(mb200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    delete ctx;
}

extern "C" int mb200_ctx_synchronize(mb200_ctx *ctx) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context");
    MB_LOCK(ctx);
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
    return MB200_OK;
}

extern "C" uint64_t mb200_ctx_stream(mb200_ctx *ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }

extern "C" int mb200_ctx_stats(mb200_ctx *ctx, int64_t out[8]) {
    if (!ctx || !out) return fail(MB200_EINVAL, "null argument");
    for (int i = 0; i < 8; ++i) out[i] = ctx->stats[i];
    return MB200_OK;
}

extern "C" int mb200_ctx_set_timing(mb200_ctx *ctx, int enable) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context");
    ctx->timing = enable != 0;
    return MB200_OK;
}

extern "C" int64_t mb200_ctx_launch_count(mb200_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int mb200_prefetch_frames_shared(mb200_ctx *ctx, const void *host_frames, size_t bytes, uint64_t tag) {
    if (!ctx) return fail(MB200_ENODEVICE, "no CUDA context");
    if (!host_frames || bytes == 0) return MB200_OK;
    MB_CUDA(cudaSetDevice(ctx->device));
    std::lock_guard<std::mutex> lock(ctx->pref_mutex);
    // MAICOS_B200_PREF_SLOTS (developer aid): use fewer than N_PREF slots
    static const int NP = std::getenv("MAICOS_B200_PREF_SLOTS")
                              ? std::min(std::max(std::atoi(std::getenv("MAICOS_B200_PREF_SLOTS")), 1), (int)mb200_ctx::N_PREF)
                              : (int)mb200_ctx::N_PREF;
    if (tag != 0)  // the same frames announced again by another analysis of a collection: one more reader, no second copy
        for (int slot = 0; slot < NP; ++slot)
            if (ctx->pref_tag[slot] == tag && ctx->pref_host[slot] == host_frames && ctx->pref_bytes[slot] == bytes) {
                ctx->pref_readers[slot] += 1;
                ctx->pref_src[slot] = host_frames;
                ctx->pref_saved += 1;
                return MB200_OK;
            }
    // a slot no running call reads: one whose announcements were all taken if there is any, else the oldest content
    // (an announcement nobody took -- its call copies by itself)
    // (measured: a single analysis that rotates through all four buffers instead of alternating between the first two
    // has runs 1.5 - 3x slower now and then, tools/r2_c4_probe.py -- hence the LOWEST free slot, not the oldest)
    int slot = -1;
    for (int i = 0; i < NP && slot < 0; ++i)
        if (ctx->pref_busy[i] == 0 && ctx->pref_readers[i] <= 0) slot = i;
    for (int i = 0; i < NP; ++i) {
        if (slot >= 0 && ctx->pref_readers[slot] <= 0) break;
        if (ctx->pref_busy[i] == 0 && (slot < 0 || ctx->pref_seq[i] < ctx->pref_seq[slot])) slot = i;
    }
    if (slot < 0) {  // every slot is being read: the call copies its frames by itself
        ctx->pref_skipped += 1;
        return MB200_OK;
    }
    ctx->pref_src[slot] = ctx->pref_host[slot] = nullptr;
    ctx->pref_tag[slot] = 0;
    ctx->pref_readers[slot] = 0;
    if (ctx->pref_done_set[slot])  // the slot's previous content may still be read by a queued kernel
        MB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->pref_done_ev[slot], 0));
    MB_TRY(ctx->d_pref[slot].ensure(bytes));
    MB_CUDA(cudaMemcpyAsync(ctx->d_pref[slot].p, host_frames, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    MB_CUDA(cudaEventRecord(ctx->pref_ev[slot], ctx->copy_stream));
    ctx->pref_src[slot] = ctx->pref_host[slot] = host_frames;
    ctx->pref_bytes[slot] = bytes;
    ctx->pref_tag[slot] = tag;
    ctx->pref_readers[slot] = 1;
    ctx->pref_seq[slot] = ++ctx->pref_counter;
    ctx->pref_copies += 1;
    return MB200_OK;
}

extern "C" int mb200_prefetch_frames(mb200_ctx *ctx, const void *host_frames, size_t bytes) {
    return mb200_prefetch_frames_shared(ctx, host_frames, bytes, 0);
}

extern "C" int mb200_ctx_prefetch_stats(mb200_ctx *ctx, int64_t out[3]) {
    if (!ctx || !out) return fail(MB200_EINVAL, "invalid argument to mb200_ctx_prefetch_stats");
    std::lock_guard<std::mutex> lock(ctx->pref_mutex);
    out[0] = ctx->pref_copies;
    out[1] = ctx->pref_saved;
    out[2] = ctx->pref_skipped;
    return MB200_OK;
}

namespace mb200 {
const void *PrefetchRead::take(const void *host, size_t bytes) {
    if (!host || slot >= 0) return nullptr;
    std::lock_guard<std::mutex> lock(ctx->pref_mutex);
    for (int i = 0; i < mb200_ctx::N_PREF; ++i)
        if (ctx->pref_src[i] == host && ctx->pref_bytes[i] >= bytes) {
            if (cudaStreamWaitEvent(ctx->stream, ctx->pref_ev[i], 0) != cudaSuccess) return nullptr;
            if (--ctx->pref_readers[i] <= 0) ctx->pref_src[i] = nullptr;  // every announcement is taken once
            ctx->pref_busy[i] += 1;
            slot = i;
            return ctx->d_pref[i].p;
        }
    return nullptr;
}

void PrefetchRead::release() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(ctx->pref_mutex);
    if (cudaEventRecord(ctx->pref_done_ev[slot], ctx->stream) == cudaSuccess) ctx->pref_done_set[slot] = true;
    ctx->pref_busy[slot] -= 1;
    slot = -1;
}
}  // namespace mb200

extern "C" int mb200_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(MB200_EINVAL, "null output pointer");
    *out = nullptr;
    if (mb200_device_count() <= 0) return fail(MB200_ENODEVICE, "no CUDA device visible: cannot pin host memory");
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        *out = nullptr;
        return fail(MB200_ENOMEM, "cudaMallocHost(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    return MB200_OK;
}

extern "C" void mb200_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// Host-side staging helper: memcpy split over a few threads.  One core copies about 10 GB/s, a 12 MB frame of 1M atoms
// therefore costs more host time than its histogram costs device time; four threads bring the copy under the PCIe time
// of the frame.  Pure host code (no device needed); the caller's ctypes binding releases the GIL.
extern "C" int mb200_host_copy(void *dst, const void *src, size_t bytes, int n_threads) {
    if (bytes == 0) return MB200_OK;
    if (!dst || !src) return fail(MB200_EINVAL, "null pointer in mb200_host_copy");
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    int nt = std::max(1, std::min<int>(n_threads > 0 ? n_threads : 6, (int)hw));  // 6: best on the 16-core B200 hosts (tools/host_copy_bw.py)
    if (bytes < ((size_t)1 << 22)) nt = 1;
    if (nt == 1) {
        std::memcpy(dst, src, bytes);
        return MB200_OK;
    }
    const size_t chunk = ((bytes / nt) + 4095) & ~(size_t)4095;
    std::vector<std::thread> workers;
    workers.reserve(nt - 1);
    for (int i = 1; i < nt; ++i) {
        const size_t off = chunk * (size_t)i;
        if (off >= bytes) break;
        const size_t len = std::min(chunk, bytes - off);
        workers.emplace_back([=]() { std::memcpy((char *)dst + off, (const char *)src + off, len); });
    }
    std::memcpy(dst, src, std::min(chunk, bytes));
    for (auto &w : workers) w.join();
    return MB200_OK;
}

// Running mean / standard error / sum of per-frame observables for a whole batch of frames, host only.  Restates the
// recurrences of the reference frame loop (src/maicos/core/base.py:460-480 with lib/math.py:325-459: new_mean,
// new_variance) operation by operation, frame after frame, so that the result is bitwise what the per-frame Python
// code gives -- a 40-frame batch costs microseconds instead of 40 x 3 x ~10 small NumPy calls:
//   old_var = sem^2 (n - 1);  mean_n = ((n - 1) mean + x) / n;  s = old_var (n - 1) + (x - mean_old)(x - mean_n), s < 0 -> 0;
//   sem_n = sqrt(s / n / n);  sum += x.
// x: [n_frames][size] float64 (x_f64) or int64 (x_i64: means / sems see it as float64, the sum stays int64).
// n0 = frames already in the statistics (>= 1: the first frame initialises them on the caller's side); scalar_pow: the
// observable is a Python / NumPy float scalar (see below).
extern "C" int mb200_running_stats(int64_t n0, int64_t n_frames, int64_t size, int scalar_pow, const double *x_f64,
                                   const int64_t *x_i64, double *mean, double *sem, double *sum_f64, int64_t *sum_i64) {
    if (n0 < 1 || n_frames < 0 || size < 0 || (!x_f64 && !x_i64) || !mean || !sem || (x_f64 && !sum_f64) || (x_i64 && !sum_i64))
        return fail(MB200_EINVAL, "invalid argument to mb200_running_stats");
    for (int64_t f = 0; f < n_frames; ++f) {
        const double n = (double)(n0 + f + 1), nm1 = (double)(n0 + f);
        for (int64_t i = 0; i < size; ++i) {
            const double x = x_f64 ? x_f64[f * size + i] : (double)x_i64[f * size + i];
            const double old_mean = mean[i];
            // sem ** 2: NumPy squares arrays with a multiplication, Python / NumPy float scalars go through libm's pow
            const double old_var = (scalar_pow ? std::pow(sem[i], 2.0) : sem[i] * sem[i]) * nm1;
            const double t = nm1 * old_mean;
            const double new_mean = (t + x) / n;
            const double a = old_var * nm1, b = (x - old_mean) * (x - new_mean);
            double sv = a + b;
            if (sv < 0) sv = 0;
            mean[i] = new_mean;
            sem[i] = std::sqrt((sv / n) / n);
            if (x_f64) sum_f64[i] = sum_f64[i] + x;
            else sum_i64[i] = sum_i64[i] + x_i64[f * size + i];
        }
    }
    return MB200_OK;
}
