// Shared host-side plumbing of libmaicos_b200: error reporting, grow-only device /
// pinned buffers and the per-device context behind the C ABI (include/maicos_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <mutex>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/maicos_b200.h"

namespace mb200 {

extern thread_local std::string g_last_error;

inline int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define MB_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ::mb200::fail(_e == cudaErrorMemoryAllocation ? MB200_ENOMEM : MB200_ECUDA, \
                                 "CUDA error '%s' in %s (%s:%d)", cudaGetErrorString(_e),  \
                                 #expr, __FILE__, __LINE__);                               \
    } while (0)

#define MB_TRY(expr)                \
    do {                            \
        int _rc = (expr);           \
        if (_rc != MB200_OK) return _rc; \
    } while (0)

// Grow-only device buffer.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return MB200_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(MB200_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return MB200_OK;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};

// Grow-only pinned host buffer (staging for async H2D / D2H).
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return MB200_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(MB200_ENOMEM, "cudaMallocHost(%zu bytes) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return MB200_OK;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
    ~PinBuf() {
        if (p) cudaFreeHost(p);
    }
};

// fn(i0, i1) over [0, n) cut into contiguous chunks on a few host threads (q-plan construction for large reciprocal
// cubes: an NPT trajectory needs a new plan for every box).  Serial below `min_parallel` items.
template <typename F>
inline void parallel_chunks(size_t n, size_t min_parallel, F fn) {
    unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if (const char *env = std::getenv("MAICOS_B200_HOST_THREADS")) nt = (unsigned)std::max(1, std::atoi(env));
    if (n < min_parallel || nt < 2) {
        fn((size_t)0, n);
        return;
    }
    nt = (unsigned)std::min<size_t>(nt, n);
    const size_t per = (n + nt - 1) / nt;
    std::vector<std::thread> workers;
    workers.reserve(nt - 1);
    for (unsigned t = 1; t < nt; ++t) {
        const size_t i0 = std::min(n, per * t), i1 = std::min(n, per * (t + 1));
        if (i0 < i1) workers.emplace_back([=]() { fn(i0, i1); });
    }
    fn((size_t)0, std::min(n, per));
    for (auto &w : workers) w.join();
}

#define MB_LOCK(ctx) std::lock_guard<std::recursive_mutex> _mb_call_lock((ctx)->call_mutex)

struct SfPlan;    // sfactor.cu
struct PlanPool;  // sfactor.cu: retired q-plans whose host vectors and device buffers the next plan reuses
// The prefetch slot a compute call reads (context.cu).  take(): device copy of `host` if it was announced with
// mb200_prefetch_frames[_shared] (the compute stream is made to wait for the copy), else nullptr.  release(): every kernel
// that reads the slot has been queued -- records the slot's "done" event on the compute stream; until then the slot is
// busy and no announcement overwrites it.  The destructor releases on every exit path.
struct PrefetchRead {
    mb200_ctx *ctx;
    int slot = -1;
    explicit PrefetchRead(mb200_ctx *c) : ctx(c) {}
    PrefetchRead(const PrefetchRead &) = delete;
    PrefetchRead &operator=(const PrefetchRead &) = delete;
    const void *take(const void *host, size_t bytes);
    void release();
    ~PrefetchRead() { release(); }
};

}  // namespace mb200

struct mb200_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // split-precision contraction: descriptor staging in two alternating page-locked buffers (a call may be queued behind a
    // running one: mb200_saxs_frames_submit), each guarded by the event recorded behind the copies that read it
    mb200::PinBuf h_tc_stage[2];
    cudaEvent_t tc_stage_ev[2] = {nullptr, nullptr};
    bool tc_stage_used[2] = {false, false};
    cudaEvent_t tc_ev0[2] = {nullptr, nullptr}, tc_ev1[2] = {nullptr, nullptr};  // kernel timing per staging parity
    int tc_parity = 0;
    mb200::PrefetchRead *release_after_tables = nullptr;  // the prefetch slot whose last reader is the next k_sf_tables launch
    cudaEvent_t t_ev0 = nullptr, t_ev1 = nullptr;  // events around the last contraction launch (finish_timing)
    int t_parity = -1;
    // deferred Saxs batches (mb200_saxs_frames_submit / _collect): results wait in page-locked memory
    struct Pending {
        bool open = false;
        size_t n_out = 0;        // doubles per output array
        int parity = 0;          // staging parity (timing events) of the batch
        int64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cudaEvent_t done = nullptr;
    };
    static constexpr int N_PENDING = 4;  // several analyses of a collection may each have a batch waiting
    Pending pending[N_PENDING];
    mb200::PinBuf h_res[N_PENDING];
    bool timing = false;
    bool sf_attr_set = false, hist_attr_set = false, tc_attr_set = false, hc_attr_set = false;
    unsigned hs_attr_set = 0;  // k_hist_sorted<B> instantiations whose shared-memory limit is raised
    int hist_mode = 0;         // weighted histograms: 0 chunked non-returning atomics (default), 1 three returning atomics, 2 counting sort
    int precision = 1;        // 1: int8 x 5 split precision on tcgen05 (default)  0: FP64 DMMA
    bool tc_a_tmem = true;    // split-precision kernel: generated A operand in tensor memory when the l tile allows it
    bool force_fp64 = false;  // this call must use the FP64 kernel (non-finite weights)
    int64_t launches = 0;
    int64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // S(q) workspaces
    mb200::DevBuf d_pos, d_w, d_idx, d_tables, d_amp, d_sval, d_jobs, d_segs, d_items, d_out, d_ff2, d_counter, d_bimg, d_tcdesc, d_owned;
    mb200::PinBuf h_stage, h_out;
    // histogram workspaces
    mb200::DevBuf d_hpos, d_hpack, d_hw, d_hw2, d_hedges, d_hpar, d_hout, d_ctopo;
    // host -> device prefetch of the next batch of frames (mb200_prefetch_frames): copy stream, two buffers
    cudaStream_t copy_stream = nullptr;
    static constexpr int N_PREF = 4;  // a collection of two analyses keeps four announcements alive (2 each)
    cudaEvent_t pref_ev[N_PREF] = {};
    mb200::DevBuf d_pref[N_PREF];
    const void *pref_src[N_PREF] = {};   // host pointer while announcements of the slot are still to be taken
    size_t pref_bytes[N_PREF] = {};
    const void *pref_host[N_PREF] = {};  // what the slot holds (pref_src is cleared by its last reader)
    uint64_t pref_tag[N_PREF] = {};      // content tag of a shared announcement, 0 = not shared
    int pref_readers[N_PREF] = {};       // announcements of the slot not taken yet
    int pref_busy[N_PREF] = {};          // compute calls that took the slot and have not queued their last reader yet
    uint64_t pref_seq[N_PREF] = {};      // announcement number of the slot's content (oldest is overwritten first)
    uint64_t pref_counter = 0;
    int64_t pref_copies = 0, pref_saved = 0, pref_skipped = 0;  // copies issued / saved by sharing / not issued (no free slot)
    std::mutex pref_mutex;
    // the device copy of a prefetch slot may be overwritten only after the calls that consumed it have read it: recorded on
    // the compute stream when a call's last reader has been queued (PrefetchRead::release), waited for by the copy stream
    cudaEvent_t pref_done_ev[N_PREF] = {};
    bool pref_done_set[N_PREF] = {};
    // Every compute entry point holds this for its whole duration: the workspaces above and the stream are shared by all
    // callers of the context (several analysis instances each drive their own worker thread, AnalysisCollection).
    std::recursive_mutex call_mutex;
    // group-index array of the last mb200_saxs_frames call that is resident in d_idx (size, fingerprint, atom count)
    bool idx_valid = false;
    size_t idx_n = 0;
    unsigned long long idx_fp = 0;
    int64_t idx_natoms = 0;
    std::list<std::shared_ptr<mb200::SfPlan>> plans;  // small LRU cache keyed by box + q window
    size_t plan_cap = 64;                              // entries kept (raised to the frames of a call: NPT trajectories)
    std::shared_ptr<mb200::PlanPool> plan_pool;        // created with the first plan
    std::mutex plan_mutex;                             // guards plans / plans_ahead (mb200_plans_ahead runs beside a compute call)
    std::vector<std::shared_ptr<mb200::SfPlan>> plans_ahead;  // built by mb200_plans_ahead, not uploaded yet
    ~mb200_ctx();
};

// index lists of atom groups inside whole frames, resident on one device (mb200_groups_create, sfactor.cu)
struct mb200_groups {
    mb200::DevBuf idx;          // int32, the groups' lists back to back
    std::vector<int64_t> ptr;   // [n_groups + 1] offsets into idx
    int64_t n_atoms_total = 0;  // atoms per frame
    int device = 0;
};
