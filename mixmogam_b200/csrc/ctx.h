// Internal context of libmixmogam_b200 (not part of the ABI).
#pragma once
#include <cublas_v2.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <time.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/mixmogam_b200.h"

struct MmgMat {
    double* d = nullptr;
    int64_t rows = 0, cols = 0;   // dense row-major, ld == cols
};

struct MmgTimer {
    double seconds = 0.0;
    int64_t calls = 0;
};

// a stage whose start / stop events are recorded but not read yet (read when the timers are queried: no host synchronisation
// at the end of every stage)
struct MmgPendingTimer {
    const char* name;
    cudaEvent_t e0, e1;
};

struct mmg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cublasHandle_t cublas = nullptr;
    cusolverDnHandle_t cusolver = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, kev0 = nullptr, kev1 = nullptr;
    cudaStream_t stream2 = nullptr;    // side stream (created on first use): the scan's linear pre-pass underneath the R'R product
    cudaEvent_t ov0 = nullptr, ov1 = nullptr;
    std::string err;
    std::map<int64_t, MmgMat> mats;
    int64_t next_mat = 1;
    int64_t launches = 0;
    std::map<std::string, MmgTimer> timers;
    std::vector<MmgPendingTimer> pending_timers;
    std::vector<cudaEvent_t> event_pool;
    int scan_slices_hint = 0;          // digit planes the previous int8 scan of this shape certified with (0 = none: run the pilot)
    int64_t scan_hint_n = 0;
    double last_gram_ms = 0.0, last_scan_ms = 0.0, last_perm_ms = 0.0, last_ibd_ms = 0.0;
    int last_scan_impl = 0;            // MMG_IMPL_* the last mmg_emmax_scan_f64 / _betas_f64 call ran (what AUTO resolved to)
    int last_gram_pair = 0;            // ... as a CTA-pair MMA (gram_pair_kernel) rather than the multicast form of the GEMM core
    int last_gram_fp4 = 0;             // the last tcgen05 Gram multiplied e2m1 operands (kind::mxf4) rather than int8
    int last_scan_slices = 0;          // digit planes used by the most recent int8 scan
    double last_scan_rho = 0.0;        // its certified relative truncation bound on x~.x~ (max over SNPs)

    // resident genotypes
    int8_t* snps = nullptr;
    int64_t snps_capacity = 0;         // bytes behind snps (a smaller block re-uses the allocation: mmg_snps_reserve)
    int64_t m = 0, n = 0, pitch = 0;
    int snps_absmax = -1;          // max |genotype| of the resident block, -1 = not measured since the last write

    // kinship state
    int32_t* G = nullptr;          // [g_pad x g_pad]
    int64_t g_pad = 0;
    bool g_zero = true;
    int8_t* pack = nullptr;        // packed K-major operand of one chunk
    int64_t pack_bytes = 0;
    void* tiles_d = nullptr;       // tile table
    int64_t tiles_bytes = 0;
    int* flag_d = nullptr;         // device error flag

    // packed (2-bit) host -> device staging of the streamed Gram (mmg_kinship_gram_i8_host): two page-locked host slots and
    // two device slots, kept across calls (page-locking 330 MB costs more than a whole kinship)
    uint8_t* stage_host[2] = {nullptr, nullptr};
    uint8_t* stage_dev[2] = {nullptr, nullptr};
    int64_t stage_bytes = 0;
    int64_t last_h2d_packed = 0, last_h2d_raw = 0;   // chunks the last streamed Gram sent over each lane
    double pack_s_per_byte = 0.0;                    // measured host packing cost (seconds per genotype byte), 0 = not measured yet

    // scratch
    void* scratch = nullptr;
    int64_t scratch_bytes = 0;

    // persistent workspaces of the scan (grow-only, one per role): the few buffers of hundreds of MB a scan needs every call are
    // kept instead of going through the stream-ordered pool each time -- the pool's occasional growth (cuMemCreate + map) showed
    // up as milliseconds of host time on some ranks of a multi-GPU step, which every other rank then waits for in the next collective
    // linear pre-pass of the scan launched ahead on the side stream (mmg_scan_prepass_begin): its outputs wait in the
    // MMG_WS_SCAN_PRE workspace for the scan over exactly this row range
    bool early_prepass = false;
    int64_t early_begin = 0, early_count = 0;
    void* ws[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int64_t ws_bytes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
enum { MMG_WS_OZAKI_PLANES = 0, MMG_WS_QUAD_A = 1, MMG_WS_QUAD_BQ = 2, MMG_WS_SCAN_PRE = 3, MMG_WS_SCAN_VEC = 4, MMG_WS_SCAN_OUT = 5, MMG_WS_SCAN_EARLY = 6 };

namespace mmg {

extern thread_local std::string g_create_error;

inline int fail(mmg_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define MMG_CUDA(ctx, call)                                                                              \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return mmg::fail(ctx, e__ == cudaErrorMemoryAllocation ? MMG_EOOM : MMG_ECUDA, "%s:%d %s: %s", \
                             __FILE__, __LINE__, #call, cudaGetErrorString(e__));                        \
    } while (0)
#define MMG_CUBLAS(ctx, call)                                                                            \
    do {                                                                                                 \
        cublasStatus_t s__ = (call);                                                                     \
        if (s__ != CUBLAS_STATUS_SUCCESS)                                                                \
            return mmg::fail(ctx, MMG_ECUBLAS, "%s:%d %s: cublas status %d", __FILE__, __LINE__, #call, (int)s__); \
    } while (0)
#define MMG_CUSOLVER(ctx, call)                                                                          \
    do {                                                                                                 \
        cusolverStatus_t s__ = (call);                                                                   \
        if (s__ != CUSOLVER_STATUS_SUCCESS)                                                              \
            return mmg::fail(ctx, MMG_ECUSOLVER, "%s:%d %s: cusolver status %d", __FILE__, __LINE__, #call, (int)s__); \
    } while (0)
#define MMG_CHECK(ctx, cond, ...)                                      \
    do {                                                               \
        if (!(cond)) return mmg::fail(ctx, MMG_EBADARG, __VA_ARGS__);  \
    } while (0)
#define MMG_TRY(expr)                 \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != MMG_OK) return rc__; \
    } while (0)

inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// Reads every pending stage timer (waits for the stop events: the stream's work up to them must complete) and returns the
// event pairs to the pool.
inline void resolve_timers(mmg_ctx* ctx) {
    for (const MmgPendingTimer& p : ctx->pending_timers) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
            MmgTimer& t = ctx->timers[p.name];
            t.seconds += ms * 1e-3;
            t.calls += 1;
        }
        ctx->event_pool.push_back(p.e0);
        ctx->event_pool.push_back(p.e1);
    }
    cudaGetLastError();
    ctx->pending_timers.clear();
}

// accumulates CUDA-event time of a stage.  stop() only records the event: the elapsed time is read by resolve_timers
// (mmg_timer_get / mmg_sync), so a stage ends without a host synchronisation.
struct StageTimer {
    mmg_ctx* ctx;
    const char* name;
    bool running;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    static cudaEvent_t take(mmg_ctx* c) {
        cudaEvent_t e = nullptr;
        if (!c->event_pool.empty()) {
            e = c->event_pool.back();
            c->event_pool.pop_back();
        } else if (cudaEventCreate(&e) != cudaSuccess) {
            cudaGetLastError();
            e = nullptr;
        }
        return e;
    }
    StageTimer(mmg_ctx* c, const char* nm) : ctx(c), name(nm), running(true) {
        if (c->pending_timers.size() >= 256) resolve_timers(c);
        e0 = take(c);
        e1 = take(c);
        if (e0) cudaEventRecord(e0, c->stream);
    }
    void stop() {
        if (!running) return;
        running = false;
        if (!e0 || !e1) return;
        cudaEventRecord(e1, ctx->stream);
        ctx->pending_timers.push_back(MmgPendingTimer{name, e0, e1});
    }
    ~StageTimer() { stop(); }
};

// wall-clock time the HOST spends inside a scope (allocator calls, synchronisations), accumulated under `name`
struct HostTimer {
    mmg_ctx* ctx;
    const char* name;
    timespec t0;
    HostTimer(mmg_ctx* c, const char* nm) : ctx(c), name(nm) { clock_gettime(CLOCK_MONOTONIC, &t0); }
    ~HostTimer() {
        timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        MmgTimer& t = ctx->timers[name];
        t.seconds += (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        t.calls += 1;
    }
};

}  // namespace mmg
