// Internal helpers shared by the translation units of libmixmogam_b200 (not part of the ABI): stream-ordered temporaries,
// tensor-map encoding, the launchers of the tcgen05 GEMM core, environment knobs.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ctx.h"
#include "kinship_kernels.cuh"
#include "tc_gemm.cuh"

namespace mmg {

// Temporary device buffer from the stream-ordered pool (cudaMallocAsync): allocation and release are ordered on the
// context's stream, cost no device synchronisation, and the pool keeps freed blocks for the next call
// (release threshold = unlimited, set in mmg_create; trimmed when a plain cudaMalloc runs out of memory).
struct DevBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, s); }
    cudaError_t alloc(cudaStream_t st, size_t bytes) {
        s = st;
        return cudaMallocAsync(&p, bytes ? bytes : 1, st);
    }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// cudaMalloc for the long-lived blocks (genotypes, Gram, packed operand); gives the async pool's cache back first if needed
inline cudaError_t persistent_malloc(int device, void** ptr, size_t bytes) {
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            cudaDeviceSynchronize();
            cudaMemPoolTrimTo(pool, 0);
        }
        e = cudaMalloc(ptr, bytes);
    }
    return e;
}

// ---- driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda) -------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// uint8 [rows x kbytes] row-major, row stride `pitch` bytes; box = 128 bytes x box_rows; 128B swizzle
static int make_tmap_u8(mmg_ctx* ctx, CUtensorMap* tm, const void* base, int64_t kbytes, int64_t rows, int64_t pitch,
                        int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(ctx, MMG_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)kbytes, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, MMG_ECUDA, "cuTensorMapEncodeTiled failed: %d (kbytes=%lld rows=%lld pitch=%lld)",
                                       (int)r, (long long)kbytes, (long long)rows, (long long)pitch);
    return MMG_OK;
}

static int ensure_scratch(mmg_ctx* ctx, int64_t bytes) {
    if (ctx->scratch_bytes >= bytes) return MMG_OK;
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    MMG_CUDA(ctx, persistent_malloc(ctx->device, &ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return MMG_OK;
}

// view of a persistent workspace with DevBuf's accessors (nothing to release)
struct WsBuf {
    void* p = nullptr;
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// persistent workspace `id` of at least `bytes` bytes (contents undefined).  Work on it is ordered by the context's stream like
// everything else; growing it waits for the stream (rare: sizes repeat from call to call).
static int ws_get(mmg_ctx* ctx, int id, int64_t bytes, void** out) {
    if (ctx->ws_bytes[id] < bytes) {
        if (ctx->ws[id]) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->ws[id]);
        }
        ctx->ws[id] = nullptr;
        ctx->ws_bytes[id] = 0;
        MMG_CUDA(ctx, persistent_malloc(ctx->device, &ctx->ws[id], (size_t)bytes));
        ctx->ws_bytes[id] = bytes;
    }
    *out = ctx->ws[id];
    return MMG_OK;
}

static MmgMat* get_mat(mmg_ctx* ctx, mmg_mat h) {
    auto it = ctx->mats.find(h);
    return it == ctx->mats.end() ? nullptr : &it->second;
}

static int launch_check(mmg_ctx* ctx, const char* what) {
    ctx->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, MMG_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return MMG_OK;
}

static int env_int(const char* var, int dflt) {
    const char* e = getenv(var);
    return e ? atoi(e) : dflt;
}

static uint64_t env_policy(const char* var, uint64_t dflt) {
    const char* e = getenv(var);
    if (!e) return dflt;
    if (!strcmp(e, "first")) return L2_EVICT_FIRST;
    if (!strcmp(e, "last")) return L2_EVICT_LAST;
    if (!strcmp(e, "normal")) return L2_EVICT_NORMAL;
    return dflt;
}

// clusters of CS CTAs of tc_gemm_i8_kernel<Epi, CS> that can be co-resident: the persistent grid of launch_tc_gemm
template <class Epi, int CS, int KIND = TC_KIND_I8>
static int tc_gemm_max_clusters(mmg_ctx* ctx) {
    int max_clusters = ctx->sm_count / CS;
    if (CS > 1) {
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = TC_SMEM_BYTES;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cfg.gridDim = dim3((unsigned)(ctx->sm_count / CS * CS));
        int q = 0;
        if (cudaOccupancyMaxActiveClusters(&q, tc_gemm_i8_kernel<Epi, CS, KIND>, &cfg) == cudaSuccess && q > 0) max_clusters = std::min(max_clusters, q);
        else cudaGetLastError();
    }
    return std::max(1, max_clusters);
}

// Launch the tcgen05 GEMM with a cluster of CS CTAs.  The persistent grid is the number of clusters that can be
// co-resident (cudaOccupancyMaxActiveClusters; clusters of 4 do not tile every GPC) times CS.
template <class Epi, int CS, int KIND = TC_KIND_I8>
static int launch_tc_gemm(mmg_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcTile* tiles_d, int num_groups,
                          int tiles_per_group, int table_stride, int group_m_step, int rank_m_step,
                          const typename Epi::Params& ep, const char* name, uint64_t hint_a = L2_EVICT_NORMAL,
                          uint64_t hint_b = L2_EVICT_NORMAL, int prefetch = 0) {
    auto kern = tc_gemm_i8_kernel<Epi, CS, KIND>;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TC_SMEM_BYTES;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int max_clusters = tc_gemm_max_clusters<Epi, CS, KIND>(ctx);
    const int cgroups = (num_groups + CS - 1) / CS;
    const int clusters = std::max(1, std::min(cgroups, max_clusters));
    cfg.gridDim = dim3((unsigned)(clusters * CS));
    // L2 eviction priority of the two operand streams: MMG_TC_HINT_A / MMG_TC_HINT_B = normal | first | last
    const uint64_t pa = env_policy("MMG_TC_HINT_A", hint_a), pb = env_policy("MMG_TC_HINT_B", hint_b);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tiles_d, num_groups, tiles_per_group, table_stride, group_m_step,
                                       rank_m_step, pa, pb, ep, prefetch);
    ctx->launches += 1;
    if (e != cudaSuccess) return fail(ctx, MMG_ECUDA, "launch of %s (cluster %d, grid %d) failed: %s", name, CS, clusters * CS,
                                      cudaGetErrorString(e));
    return MMG_OK;
}

// MMG_GRAM_IMPL / MMG_SCAN_IMPL = tcgen05 | simt | dmma select what MMG_IMPL_AUTO means (both are CUDA paths)
static int env_impl(const char* var, int dflt) {
    const char* e = getenv(var);
    if (!e) return dflt;
    if (!strcmp(e, "tcgen05")) return MMG_IMPL_TCGEN05;
    if (!strcmp(e, "simt")) return MMG_IMPL_SIMT;
    if (!strcmp(e, "dmma")) return MMG_IMPL_DMMA;
    return dflt;
}

static double lbeta_host(double a, double b) {
    return (double)(lgammal((long double)a) + lgammal((long double)b) - lgammal((long double)a + (long double)b));
}

// the side stream (pack kernels under the Gram, the scan's linear pre-pass under the R'R product) and its two join events
static int ensure_side_stream(mmg_ctx* ctx) {
    if (ctx->stream2) return MMG_OK;
    MMG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    MMG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ov0, cudaEventDisableTiming));
    MMG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ov1, cudaEventDisableTiming));
    return MMG_OK;
}

static int ensure_tiles(mmg_ctx* ctx, const std::vector<TcTile>& tiles) {
    const int64_t bytes = (int64_t)tiles.size() * sizeof(TcTile);
    if (ctx->tiles_bytes < bytes) {
        cudaFree(ctx->tiles_d);
        ctx->tiles_d = nullptr;
        ctx->tiles_bytes = 0;
        MMG_CUDA(ctx, cudaMalloc(&ctx->tiles_d, bytes));
        ctx->tiles_bytes = bytes;
    }
    MMG_CUDA(ctx, cudaMemcpyAsync(ctx->tiles_d, tiles.data(), bytes, cudaMemcpyHostToDevice, ctx->stream));
    return MMG_OK;
}

static int scale_k_device(mmg_ctx* ctx, MmgMat* K, double* scalar) {
    const int n = (int)K->rows;
    MMG_TRY(ensure_scratch(ctx, (n + 2) * sizeof(double)));
    double* rs = (double*)ctx->scratch;
    rowsum_kernel<<<n, 256, 0, ctx->stream>>>(K->d, K->cols, n, rs);
    MMG_TRY(launch_check(ctx, "rowsum_kernel"));
    scale_k_reduce_kernel<<<1, 1024, 0, ctx->stream>>>(rs, K->d, K->cols, n, rs + n);
    MMG_TRY(launch_check(ctx, "scale_k_reduce_kernel"));
    double h[2];
    MMG_CUDA(ctx, cudaMemcpyAsync(h, rs + n, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double c = h[1] - h[0] / (double)n;                 // tr(K) - sum(K)/n  (kinship.py:95)
    const double s = (double)(n - 1) / c;                     // :96
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
    scale_matrix_kernel<<<grid, 256, 0, ctx->stream>>>(K->d, K->cols, n, n, s);
    MMG_TRY(launch_check(ctx, "scale_matrix_kernel"));
    if (scalar) *scalar = s;
    return MMG_OK;
}

}  // namespace mmg
