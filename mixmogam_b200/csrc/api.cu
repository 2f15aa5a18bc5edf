// libmixmogam_b200: C ABI (include/mixmogam_b200.h) over the sm_100a kernels.
// Host-side orchestration only; every numerical stage of the hot path runs in the kernels of
// tc_gemm.cuh / scan_tc.cuh (tcgen05), scan_dmma.cuh (FP64 tensor cores), reml.cuh, kinship_kernels.cuh.
// cuBLAS / cuSOLVER appear only for plain library work outside the hot path (dgemm plumbing, syevd).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ctx.h"
#include "fdist.cuh"
#include "ibd_tc.cuh"
#include "kinship_kernels.cuh"
#include "reml.cuh"
#include "scan_dmma.cuh"
#include "scan_tc.cuh"
#include "scan_quad.cuh"
#include "tc_gemm.cuh"

namespace mmg {
thread_local std::string g_create_error;

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
static cudaError_t persistent_malloc(int device, void** ptr, size_t bytes) {
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
template <class Epi, int CS>
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
        if (cudaOccupancyMaxActiveClusters(&q, tc_gemm_i8_kernel<Epi, CS>, &cfg) == cudaSuccess && q > 0) max_clusters = std::min(max_clusters, q);
        else cudaGetLastError();
    }
    return std::max(1, max_clusters);
}

// Launch the tcgen05 GEMM with a cluster of CS CTAs.  The persistent grid is the number of clusters that can be
// co-resident (cudaOccupancyMaxActiveClusters; clusters of 4 do not tile every GPC) times CS.
template <class Epi, int CS>
static int launch_tc_gemm(mmg_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcTile* tiles_d, int num_groups,
                          int tiles_per_group, int table_stride, int group_m_step, int rank_m_step,
                          const typename Epi::Params& ep, const char* name, uint64_t hint_a = L2_EVICT_NORMAL,
                          uint64_t hint_b = L2_EVICT_NORMAL) {
    auto kern = tc_gemm_i8_kernel<Epi, CS>;
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
    const int max_clusters = tc_gemm_max_clusters<Epi, CS>(ctx);
    const int cgroups = (num_groups + CS - 1) / CS;
    const int clusters = std::max(1, std::min(cgroups, max_clusters));
    cfg.gridDim = dim3((unsigned)(clusters * CS));
    // L2 eviction priority of the two operand streams: MMG_TC_HINT_A / MMG_TC_HINT_B = normal | first | last
    const uint64_t pa = env_policy("MMG_TC_HINT_A", hint_a), pb = env_policy("MMG_TC_HINT_B", hint_b);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tiles_d, num_groups, tiles_per_group, table_stride, group_m_step,
                                       rank_m_step, pa, pb, ep);
    ctx->launches += 1;
    if (e != cudaSuccess) return fail(ctx, MMG_ECUDA, "launch of %s (cluster %d, grid %d) failed: %s", name, CS, clusters * CS,
                                      cudaGetErrorString(e));
    return MMG_OK;
}

// Launch the genotype-stationary scan kernel (scan_quad.cuh): cluster CS, panel of PKB K-blocks, STAGES digit stages.
template <int CS, int PKB, int STAGES, bool PAIR, int BN, int LDW = 16, bool TIMED = false>
static int launch_scan_quad(mmg_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, QuadShape sh, const QuadEpi::Params& ep) {
    auto kern = scan_quad_kernel<CS, PKB, STAGES, PAIR, BN, LDW, TIMED>;
    if (!TIMED) sh.dbg = nullptr;             // only the instrumented instances write the role counters
    constexpr int smem = QuadSmem<PKB, STAGES, PAIR, BN>::kBytes;
    {   // per device, cheap: set on every launch
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fail(ctx, MMG_ECUDA, "scan_quad_kernel: cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(e));
    }
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(QP_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = ctx->sm_count / CS;
    if (CS > 1) {
        cfg.gridDim = dim3((unsigned)(ctx->sm_count / CS * CS));
        int q = 0;
        if (cudaOccupancyMaxActiveClusters(&q, kern, &cfg) == cudaSuccess && q > 0) max_clusters = std::min(max_clusters, q);
        else cudaGetLastError();
    }
    const int cgroups = (sh.num_groups + CS - 1) / CS;
    const int clusters = std::max(1, std::min(cgroups, max_clusters));
    cfg.gridDim = dim3((unsigned)(clusters * CS));
    const uint64_t pa = env_policy("MMG_TC_HINT_A", L2_EVICT_FIRST), pb = env_policy("MMG_TC_HINT_B", L2_EVICT_LAST);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, sh, pa, pb, ep);
    ctx->launches += 1;
    if (e != cudaSuccess) return fail(ctx, MMG_ECUDA, "launch of scan_quad_kernel<%d,%d,%d,%d> (grid %d) failed: %s", CS, PKB, STAGES,
                                      (int)PAIR, clusters * CS, cudaGetErrorString(e));
    return MMG_OK;
}

template <int CS>
static int launch_scan_quad_cs(mmg_ctx* ctx, int panel, const CUtensorMap& tmA, const CUtensorMap& tmB, const QuadShape& sh,
                               const QuadEpi::Params& ep) {
    if (panel == 8) {
        // MMG_SCAN_LD = 16 | 32: columns per tcgen05.ld of the epilogue (clusters of 2 only)
        if constexpr (CS == 2) {
            if (env_int("MMG_SCAN_LD", 16) == 32) return launch_scan_quad<CS, 8, 3, false, 256, 32>(ctx, tmA, tmB, sh, ep);
            if (sh.dbg) return launch_scan_quad<CS, 8, 3, false, 256, 16, true>(ctx, tmA, tmB, sh, ep);
        }
        return launch_scan_quad<CS, 8, 3, false, 256>(ctx, tmA, tmB, sh, ep);
    }
    if (panel == 4) return launch_scan_quad<CS, 4, 5, false, 256>(ctx, tmA, tmB, sh, ep);
    return launch_scan_quad<CS, 6, 4, false, 256>(ctx, tmA, tmB, sh, ep);
}

// CTA-pair form (tcgen05.mma.cta_group::2): half digit tiles of 16 KB per stage
static int launch_scan_quad_pair(mmg_ctx* ctx, int panel, const CUtensorMap& tmA, const CUtensorMap& tmB, const QuadShape& sh,
                                 const QuadEpi::Params& ep) {
    if (panel == 8) {
        if (env_int("MMG_SCAN_LD", 16) == 32) return launch_scan_quad<2, 8, 6, true, 256, 32>(ctx, tmA, tmB, sh, ep);
        if (sh.dbg) return launch_scan_quad<2, 8, 6, true, 256, 16, true>(ctx, tmA, tmB, sh, ep);
        return launch_scan_quad<2, 8, 6, true, 256>(ctx, tmA, tmB, sh, ep);
    }
    if (panel == 4) return launch_scan_quad<2, 4, 10, true, 256>(ctx, tmA, tmB, sh, ep);
    return launch_scan_quad<2, 6, 8, true, 256>(ctx, tmA, tmB, sh, ep);
}

// CTA-pair form with 128-column tiles: four accumulator stages in TMEM, so the MMA may run three tiles ahead of the
// epilogue and the cross-CTA barrier latency of the pair leaves the critical path; 8 KB half tiles per stage
static int launch_scan_quad_pair128(mmg_ctx* ctx, int panel, const CUtensorMap& tmA, const CUtensorMap& tmB, const QuadShape& sh,
                                    const QuadEpi::Params& ep) {
    if (panel == 12) return launch_scan_quad<2, 12, 4, true, 128>(ctx, tmA, tmB, sh, ep);
    if (panel == 10) return launch_scan_quad<2, 10, 8, true, 128>(ctx, tmA, tmB, sh, ep);
    if (panel == 6) return launch_scan_quad<2, 6, 16, true, 128>(ctx, tmA, tmB, sh, ep);
    return launch_scan_quad<2, 8, 12, true, 128>(ctx, tmA, tmB, sh, ep);
}
// single-CTA MMA with 128-column tiles (four accumulator stages), digit tiles multicast over the cluster of 2
static int launch_scan_quad_n128(mmg_ctx* ctx, int panel, const CUtensorMap& tmA, const CUtensorMap& tmB, const QuadShape& sh,
                                 const QuadEpi::Params& ep) {
    if (panel == 10) return launch_scan_quad<2, 10, 4, false, 128>(ctx, tmA, tmB, sh, ep);
    return launch_scan_quad<2, 8, 6, false, 128>(ctx, tmA, tmB, sh, ep);
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

}  // namespace mmg

using namespace mmg;

extern "C" {

// ======================================================================================================
// context
// ======================================================================================================
int mmg_create(int device, mmg_ctx** out) {
    if (!out) return fail(nullptr, MMG_EBADARG, "mmg_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, MMG_ECUDA, "no CUDA device available (%s); mixmogam_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= count) return fail(nullptr, MMG_EBADARG, "device %d out of range [0,%d)", device, count);
    MMG_CUDA(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    MMG_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, MMG_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    mmg_ctx* ctx = new mmg_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
        cudaEventCreate(&ctx->kev0) != cudaSuccess || cudaEventCreate(&ctx->kev1) != cudaSuccess ||
        cudaMalloc(&ctx->flag_d, sizeof(int)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, MMG_ECUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (cublasCreate(&ctx->cublas) != CUBLAS_STATUS_SUCCESS || cusolverDnCreate(&ctx->cusolver) != CUSOLVER_STATUS_SUCCESS) {
        delete ctx;
        return fail(nullptr, MMG_ECUBLAS, "cublas/cusolver handle creation failed");
    }
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    cublasSetStream(ctx->cublas, ctx->stream);
    cusolverDnSetStream(ctx->cusolver, ctx->stream);
    cublasSetPointerMode(ctx->cublas, CUBLAS_POINTER_MODE_HOST);
    // opt in to large dynamic shared memory once
    cudaFuncSetAttribute(tc_gemm_i8_kernel<GramEpi, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<GramEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<QuadEpi, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<QuadEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<QuadEpi, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<IbdEpi, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<IbdEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<OzakiEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<PermEpi, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<PermEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(scan_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SD_SMEM_BYTES);
    cudaFuncSetAttribute(scan_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SD_SMEM_BYTES);
    *out = ctx;
    return MMG_OK;
}

int mmg_destroy(mmg_ctx* ctx) {
    if (!ctx) return MMG_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->mats) cudaFreeAsync(kv.second.d, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->snps);
    cudaFree(ctx->G);
    cudaFree(ctx->pack);
    cudaFree(ctx->tiles_d);
    cudaFree(ctx->flag_d);
    cudaFree(ctx->scratch);
    for (int i = 0; i < 2; ++i) {
        if (ctx->stage_host[i]) cudaFreeHost(ctx->stage_host[i]);
        cudaFree(ctx->stage_dev[i]);
    }
    if (ctx->cublas) cublasDestroy(ctx->cublas);
    if (ctx->cusolver) cusolverDnDestroy(ctx->cusolver);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->kev0);
    cudaEventDestroy(ctx->kev1);
    if (ctx->ov0) cudaEventDestroy(ctx->ov0);
    if (ctx->ov1) cudaEventDestroy(ctx->ov1);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MMG_OK;
}

const char* mmg_last_error(mmg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mmg_device_info(mmg_ctx* ctx, char* name64, int* sm_count, int* cc_major, int* cc_minor, int64_t* free_bytes,
                    int64_t* total_bytes) {
    MMG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    MMG_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    if (name64) { strncpy(name64, prop.name, 63); name64[63] = 0; }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    size_t f = 0, t = 0;
    MMG_CUDA(ctx, cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return MMG_OK;
}

int mmg_sync(mmg_ctx* ctx) {
    MMG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

int64_t mmg_launch_count(mmg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mmg_timer_get(mmg_ctx* ctx, const char* name, double* seconds, int64_t* calls) {
    MMG_CHECK(ctx, ctx && name, "bad argument");
    auto it = ctx->timers.find(name);
    if (seconds) *seconds = it == ctx->timers.end() ? 0.0 : it->second.seconds;
    if (calls) *calls = it == ctx->timers.end() ? 0 : it->second.calls;
    return MMG_OK;
}
int mmg_timer_reset(mmg_ctx* ctx) {
    MMG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    ctx->timers.clear();
    return MMG_OK;
}
int mmg_last_kernel_ms(mmg_ctx* ctx, const char* which, double* ms) {
    MMG_CHECK(ctx, ctx && which && ms, "bad argument");
    if (!strcmp(which, "gram")) *ms = ctx->last_gram_ms;
    else if (!strcmp(which, "scan")) *ms = ctx->last_scan_ms;
    else if (!strcmp(which, "perm")) *ms = ctx->last_perm_ms;
    else if (!strcmp(which, "ibd")) *ms = ctx->last_ibd_ms;
    else return fail(ctx, MMG_EBADARG, "unknown kernel '%s'", which);
    return MMG_OK;
}

int mmg_last_h2d_info(mmg_ctx* ctx, int64_t* packed_chunks, int64_t* raw_chunks, double* pack_gbs) {
    MMG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    if (packed_chunks) *packed_chunks = ctx->last_h2d_packed;
    if (raw_chunks) *raw_chunks = ctx->last_h2d_raw;
    if (pack_gbs) *pack_gbs = ctx->pack_s_per_byte > 0.0 ? 1e-9 / ctx->pack_s_per_byte : 0.0;
    return MMG_OK;
}
int mmg_last_scan_info(mmg_ctx* ctx, int* slices, double* rho) {
    MMG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    if (slices) *slices = ctx->last_scan_slices;
    if (rho) *rho = ctx->last_scan_rho;
    return MMG_OK;
}

int mmg_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return fail(nullptr, MMG_EBADARG, "mmg_host_alloc: bad argument");
    cudaError_t e = cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(nullptr, MMG_EOOM, "cudaHostAlloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
    return MMG_OK;
}
int mmg_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
    return MMG_OK;
}

// ======================================================================================================
// device matrices
// ======================================================================================================
int mmg_mat_create(mmg_ctx* ctx, int64_t rows, int64_t cols, mmg_mat* out) {
    MMG_CHECK(ctx, ctx && out && rows > 0 && cols > 0, "mmg_mat_create: bad shape %lld x %lld", (long long)rows, (long long)cols);
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    MmgMat m;
    m.rows = rows;
    m.cols = cols;
    MMG_CUDA(ctx, cudaMallocAsync((void**)&m.d, (size_t)rows * cols * sizeof(double), ctx->stream));
    MMG_CUDA(ctx, cudaMemsetAsync(m.d, 0, (size_t)rows * cols * sizeof(double), ctx->stream));
    *out = ctx->next_mat++;
    ctx->mats[*out] = m;
    return MMG_OK;
}
int mmg_mat_free(mmg_ctx* ctx, mmg_mat h) {
    MMG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    auto it = ctx->mats.find(h);
    if (it == ctx->mats.end()) return MMG_OK;
    cudaFreeAsync(it->second.d, ctx->stream);        // stream ordered: work already queued on the matrix completes first
    ctx->mats.erase(it);
    return MMG_OK;
}
int mmg_mat_shape(mmg_ctx* ctx, mmg_mat h, int64_t* rows, int64_t* cols) {
    MmgMat* m = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, m != nullptr, "unknown matrix handle %lld", (long long)h);
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    return MMG_OK;
}
int mmg_mat_upload(mmg_ctx* ctx, mmg_mat h, const double* host, int64_t ld_host) {
    MmgMat* m = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, m && host && ld_host >= m->cols, "mmg_mat_upload: bad argument");
    StageTimer tm(ctx, "h2d");
    MMG_CUDA(ctx, cudaMemcpy2DAsync(m->d, m->cols * sizeof(double), host, ld_host * sizeof(double), m->cols * sizeof(double),
                                    m->rows, cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}
int mmg_mat_download(mmg_ctx* ctx, mmg_mat h, double* host, int64_t ld_host) {
    MmgMat* m = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, m && host && ld_host >= m->cols, "mmg_mat_download: bad argument");
    StageTimer tm(ctx, "d2h");
    MMG_CUDA(ctx, cudaMemcpy2DAsync(host, ld_host * sizeof(double), m->d, m->cols * sizeof(double), m->cols * sizeof(double),
                                    m->rows, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}
int mmg_mat_device_ptr(mmg_ctx* ctx, mmg_mat h, void** dptr, int64_t* ld) {
    MmgMat* m = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, m && dptr, "mmg_mat_device_ptr: bad argument");
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *dptr = m->d;
    if (ld) *ld = m->cols;
    return MMG_OK;
}
int mmg_mat_copy(mmg_ctx* ctx, mmg_mat dst, mmg_mat src) {
    MmgMat *d = ctx ? get_mat(ctx, dst) : nullptr, *s = ctx ? get_mat(ctx, src) : nullptr;
    MMG_CHECK(ctx, d && s && d->rows == s->rows && d->cols == s->cols, "mmg_mat_copy: shape mismatch");
    MMG_CUDA(ctx, cudaMemcpyAsync(d->d, s->d, (size_t)d->rows * d->cols * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return MMG_OK;
}
int mmg_mat_gemm(mmg_ctx* ctx, int ta, int tb, double alpha, mmg_mat Ah, mmg_mat Bh, double beta, mmg_mat Ch) {
    MmgMat *A = ctx ? get_mat(ctx, Ah) : nullptr, *B = ctx ? get_mat(ctx, Bh) : nullptr, *C = ctx ? get_mat(ctx, Ch) : nullptr;
    MMG_CHECK(ctx, A && B && C, "mmg_mat_gemm: unknown handle");
    const int64_t m = ta ? A->cols : A->rows, k = ta ? A->rows : A->cols;
    const int64_t kb = tb ? B->cols : B->rows, n = tb ? B->rows : B->cols;
    MMG_CHECK(ctx, k == kb && C->rows == m && C->cols == n, "mmg_mat_gemm: shape mismatch (%lldx%lld)*(%lldx%lld)->(%lldx%lld)",
              (long long)m, (long long)k, (long long)kb, (long long)n, (long long)C->rows, (long long)C->cols);
    MMG_CHECK(ctx, C != A && C != B, "mmg_mat_gemm: output aliases an input");
    // row-major C = op(A) op(B)  <=>  column-major C' = op(B)' op(A)'
    MMG_CUBLAS(ctx, cublasDgemm(ctx->cublas, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, (int)n, (int)m,
                                (int)k, &alpha, B->d, (int)B->cols, A->d, (int)A->cols, &beta, C->d, (int)C->cols));
    return MMG_OK;
}
// A = R[row_begin : row_begin + row_count, :]' R[...]   (row-major LOWER triangle of the n x n matrix A; the strict upper
// triangle is left as it was).  Row blocks of R are contiguous, so ranks that each take a block of the n_out rows of the
// rotation and sum their A's (all-reduce) get R'R with 1/ranks of the flops each.
int mmg_mat_syrk_rows(mmg_ctx* ctx, mmg_mat Rh, int64_t row_begin, int64_t row_count, mmg_mat Ah) {
    MmgMat* R = ctx ? get_mat(ctx, Rh) : nullptr;
    MmgMat* A = ctx ? get_mat(ctx, Ah) : nullptr;
    MMG_CHECK(ctx, R && A, "mmg_mat_syrk_rows: unknown matrix handle");
    MMG_CHECK(ctx, A->rows == R->cols && A->cols == R->cols, "A must be %lld x %lld", (long long)R->cols, (long long)R->cols);
    MMG_CHECK(ctx, row_begin >= 0 && row_count >= 0 && row_begin + row_count <= R->rows, "row block out of range");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm(ctx, "scan_prep");
    const int64_t n = R->cols;
    const double one = 1.0, zero = 0.0;
    if (row_count == 0) {
        MMG_CUDA(ctx, cudaMemsetAsync(A->d, 0, (size_t)n * n * sizeof(double), ctx->stream));
        return MMG_OK;
    }
    // the row block is the column-major n x row_count matrix Rc; column-major UPPER of Rc Rc' = row-major LOWER of A
    MMG_CUBLAS(ctx, cublasDsyrk(ctx->cublas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, (int)n, (int)row_count, &one, R->d + row_begin * n, (int)n,
                                &zero, A->d, (int)n));
    return MMG_OK;
}
int mmg_mat_scale_rows(mmg_ctx* ctx, mmg_mat h, const double* d_host) {
    MmgMat* A = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, A && d_host, "mmg_mat_scale_rows: bad argument");
    MMG_TRY(ensure_scratch(ctx, A->rows * sizeof(double)));
    MMG_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, d_host, A->rows * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((unsigned)((A->cols + 255) / 256), (unsigned)A->rows);
    scale_rows_kernel<<<grid, 256, 0, ctx->stream>>>(A->d, A->cols, (int)A->rows, (int)A->cols, (const double*)ctx->scratch);
    MMG_TRY(launch_check(ctx, "scale_rows_kernel"));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}
int mmg_mat_add_diag(mmg_ctx* ctx, mmg_mat h, double alpha) {
    MmgMat* A = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, A && A->rows == A->cols, "mmg_mat_add_diag: needs a square matrix");
    add_diag_kernel<<<(unsigned)((A->rows + 255) / 256), 256, 0, ctx->stream>>>(A->d, A->cols, (int)A->rows, alpha);
    return launch_check(ctx, "add_diag_kernel");
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
int mmg_mat_scale_k(mmg_ctx* ctx, mmg_mat h, double* scalar) {
    MmgMat* K = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, K && K->rows == K->cols, "mmg_mat_scale_k: needs a square matrix");
    MMG_TRY(scale_k_device(ctx, K, scalar));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

int mmg_mat_syevd(mmg_ctx* ctx, mmg_mat h, double* w_host, double* seconds) {
    MmgMat* A = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, A && A->rows == A->cols && w_host, "mmg_mat_syevd: needs a square matrix and w_host");
    const int64_t n = A->rows;
    StageTimer tm(ctx, "syevd");
    cusolverDnParams_t params = nullptr;
    MMG_CUSOLVER(ctx, cusolverDnCreateParams(&params));
    size_t ws_dev = 0, ws_host = 0;
    double* w_dev = nullptr;
    void* buf_dev = nullptr;
    void* buf_host = nullptr;
    int* info_dev = nullptr;
    int rc = MMG_OK;
    do {
        if (persistent_malloc(ctx->device, (void**)&w_dev, n * sizeof(double)) != cudaSuccess || persistent_malloc(ctx->device, (void**)&info_dev, sizeof(int)) != cudaSuccess) {
            rc = fail(ctx, MMG_EOOM, "syevd: allocation failed");
            break;
        }
        cusolverStatus_t st = cusolverDnXsyevd_bufferSize(ctx->cusolver, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n,
                                                          CUDA_R_64F, A->d, n, CUDA_R_64F, w_dev, CUDA_R_64F, &ws_dev, &ws_host);
        if (st != CUSOLVER_STATUS_SUCCESS) { rc = fail(ctx, MMG_ECUSOLVER, "Xsyevd_bufferSize status %d", (int)st); break; }
        if (ws_dev && persistent_malloc(ctx->device, &buf_dev, ws_dev) != cudaSuccess) { rc = fail(ctx, MMG_EOOM, "syevd: workspace %zu B", ws_dev); break; }
        if (ws_host) buf_host = malloc(ws_host);
        st = cusolverDnXsyevd(ctx->cusolver, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A->d, n,
                              CUDA_R_64F, w_dev, CUDA_R_64F, buf_dev, ws_dev, buf_host, ws_host, info_dev);
        if (st != CUSOLVER_STATUS_SUCCESS) { rc = fail(ctx, MMG_ECUSOLVER, "Xsyevd status %d", (int)st); break; }
        int info = 0;
        if (cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(w_host, w_dev, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            rc = fail(ctx, MMG_ECUDA, "syevd: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (info != 0) { rc = fail(ctx, MMG_ECUSOLVER, "syevd did not converge (info=%d)", info); break; }
    } while (0);
    cudaFree(w_dev);
    cudaFree(info_dev);
    cudaFree(buf_dev);
    free(buf_host);
    cusolverDnDestroyParams(params);
    tm.stop();
    if (seconds) *seconds = ctx->timers["syevd"].seconds;
    return rc;
}

// ======================================================================================================
// genotypes
// ======================================================================================================
int mmg_snps_free(mmg_ctx* ctx) {
    MMG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->snps);
    ctx->snps = nullptr;
    ctx->m = ctx->n = ctx->pitch = 0;
    ctx->snps_absmax = -1;
    return MMG_OK;
}
int mmg_snps_reserve(mmg_ctx* ctx, int64_t m, int64_t n) {
    MMG_CHECK(ctx, ctx && m > 0 && n > 0, "mmg_snps_reserve: bad shape");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t pitch = round_up(n, 256);   // zero padded: kernels read whole 16/32-byte groups up to the next 256
    if (!(ctx->snps && ctx->m == m && ctx->n == n)) {
        MMG_TRY(mmg_snps_free(ctx));
        MMG_CUDA(ctx, persistent_malloc(ctx->device, (void**)&ctx->snps, (size_t)m * pitch));
        ctx->m = m;
        ctx->n = n;
        ctx->pitch = pitch;
        if (pitch != n) MMG_CUDA(ctx, cudaMemsetAsync(ctx->snps, 0, (size_t)m * pitch, ctx->stream));
    }
    return MMG_OK;
}
int mmg_snps_write(mmg_ctx* ctx, int64_t row0, const int8_t* snps, int64_t rows, int64_t ld) {
    MMG_CHECK(ctx, ctx && ctx->snps && snps && row0 >= 0 && rows >= 0 && row0 + rows <= ctx->m && ld >= ctx->n,
              "mmg_snps_write: bad argument");
    ctx->snps_absmax = -1;
    StageTimer tm(ctx, "h2d");
    // one strided DMA; measured at the PCIe rate (52 GB/s from page-locked memory), a staged 1-D copy + re-pitch kernel was no faster
    if (rows)
        MMG_CUDA(ctx, cudaMemcpy2DAsync(ctx->snps + row0 * ctx->pitch, ctx->pitch, snps, ld, ctx->n, rows, cudaMemcpyHostToDevice,
                                        ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}
int mmg_snps_upload(mmg_ctx* ctx, const int8_t* snps, int64_t m, int64_t n, int64_t ld) {
    MMG_TRY(mmg_snps_reserve(ctx, m, n));
    return mmg_snps_write(ctx, 0, snps, m, ld);
}
int mmg_snps_upload_rows(mmg_ctx* ctx, const int8_t* const* rows, int64_t m, int64_t n) {
    MMG_CHECK(ctx, ctx && rows, "mmg_snps_upload_rows: bad argument");
    MMG_TRY(mmg_snps_reserve(ctx, m, n));
    ctx->snps_absmax = -1;
    StageTimer tm(ctx, "h2d");
    // gather rows into two pinned staging buffers and copy them asynchronously
    const int64_t rows_per = std::max<int64_t>(1, (32ll << 20) / n);
    int8_t* stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2];
    for (int b = 0; b < 2; ++b) {
        if (cudaHostAlloc((void**)&stage[b], (size_t)rows_per * n, cudaHostAllocDefault) != cudaSuccess) {
            for (int c = 0; c < b; ++c) { cudaFreeHost(stage[c]); cudaEventDestroy(done[c]); }
            return fail(ctx, MMG_EOOM, "pinned staging allocation failed");
        }
        cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
    }
    int rc = MMG_OK;
    int b = 0;
    for (int64_t r0 = 0; r0 < m; r0 += rows_per, b ^= 1) {
        const int64_t cnt = std::min(rows_per, m - r0);
        cudaEventSynchronize(done[b]);
        for (int64_t r = 0; r < cnt; ++r) memcpy(stage[b] + r * n, rows[r0 + r], (size_t)n);
        if (cudaMemcpy2DAsync(ctx->snps + r0 * ctx->pitch, ctx->pitch, stage[b], n, n, cnt, cudaMemcpyHostToDevice, ctx->stream) !=
            cudaSuccess) {
            rc = fail(ctx, MMG_ECUDA, "row upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        cudaEventRecord(done[b], ctx->stream);
    }
    cudaStreamSynchronize(ctx->stream);
    for (int c = 0; c < 2; ++c) { cudaFreeHost(stage[c]); cudaEventDestroy(done[c]); }
    return rc;
}
int mmg_snps_shape(mmg_ctx* ctx, int64_t* m, int64_t* n) {
    MMG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    if (m) *m = ctx->m;
    if (n) *n = ctx->n;
    return MMG_OK;
}
int mmg_snps_device_ptr(mmg_ctx* ctx, void** dptr, int64_t* pitch) {
    MMG_CHECK(ctx, ctx && dptr, "bad argument");
    ctx->snps_absmax = -1;             // the caller may write through the pointer
    *dptr = ctx->snps;
    if (pitch) *pitch = ctx->pitch;
    return MMG_OK;
}
int mmg_snps_row_sums(mmg_ctx* ctx, int64_t* sums_host, int64_t* sumsq_host) {
    MMG_CHECK(ctx, ctx && ctx->snps && sums_host, "mmg_snps_row_sums: no resident genotypes");
    MMG_TRY(ensure_scratch(ctx, 2 * ctx->m * sizeof(long long)));
    long long* s = (long long*)ctx->scratch;
    long long* q = s + ctx->m;
    snp_row_sums_kernel<<<(unsigned)((ctx->m + 7) / 8), 256, 0, ctx->stream>>>(ctx->snps, ctx->pitch, ctx->m, (int)ctx->n, s, q);
    MMG_TRY(launch_check(ctx, "snp_row_sums_kernel"));
    MMG_CUDA(ctx, cudaMemcpyAsync(sums_host, s, ctx->m * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (sumsq_host) MMG_CUDA(ctx, cudaMemcpyAsync(sumsq_host, q, ctx->m * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

// ======================================================================================================
// stage 1: kinship
// ======================================================================================================
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

// Host genotypes streaming into the resident block while the Gram runs (mmg_kinship_gram_i8_host): the copies go out on
// their own stream in the Gram's 65 536-SNP chunks, one event per chunk; the pack kernel of chunk c waits for event c only.
//
// Two lanes feed the chunks.  RAW: one strided DMA of the int8 rows (page-locked source: asynchronous, ~52 GB/s).  PACKED:
// all host threads squeeze the chunk to 2 bits per genotype (host_pack.cpp; codes 0..3 only), a quarter-size copy follows
// and unpack2_kernel expands it into the resident block.  Each chunk goes to the lane that is expected to deliver it
// first (the DMA backlog against the measured pack time), so the PCIe link and the host cores work side by side:
// 10 GB arrive in ~10 / (52 + pack rate) seconds instead of 10 / 52.  MMG_H2D_PACK=0 keeps everything on the raw lane.
struct GramHostSource {
    const int8_t* snps = nullptr;      // SNP-major host rows, row stride ld
    int64_t ld = 0;
    cudaStream_t stream = nullptr;     // raw lane
    cudaStream_t stream2 = nullptr;    // packed lane (its small copies must not queue behind the raw ones)
    std::vector<cudaEvent_t> done;     // one per chunk
    std::vector<cudaEvent_t> pre;      // raw lane: recorded right before the chunk's copy (copy duration = pre -> done)
    cudaEvent_t t0 = nullptr;
    cudaEvent_t slot_free[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false};
    bool pinned = false;               // source rows are page-locked (raw copies are asynchronous)
    bool pack_ok = true;               // packed lane available (switched off by MMG_H2D_PACK=0 or a code outside 0..3)
    int threads = 1;
    int next_slot = 0;
    std::vector<int64_t> raw_queue;    // chunk ids on the raw lane in queue order
    size_t raw_done = 0;               // how many of them have been seen complete
    double raw_rate = 0.0;             // measured raw-lane rate (bytes/s) once a copy has completed
    int64_t packed_chunks = 0, raw_chunks = 0;
    ~GramHostSource() {
        for (cudaStream_t st : {stream, stream2})
            if (st) cudaStreamSynchronize(st);  // the host rows are borrowed for the duration of the call only
        for (cudaEvent_t e : done) cudaEventDestroy(e);
        for (cudaEvent_t e : pre)
            if (e) cudaEventDestroy(e);
        if (t0) cudaEventDestroy(t0);
        for (cudaEvent_t e : slot_free)
            if (e) cudaEventDestroy(e);
        for (cudaStream_t st : {stream, stream2})
            if (st) cudaStreamDestroy(st);
    }
};

extern "C" int mmg_host_pack2(const int8_t* src, int64_t rows, int64_t n, int64_t ld, uint8_t* dst, int64_t dst_ld, int threads);
extern "C" int mmg_host_threads_default();

// packed [rows x p_ld] (2 bits per genotype, code j of a row in bits 2 (j % 4) of byte j / 4) -> int8 [rows x pitch];
// one thread per 32-bit word = 16 genotypes = one 16-byte store
__global__ void unpack2_kernel(const uint8_t* __restrict__ packed, int64_t p_ld, int8_t* __restrict__ out, int64_t pitch, int64_t rows) {
    const int64_t wpr = p_ld >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * wpr) return;
    const int64_t r = idx / wpr, w = idx - r * wpr;
    if (16 * w >= pitch) return;
    const uint32_t v = *reinterpret_cast<const uint32_t*>(packed + r * p_ld + 4 * w);
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t b = (v >> (8 * k)) & 0xffu;
        o[k] = (b & 3u) | (((b >> 2) & 3u) << 8) | (((b >> 4) & 3u) << 16) | ((b >> 6) << 24);
    }
    *reinterpret_cast<uint4*>(out + r * pitch + 16 * w) = make_uint4(o[0], o[1], o[2], o[3]);
}

static double host_now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int gram_run(mmg_ctx* ctx, int coding, int impl, int64_t snp_begin, int64_t snp_count, int reset, GramHostSource* src);

int mmg_kinship_gram_i8(mmg_ctx* ctx, int coding, int impl, int64_t snp_begin, int64_t snp_count, int reset) {
    return gram_run(ctx, coding, impl, snp_begin, snp_count, reset, nullptr);
}

int mmg_kinship_gram_i8_host(mmg_ctx* ctx, int coding, int impl, const int8_t* snps, int64_t m, int64_t n, int64_t ld, int reset) {
    MMG_CHECK(ctx, ctx && snps && m > 0 && n > 0 && ld >= n, "mmg_kinship_gram_i8_host: bad argument");
    MMG_TRY(mmg_snps_reserve(ctx, m, n));
    ctx->snps_absmax = -1;
    GramHostSource src;
    src.snps = snps;
    src.ld = ld;
    MMG_CUDA(ctx, cudaStreamCreateWithFlags(&src.stream, cudaStreamNonBlocking));
    MMG_CUDA(ctx, cudaStreamCreateWithFlags(&src.stream2, cudaStreamNonBlocking));
    MMG_CUDA(ctx, cudaEventCreate(&src.t0));
    for (int i = 0; i < 2; ++i) MMG_CUDA(ctx, cudaEventCreateWithFlags(&src.slot_free[i], cudaEventDisableTiming));
    {
        cudaPointerAttributes pa{};
        src.pinned = cudaPointerGetAttributes(&pa, snps) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
    }
    src.pack_ok = env_int("MMG_H2D_PACK", 1) != 0;
    src.threads = std::max(1, env_int("MMG_HOST_THREADS", mmg_host_threads_default()));
    // the zero fill of the row padding (mmg_snps_reserve, compute stream) must not race with the copies
    MMG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    MMG_CUDA(ctx, cudaStreamWaitEvent(src.stream, ctx->ev0, 0));
    MMG_CUDA(ctx, cudaStreamWaitEvent(src.stream2, ctx->ev0, 0));
    MMG_CUDA(ctx, cudaEventRecord(src.t0, src.stream));
    const int rc = gram_run(ctx, coding, impl, 0, m, reset, &src);
    if (rc == MMG_OK) ctx->snps_absmax = coding == MMG_CODING_DIPLOID ? 2 : 1;   // the pack kernels checked every byte against the coding
    return rc;
}

static int gram_run(mmg_ctx* ctx, int coding, int impl, int64_t snp_begin, int64_t snp_count, int reset, GramHostSource* src) {
    MMG_CHECK(ctx, ctx && ctx->snps, "mmg_kinship_gram_i8: no resident genotypes");
    MMG_CHECK(ctx, coding == MMG_CODING_BINARY || coding == MMG_CODING_DIPLOID, "unknown coding %d", coding);
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count >= 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    if (impl == MMG_IMPL_AUTO) impl = env_impl("MMG_GRAM_IMPL", MMG_IMPL_TCGEN05);
    MMG_CHECK(ctx, impl == MMG_IMPL_TCGEN05 || impl == MMG_IMPL_SIMT, "unsupported impl %d for the Gram", impl);
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = (int)ctx->n;
    const int c = coding == MMG_CODING_DIPLOID ? 2 : 1;
    const int64_t g_pad = round_up(n, 256);
    if (!ctx->G || ctx->g_pad != g_pad) {
        cudaFree(ctx->G);
        ctx->G = nullptr;
        MMG_CUDA(ctx, persistent_malloc(ctx->device, (void**)&ctx->G, (size_t)g_pad * g_pad * sizeof(int32_t)));
        ctx->g_pad = g_pad;
        reset = 1;
    }
    if (reset) {
        MMG_CUDA(ctx, cudaMemsetAsync(ctx->G, 0, (size_t)g_pad * g_pad * sizeof(int32_t), ctx->stream));
        ctx->g_zero = true;
    }
    // int32 accumulator headroom: |entries| <= K-dim (values are +-1 or 0/1)
    MMG_CHECK(ctx, (double)snp_count * c < 2.0e9, "Gram K-dimension too large for int32 accumulation");

    const int64_t chunk = 65536;                         // SNPs per packed chunk (multiple of 128)
    const int64_t p_pitch = chunk * c;
    const int64_t need = (int64_t)n * p_pitch;
    if (ctx->pack_bytes < need) {
        cudaFree(ctx->pack);
        ctx->pack = nullptr;
        ctx->pack_bytes = 0;
        MMG_CUDA(ctx, persistent_malloc(ctx->device, (void**)&ctx->pack, need));
        ctx->pack_bytes = need;
    }
    MMG_CUDA(ctx, cudaMemsetAsync(ctx->flag_d, 0, sizeof(int), ctx->stream));

    // tile table: upper-triangular 128 x 256 tiles (row tile im needed for column tile jn iff im <= 2 jn + 1).
    // With a cluster of 2, one entry covers the row-tile pair (im, im+1) of a column tile (2 jn + 2 is even).
    int gram_cs = env_int("MMG_GRAM_CLUSTER", 2);
    if (gram_cs != 1) gram_cs = 2;
    std::vector<TcTile> tiles, table;
    int gram_clusters = 1;
    if (impl == MMG_IMPL_TCGEN05) {
        const int tiles_n = (n + TC_BN - 1) / TC_BN;
        for (int jn = 0; jn < tiles_n; ++jn)
            for (int im = 0; im <= 2 * jn + 1; im += gram_cs) tiles.push_back(TcTile{im * TC_BM, jn * TC_BN, 0, 0, 0, 0, 0, 0});
        gram_clusters = gram_cs == 2 ? tc_gemm_max_clusters<GramEpi, 2>(ctx) : tc_gemm_max_clusters<GramEpi, 1>(ctx);
    }
    // Tile table of one chunk.  Entry e runs on cluster e % W (W co-resident clusters), every tile costs the same, so
    // nt = q W + r tiles take q + 1 waves with only r clusters busy in the last one (n = 10k: 820 = 11 x 74 + 6, an 8 %
    // tail).  The r tail tiles are cut along K into floor(W / r) slices each, one slice per cluster, accumulated with
    // integer atomics (exact): the tail shrinks to 1 / floor(W / r) of a wave.  MMG_GRAM_SPLITK=0 turns it off.
    const bool split_tail = env_int("MMG_GRAM_SPLITK", 1) != 0;
    auto build_table = [&](int KB) {
        table.clear();
        const int nt = (int)tiles.size(), W = gram_clusters;
        const int r = nt % W, full = nt - r;
        const int splits = (split_tail && nt > W && r > 0) ? std::min(W / r, KB / 8) : 1;
        for (int e = 0; e < (splits > 1 ? full : nt); ++e) {
            TcTile t = tiles[(size_t)e];
            t.kb0 = 0;
            t.kb1 = KB;
            table.push_back(t);
        }
        if (splits > 1)
            for (int sl = 0; sl < splits; ++sl)
                for (int e = full; e < nt; ++e) {
                    TcTile t = tiles[(size_t)e];
                    t.kb0 = (int)((int64_t)KB * sl / splits);
                    t.kb1 = (int)((int64_t)KB * (sl + 1) / splits);
                    t.aux0 = 1;                                   // GramEpi: atomic accumulate
                    table.push_back(t);
                }
    };
    double gram_ms = 0.0, pack_s = 0.0;
    // host source: the chunks are staged by the two lanes described at GramHostSource (raw DMA / 2-bit packed)
    const int64_t p2_ld = round_up((ctx->n + 3) / 4, 16);            // packed row: 2 bits per genotype
    const double pcie_rate = 1e9 * std::max(1, env_int("MMG_PCIE_GBS", 50));
    const int64_t n_chunks = (snp_count + chunk - 1) / chunk;
    std::vector<char> staged((size_t)n_chunks, 0);
    if (src) {
        src->done.resize((size_t)n_chunks, nullptr);
        src->pre.resize((size_t)n_chunks, nullptr);
        for (auto& e : src->done) MMG_CUDA(ctx, cudaEventCreate(&e));
    }
    auto chunk_rows = [&](int64_t ci) { return std::min(chunk, snp_count - ci * chunk); };
    auto raw_seconds = [&](int64_t ci) { return (double)chunk_rows(ci) * (double)ctx->n / pcie_rate; };
    auto queue_raw = [&](int64_t ci) -> int {
        const int64_t s0 = ci * chunk, cnt = chunk_rows(ci);
        MMG_CUDA(ctx, cudaEventCreate(&src->pre[(size_t)ci]));
        MMG_CUDA(ctx, cudaEventRecord(src->pre[(size_t)ci], src->stream));
        MMG_CUDA(ctx, cudaMemcpy2DAsync(ctx->snps + (snp_begin + s0) * ctx->pitch, ctx->pitch, src->snps + (snp_begin + s0) * src->ld, src->ld,
                                        ctx->n, cnt, cudaMemcpyHostToDevice, src->stream));
        MMG_CUDA(ctx, cudaEventRecord(src->done[(size_t)ci], src->stream));
        src->raw_queue.push_back(ci);
        src->raw_chunks += 1;
        staged[(size_t)ci] = 1;
        return MMG_OK;
    };
    // returns MMG_OK with staged[ci] still 0 when the chunk holds a code outside 0..3 (the caller then takes the raw lane)
    auto queue_packed = [&](int64_t ci) -> int {
        const int64_t s0 = ci * chunk, cnt = chunk_rows(ci);
        const int64_t stage_need = std::min(chunk, snp_count) * p2_ld;       // one (largest) chunk of this call
        if (ctx->stage_bytes < stage_need) {
            for (int i = 0; i < 2; ++i) {
                if (ctx->stage_host[i]) cudaFreeHost(ctx->stage_host[i]);
                cudaFree(ctx->stage_dev[i]);
                ctx->stage_host[i] = ctx->stage_dev[i] = nullptr;
            }
            ctx->stage_bytes = 0;
            for (int i = 0; i < 2; ++i) {
                MMG_CUDA(ctx, cudaHostAlloc((void**)&ctx->stage_host[i], (size_t)stage_need, cudaHostAllocDefault));
                MMG_CUDA(ctx, cudaMalloc((void**)&ctx->stage_dev[i], (size_t)stage_need));
            }
            ctx->stage_bytes = stage_need;
        }
        const int sl = src->next_slot;
        if (src->slot_used[sl]) MMG_CUDA(ctx, cudaEventSynchronize(src->slot_free[sl]));
        const double t0 = host_now();
        if (mmg_host_pack2(src->snps + (snp_begin + s0) * src->ld, cnt, ctx->n, src->ld, ctx->stage_host[sl], p2_ld, src->threads) != 0) {
            src->pack_ok = false;                       // this and all later chunks take the raw lane
            return MMG_OK;
        }
        const double per_byte = (host_now() - t0) / ((double)cnt * (double)ctx->n);
        ctx->pack_s_per_byte = ctx->pack_s_per_byte > 0.0 ? 0.5 * (ctx->pack_s_per_byte + per_byte) : per_byte;
        MMG_CUDA(ctx, cudaMemcpyAsync(ctx->stage_dev[sl], ctx->stage_host[sl], (size_t)(cnt * p2_ld), cudaMemcpyHostToDevice, src->stream2));
        const int64_t words = cnt * (p2_ld >> 2);
        unpack2_kernel<<<(unsigned)((words + 255) / 256), 256, 0, src->stream2>>>(ctx->stage_dev[sl], p2_ld, ctx->snps + (snp_begin + s0) * ctx->pitch,
                                                                                ctx->pitch, cnt);
        MMG_TRY(launch_check(ctx, "unpack2_kernel"));
        MMG_CUDA(ctx, cudaEventRecord(src->slot_free[sl], src->stream2));
        MMG_CUDA(ctx, cudaEventRecord(src->done[(size_t)ci], src->stream2));
        src->slot_used[sl] = true;
        src->next_slot = sl ^ 1;
        src->packed_chunks += 1;
        staged[(size_t)ci] = 1;
        return MMG_OK;
    };
    // Stage chunk ci (if a look-ahead has not done so already).  Packed lane when the raw lane would deliver it later:
    // pageable rows always (a raw copy blocks the host at the pageable rate), page-locked rows when the DMA backlog exceeds the
    // time the host needs to pack the chunk.  Before the host disappears into a pack, the raw lane is topped up with the
    // following chunks so that the link stays busy meanwhile.
    auto issue_copy = [&](int64_t ci) -> int {
        if (staged[(size_t)ci]) return MMG_OK;
        double raw_s = raw_seconds(ci);
        const double pack_s = ctx->pack_s_per_byte > 0.0 ? ctx->pack_s_per_byte * (double)chunk_rows(ci) * (double)ctx->n : raw_s;
        // backlog of the raw lane: bytes queued and not yet seen complete, at the rate measured on the copies that are
        // (the link is shared with the packed lane's copies and the host cores' own reads, so the nominal rate is not it)
        while (src->raw_done < src->raw_queue.size() && cudaEventQuery(src->done[(size_t)src->raw_queue[src->raw_done]]) == cudaSuccess) {
            const int64_t cd = src->raw_queue[src->raw_done++];
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, src->pre[(size_t)cd], src->done[(size_t)cd]) == cudaSuccess && ms > 0.f) {
                const double r = (double)chunk_rows(cd) * (double)ctx->n / (1e-3 * ms);
                src->raw_rate = src->raw_rate > 0.0 ? 0.5 * (src->raw_rate + r) : r;
            }
        }
        cudaGetLastError();                                  // cudaErrorNotReady from the query is not an error
        double pending = 0.0;
        for (size_t qi = src->raw_done; qi < src->raw_queue.size(); ++qi) pending += (double)chunk_rows(src->raw_queue[qi]) * (double)ctx->n;
        const double rate = src->raw_rate > 0.0 ? std::min(src->raw_rate, pcie_rate * 1.2) : pcie_rate;
        const double backlog = src->pinned ? pending / rate : 0.0;
        raw_s = (double)chunk_rows(ci) * (double)ctx->n / rate;
        if (src->pack_ok && (!src->pinned || backlog + raw_s > pack_s + 0.25 * raw_s)) {
            if (src->pinned) {
                double ahead = backlog;
                for (int64_t cj = ci + 1; cj < n_chunks && ahead < pack_s; ++cj) {
                    if (staged[(size_t)cj]) continue;
                    MMG_TRY(queue_raw(cj));
                    ahead += (double)chunk_rows(cj) * (double)ctx->n / rate;
                }
            }
            MMG_TRY(queue_packed(ci));
        }
        if (!staged[(size_t)ci]) MMG_TRY(queue_raw(ci));
        return MMG_OK;
    };
    // host source: no host synchronisation inside the chunk loop (the host packs while the GPU works): per-chunk events
    std::vector<cudaEvent_t> tev;
    struct TevGuard {
        std::vector<cudaEvent_t>& v;
        ~TevGuard() { for (cudaEvent_t e : v) cudaEventDestroy(e); }
    } tev_guard{tev};
    for (int64_t s0 = 0; s0 < snp_count; s0 += chunk) {
        const int64_t cnt = std::min(chunk, snp_count - s0);
        const int64_t kbytes = round_up(cnt, 128) * c;
        cudaEvent_t e_p0 = ctx->ev0, e_p1 = ctx->ev1, e_g0 = ctx->kev0, e_g1 = ctx->kev1;
        if (src) {
            MMG_TRY(issue_copy(s0 / chunk));
            MMG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, src->done[(size_t)(s0 / chunk)], 0));
            for (cudaEvent_t* e : {&e_p0, &e_p1, &e_g0, &e_g1}) {
                MMG_CUDA(ctx, cudaEventCreate(e));
                tev.push_back(*e);
            }
        }
        // ---- pack ----
        cudaEventRecord(e_p0, ctx->stream);
        dim3 pgrid((unsigned)((cnt + 127) / 128), (unsigned)((n + 63) / 64));
        if (coding == MMG_CODING_BINARY)
            pack_kmajor_kernel<0><<<pgrid, 256, 0, ctx->stream>>>(ctx->snps, ctx->pitch, snp_begin + s0, cnt, n, ctx->pack, p_pitch, ctx->flag_d);
        else
            pack_kmajor_kernel<1><<<pgrid, 256, 0, ctx->stream>>>(ctx->snps, ctx->pitch, snp_begin + s0, cnt, n, ctx->pack, p_pitch, ctx->flag_d);
        MMG_TRY(launch_check(ctx, "pack_kmajor_kernel"));
        cudaEventRecord(e_p1, ctx->stream);
        // ---- Gram ----
        const int accumulate = ctx->g_zero ? 0 : 1;
        cudaEventRecord(e_g0, ctx->stream);
        if (impl == MMG_IMPL_TCGEN05) {
            CUtensorMap tmA, tmB;
            MMG_TRY(make_tmap_u8(ctx, &tmA, ctx->pack, kbytes, n, p_pitch, TC_BM));
            MMG_TRY(make_tmap_u8(ctx, &tmB, ctx->pack, kbytes, n, p_pitch, TC_BN / gram_cs));
            build_table((int)(kbytes / TC_BK));
            MMG_TRY(ensure_tiles(ctx, table));
            GramEpi::Params ep{ctx->G, g_pad, accumulate};
            const int ngroups = (int)table.size() * gram_cs;
            if (gram_cs == 2)
                MMG_TRY((launch_tc_gemm<GramEpi, 2>(ctx, tmA, tmB, (const TcTile*)ctx->tiles_d, ngroups, 1, 1, 0, TC_BM, ep, "tc_gemm_i8_kernel<GramEpi,2>")));
            else
                MMG_TRY((launch_tc_gemm<GramEpi, 1>(ctx, tmA, tmB, (const TcTile*)ctx->tiles_d, ngroups, 1, 1, 0, 0, ep, "tc_gemm_i8_kernel<GramEpi,1>")));
        } else {
            dim3 ggrid((unsigned)((n + 63) / 64), (unsigned)((n + 63) / 64));
            gram_simt_kernel<<<ggrid, 256, 0, ctx->stream>>>(ctx->pack, p_pitch, n, kbytes, ctx->G, g_pad, accumulate);
            MMG_TRY(launch_check(ctx, "gram_simt_kernel"));
        }
        cudaEventRecord(e_g1, ctx->stream);
        ctx->g_zero = false;
        if (src) continue;                                      // timed after the loop
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        pack_s += ms * 1e-3;
        cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        gram_ms += ms;
    }
    if (src) {
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i + 3 < tev.size(); i += 4) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, tev[i], tev[i + 1]);
            pack_s += ms * 1e-3;
            cudaEventElapsedTime(&ms, tev[i + 2], tev[i + 3]);
            gram_ms += ms;
        }
    }
    ctx->timers["pack"].seconds += pack_s;
    ctx->timers["pack"].calls += 1;
    ctx->timers["gram"].seconds += gram_ms * 1e-3;
    ctx->timers["gram"].calls += 1;
    ctx->last_gram_ms = gram_ms;
    if (src && !src->done.empty()) {
        // span from the first copy to the arrival of the last chunk (either lane): overlaps the pack / gram timers
        float ms = 0.f, span = 0.f;
        for (cudaEvent_t e : src->done) {
            cudaEventSynchronize(e);
            if (cudaEventElapsedTime(&ms, src->t0, e) == cudaSuccess) span = std::max(span, ms);
        }
        ctx->timers["h2d"].seconds += span * 1e-3;
        ctx->timers["h2d"].calls += 1;
        ctx->last_h2d_packed = src->packed_chunks;
        ctx->last_h2d_raw = src->raw_chunks;
    }
    int bad = 0;
    MMG_CUDA(ctx, cudaMemcpy(&bad, ctx->flag_d, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad)
        return fail(ctx, MMG_EVALUE, "genotype values outside the domain of the '%s' coding (%s)",
                    coding == MMG_CODING_BINARY ? "binary" : "diploid_int", coding == MMG_CODING_BINARY ? "{0,1}" : "{0,1,2}");
    return MMG_OK;
}

int mmg_kinship_gram_ptr(mmg_ctx* ctx, void** dptr, int64_t* n, int64_t* ld) {
    MMG_CHECK(ctx, ctx && ctx->G && dptr, "mmg_kinship_gram_ptr: no Gram resident");
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *dptr = ctx->G;
    if (n) *n = ctx->n;
    if (ld) *ld = ctx->g_pad;
    return MMG_OK;
}

int mmg_kinship_gram_download(mmg_ctx* ctx, int32_t* G_host) {
    MMG_CHECK(ctx, ctx && ctx->G && G_host, "mmg_kinship_gram_download: no Gram resident");
    const int n = (int)ctx->n;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
    gram_mirror_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->G, ctx->g_pad, n);
    MMG_TRY(launch_check(ctx, "gram_mirror_kernel"));
    MMG_CUDA(ctx, cudaMemcpy2DAsync(G_host, (size_t)n * 4, ctx->G, (size_t)ctx->g_pad * 4, (size_t)n * 4, n, cudaMemcpyDeviceToHost,
                                    ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

int mmg_kinship_finalize_f64(mmg_ctx* ctx, int coding, int64_t m_total, int scaled, mmg_mat K_out, double* scale_scalar) {
    MmgMat* K = ctx ? get_mat(ctx, K_out) : nullptr;
    MMG_CHECK(ctx, K && ctx->G, "mmg_kinship_finalize_f64: need a Gram and an output matrix");
    MMG_CHECK(ctx, K->rows == ctx->n && K->cols == ctx->n, "K_out must be n x n");
    MMG_CHECK(ctx, m_total > 0, "m_total must be positive");
    StageTimer tm(ctx, "finalize");
    const int n = (int)ctx->n;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
    if (coding == MMG_CODING_BINARY)
        kinship_finalize_kernel<0><<<grid, 256, 0, ctx->stream>>>(ctx->G, ctx->g_pad, n, (double)m_total, K->d, K->cols);
    else
        kinship_finalize_kernel<1><<<grid, 256, 0, ctx->stream>>>(ctx->G, ctx->g_pad, n, (double)m_total, K->d, K->cols);
    MMG_TRY(launch_check(ctx, "kinship_finalize_kernel"));
    if (scale_scalar) *scale_scalar = 1.0;
    if (scaled) MMG_TRY(scale_k_device(ctx, K, scale_scalar));
    return MMG_OK;
}

__global__ void mirror_lower_to_upper_kernel(double* __restrict__ K, int64_t ld, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c < n && c > r) K[(int64_t)r * ld + c] = K[(int64_t)c * ld + r];
}

}  // extern "C"

// FP64 library path: standardise rows, cuBLAS dsyrk.  Handles any int8 genotype coding.
static int ibd_dsyrk_run(mmg_ctx* ctx, MmgMat* K, const std::vector<long long>& rows) {
    const int n = (int)ctx->n;
    const int64_t chunk = 2048;
    const int64_t zbytes = chunk * (int64_t)n * sizeof(double);
    const int64_t rbytes = round_up((int64_t)rows.size() * sizeof(long long), 256);
    MMG_TRY(ensure_scratch(ctx, zbytes + rbytes));
    double* Z = (double*)ctx->scratch;
    long long* rows_d = (long long*)((uint8_t*)ctx->scratch + zbytes);
    MMG_CUDA(ctx, cudaMemcpyAsync(rows_d, rows.data(), rows.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, cudaMemsetAsync(ctx->flag_d, 0, sizeof(int), ctx->stream));
    const double one = 1.0;
    for (int64_t r0 = 0; r0 < (int64_t)rows.size(); r0 += chunk) {
        const int64_t cnt = std::min<int64_t>(chunk, (int64_t)rows.size() - r0);
        standardise_rows_kernel<<<(unsigned)cnt, 256, 0, ctx->stream>>>(ctx->snps, ctx->pitch, rows_d + r0, n, Z, n, ctx->flag_d);
        MMG_TRY(launch_check(ctx, "standardise_rows_kernel"));
        // K += Z' Z.  Z row-major [cnt x n] is the column-major n x cnt matrix Zc; column-major UPPER of
        // Zc Zc' is the row-major lower triangle, mirrored below.
        MMG_CUBLAS(ctx, cublasDsyrk(ctx->cublas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, n, (int)cnt, &one, Z, n, &one, K->d, n));
    }
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
    mirror_lower_to_upper_kernel<<<grid, 256, 0, ctx->stream>>>(K->d, K->cols, n);
    MMG_TRY(launch_check(ctx, "mirror_lower_to_upper_kernel"));
    int bad = 0;
    MMG_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->flag_d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad) return fail(ctx, MMG_EVALUE, "monomorphic SNP in IBD kinship (std == 0; the reference asserts at kinship.py:67)");
    return MMG_OK;
}

// int8 tensor-core path (ibd_tc.cuh).  *domain_bad is set (and nothing is added to K) when a genotype is outside {0,1,2}.
static int ibd_tc_run(mmg_ctx* ctx, MmgMat* K, const std::vector<long long>& rows, int* domain_bad) {
    *domain_bad = 0;
    const int n = (int)ctx->n;
    const int64_t count = (int64_t)rows.size();
    int S = env_int("MMG_IBD_SLICES", 6);
    S = std::max(1, std::min(S, IBD_MAX_SLICES));
    const int64_t g_pad = round_up(n, 256), n_padM = g_pad;
    const int64_t dig_pitch = round_up(count, 128) + 128;
    DevBuf st, dig, Gw, P;
    // st: sums[m] | sumsq[m] | rows[count] (int64)  then  w | mean | coef [count] | u[g_pad] | cacc | amax (8-byte words)
    const int64_t n64 = 2 * ctx->m + count;
    const int64_t nd = 3 * count + g_pad + 2;
    MMG_CUDA(ctx, st.alloc(ctx->stream, (size_t)(n64 + nd) * 8));
    long long* d_sums = st.as<long long>();
    long long* d_sumsq = d_sums + ctx->m;
    long long* d_rows = d_sumsq + ctx->m;
    double* d_w = (double*)(d_rows + count);
    double* d_mean = d_w + count;
    double* d_coef = d_mean + count;
    double* d_u = d_coef + count;
    double* d_c = d_u + g_pad;
    unsigned long long* d_amax = (unsigned long long*)(d_c + 1);
    MMG_CUDA(ctx, cudaMemsetAsync(d_u, 0, (size_t)(g_pad + 2) * 8, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(d_rows, rows.data(), (size_t)count * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, cudaMemsetAsync(ctx->flag_d, 0, sizeof(int), ctx->stream));
    snp_row_sums_kernel<<<(unsigned)((ctx->m + 7) / 8), 256, 0, ctx->stream>>>(ctx->snps, ctx->pitch, ctx->m, n, d_sums, d_sumsq);
    MMG_TRY(launch_check(ctx, "snp_row_sums_kernel"));
    ibd_weights_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(d_sums, d_sumsq, d_rows, count, n, d_w, d_mean, d_amax,
                                                                                 ctx->flag_d);
    MMG_TRY(launch_check(ctx, "ibd_weights_kernel"));
    double amax = 0.0;
    int bad = 0;
    MMG_CUDA(ctx, cudaMemcpyAsync(&amax, d_amax, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->flag_d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad & 1) return fail(ctx, MMG_EVALUE, "monomorphic SNP in IBD kinship (std == 0; the reference asserts at kinship.py:67)");
    if (!(amax > 0.0) || !std::isfinite(amax)) return fail(ctx, MMG_EVALUE, "IBD kinship: bad weights (max %g)", amax);
    const int E = ilogb(amax) + 2;
    MMG_CUDA(ctx, dig.alloc(ctx->stream, (size_t)S * dig_pitch));
    MMG_CUDA(ctx, cudaMemsetAsync(dig.p, 0, (size_t)S * dig_pitch, ctx->stream));
    ibd_digits_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(d_w, d_mean, count, ldexp(1.0, -E), ldexp(1.0, E), S,
                                                                                dig.as<int8_t>(), dig_pitch, d_coef, d_c);
    MMG_TRY(launch_check(ctx, "ibd_digits_kernel"));
    const int slab = 2048;
    snp_weighted_colsum_kernel<<<dim3((unsigned)((n + 1023) / 1024), (unsigned)((count + slab - 1) / slab)), 256, 0, ctx->stream>>>(
        ctx->snps, ctx->pitch, d_rows, d_coef, count, slab, n, d_u);
    MMG_TRY(launch_check(ctx, "snp_weighted_colsum_kernel"));

    MMG_CUDA(ctx, Gw.alloc(ctx->stream, (size_t)g_pad * g_pad * sizeof(double)));
    MMG_CUDA(ctx, cudaMemsetAsync(Gw.p, 0, (size_t)g_pad * g_pad * sizeof(double), ctx->stream));
    int64_t chunk = (int64_t)(3.0e9 / ((double)(S + 1) * (double)n_padM)) / 128 * 128;
    chunk = std::max<int64_t>(4096, std::min<int64_t>(65536, chunk));
    chunk = std::min(chunk, round_up(count, 128));
    const int64_t p_pitch = chunk;
    MMG_CUDA(ctx, P.alloc(ctx->stream, (size_t)(S + 1) * n_padM * p_pitch));
    MMG_CUDA(ctx, cudaMemsetAsync(P.p, 0, (size_t)(S + 1) * n_padM * p_pitch, ctx->stream));

    int cs = env_int("MMG_GRAM_CLUSTER", 2);
    if (cs != 1) cs = 2;
    const int tiles_n = (n + TC_BN - 1) / TC_BN;
    std::vector<TcTile> tiles;
    int entries = 0;
    for (int jn = 0; jn < tiles_n; ++jn)
        for (int im = 0; im <= 2 * jn + 1; im += cs) {
            for (int k = 0; k < S; ++k) {
                TcTile tl{};
                tl.m0 = (int)((int64_t)(k + 1) * n_padM + (int64_t)im * TC_BM);
                tl.n0 = jn * TC_BN;
                tl.aux0 = k;
                tiles.push_back(tl);
            }
            ++entries;
        }
    IbdEpi::Params ep{};
    ep.Gw = Gw.as<double>();
    ep.ld = g_pad;
    ep.n_padM = n_padM;
    for (int k = 0; k < S; ++k) ep.w[k] = ldexp(1.0, E - 6 * (k + 1));
    double ibd_ms = 0.0;
    for (int64_t r0 = 0; r0 < count; r0 += chunk) {
        const int64_t cnt = std::min(chunk, count - r0);
        const int64_t kbytes = round_up(cnt, 128);
        pack_ibd_kernel<<<dim3((unsigned)((cnt + 127) / 128), (unsigned)((n + 63) / 64)), 256, 0, ctx->stream>>>(
            ctx->snps, ctx->pitch, d_rows + r0, cnt, n, dig.as<int8_t>() + r0, dig_pitch, S, P.as<int8_t>(), p_pitch, n_padM, ctx->flag_d);
        MMG_TRY(launch_check(ctx, "pack_ibd_kernel"));
        CUtensorMap tmA, tmB;
        MMG_TRY(make_tmap_u8(ctx, &tmA, P.p, kbytes, (int64_t)(S + 1) * n_padM, p_pitch, TC_BM));
        MMG_TRY(make_tmap_u8(ctx, &tmB, P.p, kbytes, n, p_pitch, TC_BN / cs));
        for (auto& t : tiles) { t.kb0 = 0; t.kb1 = (int)(kbytes / TC_BK); }
        MMG_TRY(ensure_tiles(ctx, tiles));
        cudaEventRecord(ctx->kev0, ctx->stream);
        if (cs == 2)
            MMG_TRY((launch_tc_gemm<IbdEpi, 2>(ctx, tmA, tmB, (const TcTile*)ctx->tiles_d, entries * 2, S, S, 0, TC_BM, ep, "tc_gemm_i8_kernel<IbdEpi,2>")));
        else
            MMG_TRY((launch_tc_gemm<IbdEpi, 1>(ctx, tmA, tmB, (const TcTile*)ctx->tiles_d, entries, S, S, 0, 0, ep, "tc_gemm_i8_kernel<IbdEpi,1>")));
        cudaEventRecord(ctx->kev1, ctx->stream);
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the tile table and P are rewritten by the next chunk
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        ibd_ms += ms;
    }
    ctx->last_ibd_ms = ibd_ms;
    MMG_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->flag_d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad & 2) {
        *domain_bad = 1;
        return MMG_OK;
    }
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
    ibd_finalize_add_kernel<<<grid, 256, 0, ctx->stream>>>(Gw.as<double>(), g_pad, n, d_u, d_c, K->d, K->cols);
    MMG_TRY(launch_check(ctx, "ibd_finalize_add_kernel"));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

extern "C" {

int mmg_kinship_ibd_accumulate_f64(mmg_ctx* ctx, mmg_mat K_acc, int64_t snp_begin, int64_t snp_count, const uint8_t* snp_mask,
                                   int64_t* used) {
    MmgMat* K = ctx ? get_mat(ctx, K_acc) : nullptr;
    MMG_CHECK(ctx, K && ctx->snps, "mmg_kinship_ibd_accumulate_f64: need resident genotypes and an accumulator");
    MMG_CHECK(ctx, K->rows == ctx->n && K->cols == ctx->n, "K_acc must be n x n");
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count >= 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm(ctx, "ibd");
    std::vector<long long> rows;
    rows.reserve(snp_count);
    for (int64_t s = 0; s < snp_count; ++s)
        if (!snp_mask || snp_mask[s]) rows.push_back(snp_begin + s);
    if (used) *used = (int64_t)rows.size();
    if (rows.empty()) return MMG_OK;
    // MMG_IBD_IMPL = tcgen05 (default; genotypes in {0,1,2}) | dsyrk (FP64 library GEMM, any int8 coding)
    const char* e = getenv("MMG_IBD_IMPL");
    bool use_tc = !(e && !strcmp(e, "dsyrk"));
    if (use_tc) {
        int domain_bad = 0;
        MMG_TRY(ibd_tc_run(ctx, K, rows, &domain_bad));
        if (!domain_bad) return MMG_OK;
    }
    return ibd_dsyrk_run(ctx, K, rows);
}

// ======================================================================================================
// stage 2: REML
// ======================================================================================================
int mmg_reml_f64(mmg_ctx* ctx, const double* eig_vals, const double* sq_etas, int64_t p, int64_t T, const double* deltas, int64_t g,
                 double esp, double* lls, double* dlls, double* opt_delta, double* opt_ll, int32_t* flags) {
    MMG_CHECK(ctx, ctx && eig_vals && sq_etas && deltas && p > 0 && T > 0 && g > 1, "mmg_reml_f64: bad argument");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm(ctx, "reml");
    const int64_t nd = p + T * p + g + 2 * T * g + 2 * T;
    MMG_TRY(ensure_scratch(ctx, nd * sizeof(double) + T * sizeof(int) + 64));
    double* d_eig = (double*)ctx->scratch;
    double* d_sq = d_eig + p;
    double* d_del = d_sq + T * p;
    double* d_lls = d_del + g;
    double* d_dlls = d_lls + T * g;
    double* d_od = d_dlls + T * g;
    double* d_ol = d_od + T;
    int* d_fl = (int*)(d_ol + T);
    MMG_CUDA(ctx, cudaMemcpyAsync(d_eig, eig_vals, p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(d_sq, sq_etas, T * p * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(d_del, deltas, g * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((unsigned)g, (unsigned)T);
    reml_grid_kernel<<<grid, REML_THREADS, 0, ctx->stream>>>(d_eig, d_sq, (int)p, d_del, (int)g, d_lls, d_dlls);
    MMG_TRY(launch_check(ctx, "reml_grid_kernel"));
    reml_refine_kernel<<<(unsigned)T, REML_THREADS, 0, ctx->stream>>>(d_eig, d_sq, (int)p, d_del, (int)g, esp, d_lls, d_dlls, d_od, d_ol,
                                                                     d_fl);
    MMG_TRY(launch_check(ctx, "reml_refine_kernel"));
    if (lls) MMG_CUDA(ctx, cudaMemcpyAsync(lls, d_lls, T * g * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (dlls) MMG_CUDA(ctx, cudaMemcpyAsync(dlls, d_dlls, T * g * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (opt_delta) MMG_CUDA(ctx, cudaMemcpyAsync(opt_delta, d_od, T * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (opt_ll) MMG_CUDA(ctx, cudaMemcpyAsync(opt_ll, d_ol, T * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (flags) MMG_CUDA(ctx, cudaMemcpyAsync(flags, d_fl, T * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

// ======================================================================================================
// stage 3: scan
// ======================================================================================================
}  // extern "C"

__global__ void means_from_sums_kernel(const long long* __restrict__ sums, int64_t count, double inv_n, double* __restrict__ mu) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) mu[i] = (double)sums[i] * inv_n;
}

// zero-padded copy of R: rows -> multiple of 128, cols -> multiple of 128
static int pad_matrix(mmg_ctx* ctx, const MmgMat* R, DevBuf& out, int64_t* rows_pad, int64_t* ld) {
    *rows_pad = round_up(R->rows, 128);
    *ld = round_up(R->cols, 128);
    MMG_CUDA(ctx, out.alloc(ctx->stream, (size_t)(*rows_pad) * (*ld) * sizeof(double)));
    MMG_CUDA(ctx, cudaMemsetAsync(out.p, 0, (size_t)(*rows_pad) * (*ld) * sizeof(double), ctx->stream));
    MMG_CUDA(ctx, cudaMemcpy2DAsync(out.p, (*ld) * sizeof(double), R->d, R->cols * sizeof(double), R->cols * sizeof(double), R->rows,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
    return MMG_OK;
}

static int launch_scan_dmma(mmg_ctx* ctx, bool perm, const ScanDmmaParams& prm) {
    const int64_t blocks = (prm.row_count + SD_BM - 1) / SD_BM;
    const int grid = (int)std::min<int64_t>(blocks, ctx->sm_count);
    cudaEventRecord(ctx->kev0, ctx->stream);
    if (perm)
        scan_dmma_kernel<true><<<grid, SD_THREADS, SD_SMEM_BYTES, ctx->stream>>>(prm);
    else
        scan_dmma_kernel<false><<<grid, SD_THREADS, SD_SMEM_BYTES, ctx->stream>>>(prm);
    MMG_TRY(launch_check(ctx, "scan_dmma_kernel"));
    cudaEventRecord(ctx->kev1, ctx->stream);
    return MMG_OK;
}

// max |x| over the resident genotype block (zero padding included), 16 bytes per thread per step
__global__ void snps_absmax_kernel(const uint4* __restrict__ p, int64_t n16, int* __restrict__ out) {
    int m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 w = p[i];
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int v = (int)(int8_t)((ww[j >> 2] >> (8 * (j & 3))) & 0xffu);
            m = max(m, v < 0 ? -v : v);
        }
    }
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// The int8 scan keeps per-tile sums in int32: |acc| <= 128 PKB |x| 128 and 32 columns x |x| per chain (2^28 at PKB = 8), safe for
// |x| <= QS_MAX_ABS_GENOTYPE (genotypes are 0/1/2, kinship.py:14-56).  Measured once per resident block.
constexpr int QS_MAX_ABS_GENOTYPE = 8;
static int scan_tc_check_domain(mmg_ctx* ctx) {
    if (ctx->snps_absmax < 0) {
        MMG_CUDA(ctx, cudaMemsetAsync(ctx->flag_d, 0, sizeof(int), ctx->stream));
        snps_absmax_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((const uint4*)ctx->snps, ctx->m * ctx->pitch / 16, ctx->flag_d);
        MMG_TRY(launch_check(ctx, "snps_absmax_kernel"));
        int v = 0;
        MMG_CUDA(ctx, cudaMemcpyAsync(&v, ctx->flag_d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->snps_absmax = v;
    }
    if (ctx->snps_absmax > QS_MAX_ABS_GENOTYPE)
        return fail(ctx, MMG_EVALUE, "int8 tensor-core scan: |genotype| up to %d exceeds its exact-integer domain (<= %d); "
                    "use the FP64 tensor-core path (scan_impl='dmma')", ctx->snps_absmax, QS_MAX_ABS_GENOTYPE);
    return MMG_OK;
}

// ---- int8 tensor-core scan: x'(R'R)x on exact integer slices (scan_tc.cuh) ------------------------------
// Number of digit planes.  MMG_TC_SLICES = k fixes it; otherwise it is chosen per call from the certified truncation
// bound  |d(x~.x~)| / x~.x~ <= (64/255) 256^-S 2^E ||x||_1^2 / x~.x~ <= MMG_TC_TOL (default 1e-7, i.e. < 2e-7 relative in
// -log10 p for r^2 <= 0.5): a pilot launch over the first SNPs measures max_s 2^E ||x||_1^2 / x~.x~, the full launch
// re-measures the bound over every SNP and is repeated with one more plane if a SNP violates it.
constexpr int QS_AUTO_PLANES = 6;          // planes cut in auto mode: 48 bits of B (bound <= (64/255) 256^-6 ~ 9e-16 2^E ||x||_1^2)
constexpr int QS_PILOT_PLANES = 2;
constexpr double QS_PILOT_HEADROOM = 3.0;  // for the rows the pilot did not see; the full launch re-checks every SNP anyway
constexpr int64_t QS_PILOT_ROWS = 64 * TC_BM;

static int scan_tc_fixed_slices() {
    const char* e = getenv("MMG_TC_SLICES");
    if (!e) return 0;
    return std::max(1, std::min(QS_MAX_SLICES, atoi(e)));
}
static double scan_tc_tol() {
    const char* e = getenv("MMG_TC_TOL");
    const double t = e ? atof(e) : 1e-7;
    return t > 0.0 ? t : 1e-7;
}

// Digit planes of the strict lower triangle of A = R'R (doubled) into Bq (S planes of [n_padN x ldq]), diag(A) into
// d_dg and v = R'y into d_v; returns the binary exponent E used for the scaling.  `A` is an n x n FP64 work matrix.
// A = R'R as exact int8 digit-plane products on the tensor cores (scan_tc.cuh, "A = R'R on the int8 tensor cores"); A_work is a
// zero-filled [n_padM x n_padM] FP64 buffer whose lower-triangular tiles are written.  *err_out: absolute error bound of its entries.
static int quad_form_int8(mmg_ctx* ctx, const MmgMat* R, double* A_work, int64_t n_padM, unsigned long long* d_amax, double* err_out) {
    const int64_t n = ctx->n, n_out = R->rows;
    MMG_CHECK(ctx, n_out < 131072, "R'R on the int8 pipe: contraction too long for exact int32 accumulation");
    MMG_CUDA(ctx, cudaMemsetAsync(d_amax, 0, sizeof(unsigned long long), ctx->stream));
    mat_amax_kernel<<<dim3(8, (unsigned)n_out), 256, 0, ctx->stream>>>(R->d, n, (int)n_out, (int)n, d_amax);
    MMG_TRY(launch_check(ctx, "mat_amax_kernel"));
    double rmax = 0.0;
    MMG_CUDA(ctx, cudaMemcpyAsync(&rmax, d_amax, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!std::isfinite(rmax)) return fail(ctx, MMG_EVALUE, "scan: the rotation is not finite (max |r| = %g)", rmax);
    const int F = digit256_exponent(rmax);
    const int64_t op_pitch = round_up(n_out, TC_BK);
    DevBuf Op;
    MMG_CUDA(ctx, Op.alloc(ctx->stream, (size_t)OZ_PLANES * n_padM * op_pitch));
    MMG_CUDA(ctx, cudaMemsetAsync(Op.p, 0, (size_t)OZ_PLANES * n_padM * op_pitch, ctx->stream));
    ozaki_planes_kernel<<<dim3((unsigned)((n + 31) / 32), (unsigned)((n_out + 31) / 32)), 256, 0, ctx->stream>>>(
        R->d, n, (int)n_out, (int)n, ldexp(1.0, -F), Op.as<int8_t>(), n_padM, op_pitch);
    MMG_TRY(launch_check(ctx, "ozaki_planes_kernel"));
    MMG_CUDA(ctx, cudaMemsetAsync(A_work, 0, (size_t)n_padM * n_padM * sizeof(double), ctx->stream));
    // tile table: row-tile pairs (im, im + 1) x column tile jn that meet the lower triangle (rows >= columns), 28 (p, q) planes each
    const int tiles_n = (int)(n_padM / TC_BN), tiles_m = (int)(n_padM / TC_BM), KB = (int)(op_pitch / TC_BK);
    // Entry order = order in which the co-resident clusters pick the output tiles up.  The operands are streamed from L2 / HBM
    // by every tile (79 K-blocks only), so the clusters of one wave should share them: the lower-triangular (row-pair, column
    // tile) grid is walked in blocks of 9 x 8 (72 ~ the 74 co-resident clusters), whose operand rows for one plane pair are
    // 9 x 2.6 + 8 x 2.6 = 44 MB -- L2 resident -- instead of row by row (a whole 103 MB plane per tile step).
    std::vector<TcTile> tiles;
    std::vector<std::pair<int, int>> order;                     // (jn, im)
    {
        const int BI = std::max(1, env_int("MMG_OZAKI_BLOCK_I", 9)), BJ = std::max(1, env_int("MMG_OZAKI_BLOCK_J", 8));
        const int pairs_m = tiles_m / 2;
        for (int bi = 0; bi < pairs_m; bi += BI)
            for (int bj = 0; bj < tiles_n; bj += BJ)
                for (int ip = bi; ip < std::min(pairs_m, bi + BI); ++ip)
                    for (int jn = bj; jn < std::min(tiles_n, bj + BJ); ++jn)
                        if (ip >= jn) order.emplace_back(jn, 2 * ip);
    }
    int entries = 0, per_entry = 0;
    const bool chain = env_int("MMG_OZAKI_CHAIN", 1) != 0 && (double)OZ_LEVELS * (double)n_out * 16384.0 < 2147483647.0;
    for (const auto& ji : order) {
        const int jn = ji.first, im = ji.second;
        {
            // level by level (p + q = lv share the weight 2^2F 256^-(lv+2)): the lv + 1 plane pairs of a level are CHAINED into one
            // int32 accumulator (7 n_out 128^2 < 2^31), so an output tile is read-modify-written 7 times, not 28 -- the FP64
            // read-modify-write of the epilogue (a row per thread, 32 sectors per access) was what bounded this kernel
            per_entry = 0;
            for (int lv = 0; lv < OZ_LEVELS; ++lv)
                for (int p = 0; p <= lv; ++p) {
                    const int q = lv - p;
                    if (p >= OZ_PLANES || q >= OZ_PLANES) continue;
                    TcTile tl{};
                    tl.m0 = (int)((int64_t)p * n_padM + (int64_t)im * TC_BM);
                    tl.n0 = (int)((int64_t)q * n_padM + (int64_t)jn * TC_BN);
                    tl.kb0 = 0;
                    tl.kb1 = KB;
                    tl.aux0 = p;
                    tl.aux1 = q;
                    tl.flags = (p < lv && chain) ? TC_TILE_CHAIN : 0;       // the pair (lv, 0) ends the chain of its level
                    tiles.push_back(tl);
                    ++per_entry;
                }
            ++entries;
        }
    }
    MMG_TRY(ensure_tiles(ctx, tiles));
    CUtensorMap tmA, tmB;
    MMG_TRY(make_tmap_u8(ctx, &tmA, Op.p, op_pitch, (int64_t)OZ_PLANES * n_padM, op_pitch, TC_BM));
    MMG_TRY(make_tmap_u8(ctx, &tmB, Op.p, op_pitch, (int64_t)OZ_PLANES * n_padM, op_pitch, TC_BN / 2));
    OzakiEpi::Params ep{};
    ep.A = A_work;
    ep.ld = n_padM;
    ep.n_padM = n_padM;
    for (int sl = 0; sl < 2 * OZ_PLANES; ++sl) ep.w[sl] = ldexp(1.0, 2 * F - 8 * (sl + 2));
    MMG_TRY((launch_tc_gemm<OzakiEpi, 2>(ctx, tmA, tmB, (const TcTile*)ctx->tiles_d, entries * 2, per_entry, per_entry, 0, TC_BM, ep,
                                         "tc_gemm_i8_kernel<OzakiEpi,2>")));
    *err_out = ozaki_error_bound(n_out) * ldexp(1.0, 2 * F);   // (Op is released in stream order when this returns)
    return MMG_OK;
}

// MMG_QUAD_A = int8 (default: exact digit-plane products on the int8 tensor pipe) | dsyrk (cuBLAS, FP64 tensor pipe)
static thread_local bool g_quad_force_dsyrk = false;     // set while a scan is repeated with the FP64 product (see scan_tc_run)
static bool quad_a_int8() {
    if (g_quad_force_dsyrk) return false;
    const char* e = getenv("MMG_QUAD_A");
    return !(e && strcmp(e, "dsyrk") == 0);
}

static int quad_prepare(mmg_ctx* ctx, const MmgMat* R, const double* d_y, const double* A_given, double* A_work, int64_t lda_work,
                        unsigned long long* d_amax, int S, int8_t* Bq, int64_t n_padN, int64_t ldq, double* d_v, double* d_dg, int* E_out,
                        double* errA_out) {
    const int64_t n = ctx->n;
    const double one = 1.0, zero = 0.0;
    const double* A = A_given;
    int64_t lda = n;
    *errA_out = 0.0;
    if (!A_given) {
        const int64_t n_out = R->rows;
        if (quad_a_int8() && scan_tc_tol() >= 1e-9) {
            MMG_TRY(quad_form_int8(ctx, R, A_work, lda_work, d_amax, errA_out));
        } else {
            // A = R'R: R row-major [n_out x n] is the column-major n x n_out matrix Rc; column-major UPPER of Rc Rc'
            // is the row-major LOWER triangle A[j][i], i <= j -- exactly the operand the slices are cut from.
            MMG_CUBLAS(ctx, cublasDsyrk(ctx->cublas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, (int)n, (int)n_out, &one, R->d, (int)n, &zero, A_work,
                                        (int)lda_work));
        }
        // v = R' y~  (x~.y~ = x.v)
        if (d_y) MMG_CUBLAS(ctx, cublasDgemv(ctx->cublas, CUBLAS_OP_N, (int)n, (int)n_out, &one, R->d, (int)n, d_y, 1, &zero, d_v, 1));
        A = A_work;
        lda = lda_work;
    }
    MMG_CUDA(ctx, cudaMemsetAsync(d_amax, 0, sizeof(unsigned long long), ctx->stream));
    dim3 agrid(8, (unsigned)n);
    quad_amax_kernel<<<agrid, 256, 0, ctx->stream>>>(A, lda, (int)n, d_amax);
    MMG_TRY(launch_check(ctx, "quad_amax_kernel"));
    double amax = 0.0;
    MMG_CUDA(ctx, cudaMemcpyAsync(&amax, d_amax, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!std::isfinite(amax)) return fail(ctx, MMG_EVALUE, "scan: R'R is not finite (max |a| = %g)", amax);
    const int E = digit256_exponent(amax);                // |2 a| 2^-E <= 0.498 (a diagonal R'R has no off-diagonal digits at all)
    dim3 sgrid((unsigned)((n + 255) / 256), (unsigned)n);
    quad_slice_kernel<<<sgrid, 256, 0, ctx->stream>>>(A, lda, (int)n, ldexp(1.0, -E), S, Bq, n_padN, ldq, d_dg);
    MMG_TRY(launch_check(ctx, "quad_slice_kernel"));
    *E_out = E;
    return MMG_OK;
}

// one launch of the quadratic-form scan over resident rows [snp_begin, +snp_count) with S of the S_alloc cut planes
static int scan_tc_launch(mmg_ctx* ctx, int T, int S, int S_alloc, const void* Bq, int64_t n_padN, int64_t ldq, int64_t snp_begin,
                          int64_t snp_count, QuadEpi::Params ep, unsigned* d_wave_sync) {
    int cs = env_int("MMG_SCAN_CLUSTER", 2);
    if (cs != 1 && cs != 2 && cs != 4 && cs != 8) cs = 2;
    // MMG_SCAN_SCHED = panel (genotype-stationary schedule, scan_quad.cuh) | pair (same, MMA as a CTA pair) |
    //                  pair128 / n128 (128-column tiles, four accumulator stages) | table (tile-table kernel, tc_gemm.cuh)
    const char* sched = getenv("MMG_SCAN_SCHED");
    if (!sched) sched = "pair";         // the CTA-pair MMA halves the L2 -> SM digit traffic per SM: 138 vs 145 ms per 1M SNPs at n = 10k
    const bool pair128 = strcmp(sched, "pair128") == 0, n128 = strcmp(sched, "n128") == 0;
    const bool pair = strcmp(sched, "pair") == 0 || pair128;
    if (pair || n128) cs = 2;
    if (cs == 8 && strcmp(sched, "table") == 0) cs = 4;      // the tile-table kernel is instantiated for clusters of 1, 2, 4
    const int bn = (pair128 || n128) ? 128 : TC_BN;
    const int tiles_n = (int)(n_padN / bn), kb_total = (int)(ldq / TC_BK);
    ep.row_begin = snp_begin;
    ep.row_count = snp_count;
    ep.out_stride = snp_count;
    CUtensorMap tmA, tmB;
    MMG_TRY(make_tmap_u8(ctx, &tmA, ctx->snps + snp_begin * ctx->pitch, ctx->pitch, snp_count, ctx->pitch, TC_BM));
    MMG_TRY(make_tmap_u8(ctx, &tmB, Bq, ldq, (int64_t)T * S_alloc * n_padN, ldq, bn / cs));
    const int groups = (int)((snp_count + TC_BM - 1) / TC_BM);
    if (strcmp(sched, "table") != 0) {
        QuadShape sh{};
        sh.num_groups = groups;
        sh.T = T;
        sh.S = S;
        sh.S_stride = S_alloc;
        sh.tiles_n = tiles_n;
        sh.kb_total = kb_total;
        sh.n_padN = (int)n_padN;
        sh.prefetch = std::max(0, env_int("MMG_SCAN_PREFETCH", 8));
        sh.pf_share = std::max(1, env_int("MMG_SCAN_PF_SHARE", 1));
        if (env_int("MMG_SCAN_WAVE_SYNC", 1) && d_wave_sync) {
            MMG_CUDA(ctx, cudaMemsetAsync(d_wave_sync, 0, sizeof(unsigned), ctx->stream));
            sh.wave_sync = d_wave_sync;
        }
        // MMG_SCAN_DBG_CLOCKS=<file>: per-CTA cycle counters of the three warp roles (time spent in each barrier wait)
        const char* dbg_path = getenv("MMG_SCAN_DBG_CLOCKS");
        DevBuf dbg;
        const int dbg_ctas = ctx->sm_count;
        if (dbg_path) {
            MMG_CUDA(ctx, dbg.alloc(ctx->stream, (size_t)dbg_ctas * 16 * sizeof(long long)));
            MMG_CUDA(ctx, cudaMemsetAsync(dbg.p, 0, (size_t)dbg_ctas * 16 * sizeof(long long), ctx->stream));
            sh.dbg = dbg.as<long long>();
        }
        const int panel = env_int("MMG_SCAN_PANEL", 8);
        if (pair128) MMG_TRY(launch_scan_quad_pair128(ctx, panel, tmA, tmB, sh, ep));
        else if (n128) MMG_TRY(launch_scan_quad_n128(ctx, panel, tmA, tmB, sh, ep));
        else if (pair) MMG_TRY(launch_scan_quad_pair(ctx, panel, tmA, tmB, sh, ep));
        else if (cs == 8) MMG_TRY(launch_scan_quad_cs<8>(ctx, panel, tmA, tmB, sh, ep));
        else if (cs == 4) MMG_TRY(launch_scan_quad_cs<4>(ctx, panel, tmA, tmB, sh, ep));
        else if (cs == 2) MMG_TRY(launch_scan_quad_cs<2>(ctx, panel, tmA, tmB, sh, ep));
        else MMG_TRY(launch_scan_quad_cs<1>(ctx, panel, tmA, tmB, sh, ep));
        if (dbg_path) {
            std::vector<long long> h((size_t)dbg_ctas * 16);
            MMG_CUDA(ctx, cudaMemcpyAsync(h.data(), dbg.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (FILE* f = fopen(dbg_path, "w")) {
                fprintf(f, "# cta prod_total prod_wait_empty prod_wait_aempty - mma_total mma_wait_full mma_wait_tempty mma_wait_afull epi_total epi_wait_tfull epi_xload epi_fp64 epi_drain\n");
                for (int c = 0; c < dbg_ctas; ++c) {
                    fprintf(f, "%d", c);
                    for (int k = 0; k < 13; ++k) fprintf(f, " %lld", h[(size_t)c * 16 + k]);
                    fprintf(f, "\n");
                }
                fclose(f);
            }
        }
    } else {
        // one shared tile table: for every phenotype, for every 256-column tile jb, one tile per slice; K only up to the diagonal
        std::vector<TcTile> tiles;
        tiles.reserve((size_t)T * tiles_n * S);
        for (int t = 0; t < T; ++t)
            for (int jb = 0; jb < tiles_n; ++jb)
                for (int k = 0; k < S; ++k) {
                    TcTile tl{};
                    tl.m0 = 0;
                    tl.n0 = (int)(((int64_t)t * S_alloc + k) * n_padN + (int64_t)jb * TC_BN);
                    tl.kb0 = 0;
                    tl.kb1 = std::min(kb_total, (jb + 1) * (TC_BN / TC_BK));
                    tl.aux0 = k;
                    tl.aux1 = (t << QS_PHEN_SHIFT) | (k == 0 ? QS_FLAG_XY : 0) | ((jb == 0 && k == 0) ? QS_FLAG_FIRST : 0) |
                              ((jb == tiles_n - 1 && k == S - 1) ? QS_FLAG_LAST : 0);
                    tl.col0 = jb * TC_BN;
                    tiles.push_back(tl);
                }
        MMG_TRY(ensure_tiles(ctx, tiles));
        const TcTile* td = (const TcTile*)ctx->tiles_d;
        if (cs == 4)
            MMG_TRY((launch_tc_gemm<QuadEpi, 4>(ctx, tmA, tmB, td, groups, (int)tiles.size(), 0, TC_BM, 0, ep, "tc_gemm_i8_kernel<QuadEpi,4>", L2_EVICT_FIRST, L2_EVICT_LAST)));
        else if (cs == 2)
            MMG_TRY((launch_tc_gemm<QuadEpi, 2>(ctx, tmA, tmB, td, groups, (int)tiles.size(), 0, TC_BM, 0, ep, "tc_gemm_i8_kernel<QuadEpi,2>", L2_EVICT_FIRST, L2_EVICT_LAST)));
        else
            MMG_TRY((launch_tc_gemm<QuadEpi, 1>(ctx, tmA, tmB, td, groups, (int)tiles.size(), 0, TC_BM, 0, ep, "tc_gemm_i8_kernel<QuadEpi,1>", L2_EVICT_FIRST, L2_EVICT_LAST)));
    }
    return MMG_OK;
}

// T phenotypes (each with its own rotation R_t, residual y~_t and h0_rss_t) in one launch.  Device outputs are
// [T][snp_count]; any may be NULL.
//
// A_given / v_given (T = 1 only): the quadratic form A = R'R (row-major lower triangle valid) and v = R'y~ were formed by
// the caller -- the multi-GPU path builds A from per-rank row blocks of R and an all-reduce (parallel.py) instead of
// repeating the n^3 product on every rank; Rs and V are then unused.
static int scan_tc_run(mmg_ctx* ctx, int T, const MmgMat* const* Rs, const double* V, const double* h0_rss, double n_p, double lbeta,
                       int64_t snp_begin, int64_t snp_count, double* d_xx, double* d_xy, double* d_rss, double* d_f, double* d_p,
                       double* d_vp, const MmgMat* A_given = nullptr, const double* v_given = nullptr) {
    const int64_t n = ctx->n, n_out = A_given ? 1 : Rs[0]->rows;
    MMG_TRY(scan_tc_check_domain(ctx));
    const int S_fixed = scan_tc_fixed_slices();
    const int S_alloc = S_fixed ? S_fixed : QS_AUTO_PLANES;
    const double tol = scan_tc_tol();
    const int64_t n_padN = round_up(n, TC_BN), ldq = round_up(n, TC_BK);
    const int64_t plane = n_padN * ldq;
    MMG_CHECK(ctx, (int64_t)T * S_alloc * n_padN < (1ll << 31), "scan: too many phenotype slices for one launch");
    DevBuf A, Bq, vec;
    const int64_t lda_work = round_up(n, TC_BN);               // padded so that the int8 R'R epilogue needs no bounds checks
    if (!A_given) MMG_CUDA(ctx, A.alloc(ctx->stream, (size_t)lda_work * lda_work * sizeof(double)));
    MMG_CUDA(ctx, Bq.alloc(ctx->stream, (size_t)T * S_alloc * plane));
    // vec: v[T][n_padN] | dg[T][n_padN] | y[T][n_out] | h0[T] | escale[T] | bscale[T] | amax | rho | wave counter
    const int64_t nd = 2 * (int64_t)T * n_padN + (int64_t)T * n_out + 3 * T + 3;
    MMG_CUDA(ctx, vec.alloc(ctx->stream, (size_t)nd * sizeof(double)));
    double* d_v = vec.as<double>();
    double* d_dg = d_v + (int64_t)T * n_padN;
    double* d_y = d_dg + (int64_t)T * n_padN;
    double* d_h0 = d_y + (int64_t)T * n_out;
    double* d_es = d_h0 + T;
    double* d_bs = d_es + T;
    unsigned long long* d_amax = (unsigned long long*)(d_bs + T);
    unsigned long long* d_rho = d_amax + 1;
    unsigned* d_wave = (unsigned*)(d_rho + 1);
    MMG_CUDA(ctx, cudaMemsetAsync(vec.p, 0, (size_t)nd * sizeof(double), ctx->stream));
    MMG_CUDA(ctx, cudaMemsetAsync(Bq.p, 0, (size_t)T * S_alloc * plane, ctx->stream));
    if (A_given) MMG_CUDA(ctx, cudaMemcpyAsync(d_v, v_given, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    else MMG_CUDA(ctx, cudaMemcpyAsync(d_y, V, (size_t)T * n_out * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(d_h0, h0_rss, (size_t)T * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<double> escale((size_t)T), bscale((size_t)T), errA((size_t)T, 0.0);
    // linear pre-pass: x.v_t, sum_j A_jj x_j^2 and ||x||_1 of every SNP in range, one stream over the genotypes.  When the rotation
    // is at hand, v_t = R_t'y~_t and diag(A_t) = column sums of squares of R_t are formed first and the pre-pass runs on a side
    // stream underneath the n^3 product A = R'R (FP64 / HBM work beside int8 tensor work); MMG_SCAN_OVERLAP=0 serialises them.
    DevBuf pre;
    MMG_CUDA(ctx, pre.alloc(ctx->stream, (size_t)((2 * T + 1) * snp_count + (int64_t)T * n_padN) * sizeof(double)));
    double* d_dg_pre = pre.as<double>();                        // first: read as double2 by the pre-pass (16-byte aligned)
    double* p_xy = d_dg_pre + (int64_t)T * n_padN;
    double* p_qd = p_xy + (int64_t)T * snp_count;
    double* p_a1 = p_qd + (int64_t)T * snp_count;
    const int pre_rows_per_block = 8 * PRE_ROWS;
    const unsigned pre_grid = (unsigned)((snp_count + pre_rows_per_block - 1) / pre_rows_per_block);
    bool pre_launched = false;
    struct SideJoin {                      // an early error return must not release `pre` under a running side-stream kernel
        cudaStream_t s = nullptr;
        ~SideJoin() { if (s) cudaStreamSynchronize(s); }
    } side_join;
    if (!A_given && env_int("MMG_SCAN_OVERLAP", 1)) {
        if (!ctx->stream2) {
            MMG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
            MMG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ov0, cudaEventDisableTiming));
            MMG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ov1, cudaEventDisableTiming));
        }
        MMG_CUDA(ctx, cudaMemsetAsync(d_dg_pre, 0, (size_t)T * n_padN * sizeof(double), ctx->stream));
        const double one = 1.0, zero = 0.0;
        for (int t = 0; t < T; ++t) {
            const MmgMat* R = Rs[t];
            MMG_CUBLAS(ctx, cublasDgemv(ctx->cublas, CUBLAS_OP_N, (int)n, (int)R->rows, &one, R->d, (int)n, d_y + (int64_t)t * n_out, 1, &zero,
                                        d_v + (int64_t)t * n_padN, 1));
            col_sumsq_kernel<<<(unsigned)((n + 31) / 32), 256, 0, ctx->stream>>>(R->d, n, (int)R->rows, (int)n, d_dg_pre + (int64_t)t * n_padN);
            MMG_TRY(launch_check(ctx, "col_sumsq_kernel"));
        }
        MMG_CUDA(ctx, cudaEventRecord(ctx->ov0, ctx->stream));
        MMG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ov0, 0));
        snp_prepass_kernel<PRE_ROWS, 4, 4><<<pre_grid, 256, 0, ctx->stream2>>>(ctx->snps, ctx->pitch, snp_begin, snp_count, T, d_v, d_dg_pre, n_padN, p_xy,
                                                                            p_qd, p_a1, snp_count);
        MMG_TRY(launch_check(ctx, "snp_prepass_kernel"));
        side_join.s = ctx->stream2;
        MMG_CUDA(ctx, cudaEventRecord(ctx->ov1, ctx->stream2));
        pre_launched = true;
    }
    for (int t = 0; t < T; ++t) {
        int E = 0;
        MMG_TRY(quad_prepare(ctx, A_given ? nullptr : Rs[t], pre_launched ? nullptr : d_y + (int64_t)t * n_out, A_given ? A_given->d : nullptr, A.as<double>(), lda_work,
                             d_amax, S_alloc, Bq.as<int8_t>() + (int64_t)t * S_alloc * plane, n_padN, ldq, d_v + (int64_t)t * n_padN,
                             d_dg + (int64_t)t * n_padN, &E, &errA[(size_t)t]));
        escale[t] = ldexp(1.0, E);
    }
    MMG_CUDA(ctx, cudaMemcpyAsync(d_es, escale.data(), (size_t)T * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));

    QuadEpi::Params ep{};
    ep.snps = ctx->snps;
    ep.pitch = ctx->pitch;
    for (int k = 0; k < QS_MAX_SLICES; ++k) ep.w[k] = ldexp(1.0, -8 * (k + 1));
    ep.escale = d_es;
    ep.v = d_v;
    ep.dg = d_dg;
    ep.bscale = d_bs;
    ep.rho_max = d_rho;
    ep.v_stride = n_padN;
    ep.h0_rss = d_h0;
    if (pre_launched) {
        MMG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ov1, 0));        // join the side stream
        side_join.s = nullptr;
    } else {
        snp_prepass_kernel<PRE_ROWS, 4, 4><<<pre_grid, 256, 0, ctx->stream>>>(ctx->snps, ctx->pitch, snp_begin, snp_count, T, d_v, d_dg, n_padN, p_xy, p_qd,
                                                                           p_a1, snp_count);
        MMG_TRY(launch_check(ctx, "snp_prepass_kernel"));
    }
    ep.pre_xy = p_xy;
    ep.pre_qd = p_qd;
    ep.pre_a1 = p_a1;
    ep.pre_stride = snp_count;
    ep.n_p = n_p;
    ep.lbeta = lbeta;

    // bound scale for S planes: |remainder| <= 128/255 per entry of B, sum_{i<j} |x_i||x_j| <= ||x||_1^2 / 2
    auto set_bscale = [&](int S) -> int {
        // + the entry-wise error bound of A itself when it came from the int8 digit-plane product: |x'(dA)x| <= errA ||x||_1^2
        for (int t = 0; t < T; ++t) bscale[t] = 0.5 * DIGIT256_REM * ldexp(1.0, -8 * S) * escale[t] + errA[(size_t)t];
        MMG_CUDA(ctx, cudaMemcpyAsync(d_bs, bscale.data(), (size_t)T * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MMG_CUDA(ctx, cudaMemsetAsync(d_rho, 0, sizeof(unsigned long long), ctx->stream));
        return MMG_OK;
    };
    auto read_rho = [&](double* rho) -> int {
        MMG_CUDA(ctx, cudaMemcpyAsync(rho, d_rho, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return MMG_OK;
    };

    int S = S_fixed ? S_fixed : S_alloc;                        // short scans: no pilot, every plane that was cut
    if (!S_fixed && snp_count >= 4 * QS_PILOT_ROWS) {
        // pilot: bound of the first rows with few planes; the bound scales exactly by 256 per plane
        QuadEpi::Params pp = ep;
        pp.xx = pp.xy = pp.rss = pp.f = pp.p = pp.var_perc = nullptr;
        MMG_TRY(set_bscale(QS_PILOT_PLANES));
        MMG_TRY(scan_tc_launch(ctx, T, QS_PILOT_PLANES, S_alloc, Bq.p, n_padN, ldq, snp_begin, QS_PILOT_ROWS, pp, d_wave));
        double rho = 0.0;
        MMG_TRY(read_rho(&rho));
        // bound(S) / bound(pilot planes), worst phenotype: 256 per plane down to the floor set by the error of A itself
        auto bound_ratio = [&](int planes) {
            double r = 0.0;
            for (int t = 0; t < T; ++t) {
                const double c = 0.5 * DIGIT256_REM * escale[t];
                r = std::max(r, (c * ldexp(1.0, -8 * planes) + errA[(size_t)t]) / (c * ldexp(1.0, -8 * QS_PILOT_PLANES) + errA[(size_t)t]));
            }
            return r;
        };
        S = QS_PILOT_PLANES;
        while (S < S_alloc && rho * bound_ratio(S) * QS_PILOT_HEADROOM > tol) ++S;
    }
    double rho = 0.0;
    for (;;) {
        ep.xx = d_xx; ep.xy = d_xy; ep.rss = d_rss; ep.f = d_f; ep.p = d_p; ep.var_perc = d_vp;
        MMG_TRY(set_bscale(S));
        cudaEventRecord(ctx->kev0, ctx->stream);
        MMG_TRY(scan_tc_launch(ctx, T, S, S_alloc, Bq.p, n_padN, ldq, snp_begin, snp_count, ep, d_wave));
        cudaEventRecord(ctx->kev1, ctx->stream);
        MMG_TRY(read_rho(&rho));                                // also: Bq / A / vec are freed on return
        if (S_fixed || rho <= tol || S >= S_alloc) break;
        ++S;                                                    // a SNP outside the certified bound: one more plane, again
    }
    ctx->last_scan_slices = S;
    ctx->last_scan_rho = rho;
    // a tolerance below what the int8 product of A can certify (its own error bound is a floor of ~1e-10 relative at n = 10k):
    // once more with A from the FP64 dsyrk
    bool int8_floor = false;
    for (double e : errA) int8_floor |= e > 0.0;
    if (!S_fixed && rho > tol && int8_floor && !g_quad_force_dsyrk) {
        g_quad_force_dsyrk = true;
        const int rc = scan_tc_run(ctx, T, Rs, V, h0_rss, n_p, lbeta, snp_begin, snp_count, d_xx, d_xy, d_rss, d_f, d_p, d_vp, A_given, v_given);
        g_quad_force_dsyrk = false;
        return rc;
    }
    return MMG_OK;
}

// ---- int8 tensor-core permutation scan (linear_models.py:1157-1164) --------------------------------------------
//   pass 1: xx_s = x_c'(R'R)x_c through the quadratic-form scan of the centred rotation R C (C = I - 11'/n)
//   pass 2: PermEpi GEMM of the genotype block with the 8 digit planes of W' = Ys'R ([P x n]),
//           ratio_p = max_s (x_c.W_p)^2 / xx_s
static int perm_scan_tc(mmg_ctx* ctx, const MmgMat* R, const MmgMat* Wt, int centre, int64_t snp_begin, int64_t snp_count,
                        double* ratio_inout) {
    StageTimer tm(ctx, "scan");
    const int64_t n = ctx->n, n_out = R->rows, P = Wt->rows;
    const int64_t P_pad = round_up(P, 32), ldq = round_up(n, TC_BK);
    const double one = 1.0, zero = 0.0;
    DevBuf Rc, aux, Wq;
    // aux: ones[n] | r1[n_out] | wsum[P_pad] | mu[snp_count] | xx[snp_count] | ratio[P_pad] | amax | sums[snp_count]
    const int64_t nd = n + n_out + P_pad + 2 * snp_count + P_pad + 1;
    MMG_CUDA(ctx, aux.alloc(ctx->stream, (size_t)nd * sizeof(double) + (size_t)snp_count * sizeof(long long)));
    MMG_CUDA(ctx, cudaMemsetAsync(aux.p, 0, (size_t)nd * sizeof(double), ctx->stream));
    double* d_ones = aux.as<double>();
    double* d_r1 = d_ones + n;
    double* d_wsum = d_r1 + n_out;
    double* d_mu = d_wsum + P_pad;
    double* d_xx = d_mu + snp_count;
    unsigned long long* d_ratio = (unsigned long long*)(d_xx + snp_count);
    unsigned long long* d_amax = d_ratio + P_pad;
    long long* d_sums = (long long*)(d_amax + 1);
    {
        std::vector<double> ones((size_t)n, 1.0);
        MMG_CUDA(ctx, cudaMemcpyAsync(d_ones, ones.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // row-major [rows x n] matrices are column-major n x rows: y = A' x gives the row sums
    MMG_CUBLAS(ctx, cublasDgemv(ctx->cublas, CUBLAS_OP_T, (int)n, (int)P, &one, Wt->d, (int)n, d_ones, 1, &zero, d_wsum, 1));
    MmgMat Rcm = *R;
    if (centre) {
        MMG_CUDA(ctx, Rc.alloc(ctx->stream, (size_t)n_out * n * sizeof(double)));
        MMG_CUDA(ctx, cudaMemcpyAsync(Rc.p, R->d, (size_t)n_out * n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        MMG_CUBLAS(ctx, cublasDgemv(ctx->cublas, CUBLAS_OP_T, (int)n, (int)n_out, &one, R->d, (int)n, d_ones, 1, &zero, d_r1, 1));
        dim3 cgrid((unsigned)((n + 255) / 256), (unsigned)n_out);
        centre_cols_kernel<<<cgrid, 256, 0, ctx->stream>>>(Rc.as<double>(), n, (int)n_out, (int)n, d_r1, 1.0 / (double)n);
        MMG_TRY(launch_check(ctx, "centre_cols_kernel"));
        Rcm.d = Rc.as<double>();
        snp_row_sums_kernel<<<(unsigned)((snp_count + 7) / 8), 256, 0, ctx->stream>>>(ctx->snps + snp_begin * ctx->pitch, ctx->pitch,
                                                                                      snp_count, (int)n, d_sums, nullptr);
        MMG_TRY(launch_check(ctx, "snp_row_sums_kernel"));
        means_from_sums_kernel<<<(unsigned)((snp_count + 255) / 256), 256, 0, ctx->stream>>>(d_sums, snp_count, 1.0 / (double)n, d_mu);
        MMG_TRY(launch_check(ctx, "means_from_sums_kernel"));
    }
    // pass 1
    {
        std::vector<double> y0((size_t)n_out, 0.0);
        const double h0 = 1.0;
        const MmgMat* Rs[1] = {&Rcm};
        MMG_TRY(scan_tc_run(ctx, 1, Rs, y0.data(), &h0, 1.0, 0.0, snp_begin, snp_count, d_xx, nullptr, nullptr, nullptr, nullptr, nullptr));
    }
    // digit planes of W'
    mat_amax_kernel<<<dim3(8, (unsigned)P), 256, 0, ctx->stream>>>(Wt->d, n, (int)P, (int)n, d_amax);
    MMG_TRY(launch_check(ctx, "mat_amax_kernel"));
    double amax = 0.0;
    MMG_CUDA(ctx, cudaMemcpyAsync(&amax, d_amax, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!std::isfinite(amax)) return fail(ctx, MMG_EVALUE, "permutation scan: W is not finite");
    const int E = amax > 0.0 ? ilogb(amax) + 2 : 0;
    const int64_t wq_rows = P_pad * PS_SLICES;
    MMG_CUDA(ctx, Wq.alloc(ctx->stream, (size_t)wq_rows * ldq));
    MMG_CUDA(ctx, cudaMemsetAsync(Wq.p, 0, (size_t)wq_rows * ldq, ctx->stream));
    perm_slice_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)P), 256, 0, ctx->stream>>>(Wt->d, n, (int)P, (int)n, ldexp(1.0, -E),
                                                                                                Wq.as<int8_t>(), ldq);
    MMG_TRY(launch_check(ctx, "perm_slice_kernel"));
    // pass 2
    PermEpi::Params ep{};
    ep.row_count = snp_count;
    for (int k = 0; k < PS_SLICES; ++k) ep.w[k] = ldexp(1.0, E - 7 * (k + 1));
    ep.xx = d_xx;
    ep.mu = centre ? d_mu : nullptr;
    ep.wsum = d_wsum;
    ep.ratio = d_ratio;
    std::vector<TcTile> tiles;
    for (int b = 0; b < (int)(P_pad / 32); ++b) {
        TcTile tl{};
        tl.n0 = b * TC_BN;
        tl.kb0 = 0;
        tl.kb1 = (int)(ldq / TC_BK);
        tl.col0 = b * 32;
        tiles.push_back(tl);
    }
    MMG_TRY(ensure_tiles(ctx, tiles));
    int cs = env_int("MMG_SCAN_CLUSTER", 2);
    if (cs != 1 && cs != 2) cs = 2;
    CUtensorMap tmA, tmB;
    MMG_TRY(make_tmap_u8(ctx, &tmA, ctx->snps + snp_begin * ctx->pitch, ctx->pitch, snp_count, ctx->pitch, TC_BM));
    MMG_TRY(make_tmap_u8(ctx, &tmB, Wq.p, ldq, wq_rows, ldq, TC_BN / cs));
    const int groups = (int)((snp_count + TC_BM - 1) / TC_BM);
    const TcTile* td = (const TcTile*)ctx->tiles_d;
    cudaEventRecord(ctx->kev0, ctx->stream);
    if (cs == 2)
        MMG_TRY((launch_tc_gemm<PermEpi, 2>(ctx, tmA, tmB, td, groups, (int)tiles.size(), 0, TC_BM, 0, ep, "tc_gemm_i8_kernel<PermEpi,2>", L2_EVICT_FIRST, L2_EVICT_LAST)));
    else
        MMG_TRY((launch_tc_gemm<PermEpi, 1>(ctx, tmA, tmB, td, groups, (int)tiles.size(), 0, TC_BM, 0, ep, "tc_gemm_i8_kernel<PermEpi,1>", L2_EVICT_FIRST, L2_EVICT_LAST)));
    cudaEventRecord(ctx->kev1, ctx->stream);
    std::vector<double> ratio((size_t)P);
    MMG_CUDA(ctx, cudaMemcpyAsync(ratio.data(), d_ratio, P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
    ctx->last_perm_ms = ms;
    for (int64_t p = 0; p < P; ++p) ratio_inout[p] = std::max(ratio_inout[p], ratio[p]);
    return MMG_OK;
}

extern "C" {

int mmg_emmax_scan_f64(mmg_ctx* ctx, mmg_mat Rh, const double* V, int nv, double h0_rss, double n_p, int impl, int64_t snp_begin,
                       int64_t snp_count, double* ps, double* f_stats, double* rss, double* var_perc, double* xx, double* dots) {
    MmgMat* R = ctx ? get_mat(ctx, Rh) : nullptr;
    MMG_CHECK(ctx, R && ctx->snps, "mmg_emmax_scan_f64: need resident genotypes and R");
    MMG_CHECK(ctx, R->cols == ctx->n, "R must have n = %lld columns (has %lld)", (long long)ctx->n, (long long)R->cols);
    MMG_CHECK(ctx, V && nv >= 1 && nv <= 16, "need 1..16 rotated-space vectors (V[0] = residual phenotype)");
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count > 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    if (impl == MMG_IMPL_AUTO) impl = env_impl("MMG_SCAN_IMPL", MMG_IMPL_TCGEN05);
    MMG_CHECK(ctx, impl == MMG_IMPL_DMMA || impl == MMG_IMPL_TCGEN05, "unsupported impl %d for the scan", impl);
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = ctx->n, n_out = R->rows;
    const double lbeta = lbeta_host(0.5 * n_p, 0.5);

    DevBuf out;      // xx, xy, rss, f, p, var_perc  (6 x snp_count doubles) + dots
    MMG_CUDA(ctx, out.alloc(ctx->stream, (size_t)(6 + nv) * snp_count * sizeof(double)));
    double* d_xx = out.as<double>();
    double* d_xy = d_xx + snp_count;
    double* d_rss = d_xy + snp_count;
    double* d_f = d_rss + snp_count;
    double* d_p = d_f + snp_count;
    double* d_vp = d_p + snp_count;
    double* d_dots = d_vp + snp_count;

    {
        StageTimer tm(ctx, "scan");
        if (impl == MMG_IMPL_DMMA) {
            DevBuf Rp, Vp;
            int64_t rows_pad = 0, ld = 0;
            MMG_TRY(pad_matrix(ctx, R, Rp, &rows_pad, &ld));
            MMG_CUDA(ctx, Vp.alloc(ctx->stream, (size_t)rows_pad * sizeof(double)));
            MMG_CUDA(ctx, cudaMemsetAsync(Vp.p, 0, (size_t)rows_pad * sizeof(double), ctx->stream));
            MMG_CUDA(ctx, cudaMemcpyAsync(Vp.p, V, n_out * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            ScanDmmaParams prm{};
            prm.snps = ctx->snps;
            prm.pitch = ctx->pitch;
            prm.row_begin = snp_begin;
            prm.row_count = snp_count;
            prm.R = Rp.as<double>();
            prm.ldr = ld;
            prm.n_out_pad = (int)rows_pad;
            prm.k_pad = (int)round_up(n, SD_BK);
            prm.y = Vp.as<double>();
            prm.h0_rss = h0_rss;
            prm.n_p = n_p;
            prm.lbeta = lbeta;
            prm.xx = d_xx;
            prm.xy = d_xy;
            prm.rss = d_rss;
            prm.f = d_f;
            prm.p = d_p;
            prm.var_perc = d_vp;
            MMG_TRY(launch_scan_dmma(ctx, false, prm));
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        } else {
            const MmgMat* Rs[1] = {R};
            MMG_TRY(scan_tc_run(ctx, 1, Rs, V, &h0_rss, n_p, lbeta, snp_begin, snp_count, d_xx, d_xy, d_rss, d_f, d_p, d_vp));
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        ctx->last_scan_ms = ms;

        if (dots) {
            // x~.V[v] = x.(R' V[v]):  W = V R  ([nv x n_out] x [n_out x n]) then an HBM-bound dot kernel
            DevBuf Vd, Wd;
            MMG_CUDA(ctx, Vd.alloc(ctx->stream, (size_t)nv * n_out * sizeof(double)));
            MMG_CUDA(ctx, Wd.alloc(ctx->stream, (size_t)nv * n * sizeof(double)));
            MMG_CUDA(ctx, cudaMemcpyAsync(Vd.p, V, (size_t)nv * n_out * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            const double one = 1.0, zero = 0.0;
            // row-major W[nv x n] = V[nv x n_out] R[n_out x n]  ->  column-major W' = R' V'
            MMG_CUBLAS(ctx, cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_N, (int)n, nv, (int)n_out, &one, R->d, (int)R->cols,
                                        Vd.as<double>(), (int)n_out, &zero, Wd.as<double>(), (int)n));
            for (int v = 0; v < nv; ++v) {
                // one vector per launch keeps the kernel simple; dots is strided by nv on the host side
                snp_dots_kernel<1><<<(unsigned)((snp_count + 7) / 8), 256, 0, ctx->stream>>>(
                    ctx->snps, ctx->pitch, snp_begin, snp_count, (int)n, Wd.as<double>() + (int64_t)v * n, n, d_dots + (int64_t)v * snp_count);
                MMG_TRY(launch_check(ctx, "snp_dots_kernel"));
            }
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    StageTimer tm2(ctx, "d2h");
    const size_t bytes = snp_count * sizeof(double);
    if (ps) MMG_CUDA(ctx, cudaMemcpyAsync(ps, d_p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (f_stats) MMG_CUDA(ctx, cudaMemcpyAsync(f_stats, d_f, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (rss) MMG_CUDA(ctx, cudaMemcpyAsync(rss, d_rss, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (var_perc) MMG_CUDA(ctx, cudaMemcpyAsync(var_perc, d_vp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (xx) MMG_CUDA(ctx, cudaMemcpyAsync(xx, d_xx, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (dots) {
        // device layout is [nv][snp_count]; the ABI promises [snp_count][nv]
        std::vector<double> tmp((size_t)nv * snp_count);
        MMG_CUDA(ctx, cudaMemcpyAsync(tmp.data(), d_dots, (size_t)nv * bytes, cudaMemcpyDeviceToHost, ctx->stream));
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int v = 0; v < nv; ++v)
            for (int64_t s = 0; s < snp_count; ++s) dots[s * nv + v] = tmp[(size_t)v * snp_count + s];
    }
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

// The int8 scan when the caller already holds the quadratic form A = R'R (n x n, row-major lower triangle valid) and
// v = R'y~: the multi-GPU path forms A from per-rank row blocks of R (mmg_mat_syrk_rows) and one all-reduce instead of
// repeating the 2 n^3 / 2 flops of the product on every rank (28 ms at n = 10k, more than an 8-way shard of the scan itself).
int mmg_emmax_scan_quad_f64(mmg_ctx* ctx, mmg_mat Ah, const double* v, double h0_rss, double n_p, int64_t snp_begin,
                            int64_t snp_count, double* ps, double* f_stats, double* rss, double* var_perc, double* xx) {
    MmgMat* A = ctx ? get_mat(ctx, Ah) : nullptr;
    MMG_CHECK(ctx, A && ctx->snps && v, "mmg_emmax_scan_quad_f64: need resident genotypes, A and v");
    MMG_CHECK(ctx, A->rows == ctx->n && A->cols == ctx->n, "A must be n x n with n = %lld (is %lld x %lld)", (long long)ctx->n,
              (long long)A->rows, (long long)A->cols);
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count > 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const double lbeta = lbeta_host(0.5 * n_p, 0.5);
    DevBuf out;      // xx, xy, rss, f, p, var_perc
    MMG_CUDA(ctx, out.alloc(ctx->stream, (size_t)6 * snp_count * sizeof(double)));
    double* d_xx = out.as<double>();
    double* d_xy = d_xx + snp_count;
    double* d_rss = d_xy + snp_count;
    double* d_f = d_rss + snp_count;
    double* d_p = d_f + snp_count;
    double* d_vp = d_p + snp_count;
    {
        StageTimer tm(ctx, "scan");
        MMG_TRY(scan_tc_run(ctx, 1, nullptr, nullptr, &h0_rss, n_p, lbeta, snp_begin, snp_count, d_xx, d_xy, d_rss, d_f, d_p, d_vp, A, v));
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        ctx->last_scan_ms = ms;
    }
    StageTimer tm2(ctx, "d2h");
    const size_t bytes = snp_count * sizeof(double);
    if (ps) MMG_CUDA(ctx, cudaMemcpyAsync(ps, d_p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (f_stats) MMG_CUDA(ctx, cudaMemcpyAsync(f_stats, d_f, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (rss) MMG_CUDA(ctx, cudaMemcpyAsync(rss, d_rss, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (var_perc) MMG_CUDA(ctx, cudaMemcpyAsync(var_perc, d_vp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (xx) MMG_CUDA(ctx, cudaMemcpyAsync(xx, d_xx, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

// Phenotype-batched scan (BASELINE.json configs[2]; the reference runs one emmax() per phenotype): T rotations R_t
// (same shape), V[t] = residual phenotype of t in its rotated space, h0_rss[t]; outputs are [T x snp_count].
int mmg_emmax_scan_multi_f64(mmg_ctx* ctx, const mmg_mat* Rh, int T, const double* V, const double* h0_rss, double n_p,
                             int64_t snp_begin, int64_t snp_count, double* ps, double* f_stats, double* rss, double* var_perc,
                             double* xx) {
    MMG_CHECK(ctx, ctx && ctx->snps && Rh && V && h0_rss && T >= 1 && T <= (1 << 20), "mmg_emmax_scan_multi_f64: bad argument");
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count > 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    std::vector<const MmgMat*> Rs((size_t)T);
    for (int t = 0; t < T; ++t) {
        Rs[t] = get_mat(ctx, Rh[t]);
        MMG_CHECK(ctx, Rs[t] && Rs[t]->cols == ctx->n && Rs[t]->rows == Rs[0]->rows, "R[%d]: unknown handle or shape mismatch", t);
    }
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const double lbeta = lbeta_host(0.5 * n_p, 0.5);
    DevBuf out;      // xx, rss, f, p, var_perc: 5 x T x snp_count doubles
    const int64_t cnt = (int64_t)T * snp_count;
    MMG_CUDA(ctx, out.alloc(ctx->stream, (size_t)5 * cnt * sizeof(double)));
    double* d_xx = out.as<double>();
    double* d_rss = d_xx + cnt;
    double* d_f = d_rss + cnt;
    double* d_p = d_f + cnt;
    double* d_vp = d_p + cnt;
    {
        StageTimer tm(ctx, "scan");
        MMG_TRY(scan_tc_run(ctx, T, Rs.data(), V, h0_rss, n_p, lbeta, snp_begin, snp_count, d_xx, nullptr, d_rss, d_f, d_p, d_vp));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        ctx->last_scan_ms = ms;
    }
    StageTimer tm2(ctx, "d2h");
    const size_t bytes = (size_t)cnt * sizeof(double);
    if (ps) MMG_CUDA(ctx, cudaMemcpyAsync(ps, d_p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (f_stats) MMG_CUDA(ctx, cudaMemcpyAsync(f_stats, d_f, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (rss) MMG_CUDA(ctx, cudaMemcpyAsync(rss, d_rss, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (var_perc) MMG_CUDA(ctx, cudaMemcpyAsync(var_perc, d_vp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (xx) MMG_CUDA(ctx, cudaMemcpyAsync(xx, d_xx, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}


int mmg_emmax_perm_scan_f64(mmg_ctx* ctx, mmg_mat Rh, mmg_mat Wh, int centre, int impl, int64_t snp_begin, int64_t snp_count,
                            double* ratio_inout) {
    MmgMat *R = ctx ? get_mat(ctx, Rh) : nullptr, *Wt = ctx ? get_mat(ctx, Wh) : nullptr;
    MMG_CHECK(ctx, R && Wt && ctx->snps && ratio_inout, "mmg_emmax_perm_scan_f64: bad argument");
    MMG_CHECK(ctx, R->cols == ctx->n && Wt->cols == ctx->n, "R and W' must have n columns");
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count > 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    if (impl == MMG_IMPL_AUTO) impl = env_impl("MMG_PERM_IMPL", MMG_IMPL_TCGEN05);
    MMG_CHECK(ctx, impl == MMG_IMPL_TCGEN05 || impl == MMG_IMPL_DMMA, "unsupported impl %d for the permutation scan", impl);
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (impl == MMG_IMPL_TCGEN05) return perm_scan_tc(ctx, R, Wt, centre, snp_begin, snp_count, ratio_inout);
    StageTimer tm(ctx, "scan");
    const int64_t n = ctx->n, P = Wt->rows;
    DevBuf Rp, Wp, aux;
    int64_t r_rows = 0, r_ld = 0, w_rows = 0, w_ld = 0;
    MMG_TRY(pad_matrix(ctx, R, Rp, &r_rows, &r_ld));
    MMG_TRY(pad_matrix(ctx, Wt, Wp, &w_rows, &w_ld));
    // aux: ones[n] | r1[r_rows] | wsum[w_rows] | zeros y[r_rows] | mu[snp_count] | xx[snp_count] | ratio[w_rows] | sums[snp_count]
    const int64_t nd = r_ld + r_rows + w_rows + r_rows + 2 * snp_count + w_rows;
    MMG_CUDA(ctx, aux.alloc(ctx->stream, (size_t)nd * sizeof(double) + (size_t)snp_count * sizeof(long long)));
    MMG_CUDA(ctx, cudaMemsetAsync(aux.p, 0, (size_t)nd * sizeof(double), ctx->stream));
    double* d_ones = aux.as<double>();
    double* d_r1 = d_ones + r_ld;
    double* d_wsum = d_r1 + r_rows;
    double* d_y0 = d_wsum + w_rows;
    double* d_mu = d_y0 + r_rows;
    double* d_xx = d_mu + snp_count;
    unsigned long long* d_ratio = (unsigned long long*)(d_xx + snp_count);
    long long* d_sums = (long long*)(d_ratio + w_rows);
    {
        std::vector<double> ones((size_t)n, 1.0);
        MMG_CUDA(ctx, cudaMemcpyAsync(d_ones, ones.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const double one = 1.0, zero = 0.0;
    // r1 = R 1, wsum = W' 1 (padded matrices are column-major [ld x rows]: y = A' x)
    MMG_CUBLAS(ctx, cublasDgemv(ctx->cublas, CUBLAS_OP_T, (int)r_ld, (int)r_rows, &one, Rp.as<double>(), (int)r_ld, d_ones, 1, &zero, d_r1, 1));
    MMG_CUBLAS(ctx, cublasDgemv(ctx->cublas, CUBLAS_OP_T, (int)w_ld, (int)w_rows, &one, Wp.as<double>(), (int)w_ld, d_ones, 1, &zero, d_wsum, 1));
    if (centre) {
        snp_row_sums_kernel<<<(unsigned)((snp_count + 7) / 8), 256, 0, ctx->stream>>>(ctx->snps + snp_begin * ctx->pitch, ctx->pitch,
                                                                                      snp_count, (int)n, d_sums, nullptr);
        MMG_TRY(launch_check(ctx, "snp_row_sums_kernel"));
        means_from_sums_kernel<<<(unsigned)((snp_count + 255) / 256), 256, 0, ctx->stream>>>(d_sums, snp_count, 1.0 / (double)n, d_mu);
        MMG_TRY(launch_check(ctx, "means_from_sums_kernel"));
    }
    ScanDmmaParams prm{};
    prm.snps = ctx->snps;
    prm.pitch = ctx->pitch;
    prm.row_begin = snp_begin;
    prm.row_count = snp_count;
    prm.k_pad = (int)round_up(n, SD_BK);
    prm.mu = d_mu;                     // zeros when !centre
    // pass 1: xx of the (centred) rotated SNPs
    prm.R = Rp.as<double>();
    prm.ldr = r_ld;
    prm.n_out_pad = (int)r_rows;
    prm.y = d_y0;
    prm.r1 = d_r1;
    prm.xx = d_xx;
    prm.h0_rss = 1.0;
    prm.n_p = 1.0;
    MMG_TRY(launch_scan_dmma(ctx, false, prm));
    // pass 2: max over SNPs of (x_c . W_p)^2 / xx
    prm.R = Wp.as<double>();
    prm.ldr = w_ld;
    prm.n_out_pad = (int)w_rows;
    prm.r1 = d_wsum;
    prm.xx = nullptr;
    prm.xx_in = d_xx;
    prm.ratio_max = d_ratio;
    MMG_TRY(launch_scan_dmma(ctx, true, prm));
    std::vector<double> ratio((size_t)P);
    MMG_CUDA(ctx, cudaMemcpyAsync(ratio.data(), d_ratio, P * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int64_t p = 0; p < P; ++p) ratio_inout[p] = std::max(ratio_inout[p], ratio[p]);
    return MMG_OK;
}

int mmg_f_sf_f64(mmg_ctx* ctx, const double* f, int64_t count, double dfn, double dfd, double* out) {
    MMG_CHECK(ctx, ctx && f && out && count >= 0 && dfn > 0 && dfd > 0, "mmg_f_sf_f64: bad argument");
    if (count == 0) return MMG_OK;
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf buf;
    MMG_CUDA(ctx, buf.alloc(ctx->stream, 2 * count * sizeof(double)));
    double* d_f = buf.as<double>();
    double* d_o = d_f + count;
    MMG_CUDA(ctx, cudaMemcpyAsync(d_f, f, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    f_sf_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(d_f, count, dfn, dfd, lbeta_host(0.5 * dfd, 0.5 * dfn), d_o);
    MMG_TRY(launch_check(ctx, "f_sf_kernel"));
    MMG_CUDA(ctx, cudaMemcpyAsync(out, d_o, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

// ======================================================================================================
// diagnostics
// ======================================================================================================
__global__ void __launch_bounds__(256) bench_dmma_kernel(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma_8x8x4(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) bench_dfma_kernel(double* out, int iters) {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}
__global__ void bench_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n16) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // extern "C"

// tcgen05 int8 MMA issue rate with shared-memory-resident operands (no loads): one thread per CTA issues `iters` K blocks
// (4 x UMMA K=32, M=128 per CTA, N=256) into two alternating accumulators.  ldtm != 0: the four epilogue warps read
// the accumulators back with tcgen05.ld at the rate of one full 128x256 tile per `ldtm` K blocks, free running
// (measures whether TMEM reads take cycles from the tensor pipe).  PAIR: cta_group::2 (M = 256 over two CTAs).
template <bool PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1) bench_imma_kernel(int iters, int ldtm, unsigned* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (TC_A_BYTES + TC_B_BYTES) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = (0x9E3779B9u * (i + 1)) & 0x03030303u;      // genotype-like bytes 0..3
    if (threadIdx.x == 0) {
        mbar_init(&done_bar, 1);
        mbar_fence_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        if (PAIR) { tmem_alloc_pair(&tmem_slot, TC_TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(&tmem_slot, TC_TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const bool leader = !PAIR || cluster_ctarank() == 0;
    if (warp == 1 && lane == 0 && leader) {
        constexpr uint32_t idesc = umma_idesc_i8(PAIR ? 2 * TC_BM : TC_BM, TC_BN);
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem)), db = umma_desc_kmajor_sw128(smem_u32(smem) + TC_A_BYTES);
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem_base + ((it >> 3) & 1) * TC_BN;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (PAIR) umma_i8_pair(d, da + 2 * kk, db + 2 * kk, idesc, ((it & 7) | kk) ? 1u : 0u);
                else umma_i8(d, da + 2 * kk, db + 2 * kk, idesc, ((it & 7) | kk) ? 1u : 0u);
            }
        }
        if (PAIR) umma_commit_pair(&done_bar, 0b11); else umma_commit(&done_bar);
    }
    if (warp >= 2 && ldtm > 0) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        unsigned acc = 0;
        const int tiles = iters / ldtm;
        for (int t = 0; t < tiles; ++t) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                uint32_t v[16];
                tmem_ld_32x16(taddr + (t & 1) * TC_BN + c * 16, v);
                tmem_ld_wait_dep(v);
#pragma unroll
                for (int j = 0; j < 16; ++j) acc += v[j];
            }
        }
        if (acc == 0x12345678u) sink[0] = acc;
    }
    if (warp == 1 || warp == 0) {
        if (lane == 0) mbar_wait(&done_bar, 0);
        __syncwarp();
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, TC_TMEM_COLS); else tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

// TMEM read-back rate of the scan's epilogue pattern: WARPS epilogue warps (WARPS / 4 per lane quadrant, each 256 * 4 / WARPS
// columns of a 128 x 256 int32 accumulator tile), tcgen05.ld.32x32b.x<LDW> double buffered with a multiply-accumulate per
// element between the waits, while (with_mma) the tensor pipe runs flat out into the other accumulator.  Reports SM cycles per tile.
template <int WARPS, int LDW>
__global__ void __launch_bounds__(64 + 32 * WARPS, 1) bench_ldtm_kernel(int tiles, int with_mma, long long* out, unsigned* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int stop_flag;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (TC_A_BYTES + TC_B_BYTES) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = (0x9E3779B9u * (i + 1)) & 0x03030303u;
    if (threadIdx.x == 0) {
        mbar_init(&done_bar, 1);
        mbar_fence_init();
        stop_flag = 0;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) { tmem_alloc(&tmem_slot, TC_TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 1 && lane == 0) {
        if (with_mma) {
            constexpr uint32_t idesc = umma_idesc_i8(TC_BM, TC_BN);
            const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem)), db = umma_desc_kmajor_sw128(smem_u32(smem) + TC_A_BYTES);
            // keep the pipe busy until the readers are done: batches of 64 K-blocks, then look at the flag
            for (int it = 0; it < (1 << 22); ++it) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_i8(tmem_base + TC_BN, da + 2 * kk, db + 2 * kk, idesc, (it | kk) ? 1u : 0u);
                if ((it & 63) == 63 && *(volatile int*)&stop_flag >= WARPS) break;
            }
        }
        umma_commit(&done_bar);
    }
    if (warp >= 2) {
        constexpr int kCols = 256 * 4 / WARPS;
        const int quad = warp & 3, part = (warp - 2) >> 2;
        const uint32_t taddr = tmem_base + part * kCols + (static_cast<uint32_t>(quad * 32) << 16);
        int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        const uint32_t xw = 0x01020100u + lane;
        const long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            uint32_t va[LDW], vb[LDW];
            if constexpr (LDW == 32) tmem_ld_32x32(taddr, va); else tmem_ld_32x16(taddr, va);
#pragma unroll
            for (int c = 0; c < kCols / LDW; c += 2) {
                tmem_ld_wait_dep(va, s0, s1, s2, s3);
                if constexpr (LDW == 32) tmem_ld_32x32(taddr + (c + 1) * LDW, vb); else tmem_ld_32x16(taddr + (c + 1) * LDW, vb);
#pragma unroll
                for (int j = 0; j < LDW; j += 4) {
                    const uint32_t w = xw + j;
                    s0 += (int)va[j + 0] * (int)(int8_t)(w & 0xffu);
                    s1 += (int)va[j + 1] * (int)(int8_t)((w >> 8) & 0xffu);
                    s2 += (int)va[j + 2] * (int)(int8_t)((w >> 16) & 0xffu);
                    s3 += (int)va[j + 3] * (int)(int8_t)(w >> 24);
                }
                tmem_ld_wait_dep(vb, s0, s1, s2, s3);
                if (c + 2 < kCols / LDW) {
                    if constexpr (LDW == 32) tmem_ld_32x32(taddr + (c + 2) * LDW, va); else tmem_ld_32x16(taddr + (c + 2) * LDW, va);
                }
#pragma unroll
                for (int j = 0; j < LDW; j += 4) {
                    const uint32_t w = xw + j + 1;
                    s0 += (int)vb[j + 0] * (int)(int8_t)(w & 0xffu);
                    s1 += (int)vb[j + 1] * (int)(int8_t)((w >> 8) & 0xffu);
                    s2 += (int)vb[j + 2] * (int)(int8_t)((w >> 16) & 0xffu);
                    s3 += (int)vb[j + 3] * (int)(int8_t)(w >> 24);
                }
            }
        }
        const long long t1 = clock64();
        if (lane == 0) {
            atomicAdd(&stop_flag, 1);
            if (warp == 2) out[blockIdx.x] = t1 - t0;
        }
        if (s0 + s1 + s2 + s3 == 0x12345678) sink[0] = 1;
    }
    if (warp == 1 || warp == 0) {
        if (lane == 0) mbar_wait(&done_bar, 0);
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

template <int WARPS, int LDW>
static int run_bench_ldtm(mmg_ctx* ctx, int with_mma, double* value) {
    const int tiles = 2000, smem = TC_A_BYTES + TC_B_BYTES + 1024, grid = ctx->sm_count;
    auto kern = bench_ldtm_kernel<WARPS, LDW>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    MMG_CUDA(ctx, cudaMemsetAsync(ctx->scratch, 0, (size_t)(grid + 2) * sizeof(long long), ctx->stream));
    long long* out = (long long*)ctx->scratch;
    kern<<<grid, 64 + 32 * WARPS, smem, ctx->stream>>>(tiles, with_mma, out, (unsigned*)(out + grid));
    ctx->launches += 1;
    MMG_TRY(launch_check(ctx, "bench_ldtm_kernel"));
    std::vector<long long> h((size_t)grid);
    MMG_CUDA(ctx, cudaMemcpyAsync(h.data(), out, (size_t)grid * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double sum = 0.0;
    for (long long v : h) sum += (double)v;
    *value = sum / grid / tiles;
    return MMG_OK;
}

extern "C" {

int mmg_microbench(mmg_ctx* ctx, const char* which, double* value) {
    MMG_CHECK(ctx, ctx && which && value, "mmg_microbench: bad argument");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    MMG_TRY(ensure_scratch(ctx, 1 << 20));
    float ms = 0.f;
    if (!strcmp(which, "dmma") || !strcmp(which, "dfma")) {
        const bool dm = !strcmp(which, "dmma");
        const int iters = 20000, blocks = ctx->sm_count * 4;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(ctx->kev0, ctx->stream);
            if (dm) bench_dmma_kernel<<<blocks, 256, 0, ctx->stream>>>((double*)ctx->scratch, iters);
            else bench_dfma_kernel<<<blocks, 256, 0, ctx->stream>>>((double*)ctx->scratch, iters);
            MMG_TRY(launch_check(ctx, "bench kernel"));
            cudaEventRecord(ctx->kev1, ctx->stream);
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        }
        const double flops = dm ? (double)blocks * 8 /*warps*/ * iters * 8.0 * 512.0 : (double)blocks * 256 * iters * 16.0 * 2.0;
        *value = flops / (ms * 1e-3) / 1e12;
        return MMG_OK;
    }
    if (!strcmp(which, "copy")) {
        const int64_t bytes = 2ll << 30;
        DevBuf a, b;
        MMG_CUDA(ctx, a.alloc(ctx->stream, bytes));
        MMG_CUDA(ctx, b.alloc(ctx->stream, bytes));
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(ctx->kev0, ctx->stream);
            bench_copy_kernel<<<ctx->sm_count * 16, 512, 0, ctx->stream>>>(a.as<uint4>(), b.as<uint4>(), bytes / 16);
            MMG_TRY(launch_check(ctx, "bench_copy_kernel"));
            cudaEventRecord(ctx->kev1, ctx->stream);
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        }
        *value = 2.0 * bytes / (ms * 1e-3) / 1e9;
        return MMG_OK;
    }
    if (!strncmp(which, "prepass", 7)) {
        // "prepass_r<rows>_u<unroll>_b<min blocks>": ms of the scan's linear pre-pass over the resident genotypes (tuning aid)
        MMG_CHECK(ctx, ctx->snps != nullptr, "prepass microbench: no resident genotypes");
        const int64_t npad = round_up(ctx->n, 256), cnt = ctx->m;
        DevBuf buf;
        MMG_CUDA(ctx, buf.alloc(ctx->stream, (size_t)(2 * npad + 3 * cnt) * sizeof(double)));
        MMG_CUDA(ctx, cudaMemsetAsync(buf.p, 0, (size_t)(2 * npad) * sizeof(double), ctx->stream));
        double* v = buf.as<double>();
        double* o = v + 2 * npad;
        auto run = [&](auto kern, int rows) {
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(ctx->kev0, ctx->stream);
                kern<<<(unsigned)((cnt + 8 * rows - 1) / (8 * rows)), 256, 0, ctx->stream>>>(ctx->snps, ctx->pitch, 0, cnt, 1, v, v + npad, npad, o, o + cnt,
                                                                                        o + 2 * cnt, cnt);
                cudaEventRecord(ctx->kev1, ctx->stream);
                cudaStreamSynchronize(ctx->stream);
                cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
            }
        };
        if (!strcmp(which, "prepass_r4_u4_b3")) run(snp_prepass_kernel<4, 4, 3>, 4);
        else if (!strcmp(which, "prepass_r4_u4_b1")) run(snp_prepass_kernel<4, 4, 1>, 4);
        else if (!strcmp(which, "prepass_r4_u2_b3")) run(snp_prepass_kernel<4, 2, 3>, 4);
        else if (!strcmp(which, "prepass_r2_u4_b4")) run(snp_prepass_kernel<2, 4, 4>, 2);
        else if (!strcmp(which, "prepass_r4_u8_b1")) run(snp_prepass_kernel<4, 8, 1>, 4);
        else if (!strcmp(which, "prepass_r4_u4_b4")) run(snp_prepass_kernel<4, 4, 4>, 4);
        else if (!strcmp(which, "prepass_r4_u8_b4")) run(snp_prepass_kernel<4, 8, 4>, 4);
        else if (!strcmp(which, "prepass_r8_u4_b2")) run(snp_prepass_kernel<8, 4, 2>, 8);
        else if (!strcmp(which, "prepass_r8_u2_b3")) run(snp_prepass_kernel<8, 2, 3>, 8);
        else if (!strcmp(which, "prepass_r2_u8_b4")) run(snp_prepass_kernel<2, 8, 4>, 2);
        else if (!strcmp(which, "prepass_r2_u16_b4")) run(snp_prepass_kernel<2, 16, 4>, 2);
        else if (!strcmp(which, "prepass_r1_u16_b4")) run(snp_prepass_kernel<1, 16, 4>, 1);
        else return fail(ctx, MMG_EBADARG, "unknown microbench '%s'", which);
        MMG_TRY(launch_check(ctx, "snp_prepass_kernel"));
        *value = ms;
        return MMG_OK;
    }
    if (!strncmp(which, "ldtm", 4)) {
        // "ldtm_w<4|8|16>_x<16|32>[_mma]": SM cycles per 128 x 256 int32 tile read back by the epilogue pattern
        const int w = strstr(which, "_w16") ? 16 : strstr(which, "_w8") ? 8 : 4;
        const int x = strstr(which, "_x32") ? 32 : 16;
        const int mma = strstr(which, "_mma") ? 1 : 0;
        if (w == 4 && x == 16) return run_bench_ldtm<4, 16>(ctx, mma, value);
        if (w == 4 && x == 32) return run_bench_ldtm<4, 32>(ctx, mma, value);
        if (w == 8 && x == 16) return run_bench_ldtm<8, 16>(ctx, mma, value);
        if (w == 8 && x == 32) return run_bench_ldtm<8, 32>(ctx, mma, value);
        if (w == 16 && x == 16) return run_bench_ldtm<16, 16>(ctx, mma, value);
        return run_bench_ldtm<16, 32>(ctx, mma, value);
    }
    if (!strncmp(which, "imma", 4)) {
        // "imma_tcgen05" | "imma_pair" | "imma_tcgen05_ldtm<k>" | "imma_pair_ldtm<k>": TOP/s
        const bool pair = strstr(which, "pair") != nullptr;
        const char* l = strstr(which, "ldtm");
        const int ldtm = l ? std::max(1, atoi(l + 4)) : 0;
        const int iters = 40000, smem = TC_A_BYTES + TC_B_BYTES + 1024;
        const int grid = pair ? ctx->sm_count / 2 * 2 : ctx->sm_count;
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3(TC_THREADS);
        cfg.gridDim = dim3((unsigned)grid);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = pair ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaFuncSetAttribute(bench_imma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(bench_imma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(ctx->kev0, ctx->stream);
            cudaError_t e = pair ? cudaLaunchKernelEx(&cfg, bench_imma_kernel<true>, iters, ldtm, (unsigned*)ctx->scratch)
                                 : cudaLaunchKernelEx(&cfg, bench_imma_kernel<false>, iters, ldtm, (unsigned*)ctx->scratch);
            ctx->launches += 1;
            if (e != cudaSuccess) return fail(ctx, MMG_ECUDA, "bench_imma_kernel launch failed: %s", cudaGetErrorString(e));
            cudaEventRecord(ctx->kev1, ctx->stream);
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        }
        *value = 2.0 * (double)grid * iters * TC_BM * TC_BN * TC_BK / (ms * 1e-3) / 1e12;
        return MMG_OK;
    }
    return fail(ctx, MMG_EBADARG, "unknown microbench '%s'", which);
}

}  // extern "C"
