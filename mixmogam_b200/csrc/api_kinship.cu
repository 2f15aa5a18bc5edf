// libmixmogam_b200: stage 1 of the C ABI -- the kinship Gram (tcgen05 int8 / SIMT), its streamed host source, the FP64
// finalisation, and the IBD kinship (int8 digit planes or cuBLAS dsyrk).
#include <algorithm>
#include "common.cuh"
#include "ibd_tc.cuh"
#include "gram_pair.cuh"

using namespace mmg;

namespace mmg {
void kinship_init_attrs() {
    cudaFuncSetAttribute(tc_gemm_i8_kernel<GramEpi, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<GramEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<GramEpiF4, 1, TC_KIND_MXF4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<GramEpiF4, 2, TC_KIND_MXF4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(gram_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<IbdEpi, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<IbdEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
}
}  // namespace mmg

// CTA-pair Gram (gram_pair.cuh): launch configuration and the number of co-resident clusters (= persistent grid / 2)
static void gram_pair_config(mmg_ctx* ctx, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int clusters) {
    cfg = cudaLaunchConfig_t{};
    cfg.blockDim = dim3(GP_THREADS);
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.dynamicSmemBytes = GP_SMEM_BYTES;
    cfg.stream = ctx->stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}
static int gram_pair_max_clusters(mmg_ctx* ctx) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    gram_pair_config(ctx, cfg, attr, ctx->sm_count / 2);
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, gram_pair_kernel, &cfg) != cudaSuccess || q <= 0) {
        cudaGetLastError();
        q = ctx->sm_count / 2;
    }
    return std::max(1, std::min(q, ctx->sm_count / 2));
}

extern "C" {

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
    const uint8_t* packed2 = nullptr;  // or: rows the caller packed already (2 bits per genotype), row stride ld2 bytes -- a quarter
    int64_t ld2 = 0;                   // of the bytes cross PCIe and no host core touches them
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
// n_valid > 0: codes of columns >= n_valid are forced to 0 (rows packed by the caller: nothing is assumed about the unused bits)
static __global__ void unpack2_kernel(const uint8_t* __restrict__ packed, int64_t p_ld, int8_t* __restrict__ out, int64_t pitch, int64_t rows,
                                      int64_t n_valid = 0) {
    const int64_t wpr = p_ld >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * wpr) return;
    const int64_t r = idx / wpr, w = idx - r * wpr;
    if (16 * w >= pitch) return;
    uint32_t v = *reinterpret_cast<const uint32_t*>(packed + r * p_ld + 4 * w);
    if (n_valid > 0 && 16 * (w + 1) > n_valid) {
        const int64_t keep = n_valid - 16 * w;                       // genotypes of this word that exist
        v = keep <= 0 ? 0u : (v & (0xffffffffu >> (32 - 2 * (int)keep)));
    }
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
static int host_source_open(mmg_ctx* ctx, GramHostSource& src);

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
    {
        cudaPointerAttributes pa{};
        src.pinned = cudaPointerGetAttributes(&pa, snps) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
    }
    src.pack_ok = env_int("MMG_H2D_PACK", 1) != 0;
    src.threads = std::max(1, env_int("MMG_HOST_THREADS", mmg_host_threads_default()));
    MMG_TRY(host_source_open(ctx, src));
    const int rc = gram_run(ctx, coding, impl, 0, m, reset, &src);
    if (rc == MMG_OK) ctx->snps_absmax = coding == MMG_CODING_DIPLOID ? 2 : 1;   // the pack kernels checked every byte against the coding
    return rc;
}

static int host_source_open(mmg_ctx* ctx, GramHostSource& src) {
    MMG_CUDA(ctx, cudaStreamCreateWithFlags(&src.stream, cudaStreamNonBlocking));
    MMG_CUDA(ctx, cudaStreamCreateWithFlags(&src.stream2, cudaStreamNonBlocking));
    MMG_CUDA(ctx, cudaEventCreate(&src.t0));
    for (int i = 0; i < 2; ++i) MMG_CUDA(ctx, cudaEventCreateWithFlags(&src.slot_free[i], cudaEventDisableTiming));
    // the zero fill of the row padding (mmg_snps_reserve, compute stream) must not race with the copies
    MMG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    MMG_CUDA(ctx, cudaStreamWaitEvent(src.stream, ctx->ev0, 0));
    MMG_CUDA(ctx, cudaStreamWaitEvent(src.stream2, ctx->ev0, 0));
    MMG_CUDA(ctx, cudaEventRecord(src.t0, src.stream2));
    return MMG_OK;
}

int mmg_kinship_gram_i8_host_packed2(mmg_ctx* ctx, int coding, int impl, const uint8_t* packed, int64_t m, int64_t n, int64_t ld_bytes, int reset) {
    MMG_CHECK(ctx, ctx && packed && m > 0 && n > 0 && ld_bytes >= (n + 3) / 4, "mmg_kinship_gram_i8_host_packed2: bad argument");
    MMG_TRY(mmg_snps_reserve(ctx, m, n));
    ctx->snps_absmax = -1;
    GramHostSource src;
    src.packed2 = packed;
    src.ld2 = ld_bytes;
    MMG_TRY(host_source_open(ctx, src));
    const int rc = gram_run(ctx, coding, impl, 0, m, reset, &src);
    if (rc == MMG_OK) ctx->snps_absmax = coding == MMG_CODING_DIPLOID ? 2 : 1;
    return rc;
}

// Upload only (scan-only callers): caller-packed rows -> resident int8 block, chunk by chunk through the device slots.
int mmg_snps_upload_packed2(mmg_ctx* ctx, const uint8_t* packed, int64_t m, int64_t n, int64_t ld_bytes) {
    MMG_CHECK(ctx, ctx && packed && m > 0 && n > 0 && ld_bytes >= (n + 3) / 4, "mmg_snps_upload_packed2: bad argument");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    MMG_TRY(mmg_snps_reserve(ctx, m, n));
    ctx->snps_absmax = -1;
    StageTimer tm(ctx, "h2d");
    const int64_t p2_ld = round_up((n + 3) / 4, 16), width = (n + 3) / 4;
    const int64_t chunk = std::min<int64_t>(m, 65536);
    DevBuf slot[2];
    for (int i = 0; i < 2; ++i) {
        MMG_CUDA(ctx, slot[i].alloc(ctx->stream, (size_t)chunk * p2_ld));
        MMG_CUDA(ctx, cudaMemsetAsync(slot[i].p, 0, (size_t)chunk * p2_ld, ctx->stream));
    }
    int sl = 0;
    for (int64_t r0 = 0; r0 < m; r0 += chunk, sl ^= 1) {
        const int64_t cnt = std::min(chunk, m - r0);
        if (ld_bytes == p2_ld)                               // rows at the slot's pitch: one contiguous copy (see queue_prepacked)
            MMG_CUDA(ctx, cudaMemcpyAsync(slot[sl].p, packed + r0 * ld_bytes, (size_t)(cnt * p2_ld), cudaMemcpyHostToDevice, ctx->stream));
        else
            MMG_CUDA(ctx, cudaMemcpy2DAsync(slot[sl].p, p2_ld, packed + r0 * ld_bytes, ld_bytes, width, cnt, cudaMemcpyHostToDevice, ctx->stream));
        const int64_t words = cnt * (p2_ld >> 2);
        unpack2_kernel<<<(unsigned)((words + 255) / 256), 256, 0, ctx->stream>>>(slot[sl].as<uint8_t>(), p2_ld, ctx->snps + r0 * ctx->pitch, ctx->pitch, cnt, n);
        MMG_TRY(launch_check(ctx, "unpack2_kernel"));
    }
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the caller's rows are borrowed for the duration of the call only
    return MMG_OK;
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

    // MMG_GRAM_KIND = fp4 (default) | i8: the 0 / +-1 planes of both codings are exact e2m1 values, which the tensor cores
    // multiply at twice the int8 rate (kind::mxf4, unit block scales); sums of at most 2^17 such products per chunk are exact
    // in the FP32 accumulator.  The SIMT cross-check kernel reads int8 operands.
    bool fp4 = impl == MMG_IMPL_TCGEN05;
    if (const char* e = getenv("MMG_GRAM_KIND")) fp4 = fp4 && strcmp(e, "i8") != 0;
    ctx->last_gram_fp4 = fp4 ? 1 : 0;
    const int64_t chunk = 65536;                         // SNPs per packed chunk (multiple of 256)
    const int64_t p_pitch = fp4 ? chunk * c / 2 : chunk * c;   // bytes per individual
    // MMG_GRAM_OVERLAP=1: the pack kernel of chunk c + 1 (HBM bound, a few KB of shared memory per block) runs on the side stream
    // underneath the Gram of chunk c (one persistent CTA per SM), in two operand slots.  Measured neutral at n = 10k x 1M (209.6 vs
    // 209.8 ms per step: the packs take 23 instead of 8 ms next to the Gram and the Gram 37.7 instead of 34.0 ms -- they compete for
    // the same L2 / HBM path), so the default stays the serial order.
    const bool overlap = env_int("MMG_GRAM_OVERLAP", 0) != 0 && snp_count > chunk;
    const int64_t slot_bytes = (int64_t)n * p_pitch;
    const int64_t need = slot_bytes * (overlap ? 2 : 1);
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
    // MMG_GRAM_PAIR (default 1): the e2m1 Gram as a CTA-pair MMA (tcgen05.mma.cta_group::2, gram_pair.cuh) instead of two
    // single-CTA MMAs that multicast the shared operand
    const bool pair = fp4 && gram_cs == 2 && env_int("MMG_GRAM_PAIR", 1) != 0;
    ctx->last_gram_pair = pair ? 1 : 0;
    std::vector<TcTile> tiles, table;
    int gram_clusters = 1;
    if (impl == MMG_IMPL_TCGEN05) {
        const int tiles_n = (n + TC_BN - 1) / TC_BN;
        for (int jn = 0; jn < tiles_n; ++jn)
            for (int im = 0; im <= 2 * jn + 1; im += gram_cs) tiles.push_back(TcTile{im * TC_BM, jn * TC_BN, 0, 0, 0, 0, 0, 0});
        // MMG_GRAM_BLOCKED=1: walk the triangle in blocks of 9 column tiles x 8 row-tile pairs (about one wave of the persistent grid)
        // instead of column by column: a wave then touches 4352 operand rows instead of nearly all of them, and the operand is
        // read from HBM ~5 instead of ~9 times per chunk.  The integer result does not depend on the order.
        if (gram_cs == 2 && env_int("MMG_GRAM_BLOCKED", 0) != 0)
            std::stable_sort(tiles.begin(), tiles.end(), [](const TcTile& a, const TcTile& b) {
                const int aj = a.n0 / TC_BN, bj = b.n0 / TC_BN, ai = a.m0 / (2 * TC_BM), bi = b.m0 / (2 * TC_BM);
                if (aj / 9 != bj / 9) return aj / 9 < bj / 9;
                if (ai / 8 != bi / 8) return ai / 8 < bi / 8;
                if (aj != bj) return aj < bj;
                return ai < bi;
            });
        if (pair) gram_clusters = gram_pair_max_clusters(ctx);
        else if (fp4) gram_clusters = gram_cs == 2 ? tc_gemm_max_clusters<GramEpiF4, 2, TC_KIND_MXF4>(ctx) : tc_gemm_max_clusters<GramEpiF4, 1, TC_KIND_MXF4>(ctx);
        else gram_clusters = gram_cs == 2 ? tc_gemm_max_clusters<GramEpi, 2>(ctx) : tc_gemm_max_clusters<GramEpi, 1>(ctx);
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
    // rows packed by the caller: straight from the caller's buffer into a device slot (no host core involved), then unpack2
    auto queue_prepacked = [&](int64_t ci) -> int {
        const int64_t s0 = ci * chunk, cnt = chunk_rows(ci);
        const int sl = src->next_slot;
        const int64_t width = (ctx->n + 3) / 4;
        // rows at the slot's own pitch (what _lib.pack_genotypes writes): ONE contiguous copy at the link rate -- a strided copy of
        // 2500-byte rows reaches a third of it (138 instead of ~50 ms for 1M SNPs); unpack2_kernel masks the codes beyond n
        if (src->ld2 == p2_ld)
            MMG_CUDA(ctx, cudaMemcpyAsync(ctx->stage_dev[sl], src->packed2 + (snp_begin + s0) * src->ld2, (size_t)(cnt * p2_ld), cudaMemcpyHostToDevice,
                                          src->stream2));
        else
            MMG_CUDA(ctx, cudaMemcpy2DAsync(ctx->stage_dev[sl], p2_ld, src->packed2 + (snp_begin + s0) * src->ld2, src->ld2, width, cnt,
                                            cudaMemcpyHostToDevice, src->stream2));
        const int64_t words = cnt * (p2_ld >> 2);
        unpack2_kernel<<<(unsigned)((words + 255) / 256), 256, 0, src->stream2>>>(ctx->stage_dev[sl], p2_ld, ctx->snps + (snp_begin + s0) * ctx->pitch,
                                                                                ctx->pitch, cnt, ctx->n);
        MMG_TRY(launch_check(ctx, "unpack2_kernel"));
        MMG_CUDA(ctx, cudaEventRecord(src->done[(size_t)ci], src->stream2));
        src->next_slot = sl ^ 1;
        src->packed_chunks += 1;
        staged[(size_t)ci] = 1;
        return MMG_OK;
    };
    if (src && src->packed2) {
        const int64_t stage_need = std::min(chunk, snp_count) * p2_ld;
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
        // the copies fill the first ceil(n/4) bytes of a slot row; the rest of the row must read as zero
        for (int i = 0; i < 2; ++i) MMG_CUDA(ctx, cudaMemsetAsync(ctx->stage_dev[i], 0, (size_t)ctx->stage_bytes, src->stream2));
    }
    // Stage chunk ci (if a look-ahead has not done so already).  Packed lane when the raw lane would deliver it later:
    // pageable rows always (a raw copy blocks the host at the pageable rate), page-locked rows when the DMA backlog exceeds the
    // time the host needs to pack the chunk.  Before the host disappears into a pack, the raw lane is topped up with the
    // following chunks so that the link stays busy meanwhile.
    auto issue_copy = [&](int64_t ci) -> int {
        if (staged[(size_t)ci]) return MMG_OK;
        if (src->packed2) {
            // all the copies are queued at once, two chunks ahead of the Gram is enough to keep the link busy; the stream order
            // of the packed lane (copy, unpack, copy, ...) makes the two device slots safe without events
            for (int64_t cj = ci; cj < std::min<int64_t>(n_chunks, ci + 2); ++cj)
                if (!staged[(size_t)cj]) MMG_TRY(queue_prepacked(cj));
            return MMG_OK;
        }
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
    // No host synchronisation inside the chunk loop (a host source packs its chunks meanwhile): per-chunk events, read at the end.
    std::vector<cudaEvent_t> tev;
    struct TevGuard {
        std::vector<cudaEvent_t>& v;
        ~TevGuard() { for (cudaEvent_t e : v) cudaEventDestroy(e); }
    } tev_guard{tev};
    cudaStream_t ps = ctx->stream;                                  // stream of the pack kernels
    if (overlap) {
        MMG_TRY(ensure_side_stream(ctx));
        ps = ctx->stream2;
        // the resident block, the zeroed Gram and the cleared flag are ordered on the main stream
        MMG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        MMG_CUDA(ctx, cudaStreamWaitEvent(ps, ctx->ev0, 0));
    }
    int table_kb = -1;
    for (int64_t s0 = 0; s0 < snp_count; s0 += chunk) {
        const int64_t ci = s0 / chunk;
        const int64_t cnt = std::min(chunk, snp_count - s0);
        const int64_t kbytes = fp4 ? round_up(cnt, 256) * c / 2 : round_up(cnt, 128) * c;   // whole 128-byte K blocks
        int8_t* slot = ctx->pack + (overlap ? (ci & 1) * slot_bytes : 0);
        cudaEvent_t ev4[4];                                          // pack begin / end, Gram begin / end
        for (cudaEvent_t& e : ev4) {
            MMG_CUDA(ctx, cudaEventCreate(&e));
            tev.push_back(e);
        }
        if (src) {
            MMG_TRY(issue_copy(ci));
            MMG_CUDA(ctx, cudaStreamWaitEvent(ps, src->done[(size_t)ci], 0));
        }
        if (overlap && ci >= 2) MMG_CUDA(ctx, cudaStreamWaitEvent(ps, tev[(size_t)(4 * (ci - 2) + 3)], 0));   // the slot's previous Gram
        // ---- pack ----
        cudaEventRecord(ev4[0], ps);
        dim3 pgrid((unsigned)((fp4 ? round_up(cnt, 256) : round_up(cnt, 128)) / 128), (unsigned)((n + 63) / 64));   // SNPs >= cnt pack to zeros
        if (coding == MMG_CODING_BINARY) {
            if (fp4) pack_kmajor_kernel<0, true><<<pgrid, 256, 0, ps>>>(ctx->snps, ctx->pitch, snp_begin + s0, cnt, n, slot, p_pitch, ctx->flag_d);
            else pack_kmajor_kernel<0><<<pgrid, 256, 0, ps>>>(ctx->snps, ctx->pitch, snp_begin + s0, cnt, n, slot, p_pitch, ctx->flag_d);
        } else {
            if (fp4) pack_kmajor_kernel<1, true><<<pgrid, 256, 0, ps>>>(ctx->snps, ctx->pitch, snp_begin + s0, cnt, n, slot, p_pitch, ctx->flag_d);
            else pack_kmajor_kernel<1><<<pgrid, 256, 0, ps>>>(ctx->snps, ctx->pitch, snp_begin + s0, cnt, n, slot, p_pitch, ctx->flag_d);
        }
        MMG_TRY(launch_check(ctx, "pack_kmajor_kernel"));
        cudaEventRecord(ev4[1], ps);
        if (overlap) MMG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev4[1], 0));
        // ---- Gram ----
        const int accumulate = ctx->g_zero ? 0 : 1;
        if (impl == MMG_IMPL_TCGEN05) {
            CUtensorMap tmA, tmB;
            MMG_TRY(make_tmap_u8(ctx, &tmA, slot, kbytes, n, p_pitch, TC_BM));
            MMG_TRY(make_tmap_u8(ctx, &tmB, slot, kbytes, n, p_pitch, TC_BN / gram_cs));
            if (table_kb != (int)(kbytes / TC_BK)) {                // same table for every full chunk
                table_kb = (int)(kbytes / TC_BK);
                build_table(table_kb);
                MMG_TRY(ensure_tiles(ctx, table));
            }
            cudaEventRecord(ev4[2], ctx->stream);
            const int ngroups = (int)table.size() * gram_cs;
            const int pf = env_int("MMG_GRAM_PREFETCH", 0);         // L2 prefetch distance of the operand streams in K blocks; off: measured 34 -> 52 ms at 8
            if (pair) {                                             // tmB's box is this CTA's half of the B tile (128 rows: gram_cs == 2)
                GramEpiF4::Params ep4{ctx->G, g_pad, accumulate};
                cudaLaunchConfig_t cfg;
                cudaLaunchAttribute attr[1];
                const int nt = (int)table.size();
                gram_pair_config(ctx, cfg, attr, std::max(1, std::min(nt, gram_clusters)));
                cudaError_t e = cudaLaunchKernelEx(&cfg, gram_pair_kernel, tmA, tmB, (const TcTile*)ctx->tiles_d, nt, (uint64_t)L2_EVICT_NORMAL, ep4);
                ctx->launches += 1;
                if (e != cudaSuccess) return fail(ctx, MMG_ECUDA, "launch of gram_pair_kernel (grid %u) failed: %s", cfg.gridDim.x, cudaGetErrorString(e));
            } else if (fp4) {
                GramEpiF4::Params ep4{ctx->G, g_pad, accumulate};
                if (gram_cs == 2)
                    MMG_TRY((launch_tc_gemm<GramEpiF4, 2, TC_KIND_MXF4>(ctx, tmA, tmB, (const TcTile*)ctx->tiles_d, ngroups, 1, 1, 0, TC_BM, ep4, "tc_gemm_i8_kernel<GramEpiF4,2,mxf4>", L2_EVICT_NORMAL, L2_EVICT_NORMAL, pf)));
                else
                    MMG_TRY((launch_tc_gemm<GramEpiF4, 1, TC_KIND_MXF4>(ctx, tmA, tmB, (const TcTile*)ctx->tiles_d, ngroups, 1, 1, 0, 0, ep4, "tc_gemm_i8_kernel<GramEpiF4,1,mxf4>", L2_EVICT_NORMAL, L2_EVICT_NORMAL, pf)));
            } else {
                GramEpi::Params ep{ctx->G, g_pad, accumulate};
                if (gram_cs == 2)
                    MMG_TRY((launch_tc_gemm<GramEpi, 2>(ctx, tmA, tmB, (const TcTile*)ctx->tiles_d, ngroups, 1, 1, 0, TC_BM, ep, "tc_gemm_i8_kernel<GramEpi,2>", L2_EVICT_NORMAL, L2_EVICT_NORMAL, pf)));
                else
                    MMG_TRY((launch_tc_gemm<GramEpi, 1>(ctx, tmA, tmB, (const TcTile*)ctx->tiles_d, ngroups, 1, 1, 0, 0, ep, "tc_gemm_i8_kernel<GramEpi,1>", L2_EVICT_NORMAL, L2_EVICT_NORMAL, pf)));
            }
        } else {
            cudaEventRecord(ev4[2], ctx->stream);
            dim3 ggrid((unsigned)((n + 63) / 64), (unsigned)((n + 63) / 64));
            gram_simt_kernel<<<ggrid, 256, 0, ctx->stream>>>(slot, p_pitch, n, kbytes, ctx->G, g_pad, accumulate);
            MMG_TRY(launch_check(ctx, "gram_simt_kernel"));
        }
        cudaEventRecord(ev4[3], ctx->stream);
        ctx->g_zero = false;
    }
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));             // every pack was waited for by its Gram
    for (size_t i = 0; i + 3 < tev.size(); i += 4) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, tev[i], tev[i + 1]);
        pack_s += ms * 1e-3;
        cudaEventElapsedTime(&ms, tev[i + 2], tev[i + 3]);
        gram_ms += ms;
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

// direction 0: pack the valid 256 x 256 blocks of the Gram (block row <= block column) into one contiguous int32 buffer and
// return its device pointer and element count -- the buffer a multi-GPU caller all-reduces (half the bytes of the padded
// square); direction 1: unpack the (reduced) buffer back into the Gram.  Stream ordered, no host synchronisation.
int mmg_kinship_gram_tri(mmg_ctx* ctx, int direction, void** dptr, int64_t* count) {
    MMG_CHECK(ctx, ctx && ctx->G, "mmg_kinship_gram_tri: no Gram resident");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t T = ctx->g_pad / 256, slots = T * (T + 1) / 2;
    MMG_TRY(ensure_scratch(ctx, slots * 65536 * (int64_t)sizeof(int32_t)));
    StageTimer tm(ctx, "finalize");
    if (direction == 0)
        gram_tri_pack_kernel<false><<<(unsigned)slots, 256, 0, ctx->stream>>>(ctx->G, ctx->g_pad, (int32_t*)ctx->scratch);
    else
        gram_tri_pack_kernel<true><<<(unsigned)slots, 256, 0, ctx->stream>>>(ctx->G, ctx->g_pad, (int32_t*)ctx->scratch);
    MMG_TRY(launch_check(ctx, "gram_tri_pack_kernel"));
    if (dptr) *dptr = ctx->scratch;
    if (count) *count = slots * 65536;
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
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm(ctx, "finalize");
    const int n = (int)ctx->n;
    const int64_t T = (n + 31) / 32, tiles = T * (T + 1) / 2;
    DevBuf part;                                                 // per-tile (sum, trace) partials + {sum, trace, factor}
    MMG_CUDA(ctx, part.alloc(ctx->stream, (size_t)(2 * tiles + 3) * sizeof(double)));
    double* d_part = part.as<double>();
    double* d_sums = d_part + 2 * tiles;
    if (scaled) {
        if (coding == MMG_CODING_BINARY)
            kin_tile_sums_kernel<0><<<(unsigned)tiles, 256, 0, ctx->stream>>>(ctx->G, ctx->g_pad, n, (double)m_total, d_part);
        else
            kin_tile_sums_kernel<1><<<(unsigned)tiles, 256, 0, ctx->stream>>>(ctx->G, ctx->g_pad, n, (double)m_total, d_part);
        MMG_TRY(launch_check(ctx, "kin_tile_sums_kernel"));
        kin_final_sums_kernel<<<1, 1024, 0, ctx->stream>>>(d_part, tiles, n, d_sums);
        MMG_TRY(launch_check(ctx, "kin_final_sums_kernel"));
    }
    const double* d_scale = scaled ? d_sums + 2 : nullptr;
    if (coding == MMG_CODING_BINARY)
        kin_write_tile_kernel<0><<<(unsigned)tiles, 256, 0, ctx->stream>>>(ctx->G, ctx->g_pad, n, (double)m_total, d_scale, K->d, K->cols);
    else
        kin_write_tile_kernel<1><<<(unsigned)tiles, 256, 0, ctx->stream>>>(ctx->G, ctx->g_pad, n, (double)m_total, d_scale, K->d, K->cols);
    MMG_TRY(launch_check(ctx, "kin_write_tile_kernel"));
    if (scale_scalar) {
        *scale_scalar = 1.0;
        if (scaled) {                                           // only a caller that asks for the factor waits for it
            MMG_CUDA(ctx, cudaMemcpyAsync(scale_scalar, d_sums + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    return MMG_OK;
}

static __global__ void mirror_lower_to_upper_kernel(double* __restrict__ K, int64_t ld, int n) {
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

}  // extern "C"
