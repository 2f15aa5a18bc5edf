// libmixmogam_b200: stage 3 of the C ABI -- the SNP scan (int8 tcgen05 quadratic form, FP64 DMMA rotation), the
// phenotype-batched scan and the permutation scan.
#include "common.cuh"
#include "fdist.cuh"
#include "scan_dmma.cuh"
#include "scan_tc.cuh"
#include "scan_quad.cuh"

namespace mmg {

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
    cudaLaunchAttribute attr[2];
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
    // The per-wave barrier (sh.wave_sync) spins until every CTA of the grid has arrived: the grid is launched COOPERATIVELY, so the
    // runtime starts it only when all its CTAs can be co-resident (another kernel holding SMs -- the side-stream pre-pass, a second
    // process under MPS -- delays the launch instead of dead-locking the spin).  A driver that refuses the cooperative cluster launch
    // gets the scan without the barrier (correct, only more HBM traffic).
    cudaError_t e = cudaErrorUnknown;
    if (sh.wave_sync != nullptr && env_int("MMG_SCAN_COOP", 1) != 0) {
        attr[1].id = cudaLaunchAttributeCooperative;
        attr[1].val.cooperative = 1;
        cfg.numAttrs = 2;
        e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, sh, pa, pb, ep);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cfg.numAttrs = 1;
            sh.wave_sync = nullptr;
        }
    }
    if (e != cudaSuccess) e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, sh, pa, pb, ep);
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

void scan_init_attrs() {
    cudaFuncSetAttribute(tc_gemm_i8_kernel<QuadEpi, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<QuadEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<QuadEpi, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<OzakiEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<PermEpi, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<PermEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(scan_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SD_SMEM_BYTES);
    cudaFuncSetAttribute(scan_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SD_SMEM_BYTES);
}

}  // namespace mmg

using namespace mmg;

static __global__ void means_from_sums_kernel(const long long* __restrict__ sums, int64_t count, double inv_n, double* __restrict__ mu) {
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

static int launch_scan_dmma(mmg_ctx* ctx, bool perm, const ScanDmmaParams& prm_in) {
    ScanDmmaParams prm = prm_in;
    const int64_t blocks = (prm.row_count + SD_BM - 1) / SD_BM;
    // Short scans (the single-SNP calls of the stepwise callers, top-hit lists): with one CTA per 128-row block a 1-SNP scan at
    // n = 10k runs 2.6e10 flops on ONE SM (127 ms).  The column tiles of R are split over the idle SMs instead and the partial
    // moments are added in a fixed order by a finishing kernel (1-2 ms).
    DevBuf part;
    prm.nsplit = 0;
    prm.part = nullptr;
    const int NT = prm.n_out_pad / SD_BN;
    if (!perm && prm.mu == nullptr && blocks * 2 <= ctx->sm_count && NT > 1) {
        prm.nsplit = (int)std::min<int64_t>(NT, ctx->sm_count / blocks);
        MMG_CUDA(ctx, part.alloc(ctx->stream, (size_t)(2 * prm.nsplit) * prm.row_count * sizeof(double)));
        prm.part = part.as<double>();
    }
    const int64_t items = blocks * std::max(1, prm.nsplit);
    const int grid = (int)std::min<int64_t>(items, ctx->sm_count);
    cudaEventRecord(ctx->kev0, ctx->stream);
    if (perm)
        scan_dmma_kernel<true><<<grid, SD_THREADS, SD_SMEM_BYTES, ctx->stream>>>(prm);
    else
        scan_dmma_kernel<false><<<grid, SD_THREADS, SD_SMEM_BYTES, ctx->stream>>>(prm);
    MMG_TRY(launch_check(ctx, "scan_dmma_kernel"));
    if (prm.nsplit > 1) {
        scan_split_finish_kernel<<<(unsigned)((prm.row_count + 127) / 128), 128, 0, ctx->stream>>>(prm.part, prm.nsplit, prm.row_count, prm.h0_rss, prm.n_p,
                                                                                                 prm.lbeta, prm.xx, prm.xy, prm.rss, prm.f, prm.p, prm.var_perc);
        MMG_TRY(launch_check(ctx, "scan_split_finish_kernel"));
    }
    cudaEventRecord(ctx->kev1, ctx->stream);
    return MMG_OK;
}

// max |x| over the resident genotype block (zero padding included), 16 bytes per thread per step
static __global__ void snps_absmax_kernel(const uint4* __restrict__ p, int64_t n16, int* __restrict__ out) {
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

// max |r| of a contiguous FP64 array -> bits of a non-negative double (grid-stride, 16-byte loads)
static __global__ void __launch_bounds__(256) array_amax_kernel(const double2* __restrict__ p, int64_t n2, const double* __restrict__ last,
                                                                unsigned long long* __restrict__ amax_bits) {
    double m = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = p[i];
        m = fmax(m, fmax(fabs(v.x), fabs(v.y)));
    }
    if (last && blockIdx.x == 0 && threadIdx.x == 0) m = fmax(m, fabs(*last));     // odd element count
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(amax_bits, (unsigned long long)__double_as_longlong(m));
}

// number of 256 x 256 blocks in the lower triangle of the padded n x n quadratic form
static int64_t qa_slots(int64_t n) {
    const int64_t T = round_up(n, QA_TILE) / QA_TILE;
    return T * (T + 1) / 2;
}

// A = R'R as exact int8 digit-plane products on the tensor cores (scan_tc.cuh, "A = R'R on the int8 tensor cores"), written in
// the PACKED block layout (qa_index): blocks [slot_begin, slot_begin + slot_count) of the lower triangle go to A_slots, which
// points at block slot_begin.  Every block costs the same (the contraction runs over all n_out rows of R), so ranks that take
// equal slot ranges are balanced.  *err_out: absolute error bound of the entries.
static int quad_form_int8(mmg_ctx* ctx, const MmgMat* R, double* A_slots, int64_t slot_begin, int64_t slot_count, unsigned long long* d_amax,
                          double* err_out) {
    const int64_t n = ctx->n, n_out = R->rows;
    const int64_t n_padM = round_up(n, QA_TILE);
    HostTimer ht_total(ctx, "host_qf_total");
    MMG_CHECK(ctx, n_out < 131072, "R'R on the int8 pipe: contraction too long for exact int32 accumulation");
    MMG_CUDA(ctx, cudaMemsetAsync(d_amax, 0, sizeof(unsigned long long), ctx->stream));
    {
        const int64_t cnt = n_out * n;
        array_amax_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((const double2*)R->d, cnt / 2, (cnt & 1) ? R->d + cnt - 1 : nullptr, d_amax);
        MMG_TRY(launch_check(ctx, "array_amax_kernel"));
    }
    double rmax = 0.0;
    MMG_CUDA(ctx, cudaMemcpyAsync(&rmax, d_amax, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!std::isfinite(rmax)) return fail(ctx, MMG_EVALUE, "scan: the rotation is not finite (max |r| = %g)", rmax);
    const int F = digit256_exponent(rmax);
    *err_out = ozaki_error_bound(n_out) * ldexp(1.0, 2 * F);
    slot_count = std::max<int64_t>(0, std::min(slot_count, qa_slots(n) - slot_begin));
    if (slot_count == 0) return MMG_OK;
    const int64_t op_pitch = round_up(n_out, TC_BK);
    WsBuf Op;
    MMG_TRY(ws_get(ctx, MMG_WS_OZAKI_PLANES, OZ_PLANES * n_padM * op_pitch, &Op.p));
    MMG_CUDA(ctx, cudaMemsetAsync(Op.p, 0, (size_t)OZ_PLANES * n_padM * op_pitch, ctx->stream));
    ozaki_planes_kernel<<<dim3((unsigned)((n + 31) / 32), (unsigned)((n_out + 31) / 32)), 256, 0, ctx->stream>>>(
        R->d, n, (int)n_out, (int)n, ldexp(1.0, -F), Op.as<int8_t>(), n_padM, op_pitch);
    MMG_TRY(launch_check(ctx, "ozaki_planes_kernel"));
    MMG_CUDA(ctx, cudaMemsetAsync(A_slots, 0, (size_t)slot_count * QA_TILE_ELEMS * sizeof(double), ctx->stream));
    // tile table: one entry per block (J, I), I <= J, in slot order: the row-tile pair (2 J, 2 J + 1) x column tile I, its 28
    // (p, q) plane pairs level by level (p + q = lv share the weight 2^2F 256^-(lv+2)).  The lv + 1 pairs of a level are CHAINED
    // into one int32 accumulator (7 n_out 128^2 < 2^31), so a block is read-modify-written 7 times, not 28 -- the FP64
    // read-modify-write of the epilogue (a row per thread, 32 sectors per access) was what bounded this kernel.  (Walking the
    // blocks in L2-sized groups instead of slot order was measured and changed nothing: operand traffic is not the bound.)
    const int KB = (int)(op_pitch / TC_BK);
    std::vector<TcTile> tiles;
    int entries = 0, per_entry = 0;
    const bool chain = env_int("MMG_OZAKI_CHAIN", 1) != 0 && (double)OZ_LEVELS * (double)n_out * 16384.0 < 2147483647.0;
    int J = (int)((std::sqrt(8.0 * (double)slot_begin + 1.0) - 1.0) * 0.5);
    while ((int64_t)(J + 1) * (J + 2) / 2 <= slot_begin) ++J;
    while ((int64_t)J * (J + 1) / 2 > slot_begin) --J;
    int I = (int)(slot_begin - (int64_t)J * (J + 1) / 2);
    for (int64_t sl = 0; sl < slot_count; ++sl) {
        per_entry = 0;
        for (int lv = 0; lv < OZ_LEVELS; ++lv)
            for (int p = 0; p <= lv; ++p) {
                const int q = lv - p;
                if (p >= OZ_PLANES || q >= OZ_PLANES) continue;
                TcTile tl{};
                tl.m0 = (int)((int64_t)p * n_padM + (int64_t)(2 * J) * TC_BM);
                tl.n0 = (int)((int64_t)q * n_padM + (int64_t)I * TC_BN);
                tl.kb0 = 0;
                tl.kb1 = KB;
                tl.aux0 = p;
                tl.aux1 = q;
                tl.flags = (p < lv && chain) ? TC_TILE_CHAIN : 0;       // the pair (lv, 0) ends the chain of its level
                tiles.push_back(tl);
                ++per_entry;
            }
        ++entries;
        if (++I > J) { I = 0; ++J; }
    }
    {
        HostTimer ht(ctx, "host_qf_tiles");
        MMG_TRY(ensure_tiles(ctx, tiles));
    }
    CUtensorMap tmA, tmB;
    MMG_TRY(make_tmap_u8(ctx, &tmA, Op.p, op_pitch, (int64_t)OZ_PLANES * n_padM, op_pitch, TC_BM));
    MMG_TRY(make_tmap_u8(ctx, &tmB, Op.p, op_pitch, (int64_t)OZ_PLANES * n_padM, op_pitch, TC_BN / 2));
    OzakiEpi::Params ep{};
    ep.A = A_slots;
    ep.ld = 0;
    ep.n_padM = n_padM;
    ep.packed = 1;
    ep.slot0 = slot_begin;
    for (int sl = 0; sl < 2 * OZ_PLANES; ++sl) ep.w[sl] = ldexp(1.0, 2 * F - 8 * (sl + 2));
    {
        StageTimer tg(ctx, "qf_gemm");
        MMG_TRY((launch_tc_gemm<OzakiEpi, 2>(ctx, tmA, tmB, (const TcTile*)ctx->tiles_d, entries * 2, per_entry, per_entry, 0, TC_BM, ep,
                                             "tc_gemm_i8_kernel<OzakiEpi,2>")));
    }
    return MMG_OK;
}

// MMG_QUAD_A = int8 (default: exact digit-plane products on the int8 tensor pipe) | dsyrk (cuBLAS, FP64 tensor pipe)
static thread_local bool g_quad_force_dsyrk = false;     // set while a scan is repeated with the FP64 product (see scan_tc_run)
static bool quad_a_int8() {
    if (g_quad_force_dsyrk) return false;
    const char* e = getenv("MMG_QUAD_A");
    return !(e && strcmp(e, "dsyrk") == 0);
}
static bool quad_use_int8() { return quad_a_int8() && scan_tc_tol() >= 1e-9; }

// A quadratic form handed to the scan: dense row-major (lower triangle valid) or packed 256 x 256 blocks (qa_index)
struct QuadA {
    const double* d = nullptr;
    int64_t ld = 0;
    bool packed = false;
    double err = 0.0;         // absolute error bound of its entries (0: formed in FP64)
};

// Digit planes of the strict lower triangle of A = R'R (doubled) into Bq (S planes of [n_padN x ldq]), diag(A) into
// d_dg and (when the rotation is given and d_y set) v = R'y into d_v; returns the binary exponent E used for the scaling.
// `given` (optional): the caller already holds A.  Otherwise A is formed from R into A_work: packed blocks by the int8
// digit-plane product, or dense [lda_work x lda_work] by cuBLAS dsyrk (A_work must hold max of the two layouts).
static int quad_prepare(mmg_ctx* ctx, const MmgMat* R, const double* d_y, const QuadA* given, double* A_work, int64_t lda_work,
                        unsigned long long* d_amax, int S, int8_t* Bq, int64_t n_padN, int64_t ldq, double* d_v, double* d_dg, int* E_out,
                        double* errA_out) {
    const int64_t n = ctx->n;
    const double one = 1.0, zero = 0.0;
    QuadA A;
    if (given) {
        A = *given;
    } else {
        const int64_t n_out = R->rows;
        if (quad_use_int8()) {
            MMG_TRY(quad_form_int8(ctx, R, A_work, 0, qa_slots(n), d_amax, &A.err));
            A.packed = true;
        } else {
            // A = R'R: R row-major [n_out x n] is the column-major n x n_out matrix Rc; column-major UPPER of Rc Rc'
            // is the row-major LOWER triangle A[j][i], i <= j -- exactly the operand the slices are cut from.
            MMG_CUBLAS(ctx, cublasDsyrk(ctx->cublas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, (int)n, (int)n_out, &one, R->d, (int)n, &zero, A_work,
                                        (int)lda_work));
            A.ld = lda_work;
        }
        // v = R' y~  (x~.y~ = x.v)
        if (d_y) MMG_CUBLAS(ctx, cublasDgemv(ctx->cublas, CUBLAS_OP_N, (int)n, (int)n_out, &one, R->d, (int)n, d_y, 1, &zero, d_v, 1));
        A.d = A_work;
    }
    *errA_out = A.err;
    MMG_CUDA(ctx, cudaMemsetAsync(d_amax, 0, sizeof(unsigned long long), ctx->stream));
    dim3 agrid(8, (unsigned)n);
    if (A.packed) quad_amax_kernel<true><<<agrid, 256, 0, ctx->stream>>>(A.d, A.ld, (int)n, d_amax);
    else quad_amax_kernel<false><<<agrid, 256, 0, ctx->stream>>>(A.d, A.ld, (int)n, d_amax);
    MMG_TRY(launch_check(ctx, "quad_amax_kernel"));
    double amax = 0.0;
    MMG_CUDA(ctx, cudaMemcpyAsync(&amax, d_amax, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!std::isfinite(amax)) return fail(ctx, MMG_EVALUE, "scan: R'R is not finite (max |a| = %g)", amax);
    const int E = digit256_exponent(amax);                // |2 a| 2^-E <= 0.498 (a diagonal R'R has no off-diagonal digits at all)
    dim3 sgrid((unsigned)((n + 255) / 256), (unsigned)n);
    if (A.packed) quad_slice_kernel<true><<<sgrid, 256, 0, ctx->stream>>>(A.d, A.ld, (int)n, ldexp(1.0, -E), S, Bq, n_padN, ldq, d_dg);
    else quad_slice_kernel<false><<<sgrid, 256, 0, ctx->stream>>>(A.d, A.ld, (int)n, ldexp(1.0, -E), S, Bq, n_padN, ldq, d_dg);
    MMG_TRY(launch_check(ctx, "quad_slice_kernel"));
    *E_out = E;
    return MMG_OK;
}

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
        // MMG_SCAN_DEFER_STATS (default 1): the epilogue stores x~.x~ only and quad_finish_kernel derives RSS / F / p and the certification
        // maximum afterwards -- nothing but accumulator draining is left between two SNP groups of a CTA
        const bool defer = env_int("MMG_SCAN_DEFER_STATS", 1) != 0 && ep.xx != nullptr && ep.pre_xy != nullptr;
        QuadEpi::Params epk = ep;
        epk.defer = defer ? 1 : 0;
        if (pair128) MMG_TRY(launch_scan_quad_pair128(ctx, panel, tmA, tmB, sh, epk));
        else if (n128) MMG_TRY(launch_scan_quad_n128(ctx, panel, tmA, tmB, sh, epk));
        else if (pair) MMG_TRY(launch_scan_quad_pair(ctx, panel, tmA, tmB, sh, epk));
        else if (cs == 8) MMG_TRY(launch_scan_quad_cs<8>(ctx, panel, tmA, tmB, sh, epk));
        else if (cs == 4) MMG_TRY(launch_scan_quad_cs<4>(ctx, panel, tmA, tmB, sh, epk));
        else if (cs == 2) MMG_TRY(launch_scan_quad_cs<2>(ctx, panel, tmA, tmB, sh, epk));
        else MMG_TRY(launch_scan_quad_cs<1>(ctx, panel, tmA, tmB, sh, epk));
        if (defer) {
            QuadEpi::Params epf = ep;
            epf.defer = 0;
            const int64_t total = (int64_t)T * snp_count;
            quad_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(epf, T);
            MMG_TRY(launch_check(ctx, "quad_finish_kernel"));
        }
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
// A_given / v_given (T = 1 only): the quadratic form A = R'R (dense or packed blocks, lower triangle valid) and v = R'y~ were
// formed by the caller -- the multi-GPU path lets every rank form a range of the blocks of A on the int8 pipe and all-gathers
// them (parallel.py) instead of repeating the n^3 product on every rank; Rs and V are then unused.  v_given is a device pointer
// when v_on_device, a host pointer otherwise.
static int scan_tc_run(mmg_ctx* ctx, int T, const MmgMat* const* Rs, const double* V, const double* h0_rss, double n_p, double lbeta,
                       int64_t snp_begin, int64_t snp_count, double* d_xx, double* d_xy, double* d_rss, double* d_f, double* d_p,
                       double* d_vp, const QuadA* A_given = nullptr, const double* v_given = nullptr, bool v_on_device = false) {
    const int64_t n = ctx->n, n_out = A_given ? 1 : Rs[0]->rows;
    MMG_TRY(scan_tc_check_domain(ctx));
    const int S_fixed = scan_tc_fixed_slices();
    const int S_alloc = S_fixed ? S_fixed : QS_AUTO_PLANES;
    const double tol = scan_tc_tol();
    const int64_t n_padN = round_up(n, TC_BN), ldq = round_up(n, TC_BK);
    const int64_t plane = n_padN * ldq;
    MMG_CHECK(ctx, (int64_t)T * S_alloc * n_padN < (1ll << 31), "scan: too many phenotype slices for one launch");
    WsBuf A, Bq, vec, pre;
    const int64_t lda_work = round_up(n, TC_BN);               // dense (dsyrk) layout of A; the packed int8 product needs less
    if (!A_given)
        MMG_TRY(ws_get(ctx, MMG_WS_QUAD_A, quad_use_int8() ? qa_slots(n) * QA_TILE_ELEMS * (int64_t)sizeof(double)
                                                             : lda_work * lda_work * (int64_t)sizeof(double), &A.p));
    MMG_TRY(ws_get(ctx, MMG_WS_QUAD_BQ, (int64_t)T * S_alloc * plane, &Bq.p));
    // vec: v[T][n_padN] | dg[T][n_padN] | y[T][n_out] | h0[T] | escale[T] | bscale[T] | amax | rho | wave counter
    const int64_t nd = 2 * (int64_t)T * n_padN + (int64_t)T * n_out + 3 * T + 3;
    MMG_TRY(ws_get(ctx, MMG_WS_SCAN_VEC, nd * (int64_t)sizeof(double), &vec.p));
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
    if (A_given) {
        if (v_given) MMG_CUDA(ctx, cudaMemcpyAsync(d_v, v_given, (size_t)n * sizeof(double), v_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    } else {
        MMG_CUDA(ctx, cudaMemcpyAsync(d_y, V, (size_t)T * n_out * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    MMG_CUDA(ctx, cudaMemcpyAsync(d_h0, h0_rss, (size_t)T * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<double> escale((size_t)T), bscale((size_t)T), errA((size_t)T, 0.0);
    // linear pre-pass: x.v_t, sum_j A_jj x_j^2 and ||x||_1 of every SNP in range, one stream over the genotypes.  When the rotation
    // is at hand, v_t = R_t'y~_t and diag(A_t) = column sums of squares of R_t are formed first and the pre-pass runs on a side
    // stream underneath the n^3 product A = R'R (FP64 / HBM work beside int8 tensor work); MMG_SCAN_OVERLAP=0 serialises them.
    MMG_TRY(ws_get(ctx, MMG_WS_SCAN_PRE, ((2 * T + 1) * snp_count + (int64_t)T * n_padN) * (int64_t)sizeof(double), &pre.p));
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
    const bool early = A_given && T == 1 && ctx->early_prepass && ctx->early_begin == snp_begin && ctx->early_count == snp_count;
    ctx->early_prepass = false;
    if (early) {
        // mmg_scan_prepass_begin launched the pre-pass of exactly these rows on the side stream, underneath whatever ran since
        // (the block-wise R'R product and its all-gather in the multi-GPU path): join it, the outputs are in `pre`
        MMG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ov1, 0));
    }
    if (!A_given && env_int("MMG_SCAN_OVERLAP", 1)) {
        MMG_TRY(ensure_side_stream(ctx));
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
        MMG_TRY(quad_prepare(ctx, A_given ? nullptr : Rs[t], pre_launched ? nullptr : d_y + (int64_t)t * n_out, A_given, A.as<double>(), lda_work,
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
    } else if (early) {
        // joined above
    } else {
        MMG_CHECK(ctx, !A_given || v_given, "scan: v = R'y is needed (no pre-pass was launched ahead for these rows)");
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
    // bound(S) / bound(S0), worst phenotype: 256 per plane down to the floor set by the error of A itself
    auto bound_ratio = [&](int planes, int planes0) {
        double r = 0.0;
        for (int t = 0; t < T; ++t) {
            const double c = 0.5 * DIGIT256_REM * escale[t];
            r = std::max(r, (c * ldexp(1.0, -8 * planes) + errA[(size_t)t]) / (c * ldexp(1.0, -8 * planes0) + errA[(size_t)t]));
        }
        return r;
    };

    int S = S_fixed ? S_fixed : S_alloc;                        // short scans: no pilot, every plane that was cut
    if (!S_fixed && snp_count >= 4 * QS_PILOT_ROWS) {
        const bool use_hint = env_int("MMG_SCAN_HINT", 1) != 0 && ctx->scan_slices_hint > 0 && ctx->scan_hint_n == n;
        if (use_hint) {
            // the previous scan of this shape certified its bound with this many planes (with head-room): start there.  The full
            // launch measures the bound over every SNP anyway and is repeated with one more plane if one violates it.
            S = std::min(S_alloc, ctx->scan_slices_hint);
        } else {
            // pilot: bound of the first rows with few planes; the bound scales exactly by 256 per plane
            QuadEpi::Params pp = ep;
            pp.xx = pp.xy = pp.rss = pp.f = pp.p = pp.var_perc = nullptr;
            MMG_TRY(set_bscale(QS_PILOT_PLANES));
            MMG_TRY(scan_tc_launch(ctx, T, QS_PILOT_PLANES, S_alloc, Bq.p, n_padN, ldq, snp_begin, QS_PILOT_ROWS, pp, d_wave));
            double rho = 0.0;
            MMG_TRY(read_rho(&rho));
            S = QS_PILOT_PLANES;
            while (S < S_alloc && rho * bound_ratio(S, QS_PILOT_PLANES) * QS_PILOT_HEADROOM > tol) ++S;
        }
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
    if (!S_fixed && snp_count >= 4 * QS_PILOT_ROWS) {
        // next call: one plane fewer when the measured bound says it would still certify with the pilot's head-room
        ctx->scan_hint_n = n;
        ctx->scan_slices_hint = (S > 1 && rho <= tol && rho * bound_ratio(S - 1, S) * QS_PILOT_HEADROOM <= tol) ? S - 1 : S;
    }
    // a tolerance below what the int8 product of A can certify (its own error bound is a floor of ~1e-10 relative at n = 10k):
    // once more with A from the FP64 dsyrk
    bool int8_floor = false;
    for (double e : errA) int8_floor |= e > 0.0;
    if (!S_fixed && rho > tol && int8_floor && !g_quad_force_dsyrk && !A_given) {
        g_quad_force_dsyrk = true;
        const int rc = scan_tc_run(ctx, T, Rs, V, h0_rss, n_p, lbeta, snp_begin, snp_count, d_xx, d_xy, d_rss, d_f, d_p, d_vp);
        g_quad_force_dsyrk = false;
        return rc;
    }
    if (!S_fixed && rho > tol && !env_int("MMG_TC_ALLOW_UNCERTIFIED", 0))
        return fail(ctx, MMG_EVALUE, "int8 scan: the certified truncation bound on x~.x~ is %.3g after %d digit planes, above the tolerance %.3g "
                    "(MMG_TC_TOL); use scan_impl='dmma', a larger tolerance, or MMG_TC_ALLOW_UNCERTIFIED=1 to take the result as is", rho, S, tol);
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

// x~.x~, x~.V[0] and the statistics of SNP rows [snp_begin, +snp_count) of the resident block, left on the device in `out`:
// xx | xy | rss | f | p | var_perc (snp_count doubles each), then -- with want_dots -- x~.V[v] for v < nv ([nv][snp_count]).
// What MMG_IMPL_AUTO means for a scan of snp_count SNPs: the int8 quadratic-form scan pays a one-off n^3 product (A = R'R,
// digit planes, pilot: ~15 ms at n = 10k) before its 0.14 us per SNP; the FP64 tensor-core scan rotates each SNP directly
// (2 n^2 flops, ~8 us at n = 10k) with no set-up.  Short scans -- the single-SNP calls of the stepwise / MLMM callers
// (linear_models.py:2720,2825), the top-hit lists -- therefore take the FP64 kernel: crossover near snp_count = n / 8.
static int resolve_scan_impl(mmg_ctx* ctx, int impl, int64_t snp_count) {
    if (impl != MMG_IMPL_AUTO) return impl;
    if (getenv("MMG_SCAN_IMPL")) return env_impl("MMG_SCAN_IMPL", MMG_IMPL_TCGEN05);
    return snp_count * 8 <= ctx->n ? MMG_IMPL_DMMA : MMG_IMPL_TCGEN05;
}

static int scan_device(mmg_ctx* ctx, MmgMat* R, const double* V, int nv, double h0_rss, double n_p, int impl, int64_t snp_begin,
                       int64_t snp_count, double lbeta, DevBuf& out, bool want_dots) {
    const int64_t n = ctx->n, n_out = R->rows;
    MMG_CUDA(ctx, out.alloc(ctx->stream, (size_t)(6 + nv) * snp_count * sizeof(double)));
    double* d_xx = out.as<double>();
    double* d_xy = d_xx + snp_count;
    double* d_rss = d_xy + snp_count;
    double* d_f = d_rss + snp_count;
    double* d_p = d_f + snp_count;
    double* d_vp = d_p + snp_count;
    double* d_dots = d_vp + snp_count;
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

    if (want_dots) {
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
    return MMG_OK;
}

// lstsq([h0_X, x~], y~res) per SNP (linear_models.py:1323, with_betas=True) from its normal equations: with A = h0_X'h0_X,
// c0 = h0_X'y~res, b = x~'h0_X, the Schur complement s = x~.x~ - b'A^-1 b gives beta_x = (x~.y~ - b'A^-1 c0) / s,
// beta_0 = A^-1 c0 - A^-1 b beta_x, rss = y~.y~ - beta_0'c0 - beta_x x~.y~; then F and p (:1345-1349).  A rank-deficient
// column (s ~ 0) or an exact zero residue keeps the null fit (`if rss:`, :1325).  One thread per SNP.
struct BetasParams {
    const double *xx, *dots;     // [snp_count], [1 + q0][snp_count]
    int64_t snp_count;
    int q0;
    double Ainv[16 * 16], c0[16], a0[16], h0_betas[16];
    double yy, h0_rss, n_p, lbeta;
    double *rss, *f, *p, *var_perc, *betas;      // betas: [snp_count][q0 + 1]
};
static __global__ void __launch_bounds__(256) betas_finish_kernel(const BetasParams prm) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= prm.snp_count) return;
    const int q0 = prm.q0;
    const double xx = prm.xx[s], xy = prm.dots[s];
    double Ab[16];
    double bAb = 0.0, Abc = 0.0;
    for (int i = 0; i < q0; ++i) {
        double a = 0.0;
        for (int j = 0; j < q0; ++j) a = fma(prm.Ainv[i * 16 + j], prm.dots[(int64_t)(1 + j) * prm.snp_count + s], a);
        Ab[i] = a;
        bAb = fma(a, prm.dots[(int64_t)(1 + i) * prm.snp_count + s], bAb);
        Abc = fma(a, prm.c0[i], Abc);
    }
    const double sc = xx - bAb;
    const bool ok = sc > 1e-12 * fmax(xx, 1e-300);
    double rss = prm.h0_rss;
    bool good = false;
    double bx = 0.0;
    if (ok) {
        bx = (xy - Abc) / sc;
        double fit = bx * xy;
        for (int i = 0; i < q0; ++i) fit = fma(prm.a0[i] - Ab[i] * bx, prm.c0[i], fit);
        const double r = prm.yy - fit;
        if (r != 0.0) {
            rss = r;
            good = true;
        }
    }
    double* bo = prm.betas + s * (q0 + 1);
    if (good) {
        for (int i = 0; i < q0; ++i) bo[i] = prm.a0[i] - Ab[i] * bx;
        bo[q0] = bx;
    } else {
        for (int i = 0; i < q0; ++i) bo[i] = prm.h0_betas[i];
        bo[q0] = nan("");                           // marks "null fit kept": the host hands out h0_betas for this SNP
    }
    const double ratio = prm.h0_rss / rss;
    const double f = (ratio - 1.0) * prm.n_p;
    prm.rss[s] = rss;
    prm.f[s] = f;
    prm.var_perc[s] = 1.0 - 1.0 / ratio;
    prm.p[s] = f_sf(fmax(f, 0.0), 1.0, prm.n_p, prm.lbeta);
}

extern "C" {

int mmg_emmax_scan_f64(mmg_ctx* ctx, mmg_mat Rh, const double* V, int nv, double h0_rss, double n_p, int impl, int64_t snp_begin,
                       int64_t snp_count, double* ps, double* f_stats, double* rss, double* var_perc, double* xx, double* dots) {
    MmgMat* R = ctx ? get_mat(ctx, Rh) : nullptr;
    MMG_CHECK(ctx, R && ctx->snps, "mmg_emmax_scan_f64: need resident genotypes and R");
    MMG_CHECK(ctx, R->cols == ctx->n, "R must have n = %lld columns (has %lld)", (long long)ctx->n, (long long)R->cols);
    MMG_CHECK(ctx, V && nv >= 1 && nv <= 16, "need 1..16 rotated-space vectors (V[0] = residual phenotype)");
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count > 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    impl = resolve_scan_impl(ctx, impl, snp_count);
    MMG_CHECK(ctx, impl == MMG_IMPL_DMMA || impl == MMG_IMPL_TCGEN05, "unsupported impl %d for the scan", impl);
    ctx->last_scan_impl = impl;
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const double lbeta = lbeta_host(0.5 * n_p, 0.5);
    DevBuf out;
    MMG_TRY(scan_device(ctx, R, V, nv, h0_rss, n_p, impl, snp_begin, snp_count, lbeta, out, dots != nullptr));
    double* d_xx = out.as<double>();
    double* d_rss = d_xx + 2 * snp_count;
    double* d_f = d_rss + snp_count;
    double* d_p = d_f + snp_count;
    double* d_vp = d_p + snp_count;
    double* d_dots = d_vp + snp_count;
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

// with_betas=True (linear_models.py:1323): the scan with R = H (no projection), V = [y~res; h0_X'] and the per-SNP
// (q0 + 1) x (q0 + 1) least squares finished on the device (betas_finish_kernel).  Ainv = (h0_X'h0_X)^-1 [q0 x q0], c0 = h0_X'y~res,
// yy = y~res.y~res, h0_betas [q0].  Outputs (host): ps, f_stats, rss, var_perc [snp_count], betas [snp_count x (q0 + 1)] -- the last
// entry of a row is NaN where the SNP kept the null fit (rank-deficient column or zero residue).
int mmg_emmax_scan_betas_f64(mmg_ctx* ctx, mmg_mat Rh, const double* V, int q0, const double* Ainv, const double* c0, double yy,
                             const double* h0_betas, double h0_rss, double n_p, int impl, int64_t snp_begin, int64_t snp_count,
                             double* ps, double* f_stats, double* rss, double* var_perc, double* betas) {
    MmgMat* R = ctx ? get_mat(ctx, Rh) : nullptr;
    MMG_CHECK(ctx, R && ctx->snps, "mmg_emmax_scan_betas_f64: need resident genotypes and R");
    MMG_CHECK(ctx, R->cols == ctx->n, "R must have n = %lld columns (has %lld)", (long long)ctx->n, (long long)R->cols);
    MMG_CHECK(ctx, V && q0 >= 1 && q0 <= 15 && Ainv && c0 && h0_betas, "need V = [y~res; h0_X'] with 1..15 fixed-effect columns");
    MMG_CHECK(ctx, ps && f_stats && rss && var_perc && betas, "all outputs are required");
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count > 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    impl = resolve_scan_impl(ctx, impl, snp_count);
    MMG_CHECK(ctx, impl == MMG_IMPL_DMMA || impl == MMG_IMPL_TCGEN05, "unsupported impl %d for the scan", impl);
    ctx->last_scan_impl = impl;
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const double lbeta = lbeta_host(0.5 * n_p, 0.5);
    DevBuf out, bbuf;
    MMG_TRY(scan_device(ctx, R, V, 1 + q0, h0_rss, n_p, impl, snp_begin, snp_count, lbeta, out, true));
    MMG_CUDA(ctx, bbuf.alloc(ctx->stream, (size_t)snp_count * (q0 + 1) * sizeof(double)));
    BetasParams bp{};
    bp.xx = out.as<double>();
    double* d_rss = out.as<double>() + 2 * snp_count;
    double* d_f = d_rss + snp_count;
    double* d_p = d_f + snp_count;
    double* d_vp = d_p + snp_count;
    bp.dots = d_vp + snp_count;
    bp.snp_count = snp_count;
    bp.q0 = q0;
    for (int i = 0; i < q0; ++i) {
        bp.c0[i] = c0[i];
        bp.h0_betas[i] = h0_betas[i];
        double a = 0.0;
        for (int j = 0; j < q0; ++j) {
            bp.Ainv[i * 16 + j] = Ainv[i * q0 + j];
            a += Ainv[i * q0 + j] * c0[j];
        }
        bp.a0[i] = a;
    }
    bp.yy = yy;
    bp.h0_rss = h0_rss;
    bp.n_p = n_p;
    bp.lbeta = lbeta;
    bp.rss = d_rss;
    bp.f = d_f;
    bp.p = d_p;
    bp.var_perc = d_vp;
    bp.betas = bbuf.as<double>();
    {
        StageTimer tm(ctx, "scan");
        betas_finish_kernel<<<(unsigned)((snp_count + 255) / 256), 256, 0, ctx->stream>>>(bp);
        MMG_TRY(launch_check(ctx, "betas_finish_kernel"));
    }
    StageTimer tm2(ctx, "d2h");
    const size_t bytes = snp_count * sizeof(double);
    MMG_CUDA(ctx, cudaMemcpyAsync(ps, d_p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(f_stats, d_f, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(rss, d_rss, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(var_perc, d_vp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(betas, bbuf.p, (size_t)(q0 + 1) * bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

// Real-valued genotype rows (imputed dosages; the reference's scan takes any numeric row, linear_models.py:1317): the FP64
// tensor-core scan with the A operand staged as FP64.  The rows are not made resident: host chunks of <= 16384 rows go through
// one device buffer.
int mmg_emmax_scan_rows_f64(mmg_ctx* ctx, mmg_mat Rh, const double* V, int nv, double h0_rss, double n_p, const double* xs, int64_t m,
                            int64_t ld, double* ps, double* f_stats, double* rss, double* var_perc, double* xx, double* dots) {
    MmgMat* R = ctx ? get_mat(ctx, Rh) : nullptr;
    MMG_CHECK(ctx, R && xs && m > 0, "mmg_emmax_scan_rows_f64: need R and the genotype rows");
    MMG_CHECK(ctx, ld >= R->cols, "row stride %lld shorter than n = %lld", (long long)ld, (long long)R->cols);
    MMG_CHECK(ctx, V && nv >= 1 && nv <= 16, "need 1..16 rotated-space vectors (V[0] = residual phenotype)");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = R->cols, n_out = R->rows;
    const double lbeta = lbeta_host(0.5 * n_p, 0.5);
    const int64_t chunk = std::min<int64_t>(m, 16384);
    const int64_t pitch = round_up(n, SD_BK);
    static bool attr_set = false;
    if (!attr_set) {
        MMG_CUDA(ctx, cudaFuncSetAttribute(scan_dmma_kernel<false, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, sd_smem_bytes<double>()));
        attr_set = true;
    }
    StageTimer tm(ctx, "scan");
    DevBuf Rp, Vp, Xd, out, Vd, Wd;
    int64_t rows_pad = 0, ldr = 0;
    MMG_TRY(pad_matrix(ctx, R, Rp, &rows_pad, &ldr));
    MMG_CUDA(ctx, Vp.alloc(ctx->stream, (size_t)rows_pad * sizeof(double)));
    MMG_CUDA(ctx, cudaMemsetAsync(Vp.p, 0, (size_t)rows_pad * sizeof(double), ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(Vp.p, V, n_out * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, Xd.alloc(ctx->stream, (size_t)chunk * pitch * sizeof(double)));
    MMG_CUDA(ctx, cudaMemsetAsync(Xd.p, 0, (size_t)chunk * pitch * sizeof(double), ctx->stream));
    MMG_CUDA(ctx, out.alloc(ctx->stream, (size_t)(6 + nv) * chunk * sizeof(double)));
    double* d_xx = out.as<double>();
    double* d_xy = d_xx + chunk;
    double* d_rss = d_xy + chunk;
    double* d_f = d_rss + chunk;
    double* d_p = d_f + chunk;
    double* d_vp = d_p + chunk;
    double* d_dots = d_vp + chunk;
    if (dots) {
        // x~.V[v] = x.(R' V[v]):  W = V R
        MMG_CUDA(ctx, Vd.alloc(ctx->stream, (size_t)nv * n_out * sizeof(double)));
        MMG_CUDA(ctx, Wd.alloc(ctx->stream, (size_t)nv * n * sizeof(double)));
        MMG_CUDA(ctx, cudaMemcpyAsync(Vd.p, V, (size_t)nv * n_out * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        const double one = 1.0, zero = 0.0;
        MMG_CUBLAS(ctx, cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_N, (int)n, nv, (int)n_out, &one, R->d, (int)R->cols,
                                    Vd.as<double>(), (int)n_out, &zero, Wd.as<double>(), (int)n));
    }
    std::vector<double> tmp;
    double total_ms = 0.0;
    for (int64_t r0 = 0; r0 < m; r0 += chunk) {
        const int64_t rows = std::min<int64_t>(chunk, m - r0);
        MMG_CUDA(ctx, cudaMemcpy2DAsync(Xd.p, pitch * sizeof(double), xs + r0 * ld, ld * sizeof(double), n * sizeof(double), rows,
                                        cudaMemcpyHostToDevice, ctx->stream));
        ScanDmmaParams prm{};
        prm.snps = Xd.p;
        prm.pitch = pitch;
        prm.row_begin = 0;
        prm.row_count = rows;
        prm.R = Rp.as<double>();
        prm.ldr = ldr;
        prm.n_out_pad = (int)rows_pad;
        prm.k_pad = (int)pitch;
        prm.y = Vp.as<double>();
        prm.h0_rss = h0_rss;
        prm.n_p = n_p;
        prm.lbeta = lbeta;
        prm.xx = d_xx; prm.xy = d_xy; prm.rss = d_rss; prm.f = d_f; prm.p = d_p; prm.var_perc = d_vp;
        // short scans: column tiles of R split over the idle SMs, as in launch_scan_dmma
        const int64_t blocks = (rows + SD_BM - 1) / SD_BM;
        const int NT = (int)(rows_pad / SD_BN);
        DevBuf part;
        if (blocks * 2 <= ctx->sm_count && NT > 1) {
            prm.nsplit = (int)std::min<int64_t>(NT, ctx->sm_count / blocks);
            MMG_CUDA(ctx, part.alloc(ctx->stream, (size_t)(2 * prm.nsplit) * rows * sizeof(double)));
            prm.part = part.as<double>();
        }
        const int grid = (int)std::min<int64_t>(blocks * std::max(1, prm.nsplit), ctx->sm_count);
        cudaEventRecord(ctx->kev0, ctx->stream);
        scan_dmma_kernel<false, double><<<grid, SD_THREADS, sd_smem_bytes<double>(), ctx->stream>>>(prm);
        MMG_TRY(launch_check(ctx, "scan_dmma_kernel<double>"));
        if (prm.nsplit > 1) {
            scan_split_finish_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, ctx->stream>>>(prm.part, prm.nsplit, rows, h0_rss, n_p, lbeta, d_xx, d_xy, d_rss,
                                                                                            d_f, d_p, d_vp);
            MMG_TRY(launch_check(ctx, "scan_split_finish_kernel"));
        }
        cudaEventRecord(ctx->kev1, ctx->stream);
        if (dots)
            for (int v = 0; v < nv; ++v) {
                row_dots_f64_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(Xd.as<double>(), pitch, rows, (int)n,
                                                                                         Wd.as<double>() + (int64_t)v * n, d_dots + (int64_t)v * chunk);
                MMG_TRY(launch_check(ctx, "row_dots_f64_kernel"));
            }
        const size_t bytes = rows * sizeof(double);
        if (ps) MMG_CUDA(ctx, cudaMemcpyAsync(ps + r0, d_p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (f_stats) MMG_CUDA(ctx, cudaMemcpyAsync(f_stats + r0, d_f, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (rss) MMG_CUDA(ctx, cudaMemcpyAsync(rss + r0, d_rss, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (var_perc) MMG_CUDA(ctx, cudaMemcpyAsync(var_perc + r0, d_vp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (xx) MMG_CUDA(ctx, cudaMemcpyAsync(xx + r0, d_xx, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (dots) {
            tmp.resize((size_t)nv * chunk);
            MMG_CUDA(ctx, cudaMemcpyAsync(tmp.data(), d_dots, (size_t)nv * chunk * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (dots)
            for (int v = 0; v < nv; ++v)
                for (int64_t s = 0; s < rows; ++s) dots[(r0 + s) * nv + v] = tmp[(size_t)v * chunk + s];
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        total_ms += ms;
    }
    ctx->last_scan_ms = total_ms;
    return MMG_OK;
}

// The int8 scan when the caller already holds the quadratic form A = R'R and v = R'y~: the multi-GPU path forms A once across
// the ranks (mmg_quad_form_tiles + all-gather) instead of repeating the 2 n^3 / 2 flops of the product on every rank (12.5 ms at
// n = 10k, as long as an 8-way shard of the scan itself).  Device outputs; the two ABI entries below differ in where v comes from
// and where the results go.
static int scan_quad_given(mmg_ctx* ctx, const QuadA& A, const double* v, bool v_on_device, double h0_rss, double n_p, int64_t snp_begin,
                           int64_t snp_count, double* d_out /* [5 x ld]: p, f, rss, var_perc, xx */, int64_t ld) {
    const double lbeta = lbeta_host(0.5 * n_p, 0.5);
    DevBuf xy;
    MMG_CUDA(ctx, xy.alloc(ctx->stream, (size_t)snp_count * sizeof(double)));
    StageTimer tm(ctx, "scan");
    MMG_TRY(scan_tc_run(ctx, 1, nullptr, nullptr, &h0_rss, n_p, lbeta, snp_begin, snp_count, d_out + 4 * ld, xy.as<double>(), d_out + 2 * ld,
                        d_out + ld, d_out, d_out + 3 * ld, &A, v, v_on_device));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);           // scan_tc_run ends with a stream synchronisation
    ctx->last_scan_ms = ms;
    return MMG_OK;
}

int mmg_emmax_scan_quad_f64(mmg_ctx* ctx, mmg_mat Ah, const double* v, double h0_rss, double n_p, int64_t snp_begin,
                            int64_t snp_count, double* ps, double* f_stats, double* rss, double* var_perc, double* xx) {
    MmgMat* A = ctx ? get_mat(ctx, Ah) : nullptr;
    MMG_CHECK(ctx, A && ctx->snps && v, "mmg_emmax_scan_quad_f64: need resident genotypes, A and v");
    MMG_CHECK(ctx, A->rows == ctx->n && A->cols == ctx->n, "A must be n x n with n = %lld (is %lld x %lld)", (long long)ctx->n,
              (long long)A->rows, (long long)A->cols);
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count > 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf out;      // p, f, rss, var_perc, xx
    MMG_CUDA(ctx, out.alloc(ctx->stream, (size_t)5 * snp_count * sizeof(double)));
    QuadA qa;
    qa.d = A->d;
    qa.ld = A->cols;
    MMG_TRY(scan_quad_given(ctx, qa, v, false, h0_rss, n_p, snp_begin, snp_count, out.as<double>(), snp_count));
    StageTimer tm2(ctx, "d2h");
    const size_t bytes = snp_count * sizeof(double);
    double* host[5] = {ps, f_stats, rss, var_perc, xx};
    for (int k = 0; k < 5; ++k)
        if (host[k]) MMG_CUDA(ctx, cudaMemcpyAsync(host[k], out.as<double>() + (int64_t)k * snp_count, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

// Launches the linear pre-pass of the int8 scan over resident rows [snp_begin, +snp_count) on the SIDE stream: v = R'y~
// (x~.y~ = x.v), diag(R'R) = column sums of squares of R, then x.v, sum_j A_jj x_j^2 and ||x||_1 per SNP (snp_prepass_kernel) --
// HBM / FP64 work that runs underneath the tensor-core work the caller queues next on the main stream (mmg_quad_form_tiles and
// the all-gather of the multi-GPU path).  The following mmg_emmax_scan_quad_dev over the same rows joins it and may then be
// called with v = 0.
int mmg_scan_prepass_begin(mmg_ctx* ctx, mmg_mat Rh, const double* yres, int64_t snp_begin, int64_t snp_count) {
    MmgMat* R = ctx ? get_mat(ctx, Rh) : nullptr;
    MMG_CHECK(ctx, R && yres && ctx->snps, "mmg_scan_prepass_begin: need resident genotypes, R and the residual phenotype");
    MMG_CHECK(ctx, R->cols == ctx->n, "R must have n = %lld columns (has %lld)", (long long)ctx->n, (long long)R->cols);
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count > 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = ctx->n, n_out = R->rows, n_padN = round_up(n, TC_BN);
    ctx->early_prepass = false;
    MMG_TRY(ensure_side_stream(ctx));
    WsBuf pre, in;
    MMG_TRY(ws_get(ctx, MMG_WS_SCAN_PRE, (3 * snp_count + n_padN) * (int64_t)sizeof(double), &pre.p));       // the layout scan_tc_run expects (T = 1)
    MMG_TRY(ws_get(ctx, MMG_WS_SCAN_EARLY, (2 * n_padN + n_out) * (int64_t)sizeof(double), &in.p));
    double* d_dg = pre.as<double>();
    double* p_xy = d_dg + n_padN;
    double* p_qd = p_xy + snp_count;
    double* p_a1 = p_qd + snp_count;
    double* d_v = in.as<double>();
    double* d_y = d_v + 2 * n_padN;
    // the main stream's work so far (R) must be complete; everything below runs on the side stream
    MMG_CUDA(ctx, cudaEventRecord(ctx->ov0, ctx->stream));
    MMG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ov0, 0));
    MMG_CUDA(ctx, cudaMemsetAsync(d_v, 0, (size_t)(2 * n_padN) * sizeof(double), ctx->stream2));
    MMG_CUDA(ctx, cudaMemsetAsync(d_dg, 0, (size_t)n_padN * sizeof(double), ctx->stream2));
    MMG_CUDA(ctx, cudaMemcpyAsync(d_y, yres, (size_t)n_out * sizeof(double), cudaMemcpyHostToDevice, ctx->stream2));
    const double one = 1.0, zero = 0.0;
    cublasSetStream(ctx->cublas, ctx->stream2);
    const cublasStatus_t st = cublasDgemv(ctx->cublas, CUBLAS_OP_N, (int)n, (int)n_out, &one, R->d, (int)n, d_y, 1, &zero, d_v, 1);
    cublasSetStream(ctx->cublas, ctx->stream);
    if (st != CUBLAS_STATUS_SUCCESS) return fail(ctx, MMG_ECUBLAS, "mmg_scan_prepass_begin: cublasDgemv status %d", (int)st);
    col_sumsq_kernel<<<(unsigned)((n + 31) / 32), 256, 0, ctx->stream2>>>(R->d, n, (int)n_out, (int)n, d_dg);
    MMG_TRY(launch_check(ctx, "col_sumsq_kernel"));
    const unsigned grid = (unsigned)((snp_count + 8 * PRE_ROWS - 1) / (8 * PRE_ROWS));
    snp_prepass_kernel<PRE_ROWS, 4, 4><<<grid, 256, 0, ctx->stream2>>>(ctx->snps, ctx->pitch, snp_begin, snp_count, 1, d_v, d_dg, n_padN, p_xy, p_qd, p_a1,
                                                                    snp_count);
    MMG_TRY(launch_check(ctx, "snp_prepass_kernel"));
    MMG_CUDA(ctx, cudaEventRecord(ctx->ov1, ctx->stream2));
    ctx->early_prepass = true;
    ctx->early_begin = snp_begin;
    ctx->early_count = snp_count;
    return MMG_OK;
}

int64_t mmg_quad_form_slots(int64_t n) { return n > 0 ? qa_slots(n) : 0; }

int mmg_quad_form_tiles(mmg_ctx* ctx, mmg_mat Rh, int64_t slot_begin, int64_t slot_count, mmg_mat Ah, double* err_abs) {
    MmgMat* R = ctx ? get_mat(ctx, Rh) : nullptr;
    MmgMat* A = ctx ? get_mat(ctx, Ah) : nullptr;
    MMG_CHECK(ctx, R && A && ctx->snps, "mmg_quad_form_tiles: need resident genotypes, R and the packed output");
    MMG_CHECK(ctx, R->cols == ctx->n, "R must have n = %lld columns (has %lld)", (long long)ctx->n, (long long)R->cols);
    MMG_CHECK(ctx, A->cols == QA_TILE_ELEMS && slot_begin >= 0 && slot_count >= 0 && slot_begin + slot_count <= A->rows,
              "packed A must be [slots x 65536] and hold blocks [%lld, %lld)", (long long)slot_begin, (long long)(slot_begin + slot_count));
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm(ctx, "scan_prep");
    DevBuf amax;
    MMG_CUDA(ctx, amax.alloc(ctx->stream, sizeof(unsigned long long)));
    double err = 0.0;
    MMG_TRY(quad_form_int8(ctx, R, A->d + slot_begin * QA_TILE_ELEMS, slot_begin, slot_count, amax.as<unsigned long long>(), &err));
    if (err_abs) *err_abs = err;
    return MMG_OK;
}

int mmg_emmax_scan_quad_dev(mmg_ctx* ctx, mmg_mat Ah, int packed, double a_err, mmg_mat vh, double h0_rss, double n_p, int64_t snp_begin,
                            int64_t snp_count, mmg_mat outh) {
    MmgMat* A = ctx ? get_mat(ctx, Ah) : nullptr;
    MmgMat* v = (ctx && vh) ? get_mat(ctx, vh) : nullptr;
    MmgMat* out = ctx ? get_mat(ctx, outh) : nullptr;
    MMG_CHECK(ctx, A && out && ctx->snps, "mmg_emmax_scan_quad_dev: need resident genotypes, A and the output matrix");
    MMG_CHECK(ctx, v || (ctx->early_prepass && ctx->early_begin == snp_begin && ctx->early_count == snp_count),
              "mmg_emmax_scan_quad_dev: v = R'y is needed unless mmg_scan_prepass_begin ran for these rows");
    const int64_t n = ctx->n;
    if (packed) MMG_CHECK(ctx, A->cols == QA_TILE_ELEMS && A->rows >= qa_slots(n), "packed A must be [>= %lld x 65536]", (long long)qa_slots(n));
    else MMG_CHECK(ctx, A->rows == n && A->cols == n, "dense A must be n x n with n = %lld", (long long)n);
    MMG_CHECK(ctx, !v || v->rows * v->cols == n, "v must hold n = %lld values", (long long)n);
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count > 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    MMG_CHECK(ctx, out->rows == 5 && out->cols >= snp_count, "out must be [5 x >= snp_count]");
    MMG_CHECK(ctx, a_err >= 0.0, "a_err must be non-negative");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    QuadA qa;
    qa.d = A->d;
    qa.ld = A->cols;
    qa.packed = packed != 0;
    qa.err = a_err;
    return scan_quad_given(ctx, qa, v ? v->d : nullptr, true, h0_rss, n_p, snp_begin, snp_count, out->d, out->cols);
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

}  // extern "C"
