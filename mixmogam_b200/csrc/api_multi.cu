// libmixmogam_b200: the phenotype-batched scan with shared rotation (scan_shared.cuh) -- one rotation g = U x per SNP on the
// int8 tensor cores for ALL phenotypes, a skinny FP64 tensor-core contraction per phenotype, fused RSS / F / p.
#include "common.cuh"
#include "scan_dmma.cuh"
#include "scan_shared.cuh"

using namespace mmg;

namespace mmg {
void multi_init_attrs() {
    cudaFuncSetAttribute(tc_gemm_i8_kernel<RotEpi, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<RotEpi, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(tc_gemm_i8_kernel<RotEpi, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(scan_dmma_kernel<false, double, SD_MODE_SQUARE_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         sd_smem_bytes<double>());
}
}  // namespace mmg

// L2 policy "evict_last for `frac` of the lines, evict_first for the rest" (createpolicy.fractional): with the genotype blocks of
// all co-resident CTAs (148 x 1.3 MB at n = 10k) larger than L2, a uniform priority makes every re-read of a block miss
// (ncu: 34.6 GB of DRAM reads per 16 k SNPs = every K-block of every tile); pinning a fraction that fits keeps that fraction.
static __global__ void make_l2_policy_kernel(float frac, unsigned long long* out) {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(p) : "f"(frac));
    *out = p;
}

static double wall_now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

extern "C" {

int mmg_emmax_scan_shared_f64(mmg_ctx* ctx, mmg_mat Uh, mmg_mat Exth, const double* W, int T, int q0, const double* h0_rss, double n_p,
                              int64_t snp_begin, int64_t snp_count, double* ps, double* f_stats, double* rss, double* var_perc,
                              double* xx, double* info) {
    MmgMat* U = ctx ? get_mat(ctx, Uh) : nullptr;
    MmgMat* Ext = ctx ? get_mat(ctx, Exth) : nullptr;
    MMG_CHECK(ctx, U && Ext && ctx->snps && W && h0_rss, "mmg_emmax_scan_shared_f64: need resident genotypes, U, the extra basis rows and W");
    const int64_t n = ctx->n;
    MMG_CHECK(ctx, U->rows == n && U->cols == n, "U must be %lld x %lld (eigenvectors as rows)", (long long)n, (long long)n);
    MMG_CHECK(ctx, T >= 1 && q0 >= 0 && q0 <= 15 && Ext->rows == (int64_t)T * (1 + q0) && Ext->cols == n,
              "Ext must be [T (1 + q0) x n] = [%lld x %lld] (has %lld x %lld)", (long long)T * (1 + q0), (long long)n, (long long)Ext->rows,
              (long long)Ext->cols);
    MMG_CHECK(ctx, snp_begin >= 0 && snp_count > 0 && snp_begin + snp_count <= ctx->m, "SNP range out of bounds");
    MMG_CHECK(ctx, n < 131072, "int32 plane sums need n < 131072");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const double lbeta = lbeta_host(0.5 * n_p, 0.5);
    const double tol = [] { const char* e = getenv("MMG_TC_TOL"); const double v = e ? atof(e) : 0.0; return v > 0.0 ? v : 1e-7; }();

    const int n_e = T * (1 + q0);
    const int64_t n_ext = n + n_e;
    const int nb_u = (int)((n + 31) / 32), nb_e = (n_e + 31) / 32;       // the extra rows start a block of their own
    const int64_t n_u_pad = (int64_t)nb_u * 32;
    const int nblocks = nb_u + nb_e;
    const int64_t ldg = (int64_t)nblocks * 32;
    const int64_t ldq = round_up(n, TC_BK);                    // contraction bytes per operand row
    const int64_t k_pad = round_up(n, SD_BK);                  // contraction range of kernel B (columns of g that belong to U)
    const int64_t T_pad = round_up(T, SD_BN);
    // digit planes: the n eigenvector rows carry the cost (P_u x 2 n^2 int8 ops per SNP); the few extra rows (x.v_t feeds the
    // numerator of F directly) get more
    int P = env_int("MMG_SHARED_PLANES", 5);
    P = std::max(2, std::min(P, RS_MAX_PLANES));
    int P_e = std::min(RS_MAX_PLANES, std::max(P, env_int("MMG_SHARED_PLANES_EXT", 7)));
    const int P_fixed = getenv("MMG_SHARED_PLANES") != nullptr;
    int ksplit = std::max(1, env_int("MMG_SHARED_KSPLIT", 1));
    int cs = env_int("MMG_SHARED_CLUSTER", 2);
    if (cs != 1 && cs != 2 && cs != 4) cs = 2;
    // SNP chunk: the FP64 rotated block g [chunk x ldg] is the one large temporary (<= ~4 GB).  A whole number of waves of the
    // persistent rotation grid (one 128-SNP group per CTA and wave): a chunk of 1.35 waves costs two
    const int64_t wave = (int64_t)128 * cs * (cs == 4 ? tc_gemm_max_clusters<RotEpi, 4>(ctx) : cs == 2 ? tc_gemm_max_clusters<RotEpi, 2>(ctx)
                                                                                                      : tc_gemm_max_clusters<RotEpi, 1>(ctx));
    int64_t chunk = wave * std::max<int64_t>(1, std::min<int64_t>(4, ((int64_t)4 << 30) / (wave * ldg * 8)));
    chunk = std::min<int64_t>(chunk, round_up(snp_count, 128));
    if (const int c = env_int("MMG_SHARED_CHUNK", 0)) chunk = round_up(c, 128);
    // genotype operand: evict_last for 30 % of its lines (what fits L2 beside the basis stream), evict_first for the rest --
    // measured 52.4 -> 47.4 ms per 131 k SNPs against a uniform priority; MMG_SHARED_A_FRAC = 0 switches it off
    uint64_t policy_a = L2_EVICT_LAST;
    {
        const char* fr = getenv("MMG_SHARED_A_FRAC");
        const float f = fr ? (float)atof(fr) : 0.3f;
        static float cached_f = -1.f;
        static unsigned long long cached_p = 0;
        if (f > 0.f && f <= 1.f) {
            if (cached_f != f) {
                DevBuf pb;
                MMG_CUDA(ctx, pb.alloc(ctx->stream, 8));
                make_l2_policy_kernel<<<1, 1, 0, ctx->stream>>>(f, pb.as<unsigned long long>());
                MMG_TRY(launch_check(ctx, "make_l2_policy_kernel"));
                unsigned long long pv = 0;
                MMG_CUDA(ctx, cudaMemcpyAsync(&pv, pb.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
                MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                cached_p = pv;
                cached_f = f;
            }
            policy_a = cached_p;
        }
    }

    StageTimer tm(ctx, "scan");
    const double t_setup0 = wall_now();
    DevBuf rs, Bq, Wd, gbuf, abuf, l1buf, outbuf, small;
    MMG_CUDA(ctx, rs.alloc(ctx->stream, (size_t)ldg * sizeof(double)));
    MMG_CUDA(ctx, small.alloc(ctx->stream, 64 + (size_t)(2 * T + n_e) * sizeof(double)));
    int* d_bad = small.as<int>();
    unsigned long long* d_rho = reinterpret_cast<unsigned long long*>(small.as<char>() + 16);
    double* d_h0 = reinterpret_cast<double*>(small.as<char>() + 64);
    double* d_w1 = d_h0 + T;
    double* d_es = d_w1 + T;
    MMG_CUDA(ctx, cudaMemsetAsync(small.p, 0, 64, ctx->stream));
    {   // rscale = 1 for the padding rows of the last block
        std::vector<double> ones((size_t)ldg, 1.0);
        MMG_CUDA(ctx, cudaMemcpyAsync(rs.p, ones.data(), (size_t)ldg * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    rot_row_scale_kernel<<<(unsigned)n_ext, 256, 0, ctx->stream>>>(U->d, U->cols, (int)n, Ext->d, Ext->cols, n_e, (int)n, (int)n_u_pad,
                                                                   rs.as<double>(), d_bad);
    MMG_TRY(launch_check(ctx, "rot_row_scale_kernel"));
    std::vector<double> rscale((size_t)ldg);
    int bad = 0;
    MMG_CUDA(ctx, cudaMemcpyAsync(rscale.data(), rs.p, (size_t)ldg * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad) return fail(ctx, MMG_EVALUE, "mmg_emmax_scan_shared_f64: non-finite entries in U / Ext");
    double umax = 0.0;
    for (int64_t r = 0; r < n; ++r) umax = std::max(umax, rscale[(size_t)r]);
    // per-phenotype constants
    std::vector<double> w1((size_t)T, 0.0);
    for (int t = 0; t < T; ++t) {
        double s = 0.0;
        for (int64_t k = 0; k < n; ++k) s += W[(size_t)t * n + k];
        w1[(size_t)t] = s;
    }
    MMG_CUDA(ctx, cudaMemcpyAsync(d_h0, h0_rss, (size_t)T * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(d_w1, w1.data(), (size_t)T * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(d_es, rscale.data() + n_u_pad, (size_t)n_e * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    // weights of kernel B: [T_pad x k_pad], zero beyond T / n
    MMG_CUDA(ctx, Wd.alloc(ctx->stream, (size_t)T_pad * k_pad * sizeof(double)));
    MMG_CUDA(ctx, cudaMemsetAsync(Wd.p, 0, (size_t)T_pad * k_pad * sizeof(double), ctx->stream));
    MMG_CUDA(ctx, cudaMemcpy2DAsync(Wd.p, k_pad * sizeof(double), W, n * sizeof(double), n * sizeof(double), T, cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, gbuf.alloc(ctx->stream, (size_t)chunk * ldg * sizeof(double)));
    MMG_CUDA(ctx, abuf.alloc(ctx->stream, (size_t)chunk * T_pad * sizeof(double)));
    MMG_CUDA(ctx, l1buf.alloc(ctx->stream, (size_t)chunk * sizeof(double)));
    MMG_CUDA(ctx, outbuf.alloc(ctx->stream, (size_t)5 * T * snp_count * sizeof(double)));
    double* o_p = outbuf.as<double>();
    double* o_f = o_p + (size_t)T * snp_count;
    double* o_rss = o_f + (size_t)T * snp_count;
    double* o_vp = o_rss + (size_t)T * snp_count;
    double* o_xx = o_vp + (size_t)T * snp_count;

    double rho_xx = 0.0, rho_xy = 0.0, rot_ms = 0.0, con_ms = 0.0, fin_ms = 0.0;
    int passes = 0;
    const double t_loop0 = wall_now();
    for (;;) {
        // ---- digit planes of the extended basis ----
        const int64_t b_rows = round_up(((int64_t)nb_u * P + (int64_t)nb_e * P_e) * 32, TC_BN);
        if (Bq.p) {
            cudaFreeAsync(Bq.p, Bq.s);
            Bq.p = nullptr;
        }
        MMG_CUDA(ctx, Bq.alloc(ctx->stream, (size_t)b_rows * ldq));
        MMG_CUDA(ctx, cudaMemsetAsync(Bq.p, 0, (size_t)b_rows * ldq, ctx->stream));
        rot_slice_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)n_ext), 256, 0, ctx->stream>>>(U->d, U->cols, (int)n, Ext->d, Ext->cols, n_e,
                                                                                                      (int)n, P, P_e, nb_u, rs.as<double>(), Bq.as<int8_t>(), ldq);
        MMG_TRY(launch_check(ctx, "rot_slice_kernel"));
        MMG_CUDA(ctx, cudaMemsetAsync(d_rho, 0, 16, ctx->stream));
        // tile table of one 128-SNP group: the contraction range in `ksplit` parts, every part sweeps all 256-row tiles
        const int ntiles = (int)(b_rows / TC_BN), KB = (int)(ldq / TC_BK);
        ksplit = std::min(ksplit, KB);
        std::vector<TcTile> tiles;
        for (int h = 0; h < ksplit; ++h)
            for (int ti = 0; ti < ntiles; ++ti) {
                TcTile tl{};
                tl.n0 = ti * TC_BN;
                tl.kb0 = (int)((int64_t)KB * h / ksplit);
                tl.kb1 = (int)((int64_t)KB * (h + 1) / ksplit);
                tl.aux0 = ti * (TC_BN / 32);
                tl.aux1 = h;
                tiles.push_back(tl);
            }
        MMG_TRY(ensure_tiles(ctx, tiles));
        const TcTile* td = (const TcTile*)ctx->tiles_d;
        const double rem_u = DIGIT256_REM * ldexp(1.0, -8 * P), rem_e = DIGIT256_REM * ldexp(1.0, -8 * P_e);
        rot_ms = con_ms = 0.0;
        std::vector<cudaEvent_t> evs;
        for (int64_t c0 = 0; c0 < snp_count; c0 += chunk) {
            const int64_t rows = std::min(chunk, snp_count - c0);
            snp_l1_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ctx->stream>>>(ctx->snps, ctx->pitch, snp_begin + c0, rows, l1buf.as<double>());
            MMG_TRY(launch_check(ctx, "snp_l1_kernel"));
            // ---- kernel A: rotation ----
            RotEpi::Params ep{};
            ep.g = gbuf.as<double>();
            ep.ldg = ldg;
            ep.row_count = rows;
            ep.P_u = P;
            ep.P_e = P_e;
            ep.nb_u = nb_u;
            ep.nblocks = nblocks;
            for (int p = 0; p < RS_MAX_PLANES; ++p) ep.w[p] = ldexp(1.0, -8 * (p + 1));
            ep.rscale = rs.as<double>();
            CUtensorMap tmA, tmB;
            MMG_TRY(make_tmap_u8(ctx, &tmA, ctx->snps + (snp_begin + c0) * ctx->pitch, ctx->pitch, rows, ctx->pitch, TC_BM));
            MMG_TRY(make_tmap_u8(ctx, &tmB, Bq.p, ldq, b_rows, ldq, TC_BN / cs));
            const int groups = (int)((rows + TC_BM - 1) / TC_BM);
            cudaEvent_t e[6];
            for (auto& x : e) { MMG_CUDA(ctx, cudaEventCreate(&x)); evs.push_back(x); }
            cudaEventRecord(e[0], ctx->stream);
            if (cs == 4)
                MMG_TRY((launch_tc_gemm<RotEpi, 4>(ctx, tmA, tmB, td, groups, (int)tiles.size(), 0, TC_BM, 0, ep, "tc_gemm_i8_kernel<RotEpi,4>", policy_a, L2_EVICT_NORMAL)));
            else if (cs == 2)
                MMG_TRY((launch_tc_gemm<RotEpi, 2>(ctx, tmA, tmB, td, groups, (int)tiles.size(), 0, TC_BM, 0, ep, "tc_gemm_i8_kernel<RotEpi,2>", policy_a, L2_EVICT_NORMAL)));
            else
                MMG_TRY((launch_tc_gemm<RotEpi, 1>(ctx, tmA, tmB, td, groups, (int)tiles.size(), 0, TC_BM, 0, ep, "tc_gemm_i8_kernel<RotEpi,1>", policy_a, L2_EVICT_NORMAL)));
            cudaEventRecord(e[1], ctx->stream);
            // ---- kernel B: a[s][t] = sum_k g[s][k]^2 w[t][k] ----
            ScanDmmaParams prm{};
            prm.snps = gbuf.p;
            prm.pitch = ldg;
            prm.row_begin = 0;
            prm.row_count = rows;
            prm.R = Wd.as<double>();
            prm.ldr = k_pad;
            prm.n_out_pad = (int)T_pad;
            prm.k_pad = (int)k_pad;
            prm.cstore = abuf.as<double>();
            prm.ldc = T_pad;
            const int grid = (int)std::min<int64_t>((rows + SD_BM - 1) / SD_BM, ctx->sm_count);
            cudaEventRecord(e[2], ctx->stream);
            scan_dmma_kernel<false, double, SD_MODE_SQUARE_STORE><<<grid, SD_THREADS, sd_smem_bytes<double>(), ctx->stream>>>(prm);
            MMG_TRY(launch_check(ctx, "scan_dmma_kernel<square,store>"));
            cudaEventRecord(e[3], ctx->stream);
            // ---- kernel C: statistics ----
            SharedFinishParams fp{};
            fp.a = abuf.as<double>();
            fp.lda = T_pad;
            fp.g = gbuf.as<double>();
            fp.ldg = ldg;
            fp.l1 = l1buf.as<double>();
            fp.rows = rows;
            fp.T = T;
            fp.q0 = q0;
            fp.n_u = (int)n_u_pad;
            fp.h0_rss = d_h0;
            fp.w1 = d_w1;
            fp.escale = d_es;
            fp.eps_u = umax * rem_u;
            fp.rem = rem_e;
            fp.n_p = n_p;
            fp.lbeta = lbeta;
            fp.out_stride = snp_count;
            fp.out_row0 = c0;
            fp.xx = o_xx;
            fp.rss = o_rss;
            fp.f = o_f;
            fp.p = o_p;
            fp.var_perc = o_vp;
            fp.rho_max = d_rho;
            cudaEventRecord(e[4], ctx->stream);
            shared_finish_kernel<<<dim3((unsigned)((rows + 255) / 256), (unsigned)T), 256, 0, ctx->stream>>>(fp);
            MMG_TRY(launch_check(ctx, "shared_finish_kernel"));
            cudaEventRecord(e[5], ctx->stream);
        }
        double rho[2] = {0.0, 0.0};
        MMG_CUDA(ctx, cudaMemcpyAsync(rho, d_rho, 16, cudaMemcpyDeviceToHost, ctx->stream));
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        fin_ms = 0.0;
        for (size_t i = 0; i + 5 < evs.size(); i += 6) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, evs[i], evs[i + 1]) == cudaSuccess) rot_ms += ms;
            if (cudaEventElapsedTime(&ms, evs[i + 2], evs[i + 3]) == cudaSuccess) con_ms += ms;
            if (cudaEventElapsedTime(&ms, evs[i + 4], evs[i + 5]) == cudaSuccess) fin_ms += ms;
        }
        passes += 1;
        for (cudaEvent_t x : evs) cudaEventDestroy(x);
        rho_xx = rho[0];
        rho_xy = rho[1];
        // x~.x~ to `tol` relative; x~.y~ to tol / 100 on the scale of the t-statistic (|d(-log10 p)| = 0.35 |dz| near p = 1, where the
        // north-star tolerance is 1e-6 x 1e-3 absolute)
        const bool ok_xx = rho_xx <= tol, ok_xy = rho_xy <= 0.01 * tol;
        if ((ok_xx && ok_xy) || P_fixed) break;
        if ((!ok_xx && P == RS_MAX_PLANES) || (ok_xx && !ok_xy && P_e == RS_MAX_PLANES))
            return fail(ctx, MMG_EVALUE, "shared-rotation scan: certified bound %.3g (x~.x~) / %.3g (x~.y~) above the tolerance %.3g with %d / %d planes",
                        rho_xx, rho_xy, tol, P, P_e);
        if (!ok_xx) ++P;
        P_e = std::min(RS_MAX_PLANES, std::max(P_e + (ok_xy ? 0 : 1), P));
    }
    if (env_int("MMG_SHARED_DEBUG", 0))
        fprintf(stderr, "[shared scan] set-up %.1f ms, %d pass(es) %.1f ms wall; last pass: rotation %.1f contraction %.1f statistics %.1f ms; chunk %lld SNPs\n",
                1e3 * (t_loop0 - t_setup0), passes, 1e3 * (wall_now() - t_loop0), rot_ms, con_ms, fin_ms, (long long)chunk);
    ctx->last_scan_ms = rot_ms + con_ms;
    ctx->last_scan_slices = P;
    ctx->last_scan_rho = rho_xx;
    if (info) {
        info[0] = (double)P;
        info[1] = rho_xx;
        info[2] = rho_xy;
        info[3] = rot_ms;
        info[4] = con_ms;
    }
    tm.stop();
    StageTimer tm2(ctx, "d2h");
    const size_t bytes = (size_t)T * snp_count * sizeof(double);
    if (ps) MMG_CUDA(ctx, cudaMemcpyAsync(ps, o_p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (f_stats) MMG_CUDA(ctx, cudaMemcpyAsync(f_stats, o_f, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (rss) MMG_CUDA(ctx, cudaMemcpyAsync(rss, o_rss, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (var_perc) MMG_CUDA(ctx, cudaMemcpyAsync(var_perc, o_vp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (xx) MMG_CUDA(ctx, cudaMemcpyAsync(xx, o_xx, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

}  // extern "C"
