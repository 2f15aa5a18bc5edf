// libmixmogam_b200: exact EMMA for a batch of SNPs and the ML / REML fits that need no eig_R (emma.cuh).
#include "common.cuh"
#include "emma.cuh"

using namespace mmg;

namespace {

template <int Q1>
int emma_launch(mmg_ctx* ctx, const EmmaParams& prm, int64_t kk) {
    emma_grid_kernel<Q1><<<dim3((unsigned)prm.g, (unsigned)kk), REML_THREADS, 0, ctx->stream>>>(prm);
    MMG_TRY(launch_check(ctx, "emma_grid_kernel"));
    emma_refine_kernel<Q1><<<(unsigned)kk, REML_THREADS, 0, ctx->stream>>>(prm);
    return launch_check(ctx, "emma_refine_kernel");
}

}  // namespace

extern "C" {

int mmg_emma_f64(mmg_ctx* ctx, int method, mmg_mat ULh, const double* lam, const double* X0, int q0, const double* y, const double* xs,
                 const int64_t* snp_rows, int64_t k, const double* deltas, int g, double esp, double* out, double* lls, double* dlls) {
    MmgMat* UL = ctx ? get_mat(ctx, ULh) : nullptr;
    MMG_CHECK(ctx, UL && UL->rows == UL->cols && lam && y && deltas && out && g > 1, "mmg_emma_f64: bad argument");
    MMG_CHECK(ctx, method == 0 || method == 1, "mmg_emma_f64: method must be 0 (REML) or 1 (ML)");
    MMG_CHECK(ctx, q0 >= 0 && (q0 == 0 || X0), "mmg_emma_f64: X0 missing");
    const int64_t n = UL->rows;
    const int has_snp = (k > 0 && (xs || snp_rows)) ? 1 : 0;
    const int64_t kk = has_snp ? k : 1;
    const int q = q0 + has_snp;
    MMG_CHECK(ctx, q >= 1 && q <= EMMA_QMAX, "mmg_emma_f64: 1..%d fixed-effect columns (incl. the SNP) supported, got %d", EMMA_QMAX, q);
    MMG_CHECK(ctx, n > q, "mmg_emma_f64: more fixed effects than individuals");
    if (snp_rows && !xs) {
        MMG_CHECK(ctx, ctx->snps && ctx->n == n, "mmg_emma_f64: resident genotypes with n = %lld individuals needed", (long long)n);
        for (int64_t s = 0; s < k; ++s) MMG_CHECK(ctx, snp_rows[s] >= 0 && snp_rows[s] < ctx->m, "mmg_emma_f64: SNP row %lld out of range", (long long)snp_rows[s]);
    }
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm(ctx, "reml");
    const int ow = EMMA_OUT + q;
    // device layout: lam[n] | XY[n x (q0+1)] | Z0[n x (q0+1)] | deltas[g] | lls[kk x g] | dlls[kk x g] | out[kk x ow] | XS[k x n] | G[k x n]
    const int64_t c1 = q0 + 1;
    const int64_t nd = n + 2 * n * c1 + g + 2 * kk * g + kk * ow + (has_snp ? 2 * k * n : 0);
    DevBuf buf, rows;
    MMG_CUDA(ctx, buf.alloc(ctx->stream, (size_t)nd * sizeof(double)));
    double* d_lam = buf.as<double>();
    double* d_xy = d_lam + n;
    double* d_z0 = d_xy + n * c1;
    double* d_del = d_z0 + n * c1;
    double* d_lls = d_del + g;
    double* d_dlls = d_lls + kk * g;
    double* d_out = d_dlls + kk * g;
    double* d_xs = d_out + kk * ow;
    double* d_g = d_xs + (has_snp ? k * n : 0);
    std::vector<double> xy((size_t)(n * c1));
    for (int64_t i = 0; i < n; ++i) {
        for (int a = 0; a < q0; ++a) xy[(size_t)(i * c1 + a)] = X0[i * q0 + a];
        xy[(size_t)(i * c1 + q0)] = y[i];
    }
    MMG_CUDA(ctx, cudaMemcpyAsync(d_lam, lam, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(d_xy, xy.data(), (size_t)(n * c1) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, cudaMemcpyAsync(d_del, deltas, g * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const double one = 1.0, zero = 0.0;
    // Z0 = UL [X0 y]  (row-major [n x c1] = [n x n][n x c1]  <=>  column-major Z0' = XY' UL')
    MMG_CUBLAS(ctx, cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_N, (int)c1, (int)n, (int)n, &one, d_xy, (int)c1, UL->d, (int)n, &zero, d_z0, (int)c1));
    if (has_snp) {
        if (xs) {
            MMG_CUDA(ctx, cudaMemcpyAsync(d_xs, xs, (size_t)(k * n) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        } else {
            MMG_CUDA(ctx, rows.alloc(ctx->stream, (size_t)k * sizeof(long long)));
            std::vector<long long> r((size_t)k);
            for (int64_t s = 0; s < k; ++s) r[(size_t)s] = snp_rows[s];
            MMG_CUDA(ctx, cudaMemcpyAsync(rows.p, r.data(), (size_t)k * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // r goes out of scope
            gather_rows_f64_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)k), 256, 0, ctx->stream>>>(ctx->snps, ctx->pitch, rows.as<long long>(), (int)n, d_xs);
            MMG_TRY(launch_check(ctx, "gather_rows_f64_kernel"));
        }
        // G = XS UL'  (row-major [k x n] = [k x n][n x n]'  <=>  column-major G' = UL XS')
        MMG_CUBLAS(ctx, cublasDgemm(ctx->cublas, CUBLAS_OP_T, CUBLAS_OP_N, (int)n, (int)k, (int)n, &one, UL->d, (int)n, d_xs, (int)n, &zero, d_g, (int)n));
    }
    EmmaParams prm{};
    prm.n = (int)n;
    prm.q0 = q0;
    prm.has_snp = has_snp;
    prm.method = method;
    prm.lam = d_lam;
    prm.Z0 = d_z0;
    prm.G = d_g;
    prm.deltas = d_del;
    prm.g = g;
    prm.esp = esp;
    prm.lbeta = lbeta_host(0.5 * (double)(n - q), 0.5);
    prm.lls = d_lls;
    prm.dlls = d_dlls;
    prm.out = d_out;
    switch (q + 1) {
        case 2: MMG_TRY(emma_launch<2>(ctx, prm, kk)); break;
        case 3: MMG_TRY(emma_launch<3>(ctx, prm, kk)); break;
        case 4: MMG_TRY(emma_launch<4>(ctx, prm, kk)); break;
        case 5: MMG_TRY(emma_launch<5>(ctx, prm, kk)); break;
        case 6: MMG_TRY(emma_launch<6>(ctx, prm, kk)); break;
        case 7: MMG_TRY(emma_launch<7>(ctx, prm, kk)); break;
        case 8: MMG_TRY(emma_launch<8>(ctx, prm, kk)); break;
        default: MMG_TRY(emma_launch<9>(ctx, prm, kk)); break;
    }
    MMG_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)(kk * ow) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (lls) MMG_CUDA(ctx, cudaMemcpyAsync(lls, d_lls, (size_t)(kk * g) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (dlls) MMG_CUDA(ctx, cudaMemcpyAsync(dlls, d_dlls, (size_t)(kk * g) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

}  // extern "C"
