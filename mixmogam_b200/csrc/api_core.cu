// libmixmogam_b200: C ABI (include/mixmogam_b200.h) over the sm_100a kernels -- context, timers, device matrices, resident
// genotypes, REML, F survival function.  Host-side orchestration only; cuBLAS / cuSOLVER appear only for plain library work
// outside the hot path (dgemm plumbing, syevd).
#include "common.cuh"
#include "fdist.cuh"
#include "reml.cuh"
#include "scan_dmma.cuh"

namespace mmg {
thread_local std::string g_create_error;
void kinship_init_attrs();
void scan_init_attrs();
void multi_init_attrs();
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
    // opt in to large dynamic shared memory once (per translation unit: each knows its own kernel instances)
    kinship_init_attrs();
    scan_init_attrs();
    multi_init_attrs();
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
    for (int i = 0; i < 8; ++i) cudaFree(ctx->ws[i]);
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
    resolve_timers(ctx);
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
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
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    resolve_timers(ctx);
    return MMG_OK;
}

int mmg_stream_handle(mmg_ctx* ctx, void** stream) {
    MMG_CHECK(ctx, ctx && stream, "mmg_stream_handle: bad argument");
    *stream = (void*)ctx->stream;
    return MMG_OK;
}

int64_t mmg_launch_count(mmg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mmg_timer_get(mmg_ctx* ctx, const char* name, double* seconds, int64_t* calls) {
    MMG_CHECK(ctx, ctx && name, "bad argument");
    resolve_timers(ctx);
    auto it = ctx->timers.find(name);
    if (seconds) *seconds = it == ctx->timers.end() ? 0.0 : it->second.seconds;
    if (calls) *calls = it == ctx->timers.end() ? 0 : it->second.calls;
    return MMG_OK;
}
int mmg_timer_reset(mmg_ctx* ctx) {
    MMG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    resolve_timers(ctx);
    ctx->timers.clear();
    return MMG_OK;
}
int mmg_last_kernel_ms(mmg_ctx* ctx, const char* which, double* ms) {
    MMG_CHECK(ctx, ctx && which && ms, "bad argument");
    if (!strcmp(which, "gram")) *ms = ctx->last_gram_ms;
    else if (!strcmp(which, "gram_is_fp4")) *ms = (double)ctx->last_gram_fp4;
    else if (!strcmp(which, "gram_is_pair")) *ms = (double)ctx->last_gram_pair;
    else if (!strcmp(which, "scan_impl")) *ms = (double)ctx->last_scan_impl;
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
int mmg_mat_alloc(mmg_ctx* ctx, int64_t rows, int64_t cols, int zero, mmg_mat* out) {
    MMG_CHECK(ctx, ctx && out && rows > 0 && cols > 0, "mmg_mat_alloc: bad shape %lld x %lld", (long long)rows, (long long)cols);
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    MmgMat m;
    m.rows = rows;
    m.cols = cols;
    MMG_CUDA(ctx, cudaMallocAsync((void**)&m.d, (size_t)rows * cols * sizeof(double), ctx->stream));
    if (zero) MMG_CUDA(ctx, cudaMemsetAsync(m.d, 0, (size_t)rows * cols * sizeof(double), ctx->stream));
    *out = ctx->next_mat++;
    ctx->mats[*out] = m;
    return MMG_OK;
}
int mmg_mat_create(mmg_ctx* ctx, int64_t rows, int64_t cols, mmg_mat* out) { return mmg_mat_alloc(ctx, rows, cols, 1, out); }
int mmg_mat_free(mmg_ctx* ctx, mmg_mat h) {
    MMG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
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
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm(ctx, "h2d");
    MMG_CUDA(ctx, cudaMemcpy2DAsync(m->d, m->cols * sizeof(double), host, ld_host * sizeof(double), m->cols * sizeof(double),
                                    m->rows, cudaMemcpyHostToDevice, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}
int mmg_mat_download(mmg_ctx* ctx, mmg_mat h, double* host, int64_t ld_host) {
    MmgMat* m = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, m && host && ld_host >= m->cols, "mmg_mat_download: bad argument");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm(ctx, "d2h");
    MMG_CUDA(ctx, cudaMemcpy2DAsync(host, ld_host * sizeof(double), m->d, m->cols * sizeof(double), m->cols * sizeof(double),
                                    m->rows, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}
int mmg_mat_download_rows(mmg_ctx* ctx, mmg_mat h, int64_t row0, int64_t row_step, int64_t nrows, double* host, int64_t ld_host) {
    MmgMat* m = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, m && host && ld_host >= m->cols && row0 >= 0 && row_step >= 1 && nrows >= 1 && row0 + (nrows - 1) * row_step < m->rows,
              "mmg_mat_download_rows: bad argument");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm(ctx, "d2h");
    MMG_CUDA(ctx, cudaMemcpy2DAsync(host, ld_host * sizeof(double), m->d + row0 * m->cols, row_step * m->cols * sizeof(double),
                                    m->cols * sizeof(double), nrows, cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}
int mmg_mat_device_ptr(mmg_ctx* ctx, mmg_mat h, void** dptr, int64_t* ld) {
    MmgMat* m = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, m && dptr, "mmg_mat_device_ptr: bad argument");
    *dptr = m->d;
    if (ld) *ld = m->cols;
    return MMG_OK;
}
int mmg_mat_copy(mmg_ctx* ctx, mmg_mat dst, mmg_mat src) {
    MmgMat *d = ctx ? get_mat(ctx, dst) : nullptr, *s = ctx ? get_mat(ctx, src) : nullptr;
    MMG_CHECK(ctx, d && s && d->rows == s->rows && d->cols == s->cols, "mmg_mat_copy: shape mismatch");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm_matrix(ctx, "matrix");
    MMG_CUDA(ctx, cudaMemcpyAsync(d->d, s->d, (size_t)d->rows * d->cols * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return MMG_OK;
}
int mmg_mat_gemm(mmg_ctx* ctx, int ta, int tb, double alpha, mmg_mat Ah, mmg_mat Bh, double beta, mmg_mat Ch) {
    MmgMat *A = ctx ? get_mat(ctx, Ah) : nullptr, *B = ctx ? get_mat(ctx, Bh) : nullptr, *C = ctx ? get_mat(ctx, Ch) : nullptr;
    MMG_CHECK(ctx, A && B && C, "mmg_mat_gemm: unknown handle");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm_matrix(ctx, "matrix");
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
int mmg_mat_scale_rows(mmg_ctx* ctx, mmg_mat h, const double* d_host) {
    MmgMat* A = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, A && d_host, "mmg_mat_scale_rows: bad argument");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm_matrix(ctx, "matrix");
    MMG_TRY(ensure_scratch(ctx, A->rows * sizeof(double)));
    MMG_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, d_host, A->rows * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((unsigned)((A->cols + 255) / 256), (unsigned)A->rows);
    scale_rows_kernel<<<grid, 256, 0, ctx->stream>>>(A->d, A->cols, (int)A->rows, (int)A->cols, (const double*)ctx->scratch);
    MMG_TRY(launch_check(ctx, "scale_rows_kernel"));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}
// R = (I - QQ') diag(d) U in one pass over U (rotation_kernel): U [n x n] eigenvectors as rows, d [n] host (1/sqrt(lambda + delta),
// linear_models.py:898), Q [n x q] host with orthonormal columns (the QR of the rotated fixed effects, :1300; q = 0: R = diag(d) U,
// the H_sqrt_inv of :898 itself).  R_out [n x n].
int mmg_mat_rotation(mmg_ctx* ctx, mmg_mat Uh, const double* d_host, const double* Q_host, int q, mmg_mat Rh) {
    MmgMat *U = ctx ? get_mat(ctx, Uh) : nullptr, *R = ctx ? get_mat(ctx, Rh) : nullptr;
    MMG_CHECK(ctx, U && R && U != R && d_host && U->rows == U->cols && R->rows == U->rows && R->cols == U->cols, "mmg_mat_rotation: bad argument");
    MMG_CHECK(ctx, q >= 0 && q <= 8 && (q == 0 || Q_host), "mmg_mat_rotation: 0..8 projected columns supported");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm(ctx, "matrix");
    const int64_t n = U->rows;
    DevBuf buf;      // d[n] | Q[n x q] | Qd[n x q] (Q with rows scaled by d) | QtH[q x n]
    MMG_CUDA(ctx, buf.alloc(ctx->stream, (size_t)(n + 3 * n * std::max(q, 1)) * sizeof(double)));
    double* d_d = buf.as<double>();
    double* d_q = d_d + n;
    double* d_qd = d_q + n * std::max(q, 1);
    double* d_qth = d_qd + n * std::max(q, 1);
    MMG_CUDA(ctx, cudaMemcpyAsync(d_d, d_host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (q > 0) {
        std::vector<double> qd((size_t)(n * q));
        for (int64_t k = 0; k < n; ++k)
            for (int c = 0; c < q; ++c) qd[(size_t)(k * q + c)] = Q_host[k * q + c] * d_host[k];
        MMG_CUDA(ctx, cudaMemcpyAsync(d_q, Q_host, (size_t)(n * q) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MMG_CUDA(ctx, cudaMemcpyAsync(d_qd, qd.data(), (size_t)(n * q) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // qd goes out of scope
        const double one = 1.0, zero = 0.0;
        // QtH = Qd' U  (row-major [q x n] = [n x q]' [n x n]  <=>  column-major QtH' = U' Qd)
        MMG_CUBLAS(ctx, cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, (int)n, q, (int)n, &one, U->d, (int)n, d_qd, q, &zero, d_qth, (int)n));
    }
    rotation_kernel<<<dim3(4, (unsigned)n), 256, 0, ctx->stream>>>(U->d, d_d, q > 0 ? d_q : nullptr, d_qth, q, (int)n, R->d);
    return launch_check(ctx, "rotation_kernel");
}
int mmg_mat_add_diag(mmg_ctx* ctx, mmg_mat h, double alpha) {
    MmgMat* A = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, A && A->rows == A->cols, "mmg_mat_add_diag: needs a square matrix");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    add_diag_kernel<<<(unsigned)((A->rows + 255) / 256), 256, 0, ctx->stream>>>(A->d, A->cols, (int)A->rows, alpha);
    return launch_check(ctx, "add_diag_kernel");
}

// dst = kinship.scale_k(src) without touching src (LinearMixedModel.add_random_effect keeps the caller's K intact,
// linear_models.py:580): row sums of src, the factor on the device, one scaled copy -- instead of a copy, row sums and an
// in-place scaling.  *scalar (nullable): the factor; asking for it waits for the stream.
int mmg_mat_scale_k_copy(mmg_ctx* ctx, mmg_mat srch, mmg_mat dsth, double* scalar) {
    MmgMat *S = ctx ? get_mat(ctx, srch) : nullptr, *D = ctx ? get_mat(ctx, dsth) : nullptr;
    MMG_CHECK(ctx, S && D && S != D && S->rows == S->cols && D->rows == S->rows && D->cols == S->cols, "mmg_mat_scale_k_copy: needs two square matrices of one shape");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm(ctx, "matrix");
    const int n = (int)S->rows;
    MMG_TRY(ensure_scratch(ctx, (n + 3) * sizeof(double)));
    double* rs = (double*)ctx->scratch;
    rowsum_kernel<<<n, 256, 0, ctx->stream>>>(S->d, S->cols, n, rs);
    MMG_TRY(launch_check(ctx, "rowsum_kernel"));
    scale_k_reduce_kernel<<<1, 1024, 0, ctx->stream>>>(rs, S->d, S->cols, n, rs + n);
    MMG_TRY(launch_check(ctx, "scale_k_reduce_kernel"));
    scale_factor_kernel<<<1, 1, 0, ctx->stream>>>(rs + n, n);
    MMG_TRY(launch_check(ctx, "scale_factor_kernel"));
    const int64_t cnt = (int64_t)n * n;
    scaled_copy_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((const double2*)S->d, (double2*)D->d, cnt / 2, rs + n + 2,
                                                                  (cnt & 1) ? S->d + cnt - 1 : nullptr, (cnt & 1) ? D->d + cnt - 1 : nullptr);
    MMG_TRY(launch_check(ctx, "scaled_copy_kernel"));
    if (scalar) {
        MMG_CUDA(ctx, cudaMemcpyAsync(scalar, rs + n + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MMG_OK;
}
int mmg_mat_scale_k(mmg_ctx* ctx, mmg_mat h, double* scalar) {
    MmgMat* K = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, K && K->rows == K->cols, "mmg_mat_scale_k: needs a square matrix");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    StageTimer tm_matrix(ctx, "matrix");
    MMG_TRY(scale_k_device(ctx, K, scalar));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MMG_OK;
}

int mmg_mat_syevd(mmg_ctx* ctx, mmg_mat h, double* w_host, double* seconds) {
    MmgMat* A = ctx ? get_mat(ctx, h) : nullptr;
    MMG_CHECK(ctx, A && A->rows == A->cols && w_host, "mmg_mat_syevd: needs a square matrix and w_host");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n = A->rows;
    StageTimer tm(ctx, "syevd");
    // MMG_SYEVD_JACOBI_MAX = k: matrices up to k x k take cuSOLVER's Jacobi solver (entirely on the device, tolerance 1e-15, eigenvalues
    // sorted ascending like syevd) instead of Xsyevd, whose tridiagonal stage runs on the host for small n.  Off by default: at
    // n = 198 (configs[0]) it is the slower of the two at steady state (6.3 against 2.7 ms per decomposition); it exists for hosts whose
    // cores are busy (one run next to numpy's BLAS threads showed 0.1 - 0.8 s in Xsyevd for the same matrix).
    if (n <= env_int("MMG_SYEVD_JACOBI_MAX", 0)) {
        syevjInfo_t jp = nullptr;
        MMG_CUSOLVER(ctx, cusolverDnCreateSyevjInfo(&jp));
        cusolverDnXsyevjSetTolerance(jp, 1e-15);
        cusolverDnXsyevjSetMaxSweeps(jp, 200);
        cusolverDnXsyevjSetSortEig(jp, 1);
        DevBuf wd, work, infod;
        int lwork = 0, rc = MMG_OK, info = 0;
        do {
            if (wd.alloc(ctx->stream, n * sizeof(double)) != cudaSuccess || infod.alloc(ctx->stream, sizeof(int)) != cudaSuccess) { rc = fail(ctx, MMG_EOOM, "syevj: allocation failed"); break; }
            cusolverStatus_t st = cusolverDnDsyevj_bufferSize(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, A->d, (int)n, wd.as<double>(), &lwork, jp);
            if (st != CUSOLVER_STATUS_SUCCESS) { rc = fail(ctx, MMG_ECUSOLVER, "Dsyevj_bufferSize status %d", (int)st); break; }
            if (work.alloc(ctx->stream, (size_t)std::max(lwork, 1) * sizeof(double)) != cudaSuccess) { rc = fail(ctx, MMG_EOOM, "syevj: workspace"); break; }
            st = cusolverDnDsyevj(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, A->d, (int)n, wd.as<double>(), work.as<double>(), lwork,
                                  infod.as<int>(), jp);
            if (st != CUSOLVER_STATUS_SUCCESS) { rc = fail(ctx, MMG_ECUSOLVER, "Dsyevj status %d", (int)st); break; }
            if (cudaMemcpyAsync(&info, infod.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaMemcpyAsync(w_host, wd.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = fail(ctx, MMG_ECUDA, "syevj: %s", cudaGetErrorString(cudaGetLastError())); break; }
            if (info != 0) rc = fail(ctx, MMG_ECUSOLVER, "syevj did not converge (info=%d)", info);
        } while (0);
        cusolverDnDestroySyevjInfo(jp);
        tm.stop();
        resolve_timers(ctx);
        if (seconds) *seconds = ctx->timers["syevd"].seconds;
        return rc;
    }
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
    resolve_timers(ctx);
    if (seconds) *seconds = ctx->timers["syevd"].seconds;
    return rc;
}

// ======================================================================================================
// genotypes
// ======================================================================================================
}  // extern "C" (interrupted for a kernel)
// contiguous rows of n bytes -> rows at the resident block's pitch (grid-stride over the bytes; the padding stays zero)
static __global__ void repitch_rows_kernel(const int8_t* __restrict__ src, int64_t n, int8_t* __restrict__ dst, int64_t pitch, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n;
        dst[r * pitch + (i - r * n)] = src[i];
    }
}
extern "C" {
int mmg_snps_free(mmg_ctx* ctx) {
    MMG_CHECK(ctx, ctx != nullptr, "ctx is NULL");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->snps);
    ctx->snps = nullptr;
    ctx->snps_capacity = 0;
    ctx->m = ctx->n = ctx->pitch = 0;
    ctx->snps_absmax = -1;
    return MMG_OK;
}
int mmg_snps_reserve(mmg_ctx* ctx, int64_t m, int64_t n) {
    MMG_CHECK(ctx, ctx && m > 0 && n > 0, "mmg_snps_reserve: bad shape");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t pitch = round_up(n, 256);   // zero padded: kernels read whole 16/32-byte groups up to the next 256
    if (!(ctx->snps && ctx->m == m && ctx->n == n)) {
        // a block that fits the allocation (but not one that would strand more than 3/4 of a large one) re-uses it: the chromosome
        // loop of hdf5_data.run_emmax reserves a different m every time, and cudaFree + cudaMalloc of a GB-sized block cost ~40 ms
        const int64_t need = m * pitch;
        if (!(ctx->snps && need <= ctx->snps_capacity && (need * 4 >= ctx->snps_capacity || ctx->snps_capacity <= (256ll << 20)))) {
            MMG_TRY(mmg_snps_free(ctx));
            MMG_CUDA(ctx, persistent_malloc(ctx->device, (void**)&ctx->snps, (size_t)need));
            ctx->snps_capacity = need;
        }
        ctx->snps_absmax = -1;
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
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->snps_absmax = -1;
    StageTimer tm(ctx, "h2d");
    if (rows && ld == ctx->n && ctx->n != ctx->pitch && ctx->n < 2048) {
        // short rows (configs[0]: n = 198 accessions): a strided DMA moves them one 198-byte row at a time -- 0.8 GB/s, 52 ms for
        // 214 k SNPs.  The block is contiguous on the host: 1-D copies into a device staging buffer + a re-pitch kernel instead.
        const int64_t piece = std::max<int64_t>(1, (256ll << 20) / ctx->n);
        DevBuf tmp;
        MMG_CUDA(ctx, tmp.alloc(ctx->stream, (size_t)std::min(piece, rows) * ctx->n));
        for (int64_t r0 = 0; r0 < rows; r0 += piece) {
            const int64_t cnt = std::min(piece, rows - r0);
            MMG_CUDA(ctx, cudaMemcpyAsync(tmp.p, snps + r0 * ld, (size_t)(cnt * ctx->n), cudaMemcpyHostToDevice, ctx->stream));
            const int64_t total = cnt * ctx->n;
            repitch_rows_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1 << 20), 256, 0, ctx->stream>>>(
                tmp.as<int8_t>(), ctx->n, ctx->snps + (row0 + r0) * ctx->pitch, ctx->pitch, total);
            MMG_TRY(launch_check(ctx, "repitch_rows_kernel"));
        }
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return MMG_OK;
    }
    // one strided DMA; measured at the PCIe rate for 10 KB rows (52 GB/s from page-locked memory), where a staged 1-D copy +
    // re-pitch kernel was no faster
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
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    MMG_TRY(mmg_snps_reserve(ctx, m, n));
    ctx->snps_absmax = -1;
    StageTimer tm(ctx, "h2d");
    if (m * n <= (4ll << 20)) {
        // a few rows (the single-SNP calls of the stepwise callers, linear_models.py:2720,2825): straight from the caller's rows --
        // the driver stages small pageable copies itself; two page-locked buffers cost ~20 ms to allocate, far more than the copy
        for (int64_t r = 0; r < m; ++r)
            MMG_CUDA(ctx, cudaMemcpyAsync(ctx->snps + r * ctx->pitch, rows[r], (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return MMG_OK;
    }
    // gather rows into two pinned staging buffers and copy them asynchronously
    const int64_t rows_per = std::max<int64_t>(1, std::min<int64_t>(m, (32ll << 20) / n));
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
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
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

}  // extern "C"
