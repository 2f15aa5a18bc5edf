/*
 * mixmogam_b200 -- C ABI of the B200-native EMMAX hot path.
 *
 * The reference (bvilhjal/mixmogam) is pure Python and exposes no FFI; its hot
 * path sits behind the Python functions listed below.  This header is the
 * boundary the drop-in Python modules (mixmogam_b200/kinship.py,
 * linear_models.py, hdf5_data.py) bind through ctypes; INTEGRATION.md shows the
 * stub a maintainer of the reference would add.  Each entry point cites the
 * reference lines it replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - every function returns an int status: 0 = ok, negative = error class;
 *     the message is available through mmg_last_error().  No C++ exception
 *     crosses the ABI.
 *   - host pointers are borrowed for the duration of the call only; outputs are
 *     pre-allocated by the caller.  Matrices are row-major (C order).
 *   - device state lives in an opaque mmg_ctx (one per GPU).  Calls on one ctx
 *     are serialised by the caller; different ctxs may be driven from different
 *     threads.
 *   - there is NO CPU fallback: without a CUDA device mmg_create fails.
 */
#ifndef MIXMOGAM_B200_H
#define MIXMOGAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmg_ctx mmg_ctx;
typedef int64_t mmg_mat;          /* handle of a device-resident FP64 matrix */

enum {
    MMG_OK = 0,
    MMG_EBADARG = -1,             /* bad argument / shape / state            */
    MMG_ECUDA = -2,               /* CUDA runtime or driver error            */
    MMG_ENCCL = -3,               /* reserved (collectives run through torch.distributed) */
    MMG_ECUSOLVER = -4,
    MMG_EOOM = -5,                /* device or pinned-host allocation failed */
    MMG_ECUBLAS = -6,
    MMG_EVALUE = -7               /* input values outside the supported domain (e.g. genotype not in {0,1,2}) */
};

/* genotype codings of kinship.calc_ibs_kinship (kinship.py:14-56) */
enum { MMG_CODING_BINARY = 0, MMG_CODING_DIPLOID = 1 };
/* kernel implementations */
enum { MMG_IMPL_AUTO = 0, MMG_IMPL_TCGEN05 = 1, MMG_IMPL_SIMT = 2, MMG_IMPL_DMMA = 3 };

/* ---- context ---------------------------------------------------------------- */
int mmg_create(int device, mmg_ctx** out);
int mmg_destroy(mmg_ctx* ctx);
const char* mmg_last_error(mmg_ctx* ctx);            /* ctx may be NULL: last creation error */
int mmg_device_info(mmg_ctx* ctx, char* name64, int* sm_count, int* cc_major, int* cc_minor,
                    int64_t* free_bytes, int64_t* total_bytes);
int mmg_sync(mmg_ctx* ctx);
/* the CUDA stream (cudaStream_t) every call on ctx is ordered on.  A multi-GPU caller enqueues its NCCL collectives on it
 * (torch.cuda.ExternalStream in mixmogam_b200/parallel.py), so a collective on a library-owned buffer is ordered between the
 * library's kernels without a host synchronisation. */
int mmg_stream_handle(mmg_ctx* ctx, void** stream);
/* number of kernels of THIS library launched on ctx since creation (bench.py gpu_launches) */
int64_t mmg_launch_count(mmg_ctx* ctx);
/* named stage timers in seconds, CUDA-event based, accumulated since the last reset.
 * names: "h2d","pack","gram","finalize","ibd","syevd","reml","scan_prep","scan","d2h","matrix" (FP64 matrix plumbing: copies,
 * scalings, cuBLAS dgemm) */
int mmg_timer_get(mmg_ctx* ctx, const char* name, double* seconds, int64_t* calls);
int mmg_timer_reset(mmg_ctx* ctx);
/* duration (ms) of the most recent launch of the dominant kernels, measured with CUDA
 * events on the launching stream: which = "gram" | "scan" | "perm" | "ibd".  Two more keys report what the last call ran
 * rather than a time: "gram_is_fp4" (1: the Gram multiplied e2m1 operands with tcgen05 kind::mxf4, 0: int8 / SIMT), "gram_is_pair"
 * (1: as a CTA-pair MMA, cta_group::2) and
 * "scan_impl" (the MMG_IMPL_* value MMG_IMPL_AUTO resolved to in the last mmg_emmax_scan_f64 / _betas_f64 call). */
int mmg_last_kernel_ms(mmg_ctx* ctx, const char* which, double* ms);

/* the most recent int8 tensor-core scan: number of base-128 digit planes it used (chosen per call unless MMG_TC_SLICES
 * fixes it) and the certified bound max_s |d(x~.x~)| / x~.x~ of the truncation, measured over every SNP of that launch */
int mmg_last_scan_info(mmg_ctx* ctx, int* slices, double* rho);

/* pinned host memory for callers that want full-rate PCIe copies */
int mmg_host_alloc(void** ptr, int64_t bytes);
int mmg_host_free(void* ptr);

/* ---- device FP64 matrices (plumbing for the n x n objects of linear_models.py) -- */
int mmg_mat_create(mmg_ctx* ctx, int64_t rows, int64_t cols, mmg_mat* out);   /* zero-filled */
int mmg_mat_alloc(mmg_ctx* ctx, int64_t rows, int64_t cols, int zero, mmg_mat* out);   /* zero == 0: contents undefined */
int mmg_mat_free(mmg_ctx* ctx, mmg_mat m);
int mmg_mat_shape(mmg_ctx* ctx, mmg_mat m, int64_t* rows, int64_t* cols);
int mmg_mat_upload(mmg_ctx* ctx, mmg_mat m, const double* host, int64_t ld_host);
int mmg_mat_download(mmg_ctx* ctx, mmg_mat m, double* host, int64_t ld_host);
/* rows row0, row0 + row_step, ... (nrows of them) into host [nrows x ld_host]: one strided copy */
int mmg_mat_download_rows(mmg_ctx* ctx, mmg_mat m, int64_t row0, int64_t row_step, int64_t nrows, double* host, int64_t ld_host);
/* raw device pointer; work on it must be ordered on the context's stream (mmg_stream_handle) or follow an mmg_sync */
int mmg_mat_device_ptr(mmg_ctx* ctx, mmg_mat m, void** dptr, int64_t* ld);
int mmg_mat_copy(mmg_ctx* ctx, mmg_mat dst, mmg_mat src);
/* C = alpha * op(A) * op(B) + beta * C   (cuBLAS dgemm; a plain library GEMM, outside the hot path) */
int mmg_mat_gemm(mmg_ctx* ctx, int trans_a, int trans_b, double alpha, mmg_mat A, mmg_mat B,
                 double beta, mmg_mat C);
int mmg_mat_scale_rows(mmg_ctx* ctx, mmg_mat A, const double* d_host);         /* A[i,:] *= d[i] */
int mmg_mat_add_diag(mmg_ctx* ctx, mmg_mat A, double alpha);                   /* A += alpha*I   */
/* R = (I - QQ') diag(d) U in one pass over U: U [n x n] eigenvectors as rows (eig_L), d [n] = 1/sqrt(lambda + delta)
 * (H_sqrt_inv = diag(d) U, linear_models.py:898), Q [n x q] orthonormal columns of the rotated fixed effects (:1300); q = 0
 * gives H_sqrt_inv itself.  R is the transposed M of :1303. */
int mmg_mat_rotation(mmg_ctx* ctx, mmg_mat U, const double* d_host, const double* Q_host, int q, mmg_mat R_out);
/* kinship.scale_k (kinship.py:94-100): c = tr(K) - sum(K)/n ; K *= (n-1)/c ; returns the scalar */
int mmg_mat_scale_k(mmg_ctx* ctx, mmg_mat K, double* scalar);
/* dst = scale_k(src), src untouched (linear_models.py:580 add_random_effect scales a copy of the caller's K); scalar nullable */
int mmg_mat_scale_k_copy(mmg_ctx* ctx, mmg_mat src, mmg_mat dst, double* scalar);
/* linalg.eigh (linear_models.py:594,613) through cuSOLVER syevd, FP64.  A is overwritten by the
 * eigenvectors stored as ROWS (the reference's `evecs.T`, :596,:615), eigenvalues ascending in w_host. */
int mmg_mat_syevd(mmg_ctx* ctx, mmg_mat A, double* w_host, double* seconds);

/* ---- genotypes ------------------------------------------------------------------ */
/* snps: SNP-major int8 [m x n], row stride ld (the reference's `snps` list / 2-D array,
 * kinship.py:21-23, linear_models.py:1317).  Replaces the resident genotype block. */
int mmg_snps_upload(mmg_ctx* ctx, const int8_t* snps, int64_t m, int64_t n, int64_t ld);
/* list-of-rows form: rows[s] points at n contiguous int8 */
int mmg_snps_upload_rows(mmg_ctx* ctx, const int8_t* const* rows, int64_t m, int64_t n);
/* reserve an (m x n) resident block, then fill row ranges (streamed / chunked uploads) */
int mmg_snps_reserve(mmg_ctx* ctx, int64_t m, int64_t n);
int mmg_snps_write(mmg_ctx* ctx, int64_t row0, const int8_t* snps, int64_t rows, int64_t ld);
int mmg_snps_free(mmg_ctx* ctx);
int mmg_snps_shape(mmg_ctx* ctx, int64_t* m, int64_t* n);
int mmg_snps_device_ptr(mmg_ctx* ctx, void** dptr, int64_t* pitch);
/* per-SNP sums over individuals (int64) and sums of squares, e.g. for MAF filters (hdf5_data.py:91-96) */
int mmg_snps_row_sums(mmg_ctx* ctx, int64_t* sums_host, int64_t* sumsq_host);

/* ---- stage 1: kinship --------------------------------------------------------------- */
/* Integer Gram of the resident genotype rows [snp_begin, snp_begin+snp_count):
 *   binary  (kinship.py:43-44): G += S S', S = 2x-1 (int8), K-dim = snp_count
 *   diploid (kinship.py:33-41): G += T T', T = [x>=1 | x>=2] thermometer planes, K-dim = 2*snp_count
 * accumulated into the ctx's int32 n x n Gram (bit-exact, order independent).  reset!=0 zeroes it first.
 * impl MMG_IMPL_TCGEN05 (default): the planes hold only 0 / +-1, exact e2m1 values, and are multiplied by
 * tcgen05.mma.cta_group::2.kind::mxf4.block_scale (unit scales, FP32 accumulators holding exact integers) at twice the int8
 * rate (MMG_GRAM_KIND=i8 keeps int8 operands, kind::i8); MMG_IMPL_SIMT: dp4a cross-check kernel.  All three give the same
 * integers ("i8" in the entry-point names is the type of the resident genotypes). */
int mmg_kinship_gram_i8(mmg_ctx* ctx, int coding, int impl, int64_t snp_begin, int64_t snp_count, int reset);
/* The same Gram straight from HOST genotypes (kinship.py:29-32 walks the caller's `snps` chunk by chunk): the rows
 * stream into the resident block on a copy stream, one 65 536-SNP chunk at a time, and the pack + Gram of chunk c start
 * as soon as chunk c has landed, so the contraction hides behind the PCIe transfer.  Equivalent to mmg_snps_upload +
 * mmg_kinship_gram_i8(0, m); the genotypes stay resident for the scan.  snps is borrowed until the call returns. */
int mmg_kinship_gram_i8_host(mmg_ctx* ctx, int coding, int impl, const int8_t* snps, int64_t m, int64_t n, int64_t ld,
                             int reset);
/* Host-side helper of the streamed upload (no GPU involved): packs rows [0, rows) of SNP-major int8 genotype codes 0..3
 * (row stride ld) to 2 bits each with `threads` host threads -- code j of a row in bits 2 (j % 4) of byte j / 4, row stride
 * dst_ld >= ceil(n / 4), tail bytes zeroed.  Returns 0, or 1 when a code outside 0..3 was met (dst is then not usable).
 * mmg_kinship_gram_i8_host sends part of the chunks this way (a quarter of the PCIe bytes) while the DMA engine moves the
 * others unpacked; MMG_H2D_PACK=0 switches the packed lane off, MMG_HOST_THREADS sets the thread count. */
int mmg_host_pack2(const int8_t* src, int64_t rows, int64_t n, int64_t ld, uint8_t* dst, int64_t dst_ld, int threads);
/* The same two calls for genotype rows the CALLER holds packed already, 2 bits per genotype (code j of a row in bits
 * 2 (j % 4) .. 2 (j % 4) + 1 of byte j / 4, codes 0..3; what mmg_host_pack2 writes; row stride ld_bytes >= ceil(n / 4)): a quarter
 * of the bytes cross PCIe (n / 4 bytes per SNP) and no host core touches them; unpack2_kernel expands every chunk
 * into the resident int8 block, bits beyond column n are ignored.  hdf5 readers / .bed-like stores hand this over directly. */
int mmg_kinship_gram_i8_host_packed2(mmg_ctx* ctx, int coding, int impl, const uint8_t* packed, int64_t m, int64_t n,
                                     int64_t ld_bytes, int reset);
int mmg_snps_upload_packed2(mmg_ctx* ctx, const uint8_t* packed, int64_t m, int64_t n, int64_t ld_bytes);
int mmg_host_threads_default(void);
/* lanes of the most recent mmg_kinship_gram_i8_host: 65 536-SNP chunks sent packed / unpacked, measured host packing rate (GB/s) */
int mmg_last_h2d_info(mmg_ctx* ctx, int64_t* packed_chunks, int64_t* raw_chunks, double* pack_gbs);
/* device pointer of the int32 Gram (n x n, row stride ld elements) for an NCCL all-reduce between ranks */
int mmg_kinship_gram_ptr(mmg_ctx* ctx, void** dptr, int64_t* n, int64_t* ld);
int mmg_kinship_gram_download(mmg_ctx* ctx, int32_t* G_host);
/* The tensor-core Gram fills the 256 x 256 blocks on and above the block diagonal of the padded square.  direction 0 packs
 * those blocks back to back into one contiguous int32 buffer (*dptr, *count elements: half the bytes of the square) for the
 * all-reduce between ranks (partial Grams summed exactly in int32); direction 1 unpacks the reduced buffer into the
 * Gram again.  Stream ordered. */
int mmg_kinship_gram_tri(mmg_ctx* ctx, int direction, void** dptr, int64_t* count);
/* kinship.py:50-55: binary  K = G/(2 m) + 0.5 ; diploid K = f64(f32(m - L1/2)/f32(m)) + I with
 * L1 = G_ii + G_jj - 2 G_ij and a zero diagonal count; then scale_k if scaled.  m_total = SNPs in G. */
int mmg_kinship_finalize_f64(mmg_ctx* ctx, int coding, int64_t m_total, int scaled, mmg_mat K_out,
                             double* scale_scalar);
/* kinship.calc_ibd_kinship (kinship.py:59-75) and the hdf5 variants (hdf5_data.py:30-62,84-115):
 * K += sum_s z_s z_s' over resident rows [snp_begin, +snp_count) with z = (x-mean)/std (ddof 0),
 * FP64 accumulation.  snp_mask (nullable, one byte per row in range) selects rows (MAF filter).
 * Returns the number of rows used.  Fails with MMG_EVALUE on a monomorphic row (kinship.py:67). */
int mmg_kinship_ibd_accumulate_f64(mmg_ctx* ctx, mmg_mat K_acc, int64_t snp_begin, int64_t snp_count,
                                   const uint8_t* snp_mask, int64_t* used);

/* ---- stage 2: REML ------------------------------------------------------------------- */
/* LinearMixedModel.get_estimates, REML branch (linear_models.py:789-891): the delta grid
 * (lls, dlls over deltas[g]) for T phenotypes, the sign-change bracket, the secant refinement that
 * scipy.optimize.newton performs without fprime (tol=esp, maxiter=100), the bracket validation and
 * the grid-maximum fallback.  sq_etas is [T x p] (one phenotype per row).
 * flags[t]: bit0 = a zero interval was found, bit1 = secant converged, bit2 = refined delta accepted. */
int mmg_reml_f64(mmg_ctx* ctx, const double* eig_vals, const double* sq_etas, int64_t p, int64_t T,
                 const double* deltas, int64_t g, double esp,
                 double* lls, double* dlls, double* opt_delta, double* opt_ll, int32_t* flags);

/* Exact EMMA for a batch of k SNPs, and the ML / REML variance-component fits that need no eig_R (linear_models.py:931-968
 * expedited_REML_t_test; :771-927 get_estimates with xs and its ML branch :811-824; get_ML :672-696).  The reference runs one
 * n x n eigendecomposition of S(K+I)S per tested SNP; here every quantity of the likelihood (s1..s4 of :803-809) is evaluated
 * from weighted moments in the eigenbasis of K alone (eig_L: UL = eigenvectors as rows, lam = eigenvalues), so a batch costs one
 * rotation GEMM and O(k g n q^2) flops -- identical to the reference in exact arithmetic (csrc/emma.cuh has the algebra).
 *   method: 0 = REML, 1 = ML.   X0 [n x q0] fixed effects (row-major), y [n].
 *   SNPs: xs [k x n] host doubles, or snp_rows[k] = rows of the resident genotype block; both NULL (k = 0): fit the model
 *   without a SNP (one output row).   deltas[g] = the grid (:796), esp = secant tolerance (:847).
 *   out [k x (9 + q)], q = q0 + 1 (q0 without SNP): delta, max_ll, vg, ve, f_stat, p_val, var_perc, rss, mahalanobis_rss, beta[q]
 *   (beta: fixed effects then the SNP, :903-906); lls / dlls [k x g] (nullable): the grid values (:807-810 / :821-824). */
int mmg_emma_f64(mmg_ctx* ctx, int method, mmg_mat UL, const double* lam, const double* X0, int q0, const double* y,
                 const double* xs, const int64_t* snp_rows, int64_t k, const double* deltas, int g, double esp,
                 double* out, double* lls, double* dlls);

/* ---- stage 3: SNP scan ------------------------------------------------------------------ */
/* _emmax_f_test_ hot loops (linear_models.py:1315-1349) over the resident genotype rows
 * [snp_begin, +snp_count):
 *     x~ = R x            (R = M' = (I - QQ')H, [n_out x n]; the rotation GEMM of :1318)
 *     xx = x~.x~ ; xy[v] = x~.V[v]      (V: nv rotated-space vectors [nv x n_out]; V[0] = residual y~)
 *     rss = h0_rss - xy[0]^2/xx (kept at h0_rss when xx <= 0, the `if rss:` of :1329)
 *     f = n_p * r2/(1-r2), r2 = xy[0]^2/(xx*h0_rss) ; p = F.sf(f, 1, n_p)   (:1345-1349)
 * impl: MMG_IMPL_DMMA  = FP64 tensor-core (mma.sync m8n8k4 f64) rotation fused with the reductions;
 *       MMG_IMPL_TCGEN05 = x'(R'R)x on int8 tcgen05 tensor cores: diagonal in FP64, off-diagonal as exact base-256
 *                          digit planes of R'R (itself formed as exact int8 digit-plane products, or by FP64 dsyrk:
 *                          MMG_QUAD_A); the number of planes is chosen so that the certified truncation bound
 *                          on x~.x~ is <= MMG_TC_TOL (1e-7) for every SNP (mmg_last_scan_info); a scan that cannot certify
 *                          it fails with MMG_EVALUE.  A SNP whose x~.x~ is below 1e-8 of sum_j A_jj x_j^2 is collinear with
 *                          the fixed effects (e.g. monomorphic): it keeps the null fit (rss = h0_rss, f = 0, p = 1), like the
 *                          reference's empty-residue case (:1329), and does not enter the certification.
 * Outputs (host, length snp_count, any may be NULL): ps, f_stats, rss, var_perc, xx;
 * dots: [snp_count x nv]. */
int mmg_emmax_scan_f64(mmg_ctx* ctx, mmg_mat R, const double* V, int nv, double h0_rss, double n_p,
                       int impl, int64_t snp_begin, int64_t snp_count,
                       double* ps, double* f_stats, double* rss, double* var_perc,
                       double* xx, double* dots);
/* with_betas=True of _emmax_f_test_ (linear_models.py:1323: lstsq([h0_X, x~], y~res) per SNP) finished on the device: the scan
 * with R = H (no projection) and V = [y~res; h0_X'] ([1 + q0] x n_out), then per SNP the (q0 + 1) x (q0 + 1) normal equations
 * through the Schur complement of A = h0_X'h0_X (Ainv = A^-1 [q0 x q0], c0 = h0_X'y~res [q0], yy = y~res.y~res), F and p.
 * betas: [snp_count x (q0 + 1)]; a SNP whose column is rank deficient (or whose residue is exactly 0) keeps the null fit like
 * the reference's `if rss:` (:1325) -- rss = h0_rss, the row holds h0_betas followed by NaN. */
int mmg_emmax_scan_betas_f64(mmg_ctx* ctx, mmg_mat R, const double* V, int q0, const double* Ainv, const double* c0, double yy,
                             const double* h0_betas, double h0_rss, double n_p, int impl, int64_t snp_begin, int64_t snp_count,
                             double* ps, double* f_stats, double* rss, double* var_perc, double* betas);
/* mmg_emmax_scan_f64 for REAL-VALUED genotype rows (imputed dosages; the reference's scan accepts any numeric row, it casts the
 * chunk to float32 at linear_models.py:1317): xs = [m x ld] host FP64, SNP-major, ld >= n.  FP64 tensor-core path with the
 * genotype operand staged as FP64; the rows pass through the device in chunks and do not become the resident block.
 * Same outputs as mmg_emmax_scan_f64 (any may be NULL). */
int mmg_emmax_scan_rows_f64(mmg_ctx* ctx, mmg_mat R, const double* V, int nv, double h0_rss, double n_p,
                            const double* xs, int64_t m, int64_t ld,
                            double* ps, double* f_stats, double* rss, double* var_perc,
                            double* xx, double* dots);
/* The int8 tensor-core form of mmg_emmax_scan_f64 for a caller that already holds A = R'R (n x n, row-major lower
 * triangle valid) and v = R'y~ [n].  Same outputs. */
int mmg_emmax_scan_quad_f64(mmg_ctx* ctx, mmg_mat A, const double* v, double h0_rss, double n_p,
                            int64_t snp_begin, int64_t snp_count,
                            double* ps, double* f_stats, double* rss, double* var_perc, double* xx);
/* Multi-GPU form of the same.  The quadratic form is kept as packed 256 x 256 FP64 blocks of its lower triangle: block (J, I),
 * I <= J, in slot J (J + 1) / 2 + I, row-major inside a block; mmg_quad_form_slots(n) blocks in all.  mmg_quad_form_tiles forms
 * blocks [slot_begin, +slot_count) of A = R'R as exact int8 digit-plane products on the tensor cores into the packed matrix
 * A ([>= slots x 65536] mmg_mat) -- every block costs the same, so ranks taking equal slot ranges are balanced and one in-place
 * all-gather of A completes it on every rank.  *err_abs: rigorous absolute error bound of the entries (same on every rank).
 * mmg_emmax_scan_quad_dev then scans with device-resident inputs and outputs: A (packed != 0: the block layout, else dense
 * n x n lower triangle), a_err = the error bound of A's entries (enters the certified bound; 0 for an FP64 A), v = R'y~ as an
 * mmg_mat of n values, and out = [5 x >= snp_count] rows ps, f_stats, rss, var_perc, xx (left on the device for the all-gather
 * of the per-rank slices). */
/* Launch the scan's linear pre-pass (v = R'y~, diag(R'R), then x.v, sum_j A_jj x_j^2, ||x||_1 per SNP) over resident rows
 * [snp_begin, +snp_count) on the library's side stream, underneath the tensor-core work queued next; the following
 * mmg_emmax_scan_quad_dev over the same rows joins it (and may pass v = 0). */
int mmg_scan_prepass_begin(mmg_ctx* ctx, mmg_mat R, const double* yres, int64_t snp_begin, int64_t snp_count);
int64_t mmg_quad_form_slots(int64_t n);
int mmg_quad_form_tiles(mmg_ctx* ctx, mmg_mat R, int64_t slot_begin, int64_t slot_count, mmg_mat A, double* err_abs);
int mmg_emmax_scan_quad_dev(mmg_ctx* ctx, mmg_mat A, int packed, double a_err, mmg_mat v, double h0_rss, double n_p,
                            int64_t snp_begin, int64_t snp_count, mmg_mat out);
/* Phenotype-batched scan: T phenotypes scanned against one genotype block in ONE launch (BASELINE.json configs[2];
 * the reference calls linear_models.emmax once per phenotype, linear_models.py:1790).  R[t] is the rotation of
 * phenotype t (its own delta_t enters through H_t), V[t] its residual phenotype in the rotated space ([T x n_out]),
 * h0_rss[t] its null RSS.  int8 tcgen05 path.  Outputs are [T x snp_count], any may be NULL. */
int mmg_emmax_scan_multi_f64(mmg_ctx* ctx, const mmg_mat* R, int T, const double* V, const double* h0_rss, double n_p,
                             int64_t snp_begin, int64_t snp_count,
                             double* ps, double* f_stats, double* rss, double* var_perc, double* xx);
/* Phenotype-batched scan with SHARED work: the T phenotypes are measured on the same individuals, so they share
 * the eigenbasis U of the kinship (mmg_mat n x n, eigenvectors as ROWS, linear_models.py:596) and differ in delta_t only.
 * One rotation g = U x per SNP (int8 tensor cores, exact digit planes of U) serves all of them; per phenotype
 *     x~.x~ = sum_k W[t][k] g_k^2 - sum_j (x.c_tj)^2        x~.y~ = x.v_t
 * with W[t][k] = 1 / (lambda_k + delta_t) (host, [T x n]) and Ext (mmg_mat [T (1 + q0) x n]) holding, for each phenotype,
 * v_t = U' diag(d_t) y~res_t followed by the q0 rows c_tj = U' diag(d_t) Q_t[:, j] (d_t = sqrt(W[t]), Q_t an orthonormal basis
 * of the rotated fixed effects, :1300).  Cost: one rotation + O(n T) per SNP instead of T rotations.  The number of digit
 * planes is raised until the certified bounds hold (MMG_TC_TOL, default 1e-7 relative on x~.x~); SNPs collinear with the
 * fixed effects keep the null fit.  Outputs are [T x snp_count] (any may be NULL); info (optional, 5 doubles): planes used,
 * certified bound on x~.x~, bound on x~.y~ (t-statistic scale), rotation ms, contraction ms. */
int mmg_emmax_scan_shared_f64(mmg_ctx* ctx, mmg_mat U, mmg_mat Ext, const double* W, int T, int q0, const double* h0_rss,
                              double n_p, int64_t snp_begin, int64_t snp_count,
                              double* ps, double* f_stats, double* rss, double* var_perc, double* xx, double* info);
/* _emmax_permutations_ inner loop (linear_models.py:1157-1164): with centred SNPs x_c = x - mean(x),
 * x~ = R x_c, for each permuted phenotype column W[:,p] (already rotated back: W = R' Ys, [n x P]):
 *     ratio[p] = max over SNPs of (x_c.W[:,p])^2 / (x~.x~)
 * so that min_rss[p] = |Ys_p|^2 - ratio[p].  ratio_inout is max-accumulated (initialise to 0).
 * impl: MMG_IMPL_TCGEN05 (default) = x_c'(R'R)x_c by the int8 quadratic-form scan of R(I - 11'/n), then an int8
 *       tcgen05 GEMM of the genotype block with 8 exact digit planes of W;  MMG_IMPL_DMMA = FP64 tensor cores. */
int mmg_emmax_perm_scan_f64(mmg_ctx* ctx, mmg_mat R, mmg_mat W, int centre, int impl,
                            int64_t snp_begin, int64_t snp_count, double* ratio_inout);
/* scipy.stats.f.sf(f, dfn, dfd) (linear_models.py:1349,1172) on the device, FP64 */
int mmg_f_sf_f64(mmg_ctx* ctx, const double* f, int64_t count, double dfn, double dfd, double* out);

#ifdef __cplusplus
}
#endif
#endif /* MIXMOGAM_B200_H */
