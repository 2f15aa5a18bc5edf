// Stage 3, phenotype batch with SHARED work (BASELINE.json configs[2]): T phenotypes measured on the same
// individuals share the kinship, hence its eigenbasis U (rows = eigenvectors, linear_models.py:596), and differ only in
// delta_t.  With g = U x (the rotation, shared), w_tk = 1 / (lambda_k + delta_t), d_t = sqrt(w_t):
//     x~_t = (I - Q_t Q_t') diag(d_t) g                                              (linear_models.py:1299-1303, :1318)
//     x~_t.x~_t = sum_k w_tk g_k^2 - sum_j (x.c_tj)^2,      c_tj = U' diag(d_t) Q_t[:, j]
//     x~_t.y~_t = x.v_t,                                     v_t  = U' diag(d_t) y~res_t
// so ONE rotation per SNP serves every phenotype; what is left per phenotype is a contraction of g^2 with w_t and a few dot
// products of the raw genotype vector.  The reference runs one whole scan per phenotype (linear_models.py:1790 called T times).
//
//   kernel A  tc_gemm_i8_kernel<RotEpi>: g_ext = [U; v_1, c_1*, v_2, ...] x on the int8 tensor cores -- the rows of the
//             extended basis are cut into P exact base-256 digit planes (digits.cuh; per-row power-of-two scale), the genotype
//             block is the other (exact int8) operand, every plane product is an exact int32 in TMEM.  The B operand lists,
//             for each block of 32 basis rows, its P planes one after the other (32 rows each), so one epilogue thread (= one
//             SNP) meets the P plane sums of a block in consecutive 32-column chunks, folds them in 32 FP64 registers and
//             stores 32 finished doubles -- no read-modify-write (with K cut in parts for L2 residency of the genotype
//             blocks, later parts add to the stored value).
//   kernel B  scan_dmma_kernel<.., double, SD_MODE_SQUARE_STORE>: a[s][t] = sum_k g[s][k]^2 w[t][k] on the FP64 tensor cores
//   kernel C  shared_finish_kernel: x~.x~, x~.y~ -> RSS, F, p per (SNP, phenotype), certified error bound of the digit planes.
#pragma once
#include "digits.cuh"
#include "fdist.cuh"
#include "scan_tc.cuh"

namespace mmg {

constexpr int RS_MAX_PLANES = DIGIT256_MAX_PLANES;

struct RotEpi {
    struct Params {
        double* g;               // [rows_pad x ldg] rotated genotypes of this SNP chunk
        int64_t ldg;             // multiple of 32, >= 32 * nblocks
        int64_t row_count;       // SNPs of the chunk
        int P_u, P_e;            // digit planes of the eigenvector rows / of the extra rows (v_t, c_tj: few rows, more planes)
        int nb_u;                // blocks of 32 eigenvector rows (the extra rows start a new block)
        int nblocks;             // blocks of 32 basis rows in all
        double w[RS_MAX_PLANES]; // 256^-(p+1)
        const double* rscale;    // [32 * nblocks] 2^E_r of the basis row at that position
    };
    double d[32];
    int64_t orow;

    __device__ __forceinline__ void begin_group(const Params& p, int g, int row) { orow = (int64_t)g * TC_BM + row; }
    __device__ __forceinline__ void end_group(const Params&, int, int) {}
    __device__ __forceinline__ int tile_begin(const Params&, const TcTile&, int) { return TC_BN / 32; }
    __device__ __forceinline__ void tile_end(const Params&, const TcTile&, int, int) {}
    __device__ __forceinline__ void chunk(const Params& p, const TcTile& t, int, int c, const uint32_t (&v)[32]) {
        const int gidx = t.aux0 + c;                 // position in the (block, plane) sequence: uniform across the CTA
        const int gu = p.nb_u * p.P_u;
        int blk, pl, P;
        if (gidx < gu) {
            P = p.P_u;
            blk = gidx / P;
            pl = gidx - blk * P;
        } else {
            P = p.P_e;
            const int b = (gidx - gu) / P;
            blk = p.nb_u + b;
            pl = gidx - gu - b * P;
        }
        if (blk >= p.nblocks) return;
        const double wk = p.w[pl];
        if (pl == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) d[j] = wk * (double)(int)v[j];
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) d[j] = fma(wk, (double)(int)v[j], d[j]);
        }
        if (pl != P - 1 || orow >= p.row_count) return;
        const double2* sc = reinterpret_cast<const double2*>(p.rscale + 32 * blk);
        double2* dst = reinterpret_cast<double2*>(p.g + orow * p.ldg + 32 * blk);
        if (t.aux1 == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double2 s = __ldg(sc + j);
                dst[j] = make_double2(d[2 * j] * s.x, d[2 * j + 1] * s.y);
            }
        } else {                                     // a later part of the contraction range: add to the stored partial sum
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double2 s = __ldg(sc + j);
                double2 o = dst[j];
                o.x = fma(d[2 * j], s.x, o.x);
                o.y = fma(d[2 * j + 1], s.y, o.y);
                dst[j] = o;
            }
        }
    }
};

// max |row| of the extended basis [U (n rows); Ext (n_e rows)] -> rscale[r] = 2^E_r (digit256_exponent); one block per row
static __global__ void __launch_bounds__(256) rot_row_scale_kernel(const double* __restrict__ U, int64_t ldu, int n_u,
                                                                   const double* __restrict__ Ext, int64_t lde, int n_e, int n,
                                                                   int n_u_pad, double* __restrict__ rscale, int* __restrict__ bad) {
    __shared__ double red[8];
    const int r = blockIdx.x;
    const double* src = r < n_u ? U + (int64_t)r * ldu : Ext + (int64_t)(r - n_u) * lde;
    double m = 0.0;
    bool nonfinite = false;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double a = fabs(src[i]);
        if (!(a <= 1.79e308)) nonfinite = true;
        m = fmax(m, a);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    if (nonfinite) atomicExch(bad, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
        rscale[r < n_u ? r : n_u_pad + (r - n_u)] = ldexp(1.0, digit256_exponent(m));
    }
    (void)n_e;
}

// digit planes of the extended basis into the B operand of kernel A: the blocks of 32 basis rows follow each other, every block
// as its P planes of 32 operand rows (P_u planes for the eigenvector blocks, P_e for the blocks of extra rows, which start at
// block nb_u); contraction index (individual) contiguous
static __global__ void __launch_bounds__(256) rot_slice_kernel(const double* __restrict__ U, int64_t ldu, int n_u, const double* __restrict__ Ext,
                                                               int64_t lde, int n_e, int n, int P_u, int P_e, int nb_u,
                                                               const double* __restrict__ rscale, int8_t* __restrict__ Bq, int64_t ldq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= n || r >= n_u + n_e) return;
    const bool is_u = r < n_u;
    const int re = is_u ? r : r - n_u;
    const int P = is_u ? P_u : P_e;
    const double sc = rscale[is_u ? r : 32 * nb_u + re];
    const double x = (is_u ? U[(int64_t)r * ldu + i] : Ext[(int64_t)re * lde + i]) / sc;     // exact: power of two
    int dg[DIGIT256_MAX_PLANES];
    digit256_split(x, P, dg);
    const int64_t row0 = is_u ? (int64_t)(re >> 5) * P_u * 32 : ((int64_t)nb_u * P_u + (int64_t)(re >> 5) * P_e) * 32;
    const int64_t base = (row0 + (re & 31)) * ldq + i;
    for (int p = 0; p < P; ++p) Bq[base + (int64_t)p * 32 * ldq] = (int8_t)dg[p];
}

// ||x||_1 per SNP row (one warp per row, 16 bytes per lane and step; the row padding is zero)
static __global__ void __launch_bounds__(256) snp_l1_kernel(const int8_t* __restrict__ snps, int64_t pitch, int64_t row_begin, int64_t row_count,
                                                            double* __restrict__ l1) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= row_count) return;
    const uint4* x = reinterpret_cast<const uint4*>(snps + (row_begin + row) * pitch);
    int s = 0;
    for (int64_t i = lane; i < pitch / 16; i += 32) {
        const uint4 q = x[i];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) s += __dp4a((int)__vabs4(w[k]), 0x01010101, 0);     // |-128| wraps, but the scan's domain is |x| <= 8
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) l1[row] = (double)s;
}

struct SharedFinishParams {
    const double* a;         // [rows x lda]   a[s][t] = sum_k w_tk g_sk^2
    int64_t lda;
    const double* g;         // [rows x ldg]   columns n_u + t (1 + q0) + {0: x.v_t, 1 + j: x.c_tj}
    int64_t ldg;
    const double* l1;        // [rows] ||x||_1
    int64_t rows;
    int T, q0, n_u;
    const double* h0_rss;    // [T]
    const double* w1;        // [T] sum_k w_tk
    const double* escale;    // [T (1 + q0)] 2^E of the extra basis rows (error of the dot products per unit ||x||_1 and digit remainder)
    double eps_u;            // max_k 2^E_k * DIGIT256_REM 256^-P_u: absolute error of g_k per unit ||x||_1
    double rem;              // DIGIT256_REM 256^-P_e (extra rows)
    double n_p, lbeta;
    int64_t out_stride, out_row0;      // outputs are [T][out_stride], this chunk starts at out_row0
    double *xx, *xy, *rss, *f, *p, *var_perc;
    unsigned long long* rho_max;       // [2]: max relative bound on x~.x~, max bound on the t-statistic scale of x~.y~
};

static __global__ void __launch_bounds__(256) shared_finish_kernel(const SharedFinishParams prm) {
    const int64_t s_raw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    const bool valid = s_raw < prm.rows;               // every lane stays for the warp reductions below
    const int64_t s = valid ? s_raw : prm.rows - 1;
    const double a = prm.a[s * prm.lda + t];
    const double* ge = prm.g + s * prm.ldg + prm.n_u + (int64_t)t * (1 + prm.q0);
    const double* es = prm.escale + (int64_t)t * (1 + prm.q0);
    const double x1 = prm.l1[s];
    const double sxy = ge[0];
    double cc = 0.0, err = 0.0;
    for (int j = 0; j < prm.q0; ++j) {
        const double c = ge[1 + j];
        const double ec = es[1 + j] * prm.rem * x1;
        cc = fma(c, c, cc);
        err += 2.0 * fabs(c) * ec + ec * ec;
    }
    const double sxx = a - cc;
    {   // |d a| <= sum_k w_k (2 |g_k| e + e^2) <= 2 e sqrt(W1 a) + e^2 W1,  e = eps_u ||x||_1   (Cauchy-Schwarz)
        const double e = prm.eps_u * x1, w1 = prm.w1[t];
        err += 2.0 * e * sqrt(w1 * fmax(a, 0.0)) + e * e * w1;
    }
    const double h0 = prm.h0_rss[t];
    const bool degenerate = !(sxx > QS_DEGENERATE_REL * a);      // x in the span of the fixed effects: keeps the null fit (:1329)
    {   // certification maxima: reduced over the warp first (bit patterns of non-negative doubles order like the values), one
        // atomic per warp and quantity -- a per-thread atomic on two addresses serialised the whole kernel (4.5 ms per 16 k SNPs)
        unsigned long long r0 = 0ull, r1 = 0ull;
        if (!degenerate && valid) {
            r0 = (unsigned long long)__double_as_longlong(err / sxx);
            const double exy = es[0] * prm.rem * x1;
            r1 = (unsigned long long)__double_as_longlong(exy * sqrt(prm.n_p / (sxx * h0)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a0 = __shfl_xor_sync(0xffffffffu, r0, o), a1 = __shfl_xor_sync(0xffffffffu, r1, o);
            r0 = a0 > r0 ? a0 : r0;
            r1 = a1 > r1 ? a1 : r1;
        }
        if ((threadIdx.x & 31) == 0) {
            if (r0 > prm.rho_max[0]) atomicMax(prm.rho_max, r0);
            if (r1 > prm.rho_max[1]) atomicMax(prm.rho_max + 1, r1);
        }
    }
    if (!valid) return;
    const int64_t o = (int64_t)t * prm.out_stride + prm.out_row0 + s;
    if (prm.xx) prm.xx[o] = sxx;
    if (prm.xy) prm.xy[o] = sxy;
    double rss = h0, f = 0.0, vp = 0.0, pv = 1.0;
    if (!degenerate) {
        const double r2 = (sxy * sxy) / (sxx * h0);
        const double rs = h0 - (sxy * sxy) / sxx;
        if (rs != 0.0) {
            rss = rs;
            vp = r2;
            f = prm.n_p * r2 / (1.0 - r2);
            pv = f_sf(f, 1.0, prm.n_p, prm.lbeta);
        }
    }
    if (prm.rss) prm.rss[o] = rss;
    if (prm.f) prm.f[o] = f;
    if (prm.var_perc) prm.var_perc[o] = vp;
    if (prm.p) prm.p[o] = pv;
}

}  // namespace mmg
