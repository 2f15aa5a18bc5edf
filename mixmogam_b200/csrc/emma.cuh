// Exact EMMA for a batch of SNPs WITHOUT a per-SNP eigendecomposition (linear_models.py:931-968 `expedited_REML_t_test`,
// :771-927 `get_estimates` with xs, and the ML branch :811-824 / get_ML :672-696).
//
// The reference fits, for every tested SNP x, the variance components of  y ~ [X0, x]  by EMMA: it eigendecomposes
// S(K+I)S with S = I - X(X'X)^-1X' for X = [X0, x] -- one n x n eigh PER SNP (:788, ~1 s at n = 10k on this GPU) -- and
// evaluates, over a grid of delta and a secant refinement,
//     s1 = sum eta^2/(xi+delta) ,  s2 = sum log(xi+delta) ,  s3 = sum eta^2/(xi+delta)^2 ,  s4 = sum 1/(xi+delta)
// with xi the p = n - q non-trivial eigenvalues (minus 1) and eta = U_R y.  All four are functions of
// P(delta) = H^-1 - H^-1 X (X'H^-1 X)^-1 X'H^-1,  H = K + delta I  (S(K + delta I)S restricted to the complement of X has the
// eigenvalues xi + delta and P is its pseudo-inverse):
//     s1 = y'Py ,   s3 = y'P^2 y ,   s4 = tr P ,   s2 = log|H| + log|X'H^-1 X| - log|X'X| .
// In the eigenbasis of K ALONE (eig_L, computed once: U, lambda) H^-1 is diagonal, w_i = 1/(lambda_i + delta), so with the
// rotated columns Z = U [X0, x, y] every quantity is a handful of weighted moments
//     M1 = Z' W Z ,   M2 = Z' W^2 Z ,   sum w ,   sum log(lambda + delta)          (O(n q^2) per delta)
// followed by (q x q) Cholesky solves.  One GEMM rotates all k SNPs at once; the grid, the secant refinement (the same
// reml_logic.cuh the REML stage uses, step for step) and the final GLS fit then cost O(k g n q^2) flops: milliseconds for the
// reference's default emma_num = 100 instead of 100 eigendecompositions.  Identical to the reference in exact arithmetic;
// FP64 throughout (the reference's grid is float32).
//
// ML branch: ll = 0.5 (n (log(n/2pi) - 1 - log s1) - sum log(lambda_L + delta)),  dll ~ n s3/s1 - sum 1/(lambda_L + delta):
// the same moments with n in place of p and the sums over eig_L's values.
#pragma once
#include "fdist.cuh"
#include "reml.cuh"

namespace mmg {

constexpr int EMMA_QMAX = 8;                  // fixed-effect columns including the tested SNP
constexpr int EMMA_OUT = 9;                   // per-SNP outputs before the betas: delta, max_ll, vg, ve, f, p, var_perc, rss, mahalanobis_rss

struct EmmaParams {
    int n, q0, has_snp, method;               // method 0 = REML, 1 = ML
    const double* lam;                        // [n] eigenvalues of K (eig_L)
    const double* Z0;                         // [n x (q0 + 1)] rotated fixed effects and phenotype, U [X0, y], row-major
    const double* G;                          // [k x n] rotated SNPs U x_s (has_snp)
    const double* deltas;                     // [g]
    int g;
    double esp;
    double lbeta;                             // ln B(p/2, 1/2) of the F(1, p) survival function, p = n - q
    double *lls, *dlls;                       // [k x g]
    double* out;                              // [k x (EMMA_OUT + q)]
};

// symmetric (Q1 x Q1) moment matrices stored as packed upper triangles: index(a, b), a <= b
template <int Q1>
__device__ __forceinline__ constexpr int emma_idx(int a, int b) { return a * Q1 - a * (a - 1) / 2 + (b - a); }

template <int Q1>
struct EmmaEval {
    static constexpr int Q = Q1 - 1;          // fixed-effect columns (X0 and, if present, the SNP); column Q is y
    static constexpr int NS = Q1 * (Q1 + 1) / 2;
    const EmmaParams& prm;
    const double* gs;                         // rotated SNP of this block (or nullptr)
    double* red;                              // [(2 NS + 2) x 8] shared
    double logdet_xx;                         // log|X'X| (delta independent)
    double sum_sq_etas;                       // y'Sy = sum of the squared etas of the reference (:795)

    __device__ __forceinline__ void load_row(int i, double (&z)[Q1]) const {
        const int q0 = prm.q0;
        const double* r = prm.Z0 + (int64_t)i * (q0 + 1);
#pragma unroll
        for (int a = 0; a < Q1; ++a) {
            if (a < q0) z[a] = r[a];
            else if (a == Q) z[a] = r[q0];
            else z[a] = gs[i];                // a == q0 < Q: the SNP column
        }
    }

    // moments at delta: M1 = Z'WZ, M2 = Z'W^2 Z (packed), sw = sum w, slog = sum log(lambda + delta); power 0: M1 = Z'Z only
    __device__ void moments(double delta, bool unit_weights, double (&M1)[NS], double (&M2)[NS], double& sw, double& slog) {
        double acc[2 * NS + 2];
#pragma unroll
        for (int k = 0; k < 2 * NS + 2; ++k) acc[k] = 0.0;
        for (int i = threadIdx.x; i < prm.n; i += REML_THREADS) {
            double z[Q1];
            load_row(i, z);
            const double v = prm.lam[i] + delta;
            const double w = unit_weights ? 1.0 : 1.0 / v;
            const double w2 = w * w;
#pragma unroll
            for (int a = 0; a < Q1; ++a)
#pragma unroll
                for (int b = a; b < Q1; ++b) {
                    const double zz = z[a] * z[b];
                    acc[emma_idx<Q1>(a, b)] = fma(w, zz, acc[emma_idx<Q1>(a, b)]);
                    acc[NS + emma_idx<Q1>(a, b)] = fma(w2, zz, acc[NS + emma_idx<Q1>(a, b)]);
                }
            acc[2 * NS] += w;
            acc[2 * NS + 1] += unit_weights ? 0.0 : log(v);
        }
        block_reduce_sum<2 * NS + 2>(acc, red);
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            M1[k] = acc[k];
            M2[k] = acc[NS + k];
        }
        sw = acc[2 * NS];
        slog = acc[2 * NS + 1];
    }

    // Cholesky of the leading (m x m) block of the packed symmetric M into L (row-major lower, [Q][Q]); returns log det
    __device__ static double chol(const double (&M)[NS], int m, double (&L)[Q > 0 ? Q : 1][Q > 0 ? Q : 1]) {
        double logdet = 0.0;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            if (j >= m) break;
            double d = M[emma_idx<Q1>(j, j)];
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k < j) d -= L[j][k] * L[j][k];
            d = sqrt(d);
            L[j][j] = d;
            logdet += 2.0 * log(d);
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                if (i <= j || i >= m) continue;
                double s = M[emma_idx<Q1>(j, i)];
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    if (k < j) s -= L[i][k] * L[j][k];
                L[i][j] = s / d;
            }
        }
        return logdet;
    }
    // solve (L L') x = b for the leading m unknowns, in place
    __device__ static void chol_solve(const double (&L)[Q > 0 ? Q : 1][Q > 0 ? Q : 1], int m, double (&x)[Q > 0 ? Q : 1]) {
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            if (i >= m) break;
            double s = x[i];
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k < i) s -= L[i][k] * x[k];
            x[i] = s / L[i][i];
        }
#pragma unroll
        for (int ii = 0; ii < Q; ++ii) {
            const int i = m - 1 - ii;
            if (i < 0) break;
            double s = x[i];
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k > i && k < m) s -= L[k][i] * x[k];
            x[i] = s / L[i][i];
        }
    }

    struct Fit {
        double s1, s3, s4, slog_all, beta[Q > 0 ? Q : 1];
    };
    // GLS fit of y on the m leading columns at the moments (M1, M2): s1 = y'Py, s3 = y'P^2y, s4 = tr P, slog_all = the log
    // determinant sum of the REML likelihood  (sum log(lambda + delta) + log|X'WX| - log|X'X|)
    __device__ void fit(const double (&M1)[NS], const double (&M2)[NS], double sw, double slog, int m, Fit& f) const {
        double L[Q > 0 ? Q : 1][Q > 0 ? Q : 1];
        const double logdet = chol(M1, m, L);
        double beta[Q > 0 ? Q : 1];
#pragma unroll
        for (int a = 0; a < Q; ++a) beta[a] = a < m ? M1[emma_idx<Q1>(a, Q)] : 0.0;
        chol_solve(L, m, beta);
        double s1 = M1[emma_idx<Q1>(Q, Q)], s3 = M2[emma_idx<Q1>(Q, Q)];
#pragma unroll
        for (int a = 0; a < Q; ++a) {
            if (a >= m) break;
            s1 -= beta[a] * M1[emma_idx<Q1>(a, Q)];
            s3 -= 2.0 * beta[a] * M2[emma_idx<Q1>(a, Q)];
#pragma unroll
            for (int b = 0; b < Q; ++b)
                if (b < m) s3 += beta[a] * beta[b] * M2[a <= b ? emma_idx<Q1>(a, b) : emma_idx<Q1>(b, a)];
        }
        // tr(A^-1 A2): column by column
        double tr = 0.0;
#pragma unroll
        for (int c = 0; c < Q; ++c) {
            if (c >= m) break;
            double col[Q > 0 ? Q : 1];
#pragma unroll
            for (int a = 0; a < Q; ++a) col[a] = a < m ? M2[a <= c ? emma_idx<Q1>(a, c) : emma_idx<Q1>(c, a)] : 0.0;
            chol_solve(L, m, col);
            tr += col[c];
        }
        f.s1 = s1;
        f.s3 = s3;
        f.s4 = sw - tr;
        f.slog_all = slog + logdet - logdet_xx;
#pragma unroll
        for (int a = 0; a < Q; ++a) f.beta[a] = beta[a];
    }

    // delta-independent pieces: log|X'X| and y'Sy (unit weights)
    __device__ void init() {
        double M1[NS], M2[NS], sw, slog;
        logdet_xx = 0.0;
        moments(0.0, true, M1, M2, sw, slog);
        Fit f;
        double L[Q > 0 ? Q : 1][Q > 0 ? Q : 1];
        logdet_xx = chol(M1, Q, L);
        fit(M1, M2, sw, slog, Q, f);
        sum_sq_etas = f.s1;
    }

    __device__ void eval(double delta, double& ll, double& dll2) {      // ll and the (doubled) derivative the secant runs on
        double M1[NS], M2[NS], sw, slog;
        moments(delta, false, M1, M2, sw, slog);
        Fit f;
        fit(M1, M2, sw, slog, Q, f);
        if (prm.method == 0) {
            const double pd = (double)(prm.n - Q);
            ll = 0.5 * (pd * (log(pd / (2.0 * M_PI)) - 1.0 - log(f.s1)) - f.slog_all);      // :807 == _rell_ :618-623
            dll2 = pd * f.s3 / f.s1 - f.s4;                                                 // _redll_ :626-631
        } else {
            const double nd = (double)prm.n;
            ll = 0.5 * (nd * (log(nd / (2.0 * M_PI)) - 1.0 - log(f.s1)) - slog);            // :821 == _ll_ :634-641
            dll2 = nd * f.s3 / f.s1 - sw;                                                   // _dll_ :644-650
        }
    }
    // interface of reml_logic.cuh
    __device__ double redll(double delta) {
        double ll, d;
        eval(delta, ll, d);
        return d;
    }
    __device__ double rell(double delta) {
        double ll, d;
        eval(delta, ll, d);
        return ll;
    }
};

template <int Q1>
static __global__ void __launch_bounds__(REML_THREADS) emma_grid_kernel(const EmmaParams prm) {
    __shared__ double red[(2 * EmmaEval<Q1>::NS + 2) * 8];
    const int gi = blockIdx.x, s = blockIdx.y;
    EmmaEval<Q1> ev{prm, prm.has_snp ? prm.G + (int64_t)s * prm.n : nullptr, red, 0.0, 0.0};
    ev.init();
    double ll, d;
    ev.eval(prm.deltas[gi], ll, d);
    if (threadIdx.x == 0) {
        prm.lls[(int64_t)s * prm.g + gi] = ll;
        prm.dlls[(int64_t)s * prm.g + gi] = 0.5 * d;          // :810 / :824
    }
}

template <int Q1>
static __global__ void __launch_bounds__(REML_THREADS) emma_refine_kernel(const EmmaParams prm) {
    using Ev = EmmaEval<Q1>;
    constexpr int Q = Ev::Q, NS = Ev::NS;
    __shared__ double red[(2 * NS + 2) * 8];
    const int s = blockIdx.x;
    Ev ev{prm, prm.has_snp ? prm.G + (int64_t)s * prm.n : nullptr, red, 0.0, 0.0};
    ev.init();
    double od, ol;
    int fl;
    reml_refine(ev, prm.lls + (int64_t)s * prm.g, prm.dlls + (int64_t)s * prm.g, prm.deltas, prm.g, prm.esp, &od, &ol, &fl);
    // the fit at delta-hat (:893-927): betas and Mahalanobis RSS of the full model, the null model without the SNP column,
    // the residual sum of squares in the original space, vg / ve with the reference's broadcast (closed form, :894-896)
    double M1[NS], M2[NS], sw, slog;
    ev.moments(od, false, M1, M2, sw, slog);
    typename Ev::Fit full, null_fit;
    ev.fit(M1, M2, sw, slog, Q, full);
    double h0 = full.s1;
    if (prm.has_snp) {
        // null model: the q0 leading columns only.  The packed layout keeps (a, Q) = column y for every a, so the same fit
        // routine with m = q0 does it.
        ev.fit(M1, M2, sw, slog, prm.q0, null_fit);
        h0 = null_fit.s1;
    }
    double U1[NS], U2[NS], usw, uslog;
    ev.moments(0.0, true, U1, U2, usw, uslog);                // Z'Z for the untransformed residual  |y - X beta|^2
    double rss = U1[emma_idx<Q1>(Q, Q)];
#pragma unroll
    for (int a = 0; a < Q; ++a) {
        rss -= 2.0 * full.beta[a] * U1[emma_idx<Q1>(a, Q)];
#pragma unroll
        for (int b = 0; b < Q; ++b) rss += full.beta[a] * full.beta[b] * U1[a <= b ? emma_idx<Q1>(a, b) : emma_idx<Q1>(b, a)];
    }
    if (threadIdx.x == 0) {
        const double pd = (double)(prm.n - Q);
        double* o = prm.out + (int64_t)s * (EMMA_OUT + Q);
        const double vg = ev.sum_sq_etas * full.s4 / pd;                       // :894-896 in closed form
        const double f = (h0 / full.s1 - 1.0) * pd;                            // :920 (xs.shape[1] == 1)
        o[0] = od;
        o[1] = ol;
        o[2] = vg;
        o[3] = vg * od;
        o[4] = f;
        o[5] = prm.has_snp ? f_sf(f, 1.0, pd, prm.lbeta) : 1.0;                 // :925
        o[6] = 1.0 - full.s1 / h0;                                             // :921
        o[7] = rss;
        o[8] = full.s1;
#pragma unroll
        for (int a = 0; a < Q; ++a) o[EMMA_OUT + a] = full.beta[a];
        (void)fl;
    }
}

// int8 genotype rows -> FP64 rows (the top hits of a scan are refined straight from the resident genotype block)
static __global__ void gather_rows_f64_kernel(const int8_t* __restrict__ snps, int64_t pitch, const long long* __restrict__ rows, int n,
                                              double* __restrict__ out) {
    const int64_t r = rows[blockIdx.y];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[(int64_t)blockIdx.y * n + j] = (double)snps[r * pitch + j];
}

}  // namespace mmg
