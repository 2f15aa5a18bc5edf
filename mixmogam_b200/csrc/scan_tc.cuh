// Stage 3 (int8 tensor-core path): x~.x~ = x'(R'R)x evaluated on tcgen05 int8 tensor cores.
//
// The left operand of the EMMAX rotation (linear_models.py:1317-1318) is an exact small-integer genotype
// vector, so only the FP64 matrix has to be split.  With A = R'R (symmetric, FP64):
//     x'Ax = sum_j A_jj x_j^2  +  sum_j x_j * sum_{i<j} 2 A_ji x_i
// The diagonal term stays in FP64 (epilogue).  B[j][i] = 2 A_ji * 2^-E (i < j, |B| <= 0.498) is cut into S signed
// base-256 digits b_k in [-128,127] (digits.cuh: every value of an int8, 8 bits per plane):
//     B = sum_k 256^-(k+1) b_k  (+ error <= (128/255) * 256^-S),
// every partial product x.b_k is an exact int32 (tcgen05.mma kind::i8), and the epilogue folds
//     q_s += w_k * sum_j acc[s][j] * x[s][j],   w_k = 2^E 256^-(k+1)
// in FP64.  Because B is lower triangular, N tile jb only needs K in [0, 256(jb+1)): half the MACs of
// the full rotation.  The same epilogue accumulates x~.y~ = x.(R'y~) and, once a 128-SNP row block has
// seen every tile of a phenotype, evaluates RSS / F / p (linear_models.py:1345-1349).
//
// Phenotype batching: T phenotypes with their own delta_t (hence their own A_t = R_t'R_t) are scanned in ONE
// launch -- the tile table lists the slices of A_0, A_1, ... one after the other, the 128-SNP genotype block
// (operand A of the MMA) is reused from shared memory/L2 by all T*S*ceil(n/256) tiles of its group.
//
// Permutation scan (linear_models.py:1157-1164): PermEpi contracts the genotype block with the digit planes
// of W = R'Ys ([P x n], all permuted phenotypes rotated back), 32 permutations x 8 slices per 256-column
// tile, and keeps the per-permutation maximum of (x_c.W_p)^2 / (x~_c.x~_c) over SNPs.
#pragma once
#include <algorithm>

#include "digits.cuh"
#include "fdist.cuh"
#include "tc_gemm.cuh"

namespace mmg {

constexpr int QS_MAX_SLICES = DIGIT256_MAX_PLANES;
constexpr int QS_FLAG_XY = 1;        // TcTile.aux1: accumulate x.v on this tile (first slice of a column tile)
constexpr int QS_FLAG_FIRST = 2;     //              first tile of a phenotype: reset the running sums
constexpr int QS_FLAG_LAST = 4;      //              last tile of a phenotype: evaluate and store
constexpr int QS_PHEN_SHIFT = 8;     //              phenotype index = aux1 >> 8
constexpr double QS_DEGENERATE_REL = 1e-8;   // x~.x~ <= this fraction of sum_j A_jj x_j^2: the SNP is collinear with the fixed effects

struct QuadEpi {
    struct Params {
        const int8_t* snps;
        int64_t pitch;
        int64_t row_begin, row_count;
        double w[QS_MAX_SLICES];     // slice weights 2^(-8(k+1)); the per-phenotype 2^E_t is in escale
        const double* escale;        // [T] 2^E_t
        const double* v;             // [T][v_stride] R_t'y~_t (zero padded)
        const double* dg;            // [T][v_stride] diag(R_t'R_t): the diagonal of the quadratic form is kept in FP64
        const double* bscale;        // [T] (64/255) * 256^-S * 2^E_t: truncation bound of the off-diagonal digits per unit ||x||_1^2
        unsigned long long* rho_max; // max over SNPs of bound / (x~.x~), bits of a non-negative double (certification)
        const double* pre_xy;        // [T][pre_stride] x.v_t, [T][pre_stride] sum_j A_jj x_j^2 and [pre_stride] ||x||_1 from
        const double* pre_qd;        //   snp_prepass_kernel (scan_quad_kernel reads them; the tile-table kernel computes its own)
        const double* pre_a1;
        int64_t pre_stride;
        int64_t v_stride;
        const double* h0_rss;        // [T]
        double n_p, lbeta;
        int64_t out_stride;          // outputs are [T][out_stride]
        double *xx, *xy, *rss, *f, *p, *var_perc;
        int defer;                   // 1: store() writes x~.x~ only; quad_finish_kernel derives everything else afterwards
    };
    double q, xy, qd, xl1;
    const int8_t* xrow;
    int64_t orow;

    __device__ __forceinline__ void begin_group(const Params& p, int g, int row) {
        q = 0.0;
        xy = 0.0;
        qd = 0.0;
        xl1 = 0.0;
        orow = (int64_t)g * TC_BM + row;
        xrow = (orow < p.row_count) ? p.snps + (p.row_begin + orow) * p.pitch : nullptr;
    }
    __device__ __forceinline__ void end_group(const Params&, int, int) {}
    __device__ __forceinline__ int tile_begin(const Params&, const TcTile& t, int) {
        if (t.aux1 & QS_FLAG_FIRST) {
            q = 0.0;
            xy = 0.0;
            qd = 0.0;
            xl1 = 0.0;
        }
        return TC_BN / 32;           // uniform across the warp: tcgen05.ld is warp-collective
    }
    __device__ __forceinline__ void chunk(const Params& p, const TcTile& t, int row, int c, const uint32_t (&v)[32]) {
        if (xrow == nullptr) return;
        const int col0 = t.col0 + c * 32;
        const uint4* xp = reinterpret_cast<const uint4*>(xrow + col0);
        const uint4 x0 = xp[0], x1 = xp[1];
        const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        // the accumulator spans the whole contraction here (|acc| <= n |x| 128): 32 columns x |x| can pass 2^31 at
        // n = 10k, |x| = 8, so this (non-default) kernel sums in 64 bits; the panel kernel's per-panel sums fit int32
        long long s = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int xv = (int)(int8_t)((xw[j >> 2] >> (8 * (j & 3))) & 0xffu);
            s += (long long)((int)v[j]) * xv;
        }
        q = fma(p.w[t.aux0], (double)s, q);
        if (t.aux1 & QS_FLAG_XY) {
            const double* vv = p.v + (int64_t)(t.aux1 >> QS_PHEN_SHIFT) * p.v_stride + col0;
            const double* dd = p.dg + (int64_t)(t.aux1 >> QS_PHEN_SHIFT) * p.v_stride + col0;
            int a1 = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int xv = (int)(int8_t)((xw[j >> 2] >> (8 * (j & 3))) & 0xffu);
                xy = fma((double)xv, vv[j], xy);
                qd = fma((double)(xv * xv), dd[j], qd);
                a1 += abs(xv);
            }
            xl1 += (double)a1;
        }
    }
    __device__ __forceinline__ void tile_end(const Params& p, const TcTile& t, int row, int lane) {
        if (!(t.aux1 & QS_FLAG_LAST) || xrow == nullptr) return;
        store(p, t.aux1 >> QS_PHEN_SHIFT, orow, q, xy, qd, xl1);
    }
    // RSS / F / p of one SNP for phenotype ph (linear_models.py:1329,1345-1349) from
    //   q  = off-diagonal part of x'Ax in units of 2^E (digit planes), qd = sum_j A_jj x_j^2 (FP64), xy = x.(R'y~),
    //   x1 = ||x||_1 (for the certified truncation bound |dq| <= (64/255) 256^-S 2^E ||x||_1^2)
    static __device__ __forceinline__ void store(const Params& p, int ph, int64_t orow, double q, double xy, double qd, double x1) {
        const double sxx = fma(q, p.escale[ph], qd);
        if (p.defer) {               // the statistics (divisions, log / exp, an atomic) leave the scan's epilogue: one store per SNP here
            p.xx[(int64_t)ph * p.out_stride + orow] = sxx;
            return;
        }
        finish(p, ph, orow, sxx, xy, qd, x1);
    }
    static __device__ __forceinline__ void finish(const Params& p, int ph, int64_t orow, double sxx, double sxy, double qd, double x1) {
        const int64_t o = (int64_t)ph * p.out_stride + orow;
        const double h0 = p.h0_rss[ph];
        // x~ numerically zero -- x is (nearly) in the span of the fixed effects, e.g. a monomorphic SNP or one collinear with a
        // cofactor: x~.x~ is then the difference of two equal numbers (qd = sum_j A_jj x_j^2 bounds its scale) and the statistic
        // carries no information.  Such a SNP keeps the null fit, like the reference's empty-residue case (`if rss:`,
        // linear_models.py:1329), and stays out of the certification maximum (its relative bound is meaningless).
        const bool degenerate = !(sxx > QS_DEGENERATE_REL * qd);
        if (p.rho_max != nullptr && !degenerate) {
            const double rho = p.bscale[ph] * x1 * x1 / sxx;
            atomicMax(p.rho_max, (unsigned long long)__double_as_longlong(rho));
        }
        if (p.xx) p.xx[o] = sxx;
        if (p.xy) p.xy[o] = sxy;
        double rss = h0, f = 0.0, vp = 0.0, pv = 1.0;
        if (!degenerate) {
            const double r2 = (sxy * sxy) / (sxx * h0);
            const double rs = h0 - (sxy * sxy) / sxx;
            if (rs != 0.0) {
                rss = rs;
                vp = r2;
                f = p.n_p * r2 / (1.0 - r2);
                pv = f_sf(f, 1.0, p.n_p, p.lbeta);
            }
        }
        if (p.rss) p.rss[o] = rss;
        if (p.f) p.f[o] = f;
        if (p.var_perc) p.var_perc[o] = vp;
        if (p.p) p.p[o] = pv;
    }
};

// the statistics of a deferred scan (QuadEpi::Params::defer): one thread per (phenotype, SNP), x~.x~ from the scan, the linear
// terms from the pre-pass
static __global__ void __launch_bounds__(256) quad_finish_kernel(const QuadEpi::Params p, int T) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)T * p.row_count) return;
    const int ph = (int)(i / p.row_count);
    const int64_t orow = i - (int64_t)ph * p.row_count;
    const int64_t po = (int64_t)ph * p.pre_stride + orow;
    QuadEpi::finish(p, ph, orow, p.xx[(int64_t)ph * p.out_stride + orow], p.pre_xy[po], p.pre_qd[po], p.pre_a1[orow]);
}

// ---- permutation scan epilogue --------------------------------------------------------------------------
// B operand rows of perm block b: row (b*8 + k)*32 + j = digit plane k of permutation 32 b + j.
constexpr int PS_SLICES = 8;

struct PermEpi {
    struct Params {
        int64_t row_count;
        double w[PS_SLICES];          // 2^E 128^-(k+1)
        const double* xx;             // [row_count] x~_c.x~_c (from the quadratic-form scan of the centred rotation)
        const double* mu;             // [row_count] SNP means (nullptr when SNPs are not centred)
        const double* wsum;           // [P_pad] sum_i W[p][i]
        unsigned long long* ratio;    // [P_pad] running max, bits of non-negative doubles
    };
    double d[32];
    double inv_xx, mu;

    __device__ __forceinline__ void begin_group(const Params& p, int g, int row) {
        const int64_t r = (int64_t)g * TC_BM + row;
        inv_xx = 0.0;
        mu = 0.0;
        if (r < p.row_count) {
            const double sxx = p.xx[r];
            inv_xx = sxx > 0.0 ? 1.0 / sxx : 0.0;     // x~ = 0: lstsq returns an empty residue, the minimum is kept (:1164)
            mu = p.mu ? p.mu[r] : 0.0;
        }
    }
    __device__ __forceinline__ void end_group(const Params&, int, int) {}
    __device__ __forceinline__ int tile_begin(const Params&, const TcTile&, int) { return PS_SLICES; }
    __device__ __forceinline__ void chunk(const Params& p, const TcTile&, int, int c, const uint32_t (&v)[32]) {
        const double wk = p.w[c];
        if (c == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) d[j] = wk * (double)(int)v[j];
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) d[j] = fma(wk, (double)(int)v[j], d[j]);
        }
    }
    __device__ __forceinline__ void tile_end(const Params& p, const TcTile& t, int, int lane) {
        const double* ws = p.wsum + t.col0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const double dot = d[j] - mu * ws[j];                 // x_c.W_p = x.W_p - mean(x) sum(W_p)
            d[j] = dot * dot * inv_xx;
        }
        // transposing max-reduction over the 32 lanes (rows): lane l ends with max over rows of permutation l
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < o; ++i) {
                const double send = up ? d[i] : d[i + o];
                const double keep = up ? d[i + o] : d[i];
                const double recv = __shfl_xor_sync(0xffffffffu, send, o);
                d[i] = fmax(keep, recv);
            }
        }
        atomicMax(p.ratio + t.col0 + lane, (unsigned long long)__double_as_longlong(d[0]));
    }
};

// ---- linear pre-pass of the scan ------------------------------------------------------------------------------
// Per SNP s and phenotype t:  xy[t][s] = x_s.v_t  (= x~.y~, linear_models.py:1328),  qd[t][s] = sum_j A_t[j][j] x_sj^2  (the
// FP64 diagonal of the quadratic form),  a1[s] = ||x_s||_1 (for the certified truncation bound).  One warp per SNP, the
// genotype row read once, coalesced (16 bytes per lane), v_t and diag(A_t) from L1.  int8 -> double without the
// quarter-rate I2F: the byte is biased to 0..255, placed in the mantissa of 2^52 and the bias removed by one exact DADD.
// The scan epilogue used to do this inside its accumulator-drain loop (I2F + dependent DFMA chains, 16 k cycles per
// column tile during which the MMA ran out of accumulators); here it is an HBM-rate stream of its own.
constexpr int PRE_ROWS = 2;      // SNP rows per warp (v_t / diag(A_t) are fetched from L1 once per PRE_ROWS rows); measured on 262 144 SNPs x 10 k
                                 // (profiles/r01_microbench_prepass.txt): <2 rows, unroll 4, 4 blocks/SM> 1.82 ms, <4, 4, 3> 2.61 ms, <8, 4, 2> 2.80 ms
template <int PRE_ROWS = 4, int PRE_UNROLL = 4, int PRE_MINB = 1>
static __global__ void __launch_bounds__(256, PRE_MINB) snp_prepass_kernel(const int8_t* __restrict__ snps, int64_t pitch, int64_t row_begin, int64_t row_count,
                                                          int T, const double* __restrict__ v, const double* __restrict__ dg, int64_t v_stride,
                                                          double* __restrict__ xy, double* __restrict__ qd, double* __restrict__ a1,
                                                          int64_t out_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * PRE_ROWS;
    if (r0 >= row_count) return;
    const int nr = (int)min((int64_t)PRE_ROWS, row_count - r0);
    const int8_t* xrow = snps + (row_begin + r0) * pitch;
    const double kBias = 4503599627370496.0 + 128.0;               // 2^52 + 128
    double l1[PRE_ROWS];             // ||x||_1 in FP64 too (|.| is a free operand modifier; the byte-SIMD / dot-product
#pragma unroll                       // instructions run on the quarter-rate XU pipe and bounded the first version of this kernel)
    for (int i = 0; i < PRE_ROWS; ++i) l1[i] = 0.0;
    for (int t = 0; t < T; ++t) {
        const double* vt = v + (int64_t)t * v_stride;
        const double* dt = dg + (int64_t)t * v_stride;
        double sxy[PRE_ROWS], sqd[PRE_ROWS];
#pragma unroll
        for (int i = 0; i < PRE_ROWS; ++i) sxy[i] = sqd[i] = 0.0;
        // every load is fully coalesced: lane l takes columns c + 2l, c + 2l + 1 (one double2 of v_t and of diag(A_t), one
        // 16-bit genotype pair per row); 64 columns per warp step.  pitch is a multiple of 256, v / dg are zero padded past n.
#pragma unroll PRE_UNROLL
        for (int64_t c = 2 * lane; c < pitch; c += 64) {
            const double2 vv = __ldg(reinterpret_cast<const double2*>(vt + c));
            const double2 dd = __ldg(reinterpret_cast<const double2*>(dt + c));
#pragma unroll
            for (int i = 0; i < PRE_ROWS; ++i) {
                const uint32_t g = i < nr ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(xrow + i * pitch + c)) : 0u;
                const uint32_t b = g ^ 0x8080u;                      // bytes + 128
                const double x0 = __hiloint2double(0x43300000, (int)(b & 0xffu)) - kBias;
                const double x1 = __hiloint2double(0x43300000, (int)(b >> 8)) - kBias;
                sxy[i] = fma(x0, vv.x, sxy[i]);
                sqd[i] = fma(x0 * x0, dd.x, sqd[i]);
                sxy[i] = fma(x1, vv.y, sxy[i]);
                sqd[i] = fma(x1 * x1, dd.y, sqd[i]);
                if (t == 0) l1[i] += fabs(x0) + fabs(x1);
            }
        }
#pragma unroll
        for (int i = 0; i < PRE_ROWS; ++i) {
            double a = sxy[i], b = sqd[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (lane == 0 && i < nr) {
                xy[(int64_t)t * out_stride + r0 + i] = a;
                qd[(int64_t)t * out_stride + r0 + i] = b;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < PRE_ROWS; ++i) {
        double a = l1[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0 && i < nr) a1[r0 + i] = a;
    }
}

// Storage of the quadratic form A = R'R (only its lower triangle, i <= j, is ever read):
//   dense  : row-major [n x ld]
//   packed : the 256 x 256 blocks (J, I) with I <= J back to back, block (J, I) in slot J (J + 1) / 2 + I, row-major inside a
//            block.  Half the memory of the padded square, and a contiguous range of slots is a contiguous range of bytes --
//            which is what lets the ranks of a multi-GPU run each form a slot range and all-gather the result.
constexpr int QA_TILE = 256;
constexpr int64_t QA_TILE_ELEMS = (int64_t)QA_TILE * QA_TILE;
__host__ __device__ __forceinline__ int64_t qa_slot(int J, int I) { return (int64_t)J * (J + 1) / 2 + I; }
template <bool PACKED>
__device__ __forceinline__ int64_t qa_index(int64_t ld, int j, int i) {          // element (j, i), i <= j
    if (!PACKED) return (int64_t)j * ld + i;
    return qa_slot(j >> 8, i >> 8) * QA_TILE_ELEMS + (int64_t)(j & 255) * QA_TILE + (i & 255);
}

// max |2 A[j][i]| over the strict lower triangle (i < j) -> bits of a non-negative double
template <bool PACKED>
static __global__ void quad_amax_kernel(const double* __restrict__ A, int64_t ld, int n, unsigned long long* __restrict__ amax_bits) {
    const int j = blockIdx.y;
    double m = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j; i += gridDim.x * blockDim.x) {
        const double a = fabs(A[qa_index<PACKED>(ld, j, i)]) * 2.0;
        m = fmax(m, a);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(amax_bits, (unsigned long long)__double_as_longlong(m));
}

// digits of B[j][i] = 2 A[j][i] 2^-E (i < j) into S stacked int8 planes Bq[(k * n_padN + j) * ldq + i]; the diagonal
// goes to dg[j] = A[j][j] and stays in FP64 (it is usually the largest entry: keeping it out of the digit planes lowers E)
template <bool PACKED>
static __global__ void quad_slice_kernel(const double* __restrict__ A, int64_t ld, int n, double scale /* 2^-E */, int S,
                                         int8_t* __restrict__ Bq, int64_t n_padN, int64_t ldq, double* __restrict__ dg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i > j || i >= n) return;
    if (i == j) {
        dg[j] = A[qa_index<PACKED>(ld, j, j)];
        return;
    }
    const double r = A[qa_index<PACKED>(ld, j, i)] * 2.0 * scale;        // |r| <= 0.498, exact scaling
    int d[DIGIT256_MAX_PLANES];
    digit256_split(r, S, d);                                             // base 256, exact (digits.cuh)
    for (int k = 0; k < S; ++k) Bq[((int64_t)k * n_padN + j) * ldq + i] = (int8_t)d[k];
}

// ---- A = R'R on the int8 tensor cores (exact digit-plane products) -------------------------------------------------
// The one-off FP64 product of the scan (cuBLAS dsyrk, 2 n^3 / 2 flops at the DMMA rate: 28 ms at n = 10k) as integer GEMMs:
// R 2^-F (|.| <= 0.498) is cut into P = 7 base-256 digit planes r_p (digits.cuh, exact), every plane product
//     G_pq[j][i] = sum_k r_p[k][j] r_q[k][i]           (int32, exact: n_out 128^2 < 2^31)
// comes from tcgen05.mma kind::i8, and the epilogue folds  A[j][i] += 2^2F 256^-(p+q+2) G_pq[j][i]  in FP64 for the
// 28 pairs with p + q < L = 7 (lower-triangular tiles only).  What is left out is bounded rigorously (ozaki_error_bound):
// dropped pairs  sum_{p+q>=L} 128^2 256^-(p+q+2)  and the digit remainders rho = (128/255) 256^-P per factor, times n_out --
// ~2^-41.7 2^2F at n = 10k, added to the certified truncation bound of the scan.
constexpr int OZ_PLANES = 7;
constexpr int OZ_LEVELS = 7;          // pairs (p, q) with p + q < OZ_LEVELS

// absolute error bound of the A entries in units of 2^2F
inline double ozaki_error_bound(int64_t n_out) {
    const double rho = DIGIT256_REM * ldexp(1.0, -8 * OZ_PLANES);
    double dropped = 0.0;
    for (int s = OZ_LEVELS; s <= 2 * OZ_PLANES - 2; ++s) {
        const int cnt = std::min(s + 1, 2 * OZ_PLANES - 1 - s);
        dropped += (double)cnt * 16384.0 * ldexp(1.0, -8 * (s + 2));
    }
    return (double)n_out * (dropped + rho + rho * rho) * 1.0000001;
}

// digit planes of R' (K-major operand): Op[(p n_padM + i) op_pitch + k] = digit_p(R[k][i] 2^-F); 32 x 32 transpose through smem
static __global__ void __launch_bounds__(256) ozaki_planes_kernel(const double* __restrict__ R, int64_t ldr, int n_out, int n, double scale,
                                                           int8_t* __restrict__ Op, int64_t n_padM, int64_t op_pitch) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int k = k0 + ty + 8 * r, i = i0 + tx;
        tile[ty + 8 * r][tx] = (k < n_out && i < n) ? R[(int64_t)k * ldr + i] * scale : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty + 8 * r, k = k0 + tx;
        if (i >= n || k >= n_out) continue;
        int d[DIGIT256_MAX_PLANES];
        digit256_split(tile[tx][ty + 8 * r], OZ_PLANES, d);
#pragma unroll
        for (int p = 0; p < OZ_PLANES; ++p) Op[((int64_t)p * n_padM + i) * op_pitch + k] = (int8_t)d[p];
    }
}

// epilogue: A[out_row][out_col + 32c ..] += w[p + q] acc.  The 28 (p, q) tiles of one output tile are consecutive tiles of ONE
// group (same CTA, same epilogue thread per row), so the read-modify-write needs no atomics and stays in L2.
struct OzakiEpi {
    struct Params {
        double* A;             // FP64, zeroed; lower-triangular tiles are written: dense [n_padM x ld], or packed 256 x 256 blocks
        int64_t ld;            //   (qa_index; `A` then points at slot `slot0`, the first one this launch owns)
        int64_t n_padM;
        int packed;
        int64_t slot0;
        double w[2 * OZ_PLANES];
    };
    __device__ __forceinline__ void begin_group(const Params&, int, int) {}
    __device__ __forceinline__ void end_group(const Params&, int, int) {}
    __device__ __forceinline__ int tile_begin(const Params&, const TcTile&, int) { return TC_BN / 32; }
    __device__ __forceinline__ void tile_end(const Params&, const TcTile&, int, int) {}
    __device__ __forceinline__ void chunk(const Params& p, const TcTile& t, int row, int c, const uint32_t (&v)[32]) {
        const int64_t orow = (int64_t)t.m0 - (int64_t)t.aux0 * p.n_padM + row;       // aux0 = p, aux1 = q
        const int64_t ocol = (int64_t)t.n0 - (int64_t)t.aux1 * p.n_padM + c * 32;
        double2* dst = reinterpret_cast<double2*>(
            p.packed ? p.A + (qa_slot((int)(orow >> 8), (int)(ocol >> 8)) - p.slot0) * QA_TILE_ELEMS + (orow & 255) * QA_TILE + (ocol & 255)
                     : p.A + orow * p.ld + ocol);
        const double wk = p.w[t.aux0 + t.aux1];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            double2 o = dst[j];
            o.x = fma(wk, (double)(int)v[2 * j + 0], o.x);
            o.y = fma(wk, (double)(int)v[2 * j + 1], o.y);
            dst[j] = o;
        }
    }
};

// out[i] = sum_k R[k][i]^2 = diag(R'R) of the row-major [rows x cols] matrix (deterministic: fixed split of k over the 8 warps)
static __global__ void __launch_bounds__(256) col_sumsq_kernel(const double* __restrict__ R, int64_t ld, int rows, int cols, double* __restrict__ out) {
    __shared__ double part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + tx;
    double s = 0.0;
    if (i < cols)
        for (int k = ty; k < rows; k += 8) {
            const double r = R[(int64_t)k * ld + i];
            s = fma(r, r, s);
        }
    part[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < cols) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += part[w][tx];
        out[i] = t;
    }
}

// max |W| over a row-major [rows x cols] matrix -> bits of a non-negative double
static __global__ void mat_amax_kernel(const double* __restrict__ W, int64_t ld, int rows, int cols, unsigned long long* __restrict__ amax_bits) {
    const int r = blockIdx.y;
    double m = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cols; i += gridDim.x * blockDim.x) m = fmax(m, fabs(W[(int64_t)r * ld + i]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(amax_bits, (unsigned long long)__double_as_longlong(m));
}

// digits of W[p][i] 2^-E into the permutation operand: Wq[((p/32)*8 + k)*32 + p%32][i], k < 8
static __global__ void perm_slice_kernel(const double* __restrict__ W, int64_t ld, int P, int n, double scale, int8_t* __restrict__ Wq,
                                  int64_t ldq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (i >= n || p >= P) return;
    double r = W[(int64_t)p * ld + i] * scale;                           // |r| < 0.5
    const int64_t base = ((int64_t)(p >> 5) * PS_SLICES * 32 + (p & 31)) * ldq + i;
#pragma unroll
    for (int k = 0; k < PS_SLICES; ++k) {
        r *= 128.0;
        const double d = rint(r);
        r -= d;
        Wq[base + (int64_t)k * 32 * ldq] = (int8_t)(int)d;
    }
}

// R[r][:] -= r1[r] / n   (right-multiplication by the centring matrix C = I - 11'/n: R C = R - (R 1) 1'/n)
static __global__ void centre_cols_kernel(double* __restrict__ R, int64_t ld, int rows, int cols, const double* __restrict__ r1, double inv_n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j < cols && i < rows) R[(int64_t)i * ld + j] -= r1[i] * inv_n;
}

}  // namespace mmg
