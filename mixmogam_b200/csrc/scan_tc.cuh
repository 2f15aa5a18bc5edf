// Stage 3 (int8 tensor-core path): x~.x~ = x'(R'R)x evaluated on tcgen05 int8 tensor cores.
//
// The left operand of the EMMAX rotation (linear_models.py:1317-1318) is an exact small-integer genotype
// vector, so only the FP64 matrix has to be split.  With A = R'R (symmetric, FP64):
//     x'Ax = sum_j x_j * sum_{i<=j} c_ij A_ji x_i ,   c_ij = 2 (i<j), 1 (i==j)
// B[j][i] = c_ij A_ji * 2^-E (|B| < 1/2) is cut into S signed base-128 digits b_k in [-64,64]:
//     B = sum_k 128^-(k+1) b_k  (+ error <= 0.5 * 128^-S),
// every partial product x.b_k is an exact int32 (tcgen05.mma kind::i8), and the epilogue folds
//     q_s += w_k * sum_j acc[s][j] * x[s][j],   w_k = 2^E 128^-(k+1)
// in FP64.  Because B is lower triangular, N tile jb only needs K in [0, 256(jb+1)): half the MACs of
// the full rotation.  The same epilogue accumulates x~.y~ = x.(R'y~) and, once a 128-SNP row block has
// seen every tile, evaluates RSS / F / p (linear_models.py:1345-1349).
#pragma once
#include "fdist.cuh"
#include "tc_gemm.cuh"

namespace mmg {

constexpr int QS_MAX_SLICES = 10;

struct QuadEpi {
    struct Params {
        const int8_t* snps;
        int64_t pitch;
        int64_t row_begin, row_count;
        double w[QS_MAX_SLICES];     // slice weights
        const double* v;             // [n_padN] R'y~ (zero padded)
        double h0_rss, n_p, lbeta;
        double *xx, *xy, *rss, *f, *p, *var_perc;
    };
    double q, xy;
    const int8_t* xrow;

    __device__ __forceinline__ void begin_group(const Params& p, int g, int row) {
        q = 0.0;
        xy = 0.0;
        const int64_t r = (int64_t)g * TC_BM + row;
        xrow = (r < p.row_count) ? p.snps + (p.row_begin + r) * p.pitch : nullptr;
    }
    __device__ __forceinline__ void chunk(const Params& p, const TcTile& t, int row, int c, const uint32_t (&v)[32]) {
        if (xrow == nullptr) return;
        const int col0 = t.col0 + c * 32;
        const uint4* xp = reinterpret_cast<const uint4*>(xrow + col0);
        const uint4 x0 = xp[0], x1 = xp[1];
        const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        int s = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int xv = (int)(int8_t)((xw[j >> 2] >> (8 * (j & 3))) & 0xffu);
            s += (int)v[j] * xv;
        }
        q = fma(p.w[t.aux0], (double)s, q);
        if (t.aux1 & 1) {
            const double* vv = p.v + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int xv = (int)(int8_t)((xw[j >> 2] >> (8 * (j & 3))) & 0xffu);
                xy = fma((double)xv, vv[j], xy);
            }
        }
    }
    __device__ __forceinline__ void end_group(const Params& p, int g, int row) {
        if (xrow == nullptr) return;
        const int64_t o = (int64_t)g * TC_BM + row;
        const double sxx = q, sxy = xy;
        if (p.xx) p.xx[o] = sxx;
        if (p.xy) p.xy[o] = sxy;
        double rss = p.h0_rss, f = 0.0, vp = 0.0, pv = 1.0;
        if (sxx > 0.0) {
            const double r2 = (sxy * sxy) / (sxx * p.h0_rss);
            const double rs = p.h0_rss - (sxy * sxy) / sxx;
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

// max |c_ij A[j][i]| over the row-major lower triangle (i <= j) -> bits of a non-negative double
__global__ void quad_amax_kernel(const double* __restrict__ A, int64_t ld, int n, unsigned long long* __restrict__ amax_bits) {
    const int j = blockIdx.y;
    double m = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= j; i += gridDim.x * blockDim.x) {
        const double a = fabs(A[(int64_t)j * ld + i]) * (i < j ? 2.0 : 1.0);
        m = fmax(m, a);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(amax_bits, (unsigned long long)__double_as_longlong(m));
}

// digits of B[j][i] = c_ij A[j][i] 2^-E into S stacked int8 planes Bq[(k * n_padN + j) * ldq + i]
__global__ void quad_slice_kernel(const double* __restrict__ A, int64_t ld, int n, double scale /* 2^-E */, int S,
                                  int8_t* __restrict__ Bq, int64_t n_padN, int64_t ldq) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i > j || i >= n) return;
    double r = A[(int64_t)j * ld + i] * (i < j ? 2.0 : 1.0) * scale;     // |r| < 0.5, exact scaling
    for (int k = 0; k < S; ++k) {
        r *= 128.0;
        const double d = rint(r);                                        // in [-64, 64]
        r -= d;                                                          // exact: |r| <= 0.5
        Bq[((int64_t)k * n_padN + j) * ldq + i] = (int8_t)(int)d;
    }
}

}  // namespace mmg
