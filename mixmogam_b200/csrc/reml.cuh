// Stage 2: REML delta grid + refinement on the device (linear_models.py:789-891).
//   reml_grid_kernel   : one block per (grid point, phenotype): the four p-long column sums
//                        s1..s4 of :803-809 by warp-shuffle reduction -> lls, dlls (:807,:810)
//   reml_refine_kernel : one block per phenotype: bracket search + secant + validation (reml_logic.cuh)
#pragma once
#include "reml_logic.cuh"

namespace mmg {

constexpr int REML_THREADS = 256;

template <int NV>
__device__ __forceinline__ void block_reduce_sum(double (&v)[NV], double* red /* [NV][8] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
    __syncthreads();          // protect red from the previous use
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[k * 8 + warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < REML_THREADS / 32; ++w) s += red[k * 8 + w];   // fixed order: deterministic
        v[k] = s;
    }
}

static __global__ void __launch_bounds__(REML_THREADS)
reml_grid_kernel(const double* __restrict__ eig, const double* __restrict__ sq_etas, int p,
                 const double* __restrict__ deltas, int g, double* __restrict__ lls, double* __restrict__ dlls) {
    __shared__ double red[4 * 8];
    const int gi = blockIdx.x, t = blockIdx.y;
    const double d = deltas[gi];
    const double* sq = sq_etas + (int64_t)t * p;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < p; i += REML_THREADS) {
        const double lam = eig[i] + d;
        const double e = sq[i];
        s[0] += e / lam;              // s1 (:803)
        s[1] += log(lam);             // s2 (:806)
        s[2] += e / (lam * lam);      // s3 (:808)
        s[3] += 1.0 / lam;            // s4 (:809)
    }
    block_reduce_sum<4>(s, red);
    if (threadIdx.x == 0) {
        const double pd = (double)p;
        lls[(int64_t)t * g + gi] = 0.5 * (pd * (log(pd / (2.0 * M_PI)) - 1.0 - log(s[0])) - s[1]);
        dlls[(int64_t)t * g + gi] = 0.5 * (pd * s[2] / s[0] - s[3]);
    }
}

struct RemlDevEval {
    const double* eig;
    const double* sq;
    int p;
    double* red;
    __device__ double redll(double delta) {
        double s[3] = {0.0, 0.0, 0.0};
        for (int i = threadIdx.x; i < p; i += REML_THREADS) {
            const double v1 = eig[i] + delta;
            const double v2 = sq[i] / v1;
            s[0] += v2 / v1;
            s[1] += v2;
            s[2] += 1.0 / v1;
        }
        block_reduce_sum<3>(s, red);
        return (double)p * s[0] / s[1] - s[2];
    }
    __device__ double rell(double delta) {
        double s[2] = {0.0, 0.0};
        for (int i = threadIdx.x; i < p; i += REML_THREADS) {
            const double v = eig[i] + delta;
            s[0] += sq[i] / v;
            s[1] += log(v);
        }
        block_reduce_sum<2>(s, red);
        const double pd = (double)p;
        const double c1 = 0.5 * pd * (log(pd / (2.0 * M_PI)) - 1.0);
        return c1 - 0.5 * (pd * log(s[0]) + s[1]);
    }
};

static __global__ void __launch_bounds__(REML_THREADS)
reml_refine_kernel(const double* __restrict__ eig, const double* __restrict__ sq_etas, int p,
                   const double* __restrict__ deltas, int g, double esp, const double* __restrict__ lls,
                   const double* __restrict__ dlls, double* __restrict__ opt_delta, double* __restrict__ opt_ll,
                   int* __restrict__ flags) {
    __shared__ double red[4 * 8];
    const int t = blockIdx.x;
    RemlDevEval ev{eig, sq_etas + (int64_t)t * p, p, red};
    double od, ol;
    int fl;
    reml_refine(ev, lls + (int64_t)t * g, dlls + (int64_t)t * g, deltas, g, esp, &od, &ol, &fl);
    if (threadIdx.x == 0) {
        opt_delta[t] = od;
        opt_ll[t] = ol;
        flags[t] = fl;
    }
}

}  // namespace mmg
