// Survival function of the F distribution, FP64:  scipy.stats.f.sf(f, dfn, dfd)
// (reference call sites linear_models.py:1349, :1172, :925).
//
//   sf = I_x(dfd/2, dfn/2),  x = dfd / (dfd + dfn f)         (regularised incomplete beta)
//
// evaluated with the modified-Lentz continued fraction; ln x and ln(1-x) come from log1p so the far
// tail (p ~ 1e-300) keeps full relative accuracy.  ln B(a,b) is computed once on the host in long
// double and passed in, so the per-SNP work is one continued fraction.
// Host+device: the CPU test-suite compiles this header with g++ (tests/host_check.cpp).
#pragma once
#include <math.h>

#ifndef MMG_HD
#ifdef __CUDACC__
#define MMG_HD __host__ __device__ __forceinline__
#else
#define MMG_HD inline
#endif
#endif

namespace mmg {

MMG_HD double betacf(double a, double b, double x) {
    const double FPMIN = 1e-300;
    const double EPS = 1e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = 1.0 - qab * x / qap;
    if (fabs(d) < FPMIN) d = FPMIN;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 20000; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c;
        if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c;
        if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) <= EPS) break;
    }
    return h;
}

// 2F1(a + b, 1; b + 1; z) = sum_k t_k,  t_0 = 1,  t_{k+1} = t_k (a + b + k) z / (b + 1 + k)  -- the same quantity betacf(b, a, z)
// evaluates (I_z(b, a) = z^b (1 - z)^a / (b B(a, b)) times it), as its power series.  All terms are positive (no cancellation).
// For q = (a + b) z <= 5 and z <= 0.01 the terms fall below 1e-17 of the sum within K = 20 + 3.5 q of them (t_{k+1} / t_k <=
// q / (k + 3/2) + z); the sum is
// evaluated as ONE fraction A / B by the nested recurrence  R_k = 1 + (u_k / v_k) R_{k+1} = (v_k B_{k+1} + u_k A_{k+1}) / (v_k B_{k+1}),
// u_k = (a + b + k) z, v_k = b + 1 + k: two FMAs and a multiplication per term and a single division at the end, where the
// continued fraction spends four divisions per step.  This is the branch nearly every SNP of a scan takes (F < ~8 at n ~ 10^4:
// the phenotype-batched scan evaluates T x m of them), the continued fraction keeps the tails.
MMG_HD double beta_series_small(double a, double b, double z, double q) {
    const int K = 20 + (int)(3.5 * q);
    double A = 1.0, B = 1.0;
    for (int k = K - 1; k >= 0; --k) {
        const double u = (a + b + k) * z, v = b + 1.0 + k;
        const double vb = v * B;
        A = fma(u, A, vb);
        B = vb;
    }
    return A / B;
}

// lbeta = ln B(dfd/2, dfn/2)
MMG_HD double f_sf(double f, double dfn, double dfd, double lbeta) {
    if (f != f) return f;                 // NaN
    if (!(f > 0.0)) return 1.0;
    if (isinf(f)) return 0.0;
    const double a = 0.5 * dfd, b = 0.5 * dfn;
    const double t = dfn * f / dfd;       // (1-x)/x
    const double lx = -log1p(t);          // ln x
    const double l1x = -log1p(1.0 / t);   // ln (1-x)
    const double x = 1.0 / (1.0 + t);
    const double lbt = a * lx + b * l1x - lbeta;
    const double omx = t / (1.0 + t);
    const double q = (a + b) * omx;
    // The series of the complement first, wherever it applies -- also beyond the usual switch point x = (a+1)/(a+b+2), which at
    // dfd ~ 10^4 sends every F > 3 (8 % of the tests, i.e. a lane of nearly every warp) into ~100 iterations of the continued
    // fraction at four divisions each: ~20 k instructions for the whole warp against ~600.  q <= 5 is F <= 10 at dfn = 1 (0.16 % of
    // the tests beyond it): the complement is then still >= 1.6e-3, so 1 - (.) costs under three of the subtrahend's ~14 good
    // digits (measured <= 7e-12 relative against scipy up to dfd = 5e4; the tests hold 1e-11 there and 1e-10 in the tails).
    if (q <= 5.0 && omx <= 0.01) return 1.0 - exp(lbt) * beta_series_small(a, b, omx, q) / b;
    if (x < (a + 1.0) / (a + b + 2.0)) return exp(lbt) * betacf(a, b, x) / a;
    return 1.0 - exp(lbt) * betacf(b, a, omx) / b;
}

}  // namespace mmg
