// Control logic of LinearMixedModel.get_estimates, REML branch (linear_models.py:826-891), written
// once for host and device.  `Eval` supplies the two p-long reductions
//     redll(delta) = p * sum(v2/v1)/sum(v2) - sum(1/v1)          (_redll_, :626-631)
//     rell(delta)  = c1 - 0.5*(p*log(sum(sq/v)) + sum(log v))    (_rell_,  :618-623)
// On the device every thread of the block runs this logic redundantly and the reductions broadcast
// their result, so the control flow stays uniform; on the host (tests/host_check.cpp) Eval is serial.
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

enum { REML_FLAG_INTERVAL = 1, REML_FLAG_CONVERGED = 2, REML_FLAG_ACCEPTED = 4 };

// scipy.optimize.newton(func, x0, tol=tol, maxiter=maxiter) with fprime=None (secant), disp=True:
// returns true and the root on convergence; false where scipy raises RuntimeError (the reference
// catches it and falls back to the bracket midpoint, linear_models.py:852-858).
template <class Eval>
MMG_HD bool secant_newton(Eval& ev, double x0, double tol, int maxiter, double* root) {
    double p0 = 1.0 * x0;
    const double eps = 1e-4;
    double p1 = x0 * (1 + eps);
    p1 += (p1 >= 0 ? eps : -eps);
    double q0 = ev.redll(p0);
    double q1 = ev.redll(p1);
    if (fabs(q1) < fabs(q0)) {
        double t = p0; p0 = p1; p1 = t;
        t = q0; q0 = q1; q1 = t;
    }
    double p = p1;
    for (int itr = 0; itr < maxiter; ++itr) {
        if (q1 == q0) {
            if (p1 != p0) return false;           // "Tolerance of ... reached. Failed to converge" -> RuntimeError
            *root = (p1 + p0) / 2.0;
            return true;
        } else {
            if (fabs(q1) > fabs(q0))
                p = (-q0 / q1 * p1 + p0) / (1 - q0 / q1);
            else
                p = (-q1 / q0 * p0 + p1) / (1 - q1 / q0);
        }
        // np.isclose(p, p1, rtol=0, atol=tol): finite and |p - p1| <= tol (inf == inf also counts)
        if ((isfinite(p) && isfinite(p1) && fabs(p - p1) <= tol) || (isinf(p) && p == p1)) {
            *root = p;
            return true;
        }
        p0 = p1;
        q0 = q1;
        p1 = p;
        q1 = ev.redll(p1);
    }
    return false;                                  // "Failed to converge after %d iterations" -> RuntimeError
}

template <class Eval>
MMG_HD void reml_refine(Eval& ev, const double* lls, const double* dlls, const double* deltas, int g, double esp,
                        double* out_delta, double* out_ll, int* out_flags) {
    // argmax (first maximum, numpy semantics; NaN propagates like np.argmax: first NaN wins)
    int max_i = 0;
    for (int i = 1; i < g; ++i) {
        if (lls[max_i] != lls[max_i]) break;
        if (lls[i] > lls[max_i] || lls[i] != lls[i]) max_i = i;
    }
    const double max_ll = lls[max_i];

    // sign-change intervals (:829-836); max() over tuples (mid_ll, i)
    int opt_i = -1;
    double best_mid = 0.0;
    double last_dll = dlls[0], last_ll = lls[0];
    for (int i = 1; i < g; ++i) {
        if (dlls[i] < 0 && last_dll > 0) {
            const double mid = (lls[i] + last_ll) * 0.5;
            if (opt_i < 0 || mid > best_mid || (mid == best_mid && i > opt_i)) {
                best_mid = mid;
                opt_i = i;
            }
        }
        last_ll = lls[i];
        last_dll = dlls[i];
    }

    int flags = 0;
    double opt_delta, opt_ll;
    if (opt_i >= 0) {
        flags |= REML_FLAG_INTERVAL;
        opt_delta = 0.5 * (deltas[opt_i - 1] + deltas[opt_i]);
        double new_opt_delta = opt_delta;
        double root;
        if (secant_newton(ev, opt_delta, esp, 100, &root)) {
            new_opt_delta = root;
            flags |= REML_FLAG_CONVERGED;
        }
        if (opt_i > 1 && deltas[opt_i - 1] - esp < new_opt_delta && new_opt_delta < deltas[opt_i] + esp) {
            opt_delta = new_opt_delta;
            flags |= REML_FLAG_ACCEPTED;
        } else if (opt_i == 1 && 0.0 < new_opt_delta && new_opt_delta < deltas[opt_i] + esp) {
            opt_delta = new_opt_delta;
            flags |= REML_FLAG_ACCEPTED;
        } else if (opt_i == g - 1 && new_opt_delta > deltas[opt_i - 1] - esp && !isinf(new_opt_delta)) {
            opt_delta = new_opt_delta;
            flags |= REML_FLAG_ACCEPTED;
        }
        opt_ll = ev.rell(opt_delta);               // :881-882
        if (opt_ll < max_ll) opt_delta = deltas[max_i];     // :886-887 (opt_ll is NOT updated by the reference)
    } else {
        opt_delta = deltas[max_i];
        opt_ll = max_ll;
    }
    *out_delta = opt_delta;
    *out_ll = opt_ll;
    *out_flags = flags;
}

}  // namespace mmg
