// Host build of the product's host/device headers (fdist.cuh, reml_logic.cuh) so the CPU test-suite can
// check them against the golden vectors without a GPU.  Reads binary doubles from stdin:
//   mode "fsf":  count, then count x (f, dfn, dfd)                    -> count x sf
//   mode "reml": p, g, esp, eig_vals[p], sq_etas[p], deltas[g]         -> delta, ll, flags, lls[g], dlls[g]
//   mode "digits": count, S, then count x amax-scaled value a          -> per value: E, S digits, carry out  (digits.cuh)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "digits.cuh"
#include "fdist.cuh"
#include "reml_logic.cuh"

static double rd() {
    double v = 0;
    if (fread(&v, sizeof(double), 1, stdin) != 1) return NAN;
    return v;
}
static void wr(double v) { fwrite(&v, sizeof(double), 1, stdout); }

struct HostEval {
    const double* eig;
    const double* sq;
    int p;
    double redll(double delta) {
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < p; ++i) {
            const double v1 = eig[i] + delta, v2 = sq[i] / v1;
            a += v2 / v1;
            b += v2;
            c += 1.0 / v1;
        }
        return (double)p * a / b - c;
    }
    double rell(double delta) {
        double a = 0, b = 0;
        for (int i = 0; i < p; ++i) {
            const double v = eig[i] + delta;
            a += sq[i] / v;
            b += log(v);
        }
        const double pd = (double)p, c1 = 0.5 * pd * (log(pd / (2.0 * M_PI)) - 1.0);
        return c1 - 0.5 * (pd * log(a) + b);
    }
};

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    if (!strcmp(argv[1], "fsf")) {
        const long n = (long)rd();
        for (long i = 0; i < n; ++i) {
            const double f = rd(), dfn = rd(), dfd = rd();
            const double lb = (double)(lgammal(0.5L * dfd) + lgammal(0.5L * dfn) - lgammal(0.5L * (dfd + dfn)));
            wr(mmg::f_sf(f, dfn, dfd, lb));
        }
        return 0;
    }
    if (!strcmp(argv[1], "digits")) {
        // every value is treated as its own amax: E from digit256_exponent, r = a 2^-E, then S digits
        const long n = (long)rd();
        const int S = (int)rd();
        for (long i = 0; i < n; ++i) {
            const double a = rd();
            const int E = mmg::digit256_exponent(fabs(a));
            const double r = ldexp(a, -E);
            int d[mmg::DIGIT256_MAX_PLANES];
            const long long carry = mmg::digit256_split(r, S, d);
            wr((double)E);
            for (int k = 0; k < S; ++k) wr((double)d[k]);
            wr((double)carry);
        }
        return 0;
    }
    if (!strcmp(argv[1], "reml")) {
        const int p = (int)rd(), g = (int)rd();
        const double esp = rd();
        std::vector<double> eig(p), sq(p), del(g), lls(g), dlls(g);
        for (auto& v : eig) v = rd();
        for (auto& v : sq) v = rd();
        for (auto& v : del) v = rd();
        for (int k = 0; k < g; ++k) {
            double s1 = 0, s2 = 0, s3 = 0, s4 = 0;
            for (int i = 0; i < p; ++i) {
                const double lam = eig[i] + del[k];
                s1 += sq[i] / lam;
                s2 += log(lam);
                s3 += sq[i] / (lam * lam);
                s4 += 1.0 / lam;
            }
            const double pd = (double)p;
            lls[k] = 0.5 * (pd * (log(pd / (2.0 * M_PI)) - 1.0 - log(s1)) - s2);
            dlls[k] = 0.5 * (pd * s3 / s1 - s4);
        }
        HostEval ev{eig.data(), sq.data(), p};
        double od, ol;
        int fl;
        mmg::reml_refine(ev, lls.data(), dlls.data(), del.data(), g, esp, &od, &ol, &fl);
        wr(od);
        wr(ol);
        wr((double)fl);
        for (double v : lls) wr(v);
        for (double v : dlls) wr(v);
        return 0;
    }
    return 2;
}
