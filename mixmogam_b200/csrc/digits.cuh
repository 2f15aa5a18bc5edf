// Signed base-256 digit expansion used by the int8 tensor-core scan (scan_tc.cuh / scan_quad.cuh), written once for
// host and device (tests/host_check.cpp checks it on the CPU).
//
// A scaled FP64 value r, |r| <= DIGIT256_RMAX = 0.498, is rounded to S planes, N = rint(r 256^S) (exact: a power-of-two
// scaling and one rounding to an integer < 2^55, S <= 7), and N is written in the digit set {-128..127} -- every value
// of an int8, 8 bits per plane -- from the bottom up, exactly, in integer arithmetic:
//     N = sum_{k<S} d_k 256^(S-1-k),     d_k = ((N + 128) mod 256) - 128,  N <- (N - d_k) / 256.
// {-128..127} is a complete residue system mod 256, so the expansion exists and is unique whenever
// -128 (256^S - 1)/255 <= N <= 127 (256^S - 1)/255, which |r| <= 0.498 < 127/255 guarantees (top digit in range, no carry
// out).  Using only the first S' <= S planes leaves
//     |r - sum_{k<S'} d_k 256^-(k+1)| <= (128/255) 256^-S'                                         (DIGIT256_REM)
// (the dropped digits sum to at most (128/255)(256^-S' - 256^-S) and the rounding adds <= 256^-S / 2).
// 8 bits per int8 plane instead of the 7 of a symmetric base-128 digit in [-64, 64]: one plane fewer for the same
// bound on the bench data (5 -> 4).
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

constexpr int DIGIT256_MAX_PLANES = 7;               // 56 bits: more than any double with |r| < 1/2 carries
constexpr double DIGIT256_RMAX = 0.498;              // largest scaled magnitude accepted
constexpr double DIGIT256_REM = 128.0 / 255.0;       // truncation bound per unit 256^-S'

// digits[0..S) of r (most significant first); returns the carry out of the top digit (0 for |r| <= DIGIT256_RMAX)
MMG_HD long long digit256_split(double r, int S, int* digits) {
    long long N = (long long)rint(ldexp(r, 8 * S));
    for (int k = S - 1; k >= 0; --k) {
        const int d = (int)((N + 128) & 255) - 128;
        digits[k] = d;
        N = (N - d) >> 8;                             // exact: N - d is a multiple of 256
    }
    return N;
}

// binary exponent E with amax 2^-E <= DIGIT256_RMAX (amax > 0 finite); 0 for amax == 0
MMG_HD int digit256_exponent(double amax) {
    if (!(amax > 0.0)) return 0;
    int E = ilogb(amax) + 2;                          // amax 2^-E in [1/4, 1/2)
    if (ldexp(amax, -E) > DIGIT256_RMAX) ++E;         // the sliver (0.498, 1/2): one more bit of head-room
    return E;
}

}  // namespace mmg
