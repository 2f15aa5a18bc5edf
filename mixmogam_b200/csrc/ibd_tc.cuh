// IBD kinship (kinship.py:59-75, hdf5_data.py:30-62,84-115,205-237) on the int8 tensor cores.
//
//   K m = sum_s z_s z_s',  z_s = (x_s - mu_s)/sigma_s  (ddof 0)
//       = sum_s w_s x_s x_s'  -  u 1'  -  1 u'  +  c 11',     w_s = 1/sigma_s^2 = n^2 / (n sum x^2 - (sum x)^2),
//                                                             u = sum_s w_s mu_s x_s,   c = sum_s w_s mu_s^2.
// The genotype is an exact small integer, so only the per-SNP weight has to be split: w_s 2^-E (< 1/2) is cut into
// S signed base-64 digits d_sk in [-32,32]; operand A_k[i][s] = d_sk x_is (|.| <= 64) against operand B[j][s] = x_js
// gives the exact int32 Gram of digit plane k, and the epilogue folds G += 2^E 64^-(k+1) acc_k in FP64.
// u and c are evaluated with the SAME truncated weights w^_s = 2^E sum_k d_sk 64^-(k+1), so the result is the exact
// centred form sum_s w^_s (x_i - mu)(x_j - mu) of weights that differ from w_s by <= 2^E 64^-S / 2: the truncation is
// never amplified by the cancellation of the mean.  Every digit-plane Gram is symmetric, so only the upper-triangular
// tiles are computed.
#pragma once
#include "kinship_kernels.cuh"
#include "tc_gemm.cuh"

namespace mmg {

constexpr int IBD_MAX_SLICES = 8;

// per selected SNP: w = n^2 / (n q - s^2), mean = s/n; flags a monomorphic SNP (kinship.py:67); max w -> amax_bits
static __global__ void ibd_weights_kernel(const long long* __restrict__ sums, const long long* __restrict__ sumsq,
                                   const long long* __restrict__ rows, int64_t count, int n, double* __restrict__ w,
                                   double* __restrict__ mean, unsigned long long* __restrict__ amax_bits, int* __restrict__ bad_flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double wi = 0.0;
    if (i < count) {
        const long long r = rows[i];
        const long long s = sums[r], q = sumsq[r];
        const long long D = (long long)n * q - s * s;             // n^2 var, exact
        if (D <= 0) {
            atomicOr(bad_flag, 1);
        } else {
            wi = ((double)n * (double)n) / (double)D;
        }
        w[i] = wi;
        mean[i] = (double)s / (double)n;
    }
    for (int o = 16; o > 0; o >>= 1) wi = fmax(wi, __shfl_xor_sync(0xffffffffu, wi, o));
    if ((threadIdx.x & 31) == 0 && wi > 0.0) atomicMax(amax_bits, (unsigned long long)__double_as_longlong(wi));
}

// digits[k][i] of w_i 2^-E (base 64), coef[i] = w^_i mean_i, acc[0] += sum w^_i mean_i^2
static __global__ void __launch_bounds__(256) ibd_digits_kernel(const double* __restrict__ w, const double* __restrict__ mean, int64_t count,
                                                         double scale /* 2^-E */, double inv_scale, int S, int8_t* __restrict__ digits,
                                                         int64_t dig_pitch, double* __restrict__ coef, double* __restrict__ acc) {
    __shared__ double red[8];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double c = 0.0;
    if (i < count) {
        double r = w[i] * scale;                                  // < 1/2
        double what = 0.0, pw = 1.0;
        for (int k = 0; k < S; ++k) {
            r *= 64.0;
            const double d = rint(r);                             // [-32, 32]
            r -= d;
            pw *= 1.0 / 64.0;
            what += d * pw;                                       // exact (<= 48 significant bits)
            digits[(int64_t)k * dig_pitch + i] = (int8_t)(int)d;
        }
        what *= inv_scale;
        const double mu = mean[i];
        coef[i] = what * mu;
        c = what * mu * mu;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += red[k];
        atomicAdd(acc, t);
    }
}

// pack: selected SNP rows -> K-major operand planes P[(plane * n_padM + i) * p_pitch + s]
//   plane 0     : x_is                     (operand B)
//   plane 1 + k : d_sk x_is, k < S         (operand A of digit plane k)
// Same 128 SNP x 64 individual transpose as pack_kmajor_kernel.  Genotypes outside {0,1,2} raise *bad_flag.
static __global__ void __launch_bounds__(256) pack_ibd_kernel(const int8_t* __restrict__ snps, int64_t pitch,
                                                       const long long* __restrict__ rows, int64_t s_count, int n,
                                                       const int8_t* __restrict__ digits, int64_t dig_pitch, int S,
                                                       int8_t* __restrict__ P, int64_t p_pitch, int64_t n_padM,
                                                       int* __restrict__ bad_flag) {
    __shared__ __align__(16) uint32_t tile[64][32];
    const int t = threadIdx.x;
    const int64_t s0 = (int64_t)blockIdx.x * 128;
    const int i0 = blockIdx.y * 64;
    const int i4 = t & 15, sq = t >> 4;
    bool bad = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int quad = sq + 16 * h;
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t s = s0 + 4 * quad + j;
            r[j] = (s < s_count) ? *reinterpret_cast<const uint32_t*>(snps + rows[s] * pitch + i0 + 4 * i4) : 0u;
            bad |= (__vcmpgtu4(r[j], 0x02020202u) != 0u);
        }
        const uint32_t t0 = __byte_perm(r[0], r[1], 0x5140), t1 = __byte_perm(r[2], r[3], 0x5140);
        const uint32_t t2 = __byte_perm(r[0], r[1], 0x7362), t3 = __byte_perm(r[2], r[3], 0x7362);
        uint32_t w[4];
        w[0] = __byte_perm(t0, t1, 0x5410);
        w[1] = __byte_perm(t0, t1, 0x7632);
        w[2] = __byte_perm(t2, t3, 0x5410);
        w[3] = __byte_perm(t2, t3, 0x7632);
        const int pchunk = (quad >> 2) ^ (i4 & 7);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) tile[4 * i4 + ii][pchunk * 4 + (quad & 3)] = w[ii];
    }
    if (bad) atomicOr(bad_flag, 2);
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int u = t + 256 * h;
        const int row = u >> 3, c = u & 7;             // 16 SNPs s0 + 16c ..
        const int i = i0 + row;
        if (i >= n) continue;
        const int pchunk = c ^ ((row >> 2) & 7);
        const uint4 x = *reinterpret_cast<const uint4*>(&tile[row][pchunk * 4]);
        const int64_t so = s0 + 16 * c;
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
        *reinterpret_cast<uint4*>(P + (int64_t)i * p_pitch + so) = x;            // rows beyond s_count were read as 0
        uint32_t m1[4], m2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            m1[k] = __vcmpeq4(xs[k], 0x01010101u);
            m2[k] = __vcmpeq4(xs[k], 0x02020202u);
        }
        for (int k = 0; k < S; ++k) {
            const uint4 dg = *reinterpret_cast<const uint4*>(digits + (int64_t)k * dig_pitch + so);   // dig_pitch padded, zero filled
            const uint32_t ds[4] = {dg.x, dg.y, dg.z, dg.w};
            uint32_t o[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) o[b] = (ds[b] & m1[b]) | (__vadd4(ds[b], ds[b]) & m2[b]);
            *reinterpret_cast<uint4*>(P + ((int64_t)(1 + k) * n_padM + i) * p_pitch + so) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// u[i] += sum over selected SNPs of coef_s x_is.  Block = 256 threads x 4 individuals, slab of `slab` SNPs.
static __global__ void __launch_bounds__(256) snp_weighted_colsum_kernel(const int8_t* __restrict__ snps, int64_t pitch,
                                                                  const long long* __restrict__ rows, const double* __restrict__ coef,
                                                                  int64_t count, int slab, int n, double* __restrict__ u) {
    const int i0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (i0 >= n) return;                                         // pitch is padded to 256: a 4-byte load at i0 < n stays in the row
    const int64_t sb = (int64_t)blockIdx.y * slab;
    const int64_t se = sb + slab < count ? sb + slab : count;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int64_t s = sb; s < se; ++s) {
        const uint32_t x = *reinterpret_cast<const uint32_t*>(snps + rows[s] * pitch + i0);
        const double cf = coef[s];
        a0 = fma(cf, (double)(int8_t)(x & 0xff), a0);
        a1 = fma(cf, (double)(int8_t)((x >> 8) & 0xff), a1);
        a2 = fma(cf, (double)(int8_t)((x >> 16) & 0xff), a2);
        a3 = fma(cf, (double)(int8_t)(x >> 24), a3);
    }
    atomicAdd(u + i0, a0);
    if (i0 + 1 < n) atomicAdd(u + i0 + 1, a1);
    if (i0 + 2 < n) atomicAdd(u + i0 + 2, a2);
    if (i0 + 3 < n) atomicAdd(u + i0 + 3, a3);
}

// epilogue: Gw[out_row][n0 + 32c ..] += w[k] * acc.  The S digit-plane tiles of one output tile are consecutive tiles
// of ONE group (same CTA, same epilogue thread per row), so the read-modify-write needs no atomics.
struct IbdEpi {
    struct Params {
        double* Gw;            // [g_pad x ld] FP64, g_pad multiple of 256
        int64_t ld;
        int64_t n_padM;
        double w[IBD_MAX_SLICES];
    };
    __device__ __forceinline__ void begin_group(const Params&, int, int) {}
    __device__ __forceinline__ void end_group(const Params&, int, int) {}
    __device__ __forceinline__ int tile_begin(const Params&, const TcTile&, int) { return TC_BN / 32; }
    __device__ __forceinline__ void tile_end(const Params&, const TcTile&, int, int) {}
    __device__ __forceinline__ void chunk(const Params& p, const TcTile& t, int row, int c, const uint32_t (&v)[32]) {
        const int64_t orow = (int64_t)t.m0 - (int64_t)(t.aux0 + 1) * p.n_padM + row;
        double2* dst = reinterpret_cast<double2*>(p.Gw + orow * p.ld + t.n0 + c * 32);
        const double wk = p.w[t.aux0];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            double2 o = dst[j];
            o.x = fma(wk, (double)(int)v[2 * j + 0], o.x);
            o.y = fma(wk, (double)(int)v[2 * j + 1], o.y);
            dst[j] = o;
        }
    }
};

// K[i][j] += Gw[min][max] - u_i - u_j + c
static __global__ void ibd_finalize_add_kernel(const double* __restrict__ Gw, int64_t ldg, int n, const double* __restrict__ u,
                                        const double* __restrict__ c, double* __restrict__ K, int64_t ldk) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n) return;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    K[(int64_t)i * ldk + j] += Gw[(int64_t)lo * ldg + hi] - u[i] - u[j] + c[0];
}

}  // namespace mmg
