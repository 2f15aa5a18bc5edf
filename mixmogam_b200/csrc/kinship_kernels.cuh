// Stage 1 support kernels: genotype packing for the tensor-core Gram, a SIMT (dp4a) Gram used as the
// independent cross-check of the tcgen05 kernel, the FP64 finalisation of kinship.py:50-55, scale_k
// (kinship.py:94-100) and the per-SNP standardisation of the IBD kinship (kinship.py:66).
#pragma once
#include <cstdint>

namespace mmg {

// ---------------------------------------------------------------------------------------------------
// pack: SNP-major genotypes [m x pitch] int8  ->  individual-major, K-major operand P [n x p_pitch] int8
//   binary : P[i][s]              = 2 x - 1                       (kinship.py:43)
//   diploid: P[i][32 t + j]       = [x >= 1]  (j < 16)            thermometer planes; any fixed
//            P[i][32 t + 16 + j]  = [x >= 2]                       permutation of K is a valid Gram operand
//            for SNP s = 16 t + j
// Tile = 128 SNPs x 64 individuals; 4x4 byte blocks are transposed in registers (PRMT), staged through a
// 16-byte-chunk XOR-swizzled shared tile, and written back as coalesced 16-byte vectors.
// SNPs beyond s_count pack to 0 bytes.  Values outside the coding's domain raise *bad_flag.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bytes_lt_mask(int first_idx, int nvalid) {
    // 0xff in byte b iff first_idx + b < nvalid
    uint32_t m = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (first_idx + b < nvalid) m |= 0xffu << (8 * b);
    return m;
}

// 4 bytes (0 / 1 each, or e2m1 codes in the low nibble) -> 4 nibbles in the low 16 bits, byte b in bits 4b..4b+3
__device__ __forceinline__ uint32_t squeeze_nibbles(uint32_t w) {
    const uint32_t y = (w | (w >> 4)) & 0x00ff00ffu;
    return (y | (y >> 8)) & 0xffffu;
}

// FP4 = true: the operand of the kind::mxf4 Gram, e2m1 codes packed two per byte (p_pitch in bytes as before):
//   binary : code(2 x - 1) = 0x2 (+1.0) / 0xa (-1.0), SNP s0 + j of a 16-SNP group in nibble j
//   diploid: [x >= 1] of the 16 SNPs in the first 8 bytes of the group's 16, [x >= 2] in the second 8; code 0x2 = 1.0
template <int CODING, bool FP4 = false>
static __global__ void __launch_bounds__(256) pack_kmajor_kernel(const int8_t* __restrict__ snps, int64_t pitch,
                                                          int64_t s_begin, int64_t s_count, int n,
                                                          int8_t* __restrict__ P, int64_t p_pitch,
                                                          int* __restrict__ bad_flag) {
    __shared__ __align__(16) uint32_t tile[64][32];    // [individual][128 SNP bytes as 32 words]
    const int t = threadIdx.x;
    const int64_t s0 = (int64_t)blockIdx.x * 128;
    const int i0 = blockIdx.y * 64;
    const int i4 = t & 15, sq = t >> 4;
    bool bad = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int quad = sq + 16 * h;                  // 4 consecutive SNPs
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t s = s0 + 4 * quad + j;
            r[j] = (s < s_count) ? *reinterpret_cast<const uint32_t*>(snps + (s_begin + s) * pitch + i0 + 4 * i4) : 0u;
            // a byte outside the coding's domain: any bit above bit 0 (binary); any bit above bit 1, or the value 3 (diploid)
            bad |= (CODING == 0 ? (r[j] & 0xfefefefeu) : ((r[j] & 0xfcfcfcfcu) | (r[j] & (r[j] >> 1) & 0x01010101u))) != 0u;
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
    if (bad) atomicOr(bad_flag, 1);
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int u = t + 256 * h;
        const int row = u >> 3, c = u & 7;             // 16 SNPs s0 + 16c ..
        const int i = i0 + row;
        if (i >= n) continue;
        const int pchunk = c ^ ((row >> 2) & 7);
        const uint4 x = *reinterpret_cast<const uint4*>(&tile[row][pchunk * 4]);
        int64_t rem = s_count - (s0 + 16 * c);
        const int nvalid = rem < 0 ? 0 : (rem > 16 ? 16 : (int)rem);
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
        if (FP4) {
            uint32_t a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t m = bytes_lt_mask(4 * k, nvalid);
                if (CODING == 0) {
                    // x = 1 -> 0x2, x = 0 -> 0xa
                    a[k] = squeeze_nibbles((0x0a0a0a0au ^ ((xs[k] & 0x01010101u) << 3)) & m);
                } else {
                    // bytes in {0, 1, 2} (anything else raised the flag above): [x >= 1] = bit0 | bit1, [x >= 2] = bit1; code 0x2 = 1.0
                    a[k] = squeeze_nibbles(((xs[k] | (xs[k] << 1)) & 0x02020202u & m));
                    b[k] = squeeze_nibbles((xs[k] & 0x02020202u & m));
                }
            }
            if (CODING == 0) {
                *reinterpret_cast<uint2*>(P + (int64_t)i * p_pitch + (s0 + 16 * c) / 2) = make_uint2(a[0] | (a[1] << 16), a[2] | (a[3] << 16));
            } else {
                *reinterpret_cast<uint4*>(P + (int64_t)i * p_pitch + (s0 + 16 * c)) =
                    make_uint4(a[0] | (a[1] << 16), a[2] | (a[3] << 16), b[0] | (b[1] << 16), b[2] | (b[3] << 16));
            }
        } else if (CODING == 0) {
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o[k] = __vsub4(__vadd4(xs[k], xs[k]), 0x01010101u) & bytes_lt_mask(4 * k, nvalid);
            *reinterpret_cast<uint4*>(P + (int64_t)i * p_pitch + s0 + 16 * c) = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
            uint32_t a[4], b[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t m = bytes_lt_mask(4 * k, nvalid);
                a[k] = __vcmpgeu4(xs[k], 0x01010101u) & 0x01010101u & m;
                b[k] = __vcmpgeu4(xs[k], 0x02020202u) & 0x01010101u & m;
            }
            int8_t* dst = P + (int64_t)i * p_pitch + 2 * (s0 + 16 * c);
            *reinterpret_cast<uint4*>(dst) = make_uint4(a[0], a[1], a[2], a[3]);
            *reinterpret_cast<uint4*>(dst + 16) = make_uint4(b[0], b[1], b[2], b[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// SIMT Gram on the packed operand: G[i][j] (+)= sum_k P[i][k] P[j][k] for tiles with bj >= bi, by dp4a.
// 64 x 64 outputs per block, 4 x 4 per thread.  Slow (CUDA-core) but independent of the tcgen05 path.
// ---------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) gram_simt_kernel(const int8_t* __restrict__ P, int64_t p_pitch, int n,
                                                        int64_t kbytes, int32_t* __restrict__ G, int64_t ldg,
                                                        int accumulate) {
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    __shared__ uint32_t As[64][17], Bs[64][17];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    int acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0;
    for (int64_t k0 = 0; k0 < kbytes; k0 += 64) {
        {
            const int row = t >> 2, q = t & 3;        // 64 rows x 4 chunks of 16 B
            uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
            if (k0 + 16 * q < kbytes) {               // p_pitch is padded to 128 and zero filled
                if (bi * 64 + row < n) va = *reinterpret_cast<const uint4*>(P + (int64_t)(bi * 64 + row) * p_pitch + k0 + 16 * q);
                if (bj * 64 + row < n) vb = *reinterpret_cast<const uint4*>(P + (int64_t)(bj * 64 + row) * p_pitch + k0 + 16 * q);
            }
            As[row][4 * q + 0] = va.x; As[row][4 * q + 1] = va.y; As[row][4 * q + 2] = va.z; As[row][4 * q + 3] = va.w;
            Bs[row][4 * q + 0] = vb.x; Bs[row][4 * q + 1] = vb.y; Bs[row][4 * q + 2] = vb.z; Bs[row][4 * q + 3] = vb.w;
        }
        __syncthreads();
#pragma unroll
        for (int kw = 0; kw < 16; ++kw) {
            int a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = (int)As[4 * ty + r][kw];
#pragma unroll
            for (int c = 0; c < 4; ++c) b[c] = (int)Bs[4 * tx + c][kw];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = __dp4a(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = bi * 64 + 4 * ty + r;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = bj * 64 + 4 * tx + c;
            int32_t* g = G + (int64_t)i * ldg + j;     // G is padded to a multiple of 256: in range
            *g = accumulate ? (*g + acc[r][c]) : acc[r][c];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// finalize (kinship.py:50-53): reads the upper triangle of the integer Gram
// ---------------------------------------------------------------------------------------------------
template <int CODING>
static __global__ void kinship_finalize_kernel(const int32_t* __restrict__ G, int64_t ldg, int n, double m_total,
                                        double* __restrict__ K, int64_t ldk) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n) return;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const double g = (double)G[(int64_t)lo * ldg + hi];
    double k;
    if (CODING == 0) {
        k = g / (2.0 * m_total) + 0.5;                                   // :53
    } else {
        if (i == j) {
            k = 1.0;                                                     // :35 never fills the diagonal; 0/m + 1
        } else {
            const double l1 = (double)G[(int64_t)i * ldg + i] + (double)G[(int64_t)j * ldg + j] - 2.0 * g;
            const double cnt = m_total - 0.5 * l1;                       // count0 + 0.5 count1 (:38)
            const float q = (float)cnt / (float)m_total;                 // float32 quotient (:51)
            k = (double)q;
        }
    }
    K[(int64_t)i * ldk + j] = k;
}

// ---------------------------------------------------------------------------------------------------
// Fused finalisation (kinship.py:50-55 + scale_k :94-100) straight from the integer Gram, symmetric 32 x 32 tiles:
//   kin_tile_sums_kernel  : per tile (ti <= tj) the sum of its kinship entries (off-diagonal tiles count twice) and its
//                           share of the trace -- the two scalars of scale_k -- WITHOUT writing K; kin_final_sums_kernel adds
//                           the per-tile partials in a fixed order (deterministic)
//   kin_write_tile_kernel : K[i][j] and, through a shared-memory transpose, K[j][i], scaled by *scale (or 1): every global
//                           access is coalesced and the valid half of G is read once.
// The unfused form (finalize kernel with a column-strided read of the mirrored half, then row sums, then an in-place scale)
// moved 3.6 GB for a 0.8 GB result.
// ---------------------------------------------------------------------------------------------------
template <int CODING>
__device__ __forceinline__ double kin_value(int g, int gii, int gjj, bool diag, double m_total) {
    if (CODING == 0) return (double)g / (2.0 * m_total) + 0.5;                          // :53
    if (diag) return 1.0;                                                                // :35 never fills the diagonal; 0/m + 1
    const double l1 = (double)gii + (double)gjj - 2.0 * (double)g;
    const double cnt = m_total - 0.5 * l1;                                               // count0 + 0.5 count1 (:38)
    return (double)((float)cnt / (float)m_total);                                        // float32 quotient (:51)
}

template <int CODING>
static __global__ void __launch_bounds__(256) kin_tile_sums_kernel(const int32_t* __restrict__ G, int64_t ldg, int n, double m_total,
                                                                   double* __restrict__ partial /* [tiles][2] */) {
    __shared__ double red[2][8];
    const int T = (n + 31) / 32;
    const int tile = blockIdx.x;
    // tile index -> (ti, tj), ti <= tj, row-major over the upper triangle of tiles
    int ti = (int)(((2.0 * T + 1.0) - sqrt((2.0 * T + 1.0) * (2.0 * T + 1.0) - 8.0 * (double)tile)) * 0.5);
    while ((int64_t)(ti + 1) * T - (int64_t)(ti + 1) * ti / 2 <= tile) ++ti;
    while ((int64_t)ti * T - (int64_t)ti * (ti - 1) / 2 > tile) --ti;
    const int tj = ti + (tile - (int)((int64_t)ti * T - (int64_t)ti * (ti - 1) / 2));
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = tj * 32 + tx;
    const int gjj = (CODING == 1 && j < n) ? G[(int64_t)j * ldg + j] : 0;
    double s = 0.0, tr = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = ti * 32 + ty + 8 * r;
        if (i >= n || j >= n) continue;
        if (ti == tj && j < i) continue;                        // diagonal tile: upper part + diagonal only
        const int gii = CODING == 1 ? G[(int64_t)i * ldg + i] : 0;
        const double k = kin_value<CODING>(G[(int64_t)i * ldg + j], gii, gjj, i == j, m_total);
        if (i == j) { s += k; tr += k; }
        else s += 2.0 * k;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        tr += __shfl_xor_sync(0xffffffffu, tr, o);
    }
    if (tx == 0) { red[0][ty] = s; red[1][ty] = tr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
        partial[2 * (int64_t)tile] = a;
        partial[2 * (int64_t)tile + 1] = b;
    }
}
// out[0] = sum(K), out[1] = trace(K), out[2] = scale_k factor (n - 1) / (trace - sum / n)   (kinship.py:95-96)
static __global__ void __launch_bounds__(1024) kin_final_sums_kernel(const double* __restrict__ partial, int64_t tiles, int n,
                                                                     double* __restrict__ out) {
    __shared__ double red[2][32];
    double s = 0.0, tr = 0.0;
    for (int64_t i = threadIdx.x; i < tiles; i += 1024) {
        s += partial[2 * i];
        tr += partial[2 * i + 1];
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        tr += __shfl_xor_sync(0xffffffffu, tr, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = tr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 32; ++w) { a += red[0][w]; b += red[1][w]; }
        out[0] = a;
        out[1] = b;
        out[2] = (double)(n - 1) / (b - a / (double)n);
    }
}
template <int CODING>
static __global__ void __launch_bounds__(256) kin_write_tile_kernel(const int32_t* __restrict__ G, int64_t ldg, int n, double m_total,
                                                                    const double* __restrict__ scale /* nullable */, double* __restrict__ K,
                                                                    int64_t ldk) {
    __shared__ double tile_s[32][33];
    const int T = (n + 31) / 32;
    const int tile = blockIdx.x;
    int ti = (int)(((2.0 * T + 1.0) - sqrt((2.0 * T + 1.0) * (2.0 * T + 1.0) - 8.0 * (double)tile)) * 0.5);
    while ((int64_t)(ti + 1) * T - (int64_t)(ti + 1) * ti / 2 <= tile) ++ti;
    while ((int64_t)ti * T - (int64_t)ti * (ti - 1) / 2 > tile) --ti;
    const int tj = ti + (tile - (int)((int64_t)ti * T - (int64_t)ti * (ti - 1) / 2));
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const double sc = scale ? *scale : 1.0;
    const int j = tj * 32 + tx;
    const int gjj = (CODING == 1 && j < n) ? G[(int64_t)j * ldg + j] : 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int li = ty + 8 * r, i = ti * 32 + li;
        double k = 0.0;
        if (i < n && j < n) {
            // inside a diagonal tile the entry below the diagonal is read through its mirror (both are valid there, same value)
            const int lo = i < j ? i : j, hi = i < j ? j : i;
            const int gii = CODING == 1 ? G[(int64_t)i * ldg + i] : 0;
            k = kin_value<CODING>(G[(int64_t)lo * ldg + hi], gii, gjj, i == j, m_total) * sc;
            K[(int64_t)i * ldk + j] = k;
        }
        tile_s[li][tx] = k;
    }
    if (ti == tj) return;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int lj = ty + 8 * r;                              // row of the mirrored tile = column of this one
        const int jj = tj * 32 + lj, ii = ti * 32 + tx;
        if (jj < n && ii < n) K[(int64_t)jj * ldk + ii] = tile_s[tx][lj];
    }
}

// out-of-place scale_k: dst = src * (*scale)   (16-byte accesses; the row sums of src were taken by rowsum_kernel)
static __global__ void __launch_bounds__(256) scaled_copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, int64_t n2,
                                                                 const double* __restrict__ scale, const double* __restrict__ src_last,
                                                                 double* __restrict__ dst_last) {
    const double s = *scale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        double2 v = src[i];
        v.x *= s;
        v.y *= s;
        dst[i] = v;
    }
    if (src_last && blockIdx.x == 0 && threadIdx.x == 0) *dst_last = *src_last * s;
}
// out[2] = (n - 1) / (out[1] - out[0] / n) from out[0] = sum(K), out[1] = trace(K)
static __global__ void scale_factor_kernel(double* out, int n) { out[2] = (double)(n - 1) / (out[1] - out[0] / (double)n); }

// The tensor-core Gram fills the 256 x 256 blocks (I, J) with I <= J of the padded square.  For the all-reduce between ranks
// those blocks are packed back to back (slot J (J + 1) / 2 + I, 256 KB each) so that only the valid half of the matrix
// crosses NVLink, and unpacked again afterwards.  One CTA per block, 16-byte accesses.
template <bool UNPACK>
static __global__ void __launch_bounds__(256) gram_tri_pack_kernel(int32_t* __restrict__ G, int64_t ldg, int32_t* __restrict__ packed) {
    const int slot = blockIdx.x;
    int J = (int)((sqrt(8.0 * (double)slot + 1.0) - 1.0) * 0.5);
    while ((J + 1) * (J + 2) / 2 <= slot) ++J;
    while (J * (J + 1) / 2 > slot) --J;
    const int I = slot - J * (J + 1) / 2;
    uint4* p = reinterpret_cast<uint4*>(packed + (int64_t)slot * 65536);
    for (int idx = threadIdx.x; idx < 256 * 64; idx += 256) {
        const int r = idx >> 6, c4 = idx & 63;
        uint4* g = reinterpret_cast<uint4*>(G + (int64_t)(I * 256 + r) * ldg + J * 256) + c4;
        if (UNPACK) *g = p[idx]; else p[idx] = *g;
    }
}

// mirror the upper triangle of the integer Gram into the lower one (for downloads)
static __global__ void gram_mirror_kernel(int32_t* __restrict__ G, int64_t ldg, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n || j >= i) return;
    G[(int64_t)i * ldg + j] = G[(int64_t)j * ldg + i];
}

// ---------------------------------------------------------------------------------------------------
// scale_k (kinship.py:94-100): row sums + diagonal, then a single-block final reduction (deterministic)
// ---------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) rowsum_kernel(const double* __restrict__ K, int64_t ldk, int n,
                                                     double* __restrict__ rowsum) {
    __shared__ double red[8];
    const int i = blockIdx.x;
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += 256) s += K[(int64_t)i * ldk + j];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tsum = 0.0;
        for (int w = 0; w < 8; ++w) tsum += red[w];
        rowsum[i] = tsum;
    }
}
// out[0] = sum(rowsum), out[1] = trace
static __global__ void __launch_bounds__(1024) scale_k_reduce_kernel(const double* __restrict__ rowsum,
                                                              const double* __restrict__ K, int64_t ldk, int n,
                                                              double* __restrict__ out) {
    __shared__ double red[2][32];
    double s = 0.0, tr = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        s += rowsum[i];
        tr += K[(int64_t)i * ldk + i];
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        tr += __shfl_xor_sync(0xffffffffu, tr, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = tr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 32; ++w) { a += red[0][w]; b += red[1][w]; }
        out[0] = a;
        out[1] = b;
    }
}
static __global__ void scale_matrix_kernel(double* __restrict__ A, int64_t ld, int rows, int cols, double alpha) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j < cols && i < rows) A[(int64_t)i * ld + j] *= alpha;
}
static __global__ void scale_rows_kernel(double* __restrict__ A, int64_t ld, int rows, int cols, const double* __restrict__ d) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j < cols && i < rows) A[(int64_t)i * ld + j] *= d[i];
}
// The rotation of the EMMAX scan in one pass (linear_models.py:898 + :1299-1303): R = (I - QQ') diag(d) U, i.e.
//     R[k][i] = d[k] U[k][i] - sum_c Q[k][c] QtH[c][i],      QtH = Q' diag(d) U  ([q x n], formed by one skinny GEMM)
// read U once, write R once -- instead of copy + row scaling (H_sqrt_inv) + copy + rank-q GEMM update.
static __global__ void __launch_bounds__(256) rotation_kernel(const double* __restrict__ U, const double* __restrict__ d,
                                                              const double* __restrict__ Q /* [n x q] or null */,
                                                              const double* __restrict__ QtH /* [q x n] */, int q, int n,
                                                              double* __restrict__ R) {
    const int k = blockIdx.y;
    const double dk = d[k];
    double qk[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) qk[c] = c < q ? Q[(int64_t)k * q + c] : 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double r = dk * U[(int64_t)k * n + i];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c < q) r -= qk[c] * QtH[(int64_t)c * n + i];
        R[(int64_t)k * n + i] = r;
    }
}
static __global__ void add_diag_kernel(double* __restrict__ A, int64_t ld, int n, double alpha) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[(int64_t)i * ld + i] += alpha;
}

// ---------------------------------------------------------------------------------------------------
// per-SNP sums (int64) over individuals; one warp per row
// ---------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) snp_row_sums_kernel(const int8_t* __restrict__ snps, int64_t pitch, int64_t m,
                                                           int n, long long* __restrict__ sums,
                                                           long long* __restrict__ sumsq) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= m) return;
    const int8_t* x = snps + row * pitch;
    int s = 0, q = 0;
    for (int i0 = lane * 16; i0 < n; i0 += 512) {
        const uint4 v = *reinterpret_cast<const uint4*>(x + i0);     // zero padded beyond n
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            s = __dp4a((int)w[k], 0x01010101, s);
            q = __dp4a((int)w[k], (int)w[k], q);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
        sums[row] = s;
        if (sumsq) sumsq[row] = q;
    }
}

// ---------------------------------------------------------------------------------------------------
// IBD standardisation (kinship.py:66 / hdf5_data.py:50,103): z = (x - mean) / std, ddof = 0, two-pass.
// One block per selected SNP; rows[] lists the resident row indices.  Z is [count x ldz] FP64.
// ---------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) standardise_rows_kernel(const int8_t* __restrict__ snps, int64_t pitch,
                                                               const long long* __restrict__ rows, int n,
                                                               double* __restrict__ Z, int64_t ldz,
                                                               int* __restrict__ bad_flag) {
    __shared__ double red[8];
    __shared__ double bc[2];
    const int8_t* x = snps + rows[blockIdx.x] * pitch;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += (double)x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tsum = 0.0;
        for (int w = 0; w < 8; ++w) tsum += red[w];
        bc[0] = tsum / (double)n;
    }
    __syncthreads();
    const double mean = bc[0];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double d = (double)x[i] - mean;
        v += d * d;
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tsum = 0.0;
        for (int w = 0; w < 8; ++w) tsum += red[w];
        bc[1] = sqrt(tsum / (double)n);
        if (!(bc[1] > 0.0)) atomicOr(bad_flag, 1);
    }
    __syncthreads();
    const double sd = bc[1];
    double* z = Z + (int64_t)blockIdx.x * ldz;
    for (int i = threadIdx.x; i < n; i += 256) z[i] = ((double)x[i] - mean) / sd;
}

}  // namespace mmg
