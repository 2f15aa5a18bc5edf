// Stage 3 (FP64 tensor-core path): the EMMAX rotation GEMM of linear_models.py:1315-1318 fused with the
// per-SNP OLS reductions of :1319-1339 and the F / p-value epilogue of :1345-1349.
//
//   C[s][k] = sum_i X[s][i] * R[k][i]         X: int8 genotypes (SNP-major), R = M' in FP64, row k contiguous
//   xx[s]   = sum_k C[s][k]^2                 (= x~.x~)
//   xy[s]   = sum_k C[s][k] * y[k]            (= x~.y~, y = residual rotated phenotype)
//   -> rss, F, p  (closed form of the one-column lstsq of :1328)
//
// The m x n_out rotated matrix is never written: each CTA owns 128 SNP rows, sweeps every 128-column tile
// of R and keeps the running row sums in registers.  Math is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4, the
// only FP64 tensor shape sm_100a has); the int8 A operand is widened to FP64 in registers.  Operands are
// staged by a 4-stage cp.async ring that runs continuously across the N tiles of a row block.
//
// PERM mode (linear_models.py:1157-1164): R is W' ([P x n], the permuted phenotypes rotated back) and the
// epilogue reduces max_s (C[s][p] - mu_s*colsum[p])^2 / xx[s] per column p.
#pragma once
#include "fdist.cuh"
#include "ptx.cuh"

namespace mmg {

constexpr int SD_BM = 128, SD_BN = 128, SD_BK = 32, SD_STAGES = 4, SD_THREADS = 256;
constexpr int SD_A_PITCH = 48;                       // bytes per A row in smem (32 data + 16 pad: conflict-free)
constexpr int SD_B_PITCH = 36;                       // doubles per B row (32 + 4: conflict-free LDS.64)
constexpr int SD_A_STAGE = SD_BM * SD_A_PITCH;       // 6144 B
constexpr int SD_B_STAGE = SD_BN * SD_B_PITCH * 8;   // 36864 B
constexpr int SD_STAGE_BYTES = SD_A_STAGE + SD_B_STAGE;
constexpr int SD_SMEM_BYTES = SD_STAGES * SD_STAGE_BYTES + 2 * 4 * SD_BM * 8;   // + reduction scratch
// real-valued genotype rows (dosages, linear_models.py:1317 takes any numeric row): the A operand is staged as FP64 with the
// pitch of B, three stages instead of four (3 x 72 KB + scratch = 224 KB)
constexpr int SD_AD_PITCH = 36;                      // doubles per A row
constexpr int SD_AD_STAGE = SD_BM * SD_AD_PITCH * 8; // 36864 B

template <typename XT> struct SdShape;
template <> struct SdShape<int8_t> {
    static constexpr int STAGES = SD_STAGES, A_STAGE = SD_A_STAGE;
};
template <> struct SdShape<double> {
    static constexpr int STAGES = 3, A_STAGE = SD_AD_STAGE;
};
template <typename XT> constexpr int sd_stage_bytes() { return SdShape<XT>::A_STAGE + SD_B_STAGE; }
template <typename XT> constexpr int sd_smem_bytes() { return SdShape<XT>::STAGES * sd_stage_bytes<XT>() + 2 * 4 * SD_BM * 8; }

struct ScanDmmaParams {
    const void* snps;        // genotypes: int8 (resident block) or FP64 rows; row pitch `pitch` ELEMENTS, zero padded to k_pad
    int64_t pitch;
    int64_t row_begin;       // first SNP row of this call
    int64_t row_count;
    const double* R;         // [n_out_pad x ldr], zero padded (rows to 128, cols to 32)
    int64_t ldr;
    int n_out_pad;           // multiple of 128
    int k_pad;               // K extent (individuals) rounded up to a multiple of 32, <= pitch and <= ldr
    const double* y;         // [n_out_pad] rotated residual phenotype (zero padded); SUMSQ mode
    const double* mu;        // [row_count] SNP means or nullptr (centre SNPs: x_c = x - mu)
    const double* r1;        // [n_out_pad] R*1 (needed with mu)
    double h0_rss, n_p, lbeta;
    // outputs (device, length row_count; nullable)
    double *xx, *xy, *rss, *f, *p, *var_perc;
    // PERM mode
    const double* xx_in;     // [row_count] x~.x~ of the centred SNPs
    unsigned long long* ratio_max;   // [n_out_pad] bit patterns of non-negative doubles
    // SD_MODE_SQUARE_STORE: C[s][k'] = sum_i X[s][i]^2 R[k'][i] is stored, not reduced (scan_shared.cuh: the contraction of the
    // squared rotated genotypes with the per-phenotype weights)
    double* cstore;          // [row_count x ldc]
    int64_t ldc;
    // short scans (fewer 128-row blocks than SMs): the column tiles of R are split over nsplit CTAs per row block, each writes its
    // partial (x~.x~, x~.y~) to part[(which * nsplit + split) * row_count + row]; scan_split_finish_kernel adds them in a fixed order
    int nsplit;              // 0 / 1: off
    double* part;
};
constexpr int SD_MODE_SCAN = 0, SD_MODE_SQUARE_STORE = 1;

__device__ __forceinline__ double i8_to_f64(int v) {
    // exact int -> double without the conversion pipe: 2^52 + 2^31 + v has v ^ 0x80000000 in its low word
    return __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)v)) - 4503601774854144.0;
}

template <bool PERM, typename XT = int8_t, int MODE = SD_MODE_SCAN>
static __global__ void __launch_bounds__(SD_THREADS, 1) scan_dmma_kernel(const ScanDmmaParams prm) {
    constexpr bool REAL = sizeof(XT) == 8;
    constexpr int SD_STAGES = SdShape<XT>::STAGES, SD_A_STAGE = SdShape<XT>::A_STAGE, SD_STAGE_BYTES = SdShape<XT>::A_STAGE + SD_B_STAGE;
    extern __shared__ __align__(16) uint8_t sd_smem[];
    double* red = reinterpret_cast<double*>(sd_smem + SD_STAGES * SD_STAGE_BYTES);   // [2][4][128]
    const XT* snps = static_cast<const XT*>(prm.snps);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;           // 2 x 4 warps, warp tile 64 x 32
    const int lr = lane >> 2, lc = lane & 3;
    const int KT = prm.k_pad / SD_BK;
    const int NT = prm.n_out_pad / SD_BN;
    const int64_t num_blocks = (prm.row_count + SD_BM - 1) / SD_BM;
    const int nsplit = (!PERM && MODE == SD_MODE_SCAN && prm.nsplit > 1) ? prm.nsplit : 1;
    const int nt_per = (NT + nsplit - 1) / nsplit;

    for (int64_t item = blockIdx.x; item < num_blocks * nsplit; item += gridDim.x) {
        const int64_t mb = item / nsplit;
        const int sp = (int)(item - mb * nsplit);
        const int nt_begin = sp * nt_per, nt_count = max(0, min(NT, nt_begin + nt_per) - nt_begin);
        const int64_t row0 = mb * SD_BM;               // relative to row_begin
        const int total_it = nt_count * KT;

        auto load_stage = [&](int it) {
            const int stage = it % SD_STAGES;
            const int ntl = it / KT, kt = it - ntl * KT;
            const int nt = nt_begin + ntl;
            uint8_t* sa = sd_smem + stage * SD_STAGE_BYTES;
            double* sb = reinterpret_cast<double*>(sa + SD_A_STAGE);
            if constexpr (!REAL) {   // A: 128 rows x 32 B -> one 16 B chunk per thread
                const int r = tid >> 1, h = tid & 1;
                const bool valid = (row0 + r) < prm.row_count;
                const int64_t grow = prm.row_begin + (valid ? row0 + r : 0);
                cp_async16_zfill(sa + r * SD_A_PITCH + 16 * h, snps + grow * prm.pitch + (int64_t)kt * SD_BK + 16 * h,
                                 valid);
            } else {
                double* sad = reinterpret_cast<double*>(sa);
#pragma unroll
                for (int j = 0; j < 8; ++j) {   // A: 128 rows x 256 B -> eight 16 B chunks per thread
                    const int c = tid + SD_THREADS * j;
                    const int r = c >> 4, part = c & 15;
                    const bool valid = (row0 + r) < prm.row_count;
                    const int64_t grow = prm.row_begin + (valid ? row0 + r : 0);
                    cp_async16_zfill(sad + r * SD_AD_PITCH + 2 * part, snps + grow * prm.pitch + (int64_t)kt * SD_BK + 2 * part, valid);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {   // B: 128 rows x 256 B -> eight 16 B chunks per thread
                const int c = tid + SD_THREADS * j;
                const int r = c >> 4, part = c & 15;
                cp_async16(sb + r * SD_B_PITCH + 2 * part,
                           prm.R + (int64_t)(nt * SD_BN + r) * prm.ldr + (int64_t)kt * SD_BK + 2 * part);
            }
        };

        double acc[8][4][2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        double xx[8], xy[8], mus[8], xxin[8];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi) {
            xx[mi] = 0.0;
            xy[mi] = 0.0;
            const int64_t r = row0 + wm * 64 + mi * 8 + lr;
            const bool valid = r < prm.row_count;
            mus[mi] = (prm.mu != nullptr && valid) ? prm.mu[r] : 0.0;
            xxin[mi] = (PERM && valid) ? prm.xx_in[r] : 0.0;
        }

#pragma unroll
        for (int s = 0; s < SD_STAGES - 1; ++s) {
            if (s < total_it) load_stage(s);
            cp_async_commit();
        }

        for (int it = 0; it < total_it; ++it) {
            cp_async_wait<SD_STAGES - 2>();
            __syncthreads();
            if (it + SD_STAGES - 1 < total_it) load_stage(it + SD_STAGES - 1);
            cp_async_commit();

            const int stage = it % SD_STAGES;
            const uint8_t* sa = sd_smem + stage * SD_STAGE_BYTES;
            const double* sb = reinterpret_cast<const double*>(sa + SD_A_STAGE);

            // A fragments for the whole stage: row's 32 bytes -> byte (lane&3) of each word
            uint32_t aw[REAL ? 1 : 8][8];
            if constexpr (!REAL) {
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) {
                    const uint4* ap = reinterpret_cast<const uint4*>(sa + (wm * 64 + mi * 8 + lr) * SD_A_PITCH);
                    const uint4 lo = ap[0], hi = ap[1];
                    aw[mi][0] = lo.x; aw[mi][1] = lo.y; aw[mi][2] = lo.z; aw[mi][3] = lo.w;
                    aw[mi][4] = hi.x; aw[mi][5] = hi.y; aw[mi][6] = hi.z; aw[mi][7] = hi.w;
                }
            }
#pragma unroll
            for (int kk = 0; kk < SD_BK / 4; ++kk) {
                double b[4];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = sb[(wn * 32 + ni * 8 + lr) * SD_B_PITCH + kk * 4 + lc];
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) {
                    double a;
                    if constexpr (REAL) {
                        a = reinterpret_cast<const double*>(sa)[(wm * 64 + mi * 8 + lr) * SD_AD_PITCH + kk * 4 + lc];
                        if constexpr (MODE == SD_MODE_SQUARE_STORE) a *= a;
                    } else {
                        const int v = (int)(int8_t)((aw[mi][kk] >> (8 * lc)) & 0xffu);
                        a = i8_to_f64(v);
                    }
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma_8x8x4(acc[mi][ni][0], acc[mi][ni][1], a, b[ni]);
                }
            }

            const int ntl = it / KT;
            const int nt = nt_begin + ntl;
            if (it - ntl * KT == KT - 1) {
                // ---- epilogue of this N tile: fold the 128 x 128 block of C into the running row sums ----
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const int col = nt * SD_BN + wn * 32 + ni * 8 + 2 * lc;
                    if constexpr (MODE == SD_MODE_SQUARE_STORE) {
#pragma unroll
                        for (int mi = 0; mi < 8; ++mi) {
                            const int64_t r = row0 + wm * 64 + mi * 8 + lr;
                            if (r < prm.row_count)
                                *reinterpret_cast<double2*>(prm.cstore + r * prm.ldc + col) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                            acc[mi][ni][0] = 0.0;
                            acc[mi][ni][1] = 0.0;
                        }
                    } else if (!PERM) {
                        const double y0 = prm.y[col], y1 = prm.y[col + 1];
                        double r10 = 0.0, r11 = 0.0;
                        if (prm.mu != nullptr) { r10 = prm.r1[col]; r11 = prm.r1[col + 1]; }
#pragma unroll
                        for (int mi = 0; mi < 8; ++mi) {
                            const double c0 = acc[mi][ni][0] - mus[mi] * r10;
                            const double c1 = acc[mi][ni][1] - mus[mi] * r11;
                            xx[mi] = fma(c0, c0, fma(c1, c1, xx[mi]));
                            xy[mi] = fma(c0, y0, fma(c1, y1, xy[mi]));
                            acc[mi][ni][0] = 0.0;
                            acc[mi][ni][1] = 0.0;
                        }
                    } else {
                        const double r10 = prm.r1[col], r11 = prm.r1[col + 1];
                        double m0 = 0.0, m1 = 0.0;
#pragma unroll
                        for (int mi = 0; mi < 8; ++mi) {
                            const double c0 = acc[mi][ni][0] - mus[mi] * r10;
                            const double c1 = acc[mi][ni][1] - mus[mi] * r11;
                            if (xxin[mi] > 0.0) {
                                m0 = fmax(m0, c0 * c0 / xxin[mi]);
                                m1 = fmax(m1, c1 * c1 / xxin[mi]);
                            }
                            acc[mi][ni][0] = 0.0;
                            acc[mi][ni][1] = 0.0;
                        }
                        // max over the 8 row-lanes that share these two columns
#pragma unroll
                        for (int o = 4; o < 32; o <<= 1) {
                            m0 = fmax(m0, __shfl_xor_sync(0xffffffffu, m0, o));
                            m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
                        }
                        if (lr == 0) {
                            atomicMax(prm.ratio_max + col, (unsigned long long)__double_as_longlong(m0));
                            atomicMax(prm.ratio_max + col + 1, (unsigned long long)__double_as_longlong(m1));
                        }
                    }
                }
            }
        }
        cp_async_wait<0>();

        if (!PERM && MODE == SD_MODE_SCAN) {
            // ---- row reduction: 4 lanes per row inside the warp, then the 4 warps along N ----
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) {
                xx[mi] += __shfl_xor_sync(0xffffffffu, xx[mi], 1);
                xx[mi] += __shfl_xor_sync(0xffffffffu, xx[mi], 2);
                xy[mi] += __shfl_xor_sync(0xffffffffu, xy[mi], 1);
                xy[mi] += __shfl_xor_sync(0xffffffffu, xy[mi], 2);
            }
            __syncthreads();
            if (lc == 0) {
#pragma unroll
                for (int mi = 0; mi < 8; ++mi) {
                    const int r = wm * 64 + mi * 8 + lr;
                    red[(0 * 4 + wn) * SD_BM + r] = xx[mi];
                    red[(1 * 4 + wn) * SD_BM + r] = xy[mi];
                }
            }
            __syncthreads();
            if (tid < SD_BM && row0 + tid < prm.row_count) {
                const double sxx = red[0 * SD_BM + tid] + red[1 * SD_BM + tid] + red[2 * SD_BM + tid] + red[3 * SD_BM + tid];
                const double sxy = red[4 * SD_BM + tid] + red[5 * SD_BM + tid] + red[6 * SD_BM + tid] + red[7 * SD_BM + tid];
                const int64_t o = row0 + tid;
                if (nsplit > 1) {
                    prm.part[(int64_t)sp * prm.row_count + o] = sxx;
                    prm.part[(int64_t)(nsplit + sp) * prm.row_count + o] = sxy;
                } else {
                if (prm.xx) prm.xx[o] = sxx;
                if (prm.xy) prm.xy[o] = sxy;
                if (prm.p || prm.f || prm.rss || prm.var_perc) {
                    // one-column OLS in closed form (linear_models.py:1328,1345-1349); `if rss:` guard of :1329
                    double rss = prm.h0_rss, f = 0.0, vp = 0.0, pv = 1.0;
                    if (sxx > 0.0) {
                        const double r2 = (sxy * sxy) / (sxx * prm.h0_rss);
                        const double rs = prm.h0_rss - (sxy * sxy) / sxx;
                        if (rs != 0.0) {
                            rss = rs;
                            vp = r2;
                            f = prm.n_p * r2 / (1.0 - r2);
                            pv = f_sf(f, 1.0, prm.n_p, prm.lbeta);
                        }
                    }
                    if (prm.rss) prm.rss[o] = rss;
                    if (prm.f) prm.f[o] = f;
                    if (prm.var_perc) prm.var_perc[o] = vp;
                    if (prm.p) prm.p[o] = pv;
                }
                }
            }
        }
        __syncthreads();
    }
}

// x . W[:, v] for a few FP64 vectors (with_betas columns, means): one warp per SNP row, HBM-bound on X.
// W is [nv x ldw] (vector v contiguous).  dots is [row_count x nv].
// FP64 rows: x . W[:, v], one warp per row (coalesced 8-byte loads)
static __global__ void __launch_bounds__(256) row_dots_f64_kernel(const double* __restrict__ xs, int64_t pitch, int64_t row_count, int n,
                                                                  const double* __restrict__ w, double* __restrict__ dots) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= row_count) return;
    const double* x = xs + row * pitch;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s = fma(x[i], w[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) dots[row] = s;
}

template <int NV>
static __global__ void __launch_bounds__(256) snp_dots_kernel(const int8_t* __restrict__ snps, int64_t pitch, int64_t row_begin,
                                                       int64_t row_count, int n, const double* __restrict__ W,
                                                       int64_t ldw, double* __restrict__ dots) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= row_count) return;
    const int8_t* x = snps + (row_begin + row) * pitch;
    double s[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) s[v] = 0.0;
    for (int i0 = lane * 16; i0 < n; i0 += 32 * 16) {
        const uint4 q = *reinterpret_cast<const uint4*>(x + i0);   // pitch is padded: always readable, zeros beyond n
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int xv = (int)(int8_t)((w[j >> 2] >> (8 * (j & 3))) & 0xffu);
            if (i0 + j < n) {
                const double xd = (double)xv;
#pragma unroll
                for (int v = 0; v < NV; ++v) s[v] = fma(xd, W[(int64_t)v * ldw + i0 + j], s[v]);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[v] += __shfl_xor_sync(0xffffffffu, s[v], o);
        if (lane == 0) dots[row * NV + v] = s[v];
    }
}

// per-SNP F statistics from (xx, xy): the epilogue alone, for paths that computed the moments elsewhere
static __global__ void scan_stats_kernel(const double* __restrict__ xx, const double* __restrict__ xy, int64_t count,
                                  double h0_rss, double n_p, double lbeta, double* rss_o, double* f_o, double* p_o,
                                  double* vp_o) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= count) return;
    const double sxx = xx[o], sxy = xy[o];
    double rss = h0_rss, f = 0.0, vp = 0.0, pv = 1.0;
    if (sxx > 0.0) {
        const double r2 = (sxy * sxy) / (sxx * h0_rss);
        const double rs = h0_rss - (sxy * sxy) / sxx;
        if (rs != 0.0) {
            rss = rs;
            vp = r2;
            f = n_p * r2 / (1.0 - r2);
            pv = f_sf(f, 1.0, n_p, lbeta);
        }
    }
    if (rss_o) rss_o[o] = rss;
    if (f_o) f_o[o] = f;
    if (vp_o) vp_o[o] = vp;
    if (p_o) p_o[o] = pv;
}

// short scans: the partial moments of the column splits (ScanDmmaParams::part) added in split order, then the same epilogue
static __global__ void scan_split_finish_kernel(const double* __restrict__ part, int nsplit, int64_t count, double h0_rss, double n_p,
                                                double lbeta, double* xx_o, double* xy_o, double* rss_o, double* f_o, double* p_o, double* vp_o) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= count) return;
    double sxx = 0.0, sxy = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) {
        sxx += part[(int64_t)sp * count + o];
        sxy += part[(int64_t)(nsplit + sp) * count + o];
    }
    if (xx_o) xx_o[o] = sxx;
    if (xy_o) xy_o[o] = sxy;
    double rss = h0_rss, f = 0.0, vp = 0.0, pv = 1.0;
    if (sxx > 0.0) {
        const double r2 = (sxy * sxy) / (sxx * h0_rss);
        const double rs = h0_rss - (sxy * sxy) / sxx;
        if (rs != 0.0) {
            rss = rs;
            vp = r2;
            f = n_p * r2 / (1.0 - r2);
            pv = f_sf(f, 1.0, n_p, lbeta);
        }
    }
    if (rss_o) rss_o[o] = rss;
    if (f_o) f_o[o] = f;
    if (vp_o) vp_o[o] = vp;
    if (p_o) p_o[o] = pv;
}

static __global__ void f_sf_kernel(const double* __restrict__ f, int64_t count, double dfn, double dfd, double lbeta,
                            double* __restrict__ out) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o < count) out[o] = f_sf(f[o], dfn, dfd, lbeta);
}

}  // namespace mmg
