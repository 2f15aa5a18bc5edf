// int8 x int8 -> int32 GEMM core on the 5th-gen tensor cores (tcgen05.mma kind::i8).
//
//   D[128 x 256] (TMEM, int32) = A[128 x K] * B[256 x K]'      both operands K-major, 128B swizzle
//
// Warp-specialised, persistent: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> registers -> Epi functor).  Two 256-column accumulators in TMEM
// (512 columns) let the epilogue of tile t overlap the mainloop of tile t+1; a 4-stage smem ring
// (4 x 48 KB) decouples TMA from the MMA issue.  Used by
//   - the kinship Gram  G += P P'   (kinship.py:44 / :33-41)          -> GramEpi
//   - the EMMAX scan    x'(R'R)x on integer slices of R'R             -> QuadEpi (scan_tc.cuh)
#pragma once
#include "ptx.cuh"

namespace mmg {

constexpr int TC_BM = 128;            // UMMA M (rows of A per tile)
constexpr int TC_BN = 256;            // UMMA N (rows of B per tile)
constexpr int TC_BK = 128;            // bytes of K per pipeline stage = one 128B swizzle atom
constexpr int TC_UMMA_K = 32;         // K per tcgen05.mma for 8-bit operands
constexpr int TC_STAGES = 4;
constexpr int TC_A_BYTES = TC_BM * TC_BK;                 // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK;                 // 32 KB
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;   // 48 KB
constexpr int TC_THREADS = 192;
constexpr int TC_ACC_STAGES = 2;
constexpr int TC_TMEM_COLS = TC_ACC_STAGES * TC_BN;       // 512
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct TcTile {
    int m0;    // first row of A (TMA coordinate)
    int n0;    // first row of B (TMA coordinate)
    int kb0;   // K range in units of TC_BK bytes: [kb0, kb1)
    int kb1;
    int aux0;  // epilogue-defined
    int aux1;
    int col0;  // epilogue-defined column offset
    int pad_;
};

template <class Epi>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const TcTile* __restrict__ tiles, int num_groups, int tiles_per_group, int table_stride,
                  int group_m_step, const typename Epi::Params ep) {
    // group g processes tiles[g * table_stride + ti], ti < tiles_per_group, with m0 advanced by g * group_m_step
    // (table_stride = 0: every group shares one table, e.g. the N tiles x slices of a 128-SNP row block).
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle wants 1024-byte aligned stage bases
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* full_bar = bars;                               // [TC_STAGES]
    uint64_t* empty_bar = bars + TC_STAGES;                  // [TC_STAGES]
    uint64_t* tfull_bar = bars + 2 * TC_STAGES;              // [TC_ACC_STAGES]
    uint64_t* tempty_bar = bars + 2 * TC_STAGES + TC_ACC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 2 * TC_ACC_STAGES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < TC_ACC_STAGES; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], 4);   // one elected lane of each epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TC_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int g = blockIdx.x; g < num_groups; g += gridDim.x) {
                for (int ti = 0; ti < tiles_per_group; ++ti) {
                    TcTile t = tiles[(int64_t)g * table_stride + ti];
                    t.m0 += g * group_m_step;
                    for (int kb = t.kb0; kb < t.kb1; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], TC_STAGE_BYTES);
                        uint8_t* sa = smem + stage * TC_STAGE_BYTES;
                        tma_load_2d(sa, &tmA, &full_bar[stage], kb * TC_BK, t.m0);
                        tma_load_2d(sa + TC_A_BYTES, &tmB, &full_bar[stage], kb * TC_BK, t.n0);
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(TC_BM, TC_BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int g = blockIdx.x; g < num_groups; g += gridDim.x) {
                for (int ti = 0; ti < tiles_per_group; ++ti) {
                    const TcTile t = tiles[(int64_t)g * table_stride + ti];
                    mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * TC_BN;
                    for (int kb = t.kb0; kb < t.kb1; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * TC_STAGE_BYTES);
                        const uint64_t da = umma_desc_kmajor_sw128(sa);
                        const uint64_t db = umma_desc_kmajor_sw128(sa + TC_A_BYTES);
#pragma unroll
                        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                            // advance 32 bytes of K inside the swizzle atom: +2 in the (addr >> 4) field
                            umma_i8(d_tmem, da + (uint64_t)(k * (TC_UMMA_K >> 4)), db + (uint64_t)(k * (TC_UMMA_K >> 4)),
                                    idesc, (kb > t.kb0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit(&empty_bar[stage]);           // frees the smem slot when these MMAs retire
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tfull_bar[acc]);                 // accumulator complete -> epilogue
                    if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps 2..5 =====================
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
        const int row = quad * 32 + lane;          // row of the 128-row tile held by this thread
        Epi epi;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int g = blockIdx.x; g < num_groups; g += gridDim.x) {
            epi.begin_group(ep, g, row);
            for (int ti = 0; ti < tiles_per_group; ++ti) {
                TcTile t = tiles[(int64_t)g * table_stride + ti];
                t.m0 += g * group_m_step;
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * TC_BN + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
                for (int c = 0; c < TC_BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    epi.chunk(ep, t, row, c, v);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
            epi.end_group(ep, g, row);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

// ---- epilogue of the kinship Gram: G[m0+row][n0 + 32c .. +31] (+)= acc ------------------------------
struct GramEpi {
    struct Params {
        int32_t* G;        // [n_pad x ld] int32, n_pad multiple of 256 so no bounds checks are needed
        int64_t ld;
        int accumulate;
    };
    __device__ __forceinline__ void begin_group(const Params&, int, int) {}
    __device__ __forceinline__ void end_group(const Params&, int, int) {}
    __device__ __forceinline__ void chunk(const Params& p, const TcTile& t, int row, int c, const uint32_t (&v)[32]) {
        int4* dst = reinterpret_cast<int4*>(p.G + (int64_t)(t.m0 + row) * p.ld + t.n0 + c * 32);
        if (p.accumulate) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int4 o = dst[j];
                o.x += (int)v[4 * j + 0];
                o.y += (int)v[4 * j + 1];
                o.z += (int)v[4 * j + 2];
                o.w += (int)v[4 * j + 3];
                dst[j] = o;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                dst[j] = make_int4((int)v[4 * j + 0], (int)v[4 * j + 1], (int)v[4 * j + 2], (int)v[4 * j + 3]);
        }
    }
};

}  // namespace mmg
