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
constexpr int TC_THREADS = 224;          // producer, MMA issuer, 4 epilogue warps, L2 prefetcher
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
    int flags; // TC_TILE_CHAIN: keep accumulating into the same TMEM accumulator with the NEXT tile (no hand-over, no epilogue)
};
constexpr int TC_TILE_CHAIN = 1;

// CS = cluster size (1, 2 or 4).  With CS > 1 the CS CTAs of a cluster walk the same tile sequence in lockstep
// on CS different 128-row A blocks that share the B tile: each CTA fetches 1/CS of the B rows and TMA-multicasts
// them into the shared memory of every CTA of the cluster, so the L2 -> SM operand traffic per CTA drops from
// 16+32 KB to 16+32/CS KB per K block.  A smem stage may only be refilled once every CTA of the cluster has
// consumed it, hence the MMA issuer's tcgen05.commit arrives on the empty barrier of all CS CTAs.
//
// Work assignment: cluster c, rank r processes group g = (c + i * num_clusters) * CS + r, i = 0, 1, ...;
// tile ti of that group is tiles[(g / CS) * table_stride + ti] with m0 += g * group_m_step + r * rank_m_step.
//   scan: one shared table (table_stride 0), group_m_step = 128 (consecutive SNP row blocks), rank_m_step = 0
//   Gram: one entry per cluster group (table_stride 1), group_m_step = 0, rank_m_step = 128 (adjacent row tiles)
//
// KIND = TC_KIND_MXF4: the operands are packed e2m1 (4-bit) values, 256 per 128-byte K block, multiplied with
// tcgen05.mma.kind::mxf4.block_scale (K = 64 per instruction: the same 32 bytes of K per MMA as int8, twice the elements, at
// twice the int8 issue rate).  The block scale factors (UE8M0, one per 32 elements) are all 2^0: TMEM columns 256..511 are
// filled with 0x7f bytes once and both scale-factor operands point there, which leaves ONE 256-column FP32 accumulator
// (the tiles of the Gram are hundreds of K blocks long, the missing epilogue overlap is ~1 % of a tile).  Sums of products of
// 0 / +-1 are integers < 2^24, so the FP32 accumulator is exact; the epilogue converts it to int32.
constexpr int TC_KIND_I8 = 0;
constexpr int TC_KIND_MXF4 = 1;
constexpr int TC_SF_COL = 256;        // first TMEM column of the constant scale factors (KIND = MXF4)

template <class Epi, int CS, int KIND = TC_KIND_I8>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const TcTile* __restrict__ tiles, int num_groups, int tiles_per_group, int table_stride,
                  int group_m_step, int rank_m_step, uint64_t policy_a, uint64_t policy_b, const typename Epi::Params ep, int prefetch = 0) {
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle wants 1024-byte aligned stage bases (the dynamic smem base is the same in every CTA of a
    // cluster, so CTA-relative offsets agree across the cluster)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* full_bar = bars;                               // [TC_STAGES]
    uint64_t* empty_bar = bars + TC_STAGES;                  // [TC_STAGES]
    constexpr int kAcc = KIND == TC_KIND_MXF4 ? 1 : TC_ACC_STAGES;
    uint64_t* tfull_bar = bars + 2 * TC_STAGES;              // [kAcc]
    uint64_t* tempty_bar = bars + 2 * TC_STAGES + kAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 2 * TC_ACC_STAGES);
    volatile uint32_t* progress = tmem_slot + 1;             // K blocks the producer has requested so far (read by the prefetch warp)

    // warp index through a shuffle: the role branches are then provably warp-uniform, so loop counters, addresses and
    // descriptors live in uniform registers and UTCIMMA / UTMALDG issue without per-lane ELECT loops
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int crank = (CS > 1) ? (int)cluster_ctarank() : 0;
    const int cluster_id = blockIdx.x / CS;
    const int num_clusters = gridDim.x / CS;
    const int num_cgroups = (num_groups + CS - 1) / CS;      // cluster-level groups (the last may be ragged)
    constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);
    constexpr int kBRows = TC_BN / CS;                       // B rows fetched by one CTA

    if (warp == 0 && lane == 0) {
        *progress = 0u;
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CS);   // one tcgen05.commit arrival from every CTA of the cluster
        }
        for (int a = 0; a < kAcc; ++a) {
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
    if (CS > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (KIND == TC_KIND_MXF4) {
        // constant scale factors 2^0 (UE8M0 0x7f) in every byte of columns 256..511, all 128 lanes: whatever layout the MMA
        // reads its A / B scale vectors in, it reads ones
        if (warp >= 2 && warp < 6) {
            const uint32_t taddr = tmem_base + TC_SF_COL + (static_cast<uint32_t>((warp & 3) * 32) << 16);
            for (int c = 0; c < (TC_TMEM_COLS - TC_SF_COL) / 8; ++c) tmem_st_32x8_const(taddr + c * 8, 0x7f7f7f7fu);
            tmem_st_wait();
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    if (warp == 0) {
        // ===================== TMA producer (all lanes run the loops, one elected lane issues) =====================
        {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t requested = 0;
            for (int cg = cluster_id; cg < num_cgroups; cg += num_clusters) {
                const int g = cg * CS + crank;
                for (int ti = 0; ti < tiles_per_group; ++ti) {
                    TcTile t = tiles[(int64_t)cg * table_stride + ti];
                    t.m0 += g * group_m_step + crank * rank_m_step;
                    for (int kb = t.kb0; kb < t.kb1; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&full_bar[stage], TC_STAGE_BYTES);
                            uint8_t* sa = smem + stage * TC_STAGE_BYTES;
                            tma_load_2d(sa, &tmA, &full_bar[stage], kb * TC_BK, t.m0, policy_a);
                            if (CS == 1) {
                                tma_load_2d(sa + TC_A_BYTES, &tmB, &full_bar[stage], kb * TC_BK, t.n0, policy_b);
                            } else {
                                tma_load_2d_mcast(sa + TC_A_BYTES + crank * kBRows * TC_BK, &tmB, &full_bar[stage], kb * TC_BK,
                                                  t.n0 + crank * kBRows, kMask, policy_b);
                            }
                        }
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                        ++requested;
                        if (prefetch > 0 && lane == 0) *progress = requested;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 6) {
        // ===================== L2 prefetch warp =====================
        // walks the producer's K-block sequence `prefetch` blocks ahead of what it has requested (cp.async.bulk.prefetch.tensor):
        // the 4-stage ring covers ~2000 cycles, less than an HBM miss under load, so the lines are brought into L2 beforehand
        if (prefetch > 0) {
            uint32_t done = 0;
            for (int cg = cluster_id; cg < num_cgroups; cg += num_clusters) {
                const int g = cg * CS + crank;
                for (int ti = 0; ti < tiles_per_group; ++ti) {
                    TcTile t = tiles[(int64_t)cg * table_stride + ti];
                    t.m0 += g * group_m_step + crank * rank_m_step;
                    for (int kb = t.kb0; kb < t.kb1; ++kb) {
                        while ((int)(done - *progress) >= prefetch) __nanosleep(64);
                        if (elect_one()) {
                            tma_prefetch_2d(&tmA, kb * TC_BK, t.m0);
                            tma_prefetch_2d(&tmB, kb * TC_BK, t.n0 + crank * kBRows);
                        }
                        ++done;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (all lanes run the loops, one elected lane issues) =====================
        {
            constexpr uint32_t idesc = KIND == TC_KIND_MXF4 ? umma_idesc_mxf4(TC_BM, TC_BN) : umma_idesc_i8(TC_BM, TC_BN);
            const uint32_t sf_a = tmem_base + TC_SF_COL + 64, sf_b = tmem_base + TC_SF_COL + 128;
            const uint64_t da0 = umma_desc_kmajor_sw128(smem_u32(smem)), db0 = umma_desc_kmajor_sw128(smem_u32(smem) + TC_A_BYTES);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            bool chained = false;                  // the previous tile left its sums in this accumulator (TC_TILE_CHAIN)
            for (int cg = cluster_id; cg < num_cgroups; cg += num_clusters) {
                for (int ti = 0; ti < tiles_per_group; ++ti) {
                    const TcTile t = tiles[(int64_t)cg * table_stride + ti];
                    if (!chained) {
                        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                        tc_fence_after();
                    }
                    const uint32_t d_tmem = tmem_base + acc * TC_BN;
                    for (int kb = t.kb0; kb < t.kb1; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (elect_one()) {
                            // the descriptor's start-address field is (addr >> 4); stages are whole multiples of 1024 B
                            const uint64_t da = da0 + (uint64_t)(stage * (TC_STAGE_BYTES >> 4));
                            const uint64_t db = db0 + (uint64_t)(stage * (TC_STAGE_BYTES >> 4));
#pragma unroll
                            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                                // advance 32 bytes of K inside the swizzle atom: +2 in the (addr >> 4) field
                                const uint64_t ka = da + (uint64_t)(k * (TC_UMMA_K >> 4)), kb_ = db + (uint64_t)(k * (TC_UMMA_K >> 4));
                                const uint32_t accum = (chained || kb > t.kb0 || k > 0) ? 1u : 0u;
                                if (KIND == TC_KIND_MXF4) umma_mxf4(d_tmem, ka, kb_, idesc, accum, sf_a, sf_b);
                                else umma_i8(d_tmem, ka, kb_, idesc, accum);
                            }
                            // frees the smem slot (in every CTA of the cluster) when these MMAs retire
                            if (CS == 1) umma_commit(&empty_bar[stage]); else umma_commit_mcast(&empty_bar[stage], kMask);
                        }
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                    }
                    chained = (t.flags & TC_TILE_CHAIN) != 0;
                    if (chained) continue;                           // the next tile adds to the same accumulator
                    if (elect_one()) umma_commit(&tfull_bar[acc]);   // accumulator complete -> epilogue
                    if (++acc == kAcc) { acc = 0; acc_phase ^= 1; }
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
        for (int cg = cluster_id; cg < num_cgroups; cg += num_clusters) {
            const int g = cg * CS + crank;
            const bool live = g < num_groups;      // ragged last cluster group: take part in the protocol, skip the output
            if (live) epi.begin_group(ep, g, row);
            for (int ti = 0; ti < tiles_per_group; ++ti) {
                TcTile t = tiles[(int64_t)cg * table_stride + ti];
                if (t.flags & TC_TILE_CHAIN) continue;               // folded into the tile that ends the chain
                t.m0 += g * group_m_step + crank * rank_m_step;
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * TC_BN + (static_cast<uint32_t>(quad * 32) << 16);
                if (live) {
                    const int nchunks = epi.tile_begin(ep, t, row);     // 32-column chunks of this tile that carry data
#pragma unroll 1
                    for (int c = 0; c < nchunks; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32(taddr + c * 32, v);
                        tmem_ld_wait();
                        epi.chunk(ep, t, row, c, v);
                    }
                    epi.tile_end(ep, t, row, lane);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                if (++acc == kAcc) { acc = 0; acc_phase ^= 1; }
            }
            if (live) epi.end_group(ep, g, row);
        }
    }

    tc_fence_before();
    if (CS > 1) cluster_sync_all(); else __syncthreads();   // no CTA leaves while a peer may still signal its barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

// ---- epilogue of the kinship Gram: G[m0+row][n0 + 32c .. +31] (+)= acc ------------------------------
template <bool F32ACC>
struct GramEpiT {
    struct Params {
        int32_t* G;        // [n_pad x ld] int32, n_pad multiple of 256 so no bounds checks are needed
        int64_t ld;
        int accumulate;
    };
    __device__ __forceinline__ void begin_group(const Params&, int, int) {}
    __device__ __forceinline__ void end_group(const Params&, int, int) {}
    __device__ __forceinline__ int tile_begin(const Params&, const TcTile&, int) { return TC_BN / 32; }
    __device__ __forceinline__ void tile_end(const Params&, const TcTile&, int, int) {}
    __device__ __forceinline__ void chunk(const Params& p, const TcTile& t, int row, int c, const uint32_t (&raw)[32]) {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = F32ACC ? (uint32_t)__float2int_rn(__uint_as_float(raw[j])) : raw[j];   // exact: integers < 2^24
        int4* dst = reinterpret_cast<int4*>(p.G + (int64_t)(t.m0 + row) * p.ld + t.n0 + c * 32);
        if (t.aux0) {
            // split-K slice of a tail tile (several clusters share one output tile): integer atomics are exact and order
            // independent; G was zeroed at reset, so the first chunk needs no special case
            int* d = reinterpret_cast<int*>(dst);
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(d + j, (int)v[j]);
        } else if (p.accumulate) {
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

using GramEpi = GramEpiT<false>;      // int8 operands, int32 accumulators
using GramEpiF4 = GramEpiT<true>;     // e2m1 operands, FP32 accumulators holding exact integers

}  // namespace mmg
