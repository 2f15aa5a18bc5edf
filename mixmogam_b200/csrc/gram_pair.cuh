// Kinship Gram, CTA-pair form: G[256 x 256] (+)= P[rows m0..m0+255] P[rows n0..n0+255]' as ONE tcgen05.mma.cta_group::2 of
// M = 256 over the two SMs of a cluster, e2m1 operands (kind::mxf4.block_scale, unit scales -- see tc_gemm.cuh).
//
// The cluster-of-2 form of tc_gemm_i8_kernel multicasts the shared B tile: each CTA fetches half of it and the crossbar
// delivers both halves to both SMs -- 16 KB (own A) + 32 KB (whole B) arrive in every SM per 128-byte K block, 48 KB of
// shared-memory fill for 512 tensor cycles (ncu: 20 TB/s of L2 -> SM traffic, tensor pipe 79 % active).  As a pair each CTA
// stages its own A tile and only ITS half of B (the MMA reads the other half from the peer's shared memory): 32 KB per SM and
// K block, a third less fabric traffic and shared-memory fill, and the 192 KB ring holds 6 stages instead of 4.
//
// Warp roles (192 threads): warp 0 TMA producer (both CTAs, each for its own shared memory, completing on the LEADER's full
// barrier), warp 1 MMA issuer (leader CTA only) + TMEM owner, warps 2..5 epilogue (each CTA drains its own 128 rows).
// One FP32 accumulator of 256 columns per CTA; TMEM columns 256..511 hold the constant scale factors.
#pragma once
#include "tc_gemm.cuh"

namespace mmg {

constexpr int GP_STAGES = 6;
constexpr int GP_A_BYTES = TC_BM * TC_BK;                       // 16 KB: this CTA's 128 rows of the M = 256 operand
constexpr int GP_B_BYTES = (TC_BN / 2) * TC_BK;                 // 16 KB: this CTA's half of the 256-row B tile
constexpr int GP_STAGE_BYTES = GP_A_BYTES + GP_B_BYTES;         // 32 KB
constexpr int GP_THREADS = 192;
constexpr int GP_SMEM_BYTES = GP_STAGES * GP_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

// D[tmem of both CTAs, 256 x N, FP32] (+)= A[128 rows from each CTA] * B[N/2 rows from each CTA], e2m1 operands, issued by the leader
__device__ __forceinline__ void umma_mxf4_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                               uint32_t sf_a, uint32_t sf_b) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(sf_a), "r"(sf_b)
        : "memory");
}

// tiles[e]: m0 = first row of the 256-row A block, n0 = first row of the 256-row B block, K range [kb0, kb1) in 128-byte
// blocks, aux0 != 0: split-K slice (atomic accumulation).  Cluster c runs entries c, c + num_clusters, ...  Launched with a
// cluster dimension of 2 (launch attribute).
__global__ void __launch_bounds__(GP_THREADS, 1)
gram_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcTile* __restrict__ tiles,
                 int num_tiles, uint64_t policy, const GramEpiF4::Params ep) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GP_STAGES * GP_STAGE_BYTES);
    uint64_t* full_bar = bars;                         // [GP_STAGES]  leader: both CTAs' loads of the stage have landed
    uint64_t* empty_bar = bars + GP_STAGES;            // [GP_STAGES]  each CTA: the MMAs reading this stage have retired
    uint64_t* tfull_bar = bars + 2 * GP_STAGES;        // each CTA: accumulator complete
    uint64_t* tempty_bar = tfull_bar + 1;              // leader: accumulator drained by the epilogue warps of both CTAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int crank = (int)cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const bool leader = crank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < GP_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);               // one multicast commit from the leader
        }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 8);                      // four epilogue warps in each of the two CTAs
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_slot, TC_TMEM_COLS);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 2) {                                   // unit scale factors (UE8M0 0x7f) in columns 256..511 of this CTA
        const uint32_t taddr = tmem_base + TC_SF_COL + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        for (int c = 0; c < (TC_TMEM_COLS - TC_SF_COL) / 8; ++c) tmem_st_32x8_const(taddr + c * 8, 0x7f7f7f7fu);
        tmem_st_wait();
    }
    tc_fence_before();
    cluster_sync_all();                                // the leader's MMAs read the scale factors of BOTH CTAs
    tc_fence_after();

    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int e = cluster_id; e < num_tiles; e += num_clusters) {
            const TcTile t = tiles[e];
            const int rowA = t.m0 + crank * TC_BM, rowB = t.n0 + crank * (TC_BN / 2);
            for (int kb = t.kb0; kb < t.kb1; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * GP_STAGE_BYTES);     // both CTAs' A tiles and B halves
                    const uint32_t fb = mapa_u32(&full_bar[stage], 0);
                    uint8_t* sa = smem + stage * GP_STAGE_BYTES;
                    tma_load_2d_pair(sa, &tmA, fb, kb * TC_BK, rowA, policy);
                    tma_load_2d_pair(sa + GP_A_BYTES, &tmB, fb, kb * TC_BK, rowB, policy);
                }
                if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (leader) {
            constexpr uint32_t idesc = umma_idesc_mxf4(2 * TC_BM, TC_BN);
            const uint32_t sf_a = tmem_base + TC_SF_COL + 64, sf_b = tmem_base + TC_SF_COL + 128;
            const uint64_t da0 = umma_desc_kmajor_sw128(smem_u32(smem)), db0 = umma_desc_kmajor_sw128(smem_u32(smem) + GP_A_BYTES);
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int e = cluster_id; e < num_tiles; e += num_clusters) {
                const TcTile t = tiles[e];
                mbar_wait(tempty_bar, acc_phase ^ 1);
                tc_fence_after();
                for (int kb = t.kb0; kb < t.kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t da = da0 + (uint64_t)(stage * (GP_STAGE_BYTES >> 4));
                        const uint64_t db = db0 + (uint64_t)(stage * (GP_STAGE_BYTES >> 4));
#pragma unroll
                        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
                            umma_mxf4_pair(tmem_base, da + (uint64_t)(k * (TC_UMMA_K >> 4)), db + (uint64_t)(k * (TC_UMMA_K >> 4)), idesc,
                                           (kb > t.kb0 || k > 0) ? 1u : 0u, sf_a, sf_b);
                        umma_commit_pair(&empty_bar[stage], 0b11);           // frees the stage in both CTAs
                    }
                    if (++stage == GP_STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit_pair(tfull_bar, 0b11);          // accumulator complete -> both epilogues
                acc_phase ^= 1;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps 2..5 =====================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        GramEpiF4 epi;
        uint32_t acc_phase = 0;
        for (int e = cluster_id; e < num_tiles; e += num_clusters) {
            TcTile t = tiles[e];
            t.m0 += crank * TC_BM;                                           // this CTA's 128 rows of the pair's 256
            mbar_wait(tfull_bar, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < TC_BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + c * 32, v);
                tmem_ld_wait();
                epi.chunk(ep, t, row, c, v);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(tempty_bar); else mbar_arrive_cluster(mapa_u32(tempty_bar, 0));
            }
            acc_phase ^= 1;
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, TC_TMEM_COLS);
    }
}

}  // namespace mmg
