// Stage 3, int8 tensor-core scan, genotype-stationary schedule.
//
// Same arithmetic as QuadEpi in scan_tc.cuh (x~.x~ = x'(R'R)x on exact base-256 digit planes of the lower-triangular
// B = c o R'R 2^-E, x~.y~ = x.(R'y~), then RSS / F / p, linear_models.py:1315-1349), different data movement.
//
// The table-driven kernel re-reads the 128-SNP genotype block (128 x n bytes = 1.3 MB at n = 10k) from L2 once per
// (digit plane, column tile): 148 CTAs x 1.3 MB does not fit the 126 MB L2, so the genotypes are streamed from HBM
// ~280 times (ncu: 165 GB of DRAM reads for 1.3 GB of genotypes) and every 128-byte K block costs 16 KB (genotypes)
// + 32 KB (digits) of L2 -> SM traffic, which is what bounds it (LTS cap ~6300 B/clk).
//
// Here the contraction index (individuals) is cut into panels of PKB K-blocks (PKB x 128 individuals).  A CTA keeps
// the genotype panel [128 SNPs x PKB*128 B] RESIDENT in shared memory and streams only digit-plane tiles past it:
//
//   for phenotype t:  for panel kp:  load genotype panel (once: the genotypes are read from HBM exactly once per phenotype)
//       for column tile jb >= panel (B is lower triangular):   x[s][256 jb .. +255] -> registers (epilogue threads)
//           for digit plane k:   acc[128 x 256] = panel . B_k[jb tile, panel]'   (<= PKB K-blocks of tcgen05.mma)
//                                q_s += w_k sum_j acc[s][j] x[s][j]              (the epilogue is linear in acc, so
//                                                                                  partial K sums fold in directly)
//
// L2 -> SM traffic per K block drops from 16 + 32/CS KB to 32/CS KB (CS = cluster size; the digit tiles are
// TMA-multicast to the CS CTAs of a cluster, which hold CS different SNP blocks), and all CTAs sweep the same digit
// stream in near lockstep so HBM sees it about once per wave.
//
// PAIR = true (the default schedule) runs the MMA as a CTA pair (tcgen05.mma.cta_group::2, M = 256 SNPs over the two SMs
// of a TPC): each CTA stages its own 128-SNP genotype panel and HALF of every digit tile (128 of the 256 columns), the
// leader CTA issues one MMA for both.  Shared-memory traffic per SM and K block drops from 32 KB written + 48 KB read to
// 16 + 32 KB and the L2 -> SM traffic halves without multicast (the single-CTA form sits on the L2 output cap).
//
// Warp roles (352 threads): warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), warp 2 L2 prefetcher (walks the digit
// stream `prefetch` K-blocks ahead of the producer's published progress), warps 3..10 epilogue (two per TMEM lane quadrant).
// The linear terms of the statistic (x.v, the FP64 diagonal of the quadratic form, ||x||_1) come from snp_prepass_kernel
// (scan_tc.cuh): the epilogue only drains accumulators.
#pragma once
#include "scan_tc.cuh"

namespace mmg {

struct QuadShape {
    int num_groups;     // 128-SNP row blocks
    int T;              // phenotypes
    int S;              // digit planes used per phenotype
    int S_stride;       // digit planes allocated per phenotype (plane k of phenotype t starts at row (t S_stride + k) n_padN)
    int tiles_n;        // BN-column tiles of B (n_padN / BN)
    int kb_total;       // 128-byte K blocks (ldq / 128)
    int n_padN;         // rows per digit plane
    int prefetch;       // L2 prefetch distance in K blocks (0 = off)
    int pf_share;       // every pf_share-th K block is prefetched by this cluster (all clusters sweep the same digit stream in
                        // lockstep, so they can split the prefetch stream between them); 1 = every cluster prefetches everything
    long long* dbg;     // nullptr, or [grid x 16] cycle counters of the three roles (MMG_SCAN_DBG_CLOCKS)
    unsigned* wave_sync;  // nullptr, or a zeroed counter: CTAs start each wave of SNP groups together (see below)
};

// mbarrier wait that adds the cycles spent waiting to *acc when counters are requested
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, bool timed, long long& acc) {
    if (timed) {
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        acc += clock64() - t0;
    } else {
        mbar_wait(bar, parity);
    }
}

template <int PKB, int STAGES, bool PAIR = false, int BN = TC_BN>
struct QuadSmem {
    static constexpr int kBTile = BN * TC_BK;                              // one digit K-block of a BN-column tile
    static constexpr int kAccStages = TC_TMEM_COLS / BN;                   // 2 (BN = 256) or 4 (BN = 128)
    static constexpr int kBStage = PAIR ? kBTile / 2 : kBTile;
    static constexpr int kABytes = PKB * TC_A_BYTES;
    static constexpr int kBBytes = STAGES * kBStage;
    static constexpr int kBars = 2 * STAGES + 2 * PKB + 2 * kAccStages;
    static_assert(kBars + 1 <= 64, "barrier block");
    static constexpr int kBytes = kABytes + kBBytes + 1024 /*align slack*/ + 512 /*barriers + tmem slot*/ + 1024 /*q, xy exchange*/;
    static_assert(kBytes <= 232448, "shared memory budget (227 KB)");
};

constexpr int QP_EPI_WARPS = 8;
constexpr int QP_FIRST_EPI_WARP = 3;
constexpr int QP_THREADS = 32 * (QP_FIRST_EPI_WARP + QP_EPI_WARPS);   // producer warp, MMA warp, L2-prefetch warp, 8 epilogue warps

// LDW = columns per tcgen05.ld of the epilogue (16 or 32), double buffered either way.  TIMED = per-role cycle counters
// (MMG_SCAN_DBG_CLOCKS) compiled in; the production instance carries none of the clock reads (predicated-off CS2R lines showed up
// as stall sites in the ncu source view of the run-time-flag version).
template <int CS, int PKB, int STAGES, bool PAIR, int BN, int LDW = 16, bool TIMED = false>
__global__ void __launch_bounds__(QP_THREADS, 1)
scan_quad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const QuadShape sh,
                 uint64_t policy_a, uint64_t policy_b, const QuadEpi::Params ep) {
    static_assert(PKB <= 12, "per-tile integer sums: |acc| <= 128 PKB |x| 128 (|x| <= 8) must leave room for 32 columns x |x| in int32");
    static_assert(!PAIR || CS == 2, "a CTA pair is a cluster of 2");
    static_assert(BN == 256 || BN == 128, "column tile");
    static_assert(PKB % (BN / TC_BK) == 0, "panel = whole column tiles");
    using SM = QuadSmem<PKB, STAGES, PAIR, BN>;
    constexpr int kBStage = SM::kBStage;
    constexpr int kBTile = SM::kBTile;
    constexpr int kAcc = SM::kAccStages;               // accumulator stages in the 512 TMEM columns
    constexpr int kKbPerTile = BN / TC_BK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smA = smem;                              // PKB genotype K-blocks  [128 x 128 B], 128B swizzle
    uint8_t* smB = smem + SM::kABytes;                // STAGES digit K-blocks  [256 x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kABytes + SM::kBBytes);
    uint64_t* full_bar = bars;                        // [STAGES]  B block landed
    uint64_t* empty_bar = full_bar + STAGES;          // [STAGES]  B block consumed by every CTA of the cluster
    uint64_t* afull_bar = empty_bar + STAGES;         // [PKB]     genotype K-block landed
    uint64_t* aempty_bar = afull_bar + PKB;           // [PKB]     last MMA of the panel on this K-block retired
    uint64_t* tfull_bar = aempty_bar + PKB;           // [2]       accumulator complete
    uint64_t* tempty_bar = tfull_bar + kAcc;          // [kAcc]    accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + kAcc);
    volatile uint32_t* progress = tmem_slot + 1;       // digit K-blocks the producer has requested so far (read by the prefetch warp)
    double* xchg = reinterpret_cast<double*>(bars + 64);    // [128] q, then xy, of the upper column half (1 KB)

    // warp index through a shuffle: the compiler then knows the role branches are warp-uniform, keeps loop counters,
    // addresses and descriptors in uniform registers and issues UTCIMMA / UTMALDG without per-lane ELECT loops
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const int crank = (CS > 1) ? (int)cluster_ctarank() : 0;
    const int cluster_id = blockIdx.x / CS;
    const int num_clusters = gridDim.x / CS;
    const int num_cgroups = (sh.num_groups + CS - 1) / CS;
    constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1u);
    constexpr int kBRows = BN / CS;                   // digit-tile rows this CTA fetches
    const bool leader = !PAIR || crank == 0;

    if (warp == 0 && lane == 0) {
        *progress = 0u;
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], PAIR ? 1 : CS);  // pair: one multicast commit; multicast: one commit per CTA
        }
        for (int i = 0; i < PKB; ++i) {
            mbar_init(&afull_bar[i], 1);
            mbar_init(&aempty_bar[i], 1);
        }
        for (int a = 0; a < kAcc; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], PAIR ? 2 * QP_EPI_WARPS : QP_EPI_WARPS);  // pair: both CTAs' epilogue warps release the leader's MMA
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        if (PAIR) {
            tmem_alloc_pair(tmem_slot, TC_TMEM_COLS);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_slot, TC_TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (CS > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (all lanes run the loops, one elected lane issues) =====================
        {
            const bool one = lane == 0;
            int stage = 0;
            uint32_t phase = 0;
            uint32_t abits = 0;                        // per K-block parity of the aempty barriers
            constexpr bool timed = TIMED;
            long long w_aempty = 0, w_empty = 0;
            const long long t_start = timed ? clock64() : 0;
            uint32_t requested = 0;                     // digit K-blocks requested so far, published for the prefetch warp
            unsigned wave_target = 0;
            int wave = 0;
            uint32_t pre = 0;                           // genotype K-blocks of the coming panel that are already requested
            auto issue_a = [&](int i, int kb0, int row0) {
                mbar_wait_timed(&aempty_bar[i], ((abits >> i) & 1u) ^ 1u, timed, w_aempty);
                abits ^= 1u << i;
                if (elect_one()) {
                    if (PAIR) {
                        if (leader) mbar_expect_tx(&afull_bar[i], 2 * TC_A_BYTES);
                        tma_load_2d_pair(smA + i * TC_A_BYTES, &tmA, mapa_u32(&afull_bar[i], 0), (kb0 + i) * TC_BK, row0, policy_a);
                    } else {
                        mbar_expect_tx(&afull_bar[i], TC_A_BYTES);
                        tma_load_2d(smA + i * TC_A_BYTES, &tmA, &afull_bar[i], (kb0 + i) * TC_BK, row0, policy_a);
                    }
                }
            };
            for (int cg = cluster_id; cg < num_cgroups; cg += num_clusters, ++wave) {
                const int m0 = (cg * CS + crank) * TC_BM;
                if (sh.wave_sync != nullptr) {
                    // Every CTA sweeps the same digit-plane stream once per SNP group.  Left alone the CTAs drift apart over
                    // the ~50 waves of a 1M-SNP scan until the stream (0.26 GB) is re-read from HBM by each of them
                    // (ncu: 474 GB of DRAM reads); starting each wave together keeps the stream L2-resident: one HBM
                    // pass per wave.  All CTAs are co-resident (persistent grid <= SM count), so the spin cannot deadlock.
                    const int clusters_in_wave = min(num_clusters, num_cgroups - wave * num_clusters);
                    wave_target += (unsigned)(clusters_in_wave * CS);
                    if (elect_one()) {
                        atomicAdd(sh.wave_sync, 1u);
                        unsigned seen;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(sh.wave_sync) : "memory");
                            if (seen < wave_target) __nanosleep(200);
                        } while (seen < wave_target);
                    }
                    __syncwarp();
                }
                for (int t = 0; t < sh.T; ++t) {
                    for (int kbase = 0; kbase < sh.kb_total; kbase += PKB) {
                        const int nka = min(PKB, sh.kb_total - kbase);
                        for (int i = 0; i < nka; ++i)
                            if (!((pre >> i) & 1u)) issue_a(i, kbase, m0);             // not loaded ahead (first panel of the launch)
                        pre = 0;
                        // the panel after this one (next panel / phenotype / SNP group of this CTA): its genotype K-blocks are
                        // requested while the LAST tile of this panel is still being multiplied, K-block by K-block as the
                        // MMA retires them, so a panel switch does not drain the pipeline
                        int nk = kbase + PKB, ncg = cg;
                        if (nk >= sh.kb_total) {
                            nk = 0;
                            if (t + 1 >= sh.T) ncg = cg + num_clusters;
                        }
                        const bool has_next = ncg < num_cgroups;
                        const int nka_next = has_next ? min(PKB, sh.kb_total - nk) : 0;
                        const int m0_next = (ncg * CS + crank) * TC_BM;
                        for (int jb = kbase / kKbPerTile; jb < sh.tiles_n; ++jb) {
                            const int nkb = min(nka, kKbPerTile * (jb + 1) - kbase);
                            for (int k = 0; k < sh.S; ++k) {
                                const int rowB = (t * sh.S_stride + k) * sh.n_padN + jb * BN + crank * kBRows;
                                const bool last_tile = (jb == sh.tiles_n - 1) && (k == sh.S - 1);
                                for (int i = 0; i < nkb; ++i) {
                                    mbar_wait_timed(&empty_bar[stage], phase ^ 1, timed, w_empty);
                                    if (elect_one()) {
                                        if (PAIR) {
                                            if (leader) mbar_expect_tx(&full_bar[stage], kBTile);      // both halves
                                            tma_load_2d_pair(smB + stage * kBStage, &tmB, mapa_u32(&full_bar[stage], 0), (kbase + i) * TC_BK,
                                                             rowB, policy_b);
                                        } else {
                                            mbar_expect_tx(&full_bar[stage], kBTile);
                                            uint8_t* sb = smB + stage * kBTile + crank * kBRows * TC_BK;
                                            if (CS == 1)
                                                tma_load_2d(sb, &tmB, &full_bar[stage], (kbase + i) * TC_BK, rowB, policy_b);
                                            else
                                                tma_load_2d_mcast(sb, &tmB, &full_bar[stage], (kbase + i) * TC_BK, rowB, kMask, policy_b);
                                        }
                                    }
                                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                                    ++requested;
                                    if (sh.prefetch && one) *progress = requested;
                                    // the slot just refilled was freed by the MMA of K-block i - STAGES of this tile, which also
                                    // released genotype K-block i - STAGES of the panel: reload it for the next panel now
                                    if (last_tile && i >= STAGES && i - STAGES < nka_next) {
                                        issue_a(i - STAGES, nk, m0_next);
                                        pre |= 1u << (i - STAGES);
                                    }
                                }
                            }
                        }
                        for (int i = max(0, nka - STAGES); i < nka_next; ++i)
                            if (!((pre >> i) & 1u)) {
                                issue_a(i, nk, m0_next);
                                pre |= 1u << i;
                            }
                    }
                }
            }
            if (timed && one) {
                long long* d = sh.dbg + (int64_t)blockIdx.x * 16;
                d[0] = clock64() - t_start;
                d[1] = w_empty;
                d[2] = w_aempty;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // ===================== MMA issuer (all lanes run the loops, one elected lane issues) =====================
        if (leader) {
            const bool one = lane == 0;
            constexpr uint32_t idesc = umma_idesc_i8(PAIR ? 2 * TC_BM : TC_BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            uint32_t abits = 0;                        // per K-block parity of the afull barriers
            constexpr bool timed = TIMED;
            long long w_tempty = 0, w_afull = 0, w_full = 0;
            const long long t_start = timed ? clock64() : 0;
            const uint64_t da0 = umma_desc_kmajor_sw128(smem_u32(smA)), db0 = umma_desc_kmajor_sw128(smem_u32(smB));
            for (int cg = cluster_id; cg < num_cgroups; cg += num_clusters) {
                for (int t = 0; t < sh.T; ++t) {
                    for (int kbase = 0; kbase < sh.kb_total; kbase += PKB) {
                        const int nka = min(PKB, sh.kb_total - kbase);
                        uint32_t seen = 0;
                        for (int jb = kbase / kKbPerTile; jb < sh.tiles_n; ++jb) {
                            const int nkb = min(nka, kKbPerTile * (jb + 1) - kbase);
                            for (int k = 0; k < sh.S; ++k) {
                                const bool last_tile = (jb == sh.tiles_n - 1) && (k == sh.S - 1);
                                mbar_wait_timed(&tempty_bar[acc], acc_phase ^ 1, timed, w_tempty);
                                tc_fence_after();
                                const uint32_t d_tmem = tmem_base + acc * BN;
                                for (int i = 0; i < nkb; ++i) {
                                    if (!((seen >> i) & 1u)) {             // first use of this genotype K-block in the panel
                                        mbar_wait_timed(&afull_bar[i], (abits >> i) & 1u, timed, w_afull);
                                        abits ^= 1u << i;
                                        seen |= 1u << i;
                                    }
                                    mbar_wait_timed(&full_bar[stage], phase, timed, w_full);
                                    tc_fence_after();
                                    if (elect_one()) {
                                        // descriptor start-address field is (addr >> 4): K-blocks and stages are whole multiples
                                        const uint64_t da = da0 + (uint64_t)(i * (TC_A_BYTES >> 4));
                                        const uint64_t db = db0 + (uint64_t)(stage * (kBStage >> 4));
#pragma unroll
                                        for (int kk = 0; kk < TC_BK / TC_UMMA_K; ++kk) {
                                            const uint64_t ka = da + (uint64_t)(kk * (TC_UMMA_K >> 4)), kb = db + (uint64_t)(kk * (TC_UMMA_K >> 4));
                                            if (PAIR) umma_i8_pair(d_tmem, ka, kb, idesc, (i > 0 || kk > 0) ? 1u : 0u);
                                            else umma_i8(d_tmem, ka, kb, idesc, (i > 0 || kk > 0) ? 1u : 0u);
                                        }
                                        // frees the digit stage (in every CTA that holds a copy / a half) when these MMAs retire
                                        if (PAIR) umma_commit_pair(&empty_bar[stage], 0b11);
                                        else if (CS == 1) umma_commit(&empty_bar[stage]);
                                        else umma_commit_mcast(&empty_bar[stage], kMask);
                                        if (last_tile) {                          // the panel's K-block i may be refilled
                                            if (PAIR) umma_commit_pair(&aempty_bar[i], 0b11); else umma_commit(&aempty_bar[i]);
                                        }
                                    }
                                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                                }
                                if (elect_one()) {
                                    if (PAIR) umma_commit_pair(&tfull_bar[acc], 0b11); else umma_commit(&tfull_bar[acc]);
                                }
                                if (++acc == kAcc) { acc = 0; acc_phase ^= 1; }
                            }
                        }
                    }
                }
            }
            if (timed && one) {
                long long* d = sh.dbg + (int64_t)blockIdx.x * 16;
                d[4] = clock64() - t_start;
                d[5] = w_full;
                d[6] = w_tempty;
                d[7] = w_afull;
            }
        }
        __syncwarp();
    } else if (warp == 2) {
        // ===================== L2 prefetch warp =====================
        // Walks the same digit-plane stream as the producer, `sh.prefetch` K-blocks ahead of it (cp.async.bulk.prefetch.tensor
        // into L2).  In its own warp: inside the producer loop the extra ~45 instructions per K-block made the single-thread
        // issue chain (~700 cycles per K-block) slower than the tensor pipe (512) -- the MMA spent half its time waiting for
        // tiles that had not been requested yet.
        if (sh.prefetch) {
            uint32_t done = 0;
            const int rota = cluster_id % sh.pf_share;
            int pf_count = rota;
            for (int cg = cluster_id; cg < num_cgroups; cg += num_clusters) {
                for (int t = 0; t < sh.T; ++t) {
                    for (int kbase = 0; kbase < sh.kb_total; kbase += PKB) {
                        const int nka = min(PKB, sh.kb_total - kbase);
                        for (int jb = kbase / kKbPerTile; jb < sh.tiles_n; ++jb) {
                            const int nkb = min(nka, kKbPerTile * (jb + 1) - kbase);
                            for (int k = 0; k < sh.S; ++k) {
                                const int rowB = (t * sh.S_stride + k) * sh.n_padN + jb * BN + crank * kBRows;
                                for (int i = 0; i < nkb; ++i) {
                                    // stay at most sh.prefetch K-blocks ahead of what the producer has requested
                                    while ((int)(done - *progress) >= sh.prefetch) __nanosleep(64);
                                    if (pf_count == 0) {
                                        if (elect_one()) tma_prefetch_2d(&tmB, (kbase + i) * TC_BK, rowB);
                                    }
                                    if (++pf_count == sh.pf_share) pf_count = 0;
                                    ++done;
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps 3..10 =====================
        // two warps per TMEM lane quadrant (a warp may only read lanes 32 (warp % 4) .. +31): each takes 128 of the tile's
        // 256 columns, so every SM sub-partition runs two epilogue warps that hide each other's tcgen05.ld / IMAD latency
        constexpr bool timed = TIMED;
        long long w_tfull = 0, w_x = 0, w_fp = 0, w_drain = 0;      // cycles: waiting for accumulators | x register loads | FP64 x.v pass | drain
        const long long t_start = timed ? clock64() : 0;
        const int quad = warp & 3;
        const int half = (warp - QP_FIRST_EPI_WARP) >> 2;       // 0: columns 0..127 of a tile, 1: columns 128..255
        const int row = quad * 32 + lane;
        constexpr int kCols = BN / 2;
        auto load_x = [](const int8_t* xrow, int col, uint32_t (&dst)[kCols / 4]) {
            if (xrow != nullptr) {
                const uint4* xp = reinterpret_cast<const uint4*>(xrow + col);
#pragma unroll
                for (int u = 0; u < kCols / 16; ++u) {
                    const uint4 w = __ldg(xp + u);
                    dst[4 * u + 0] = w.x; dst[4 * u + 1] = w.y; dst[4 * u + 2] = w.z; dst[4 * u + 3] = w.w;
                }
            } else {
#pragma unroll
                for (int u = 0; u < kCols / 4; ++u) dst[u] = 0u;
            }
        };
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int cg = cluster_id; cg < num_cgroups; cg += num_clusters) {
            const int g = cg * CS + crank;
            const int64_t orow = (int64_t)g * TC_BM + row;
            const int8_t* xrow = (g < sh.num_groups && orow < ep.row_count) ? ep.snps + (ep.row_begin + orow) * ep.pitch : nullptr;
            for (int t = 0; t < sh.T; ++t) {
                double q = 0.0;
                uint32_t xn[kCols / 4];
                bool have_next = false;
                for (int kbase = 0; kbase < sh.kb_total; kbase += PKB) {
                    for (int jb = kbase / kKbPerTile; jb < sh.tiles_n; ++jb) {
                        const int col0 = jb * BN + half * kCols;
                        const long long tc0 = timed ? clock64() : 0;
                        const long long tc1 = timed ? clock64() : 0;
                        // this SNP's genotypes at the warp's 128 columns, reused by the S digit planes; the loads for the
                        // next column tile are issued now and land while this tile's S accumulators are drained
                        uint32_t xr[kCols / 4];
                        if (have_next) {
#pragma unroll
                            for (int u = 0; u < kCols / 4; ++u) xr[u] = xn[u];
                        } else {
                            load_x(xrow, col0, xr);
                        }
                        {
                            int nk = kbase, nj = jb + 1;
                            if (nj >= sh.tiles_n) { nk = kbase + PKB; nj = nk / kKbPerTile; }
                            have_next = nk < sh.kb_total;
                            if (have_next) load_x(xrow, nj * BN + half * kCols, xn);
                        }
                        if (timed) {
                            const long long tc2 = clock64();
                            w_fp += tc1 - tc0;
                            w_x += tc2 - tc1;
                        }
                        for (int k = 0; k < sh.S; ++k) {
                            // keep the packed bytes opaque per digit plane: otherwise the sign-extended genotypes are
                            // hoisted out of this loop and spill
#pragma unroll
                            for (int u = 0; u < kCols / 4; ++u) asm volatile("" : "+r"(xr[u]));
                            mbar_wait_timed(&tfull_bar[acc], acc_phase, timed, w_tfull);
                            tc_fence_after();
                            const long long td0 = timed ? clock64() : 0;
                            const uint32_t taddr = tmem_base + acc * BN + half * kCols + (static_cast<uint32_t>(quad * 32) << 16);
                            int s0 = 0, s1 = 0, s2 = 0, s3 = 0;                // four chains
                            if constexpr (LDW == 32) {
                                uint32_t va[32], vb[32];
                                tmem_ld_32x32(taddr, va);
#pragma unroll
                                for (int c = 0; c < kCols / 32; c += 2) {      // 32-column chunks, loads double buffered
                                    tmem_ld_wait_dep(va, s0, s1, s2, s3);      // after the arithmetic on the previous chunk
                                    tmem_ld_32x32(taddr + (c + 1) * 32, vb);
#pragma unroll
                                    for (int j = 0; j < 32; j += 4) {
                                        const uint32_t w = xr[8 * c + (j >> 2)];
                                        s0 += (int)va[j + 0] * (int)(int8_t)(w & 0xffu);
                                        s1 += (int)va[j + 1] * (int)(int8_t)((w >> 8) & 0xffu);
                                        s2 += (int)va[j + 2] * (int)(int8_t)((w >> 16) & 0xffu);
                                        s3 += (int)va[j + 3] * (int)(int8_t)(w >> 24);
                                    }
                                    tmem_ld_wait_dep(vb, s0, s1, s2, s3);
                                    if (c + 2 < kCols / 32) tmem_ld_32x32(taddr + (c + 2) * 32, va);
#pragma unroll
                                    for (int j = 0; j < 32; j += 4) {
                                        const uint32_t w = xr[8 * (c + 1) + (j >> 2)];
                                        s0 += (int)vb[j + 0] * (int)(int8_t)(w & 0xffu);
                                        s1 += (int)vb[j + 1] * (int)(int8_t)((w >> 8) & 0xffu);
                                        s2 += (int)vb[j + 2] * (int)(int8_t)((w >> 16) & 0xffu);
                                        s3 += (int)vb[j + 3] * (int)(int8_t)(w >> 24);
                                    }
                                }
                            } else {
                            uint32_t va[16], vb[16];
                            tmem_ld_32x16(taddr, va);
#pragma unroll
                            for (int c = 0; c < kCols / 16; c += 2) {          // 16-column chunks, loads double buffered
                                tmem_ld_wait_dep(va, s0, s1, s2, s3);          // after the arithmetic on the previous chunk
                                tmem_ld_32x16(taddr + (c + 1) * 16, vb);
#pragma unroll
                                for (int j = 0; j < 16; j += 4) {
                                    const uint32_t w = xr[4 * c + (j >> 2)];
                                    s0 += (int)va[j + 0] * (int)(int8_t)(w & 0xffu);
                                    s1 += (int)va[j + 1] * (int)(int8_t)((w >> 8) & 0xffu);
                                    s2 += (int)va[j + 2] * (int)(int8_t)((w >> 16) & 0xffu);
                                    s3 += (int)va[j + 3] * (int)(int8_t)(w >> 24);
                                }
                                tmem_ld_wait_dep(vb, s0, s1, s2, s3);
                                if (c + 2 < kCols / 16) tmem_ld_32x16(taddr + (c + 2) * 16, va);
#pragma unroll
                                for (int j = 0; j < 16; j += 4) {
                                    const uint32_t w = xr[4 * (c + 1) + (j >> 2)];
                                    s0 += (int)vb[j + 0] * (int)(int8_t)(w & 0xffu);
                                    s1 += (int)vb[j + 1] * (int)(int8_t)((w >> 8) & 0xffu);
                                    s2 += (int)vb[j + 2] * (int)(int8_t)((w >> 16) & 0xffu);
                                    s3 += (int)vb[j + 3] * (int)(int8_t)(w >> 24);
                                }
                            }
                            }
                            // |s_k| <= 32 columns x |acc| x |x|, |acc| <= 128 PKB K x |x| x 128 (base-256 digits): 2^5 2^20 2^3 = 2^28 at PKB = 8, |x| <= 8
                            // (QS_MAX_ABS_GENOTYPE, enforced on the host): exact in int32, and so is the sum of the four chains
                            // (< 2^30; 1.5 2^30 at PKB = 12).  ONE int -> double conversion per tile: the ncu source view of the
                            // 1M-SNP scan showed a third of all warp stalls (stall_math) on the four conversions + three DADDs
                            // that used to stand here
                            const double qt = (double)((s0 + s1) + (s2 + s3));
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) {
                                if (PAIR && crank != 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[acc], 0)); else mbar_arrive(&tempty_bar[acc]);
                            }
                            if (++acc == kAcc) { acc = 0; acc_phase ^= 1; }
                            q = fma(ep.w[k], qt, q);
                            if (timed) w_drain += clock64() - td0;
                        }
                    }
                }
                // fold the two column halves of each SNP, then RSS / F / p by the lower half's thread; x.v, the FP64
                // diagonal of the quadratic form and ||x||_1 come from the linear pre-pass (snp_prepass_kernel)
                if (half == 1) xchg[row] = q;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (half == 0) q += xchg[row];
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (half == 0 && xrow != nullptr) {
                    const int64_t po = (int64_t)t * ep.pre_stride + orow;
                    QuadEpi::store(ep, t, orow, q, ep.pre_xy[po], ep.pre_qd[po], ep.pre_a1[orow]);
                }
            }
        }
        if (timed && warp == QP_FIRST_EPI_WARP && lane == 0) {
            long long* d = sh.dbg + (int64_t)blockIdx.x * 16;
            d[8] = clock64() - t_start;
            d[9] = w_tfull;
            d[10] = w_x;
            d[11] = w_fp;
            d[12] = w_drain;
        }
    }

    tc_fence_before();
    if (CS > 1) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, TC_TMEM_COLS); else tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

}  // namespace mmg
