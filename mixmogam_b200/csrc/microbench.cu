// libmixmogam_b200_bench: pipe-rate microbenchmarks (include/mixmogam_b200_bench.h).  A separate shared library -- the
// product library carries no benchmark kernels; this one borrows the product's context (stream, scratch, events).
#include "common.cuh"
#include "scan_tc.cuh"
#include "../../include/mixmogam_b200_bench.h"

using namespace mmg;

__global__ void __launch_bounds__(256) bench_dmma_kernel(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma_8x8x4(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) bench_dfma_kernel(double* out, int iters) {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
    const double a = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}
__global__ void bench_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n16) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// tcgen05 int8 MMA issue rate with shared-memory-resident operands (no loads): one thread per CTA issues `iters` K blocks
// (4 x UMMA K=32, M=128 per CTA, N=256) into two alternating accumulators.  ldtm != 0: the four epilogue warps read
// the accumulators back with tcgen05.ld at the rate of one full 128x256 tile per `ldtm` K blocks, free running
// (measures whether TMEM reads take cycles from the tensor pipe).  PAIR: cta_group::2 (M = 256 over two CTAs).
// F4: e2m1 operands, kind::mxf4.block_scale with unit scale factors (K = 64 per instruction), one accumulator (the Gram's form).
template <bool PAIR, bool F4 = false>
__global__ void __launch_bounds__(192, 1) bench_imma_kernel(int iters, int ldtm, unsigned* sink, int bfull = 0) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (TC_A_BYTES + TC_B_BYTES) / 4; i += blockDim.x) {
        uint32_t h = 0x9E3779B9u * (i + 1);
        if (bfull && i >= TC_A_BYTES / 4) {
            // the scan's B operand: base-256 digits, every byte value -- the multiplier arrays toggle far more than on genotype-like
            // bytes, which is what the power-capped rate of a real int8 product depends on
            h ^= h >> 15; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
            reinterpret_cast<uint32_t*>(smem)[i] = h;
        } else {
            reinterpret_cast<uint32_t*>(smem)[i] = h & (F4 ? 0x22222222u : 0x03030303u);   // genotype-like bytes 0..3 | e2m1 0 / 1.0
        }
    }
    if (threadIdx.x == 0) {
        mbar_init(&done_bar, 1);
        mbar_fence_init();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        if (PAIR) { tmem_alloc_pair(&tmem_slot, TC_TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(&tmem_slot, TC_TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const bool leader = !PAIR || cluster_ctarank() == 0;
    if (F4) {
        if (warp >= 2) {
            const uint32_t ta = tmem_base + TC_SF_COL + (static_cast<uint32_t>((warp & 3) * 32) << 16);
            for (int c = 0; c < (TC_TMEM_COLS - TC_SF_COL) / 8; ++c) tmem_st_32x8_const(ta + c * 8, 0x7f7f7f7fu);
            tmem_st_wait();
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (F4 && warp == 1 && lane == 0) {
        constexpr uint32_t idesc4 = umma_idesc_mxf4(TC_BM, TC_BN);
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem)), db = umma_desc_kmajor_sw128(smem_u32(smem) + TC_A_BYTES);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                umma_mxf4(tmem_base, da + 2 * kk, db + 2 * kk, idesc4, ((it & 7) | kk) ? 1u : 0u, tmem_base + TC_SF_COL + 64, tmem_base + TC_SF_COL + 128);
        }
        umma_commit(&done_bar);
    }
    if (!F4 && warp == 1 && lane == 0 && leader) {
        constexpr uint32_t idesc = umma_idesc_i8(PAIR ? 2 * TC_BM : TC_BM, TC_BN);
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem)), db = umma_desc_kmajor_sw128(smem_u32(smem) + TC_A_BYTES);
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem_base + ((it >> 3) & 1) * TC_BN;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (PAIR) umma_i8_pair(d, da + 2 * kk, db + 2 * kk, idesc, ((it & 7) | kk) ? 1u : 0u);
                else umma_i8(d, da + 2 * kk, db + 2 * kk, idesc, ((it & 7) | kk) ? 1u : 0u);
            }
        }
        if (PAIR) umma_commit_pair(&done_bar, 0b11); else umma_commit(&done_bar);
    }
    if (warp >= 2 && ldtm > 0) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        unsigned acc = 0;
        const int tiles = iters / ldtm;
        for (int t = 0; t < tiles; ++t) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                uint32_t v[16];
                tmem_ld_32x16(taddr + (t & 1) * TC_BN + c * 16, v);
                tmem_ld_wait_dep(v);
#pragma unroll
                for (int j = 0; j < 16; ++j) acc += v[j];
            }
        }
        if (acc == 0x12345678u) sink[0] = acc;
    }
    if (warp == 1 || warp == 0) {
        if (lane == 0) mbar_wait(&done_bar, 0);
        __syncwarp();
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, TC_TMEM_COLS); else tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

// TMEM read-back rate of the scan's epilogue pattern: WARPS epilogue warps (WARPS / 4 per lane quadrant, each 256 * 4 / WARPS
// columns of a 128 x 256 int32 accumulator tile), tcgen05.ld.32x32b.x<LDW> double buffered with a multiply-accumulate per
// element between the waits, while (with_mma) the tensor pipe runs flat out into the other accumulator.  Reports SM cycles per tile.
template <int WARPS, int LDW>
__global__ void __launch_bounds__(64 + 32 * WARPS, 1) bench_ldtm_kernel(int tiles, int with_mma, long long* out, unsigned* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int stop_flag;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (TC_A_BYTES + TC_B_BYTES) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = (0x9E3779B9u * (i + 1)) & 0x03030303u;
    if (threadIdx.x == 0) {
        mbar_init(&done_bar, 1);
        mbar_fence_init();
        stop_flag = 0;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) { tmem_alloc(&tmem_slot, TC_TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 1 && lane == 0) {
        if (with_mma) {
            constexpr uint32_t idesc = umma_idesc_i8(TC_BM, TC_BN);
            const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem)), db = umma_desc_kmajor_sw128(smem_u32(smem) + TC_A_BYTES);
            // keep the pipe busy until the readers are done: batches of 64 K-blocks, then look at the flag
            for (int it = 0; it < (1 << 22); ++it) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_i8(tmem_base + TC_BN, da + 2 * kk, db + 2 * kk, idesc, (it | kk) ? 1u : 0u);
                if ((it & 63) == 63 && *(volatile int*)&stop_flag >= WARPS) break;
            }
        }
        umma_commit(&done_bar);
    }
    if (warp >= 2) {
        constexpr int kCols = 256 * 4 / WARPS;
        const int quad = warp & 3, part = (warp - 2) >> 2;
        const uint32_t taddr = tmem_base + part * kCols + (static_cast<uint32_t>(quad * 32) << 16);
        int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        const uint32_t xw = 0x01020100u + lane;
        const long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            uint32_t va[LDW], vb[LDW];
            if constexpr (LDW == 32) tmem_ld_32x32(taddr, va); else tmem_ld_32x16(taddr, va);
#pragma unroll
            for (int c = 0; c < kCols / LDW; c += 2) {
                tmem_ld_wait_dep(va, s0, s1, s2, s3);
                if constexpr (LDW == 32) tmem_ld_32x32(taddr + (c + 1) * LDW, vb); else tmem_ld_32x16(taddr + (c + 1) * LDW, vb);
#pragma unroll
                for (int j = 0; j < LDW; j += 4) {
                    const uint32_t w = xw + j;
                    s0 += (int)va[j + 0] * (int)(int8_t)(w & 0xffu);
                    s1 += (int)va[j + 1] * (int)(int8_t)((w >> 8) & 0xffu);
                    s2 += (int)va[j + 2] * (int)(int8_t)((w >> 16) & 0xffu);
                    s3 += (int)va[j + 3] * (int)(int8_t)(w >> 24);
                }
                tmem_ld_wait_dep(vb, s0, s1, s2, s3);
                if (c + 2 < kCols / LDW) {
                    if constexpr (LDW == 32) tmem_ld_32x32(taddr + (c + 2) * LDW, va); else tmem_ld_32x16(taddr + (c + 2) * LDW, va);
                }
#pragma unroll
                for (int j = 0; j < LDW; j += 4) {
                    const uint32_t w = xw + j + 1;
                    s0 += (int)vb[j + 0] * (int)(int8_t)(w & 0xffu);
                    s1 += (int)vb[j + 1] * (int)(int8_t)((w >> 8) & 0xffu);
                    s2 += (int)vb[j + 2] * (int)(int8_t)((w >> 16) & 0xffu);
                    s3 += (int)vb[j + 3] * (int)(int8_t)(w >> 24);
                }
            }
        }
        const long long t1 = clock64();
        if (lane == 0) {
            atomicAdd(&stop_flag, 1);
            if (warp == 2) out[blockIdx.x] = t1 - t0;
        }
        if (s0 + s1 + s2 + s3 == 0x12345678) sink[0] = 1;
    }
    if (warp == 1 || warp == 0) {
        if (lane == 0) mbar_wait(&done_bar, 0);
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

template <int WARPS, int LDW>
static int run_bench_ldtm(mmg_ctx* ctx, int with_mma, double* value) {
    const int tiles = 2000, smem = TC_A_BYTES + TC_B_BYTES + 1024, grid = ctx->sm_count;
    auto kern = bench_ldtm_kernel<WARPS, LDW>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    MMG_CUDA(ctx, cudaMemsetAsync(ctx->scratch, 0, (size_t)(grid + 2) * sizeof(long long), ctx->stream));
    long long* out = (long long*)ctx->scratch;
    kern<<<grid, 64 + 32 * WARPS, smem, ctx->stream>>>(tiles, with_mma, out, (unsigned*)(out + grid));
    ctx->launches += 1;
    MMG_TRY(launch_check(ctx, "bench_ldtm_kernel"));
    std::vector<long long> h((size_t)grid);
    MMG_CUDA(ctx, cudaMemcpyAsync(h.data(), out, (size_t)grid * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double sum = 0.0;
    for (long long v : h) sum += (double)v;
    *value = sum / grid / tiles;
    return MMG_OK;
}

extern "C" {

int mmg_microbench(mmg_ctx* ctx, const char* which, double* value) {
    MMG_CHECK(ctx, ctx && which && value, "mmg_microbench: bad argument");
    MMG_CUDA(ctx, cudaSetDevice(ctx->device));
    MMG_TRY(ensure_scratch(ctx, 1 << 20));
    float ms = 0.f;
    if (!strcmp(which, "dmma") || !strcmp(which, "dfma")) {
        const bool dm = !strcmp(which, "dmma");
        const int iters = 20000, blocks = ctx->sm_count * 4;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(ctx->kev0, ctx->stream);
            if (dm) bench_dmma_kernel<<<blocks, 256, 0, ctx->stream>>>((double*)ctx->scratch, iters);
            else bench_dfma_kernel<<<blocks, 256, 0, ctx->stream>>>((double*)ctx->scratch, iters);
            MMG_TRY(launch_check(ctx, "bench kernel"));
            cudaEventRecord(ctx->kev1, ctx->stream);
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        }
        const double flops = dm ? (double)blocks * 8 /*warps*/ * iters * 8.0 * 512.0 : (double)blocks * 256 * iters * 16.0 * 2.0;
        *value = flops / (ms * 1e-3) / 1e12;
        return MMG_OK;
    }
    if (!strcmp(which, "copy")) {
        const int64_t bytes = 2ll << 30;
        DevBuf a, b;
        MMG_CUDA(ctx, a.alloc(ctx->stream, bytes));
        MMG_CUDA(ctx, b.alloc(ctx->stream, bytes));
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(ctx->kev0, ctx->stream);
            bench_copy_kernel<<<ctx->sm_count * 16, 512, 0, ctx->stream>>>(a.as<uint4>(), b.as<uint4>(), bytes / 16);
            MMG_TRY(launch_check(ctx, "bench_copy_kernel"));
            cudaEventRecord(ctx->kev1, ctx->stream);
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        }
        *value = 2.0 * bytes / (ms * 1e-3) / 1e9;
        return MMG_OK;
    }
    if (!strncmp(which, "prepass", 7)) {
        // "prepass_r<rows>_u<unroll>_b<min blocks>": ms of the scan's linear pre-pass over the resident genotypes (tuning aid)
        MMG_CHECK(ctx, ctx->snps != nullptr, "prepass microbench: no resident genotypes");
        const int64_t npad = round_up(ctx->n, 256), cnt = ctx->m;
        DevBuf buf;
        MMG_CUDA(ctx, buf.alloc(ctx->stream, (size_t)(2 * npad + 3 * cnt) * sizeof(double)));
        MMG_CUDA(ctx, cudaMemsetAsync(buf.p, 0, (size_t)(2 * npad) * sizeof(double), ctx->stream));
        double* v = buf.as<double>();
        double* o = v + 2 * npad;
        auto run = [&](auto kern, int rows) {
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(ctx->kev0, ctx->stream);
                kern<<<(unsigned)((cnt + 8 * rows - 1) / (8 * rows)), 256, 0, ctx->stream>>>(ctx->snps, ctx->pitch, 0, cnt, 1, v, v + npad, npad, o, o + cnt,
                                                                                        o + 2 * cnt, cnt);
                cudaEventRecord(ctx->kev1, ctx->stream);
                cudaStreamSynchronize(ctx->stream);
                cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
            }
        };
        if (!strcmp(which, "prepass_r4_u4_b3")) run(snp_prepass_kernel<4, 4, 3>, 4);
        else if (!strcmp(which, "prepass_r4_u4_b1")) run(snp_prepass_kernel<4, 4, 1>, 4);
        else if (!strcmp(which, "prepass_r4_u2_b3")) run(snp_prepass_kernel<4, 2, 3>, 4);
        else if (!strcmp(which, "prepass_r2_u4_b4")) run(snp_prepass_kernel<2, 4, 4>, 2);
        else if (!strcmp(which, "prepass_r4_u8_b1")) run(snp_prepass_kernel<4, 8, 1>, 4);
        else if (!strcmp(which, "prepass_r4_u4_b4")) run(snp_prepass_kernel<4, 4, 4>, 4);
        else if (!strcmp(which, "prepass_r4_u8_b4")) run(snp_prepass_kernel<4, 8, 4>, 4);
        else if (!strcmp(which, "prepass_r8_u4_b2")) run(snp_prepass_kernel<8, 4, 2>, 8);
        else if (!strcmp(which, "prepass_r8_u2_b3")) run(snp_prepass_kernel<8, 2, 3>, 8);
        else if (!strcmp(which, "prepass_r2_u8_b4")) run(snp_prepass_kernel<2, 8, 4>, 2);
        else if (!strcmp(which, "prepass_r2_u16_b4")) run(snp_prepass_kernel<2, 16, 4>, 2);
        else if (!strcmp(which, "prepass_r1_u16_b4")) run(snp_prepass_kernel<1, 16, 4>, 1);
        else return fail(ctx, MMG_EBADARG, "unknown microbench '%s'", which);
        MMG_TRY(launch_check(ctx, "snp_prepass_kernel"));
        *value = ms;
        return MMG_OK;
    }
    if (!strncmp(which, "ldtm", 4)) {
        // "ldtm_w<4|8|16>_x<16|32>[_mma]": SM cycles per 128 x 256 int32 tile read back by the epilogue pattern
        const int w = strstr(which, "_w16") ? 16 : strstr(which, "_w8") ? 8 : 4;
        const int x = strstr(which, "_x32") ? 32 : 16;
        const int mma = strstr(which, "_mma") ? 1 : 0;
        if (w == 4 && x == 16) return run_bench_ldtm<4, 16>(ctx, mma, value);
        if (w == 4 && x == 32) return run_bench_ldtm<4, 32>(ctx, mma, value);
        if (w == 8 && x == 16) return run_bench_ldtm<8, 16>(ctx, mma, value);
        if (w == 8 && x == 32) return run_bench_ldtm<8, 32>(ctx, mma, value);
        if (w == 16 && x == 16) return run_bench_ldtm<16, 16>(ctx, mma, value);
        return run_bench_ldtm<16, 32>(ctx, mma, value);
    }
    if (!strncmp(which, "imma", 4) || !strncmp(which, "mxf4", 4)) {
        // "imma_tcgen05" | "imma_pair" | "imma_tcgen05_ldtm<k>" | "imma_pair_ldtm<k>" | "mxf4_tcgen05": TOP/s; "..._digits": the B operand
        // holds full-range bytes (the scan's digit planes) instead of genotype-like ones
        const bool f4 = !strncmp(which, "mxf4", 4);
        const bool pair = !f4 && strstr(which, "pair") != nullptr;
        const int bfull = (!f4 && strstr(which, "digits") != nullptr) ? 1 : 0;
        const char* l = strstr(which, "ldtm");
        const int ldtm = l ? std::max(1, atoi(l + 4)) : 0;
        const int iters = 40000, smem = TC_A_BYTES + TC_B_BYTES + 1024;
        const int grid = pair ? ctx->sm_count / 2 * 2 : ctx->sm_count;
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3(192);            // MMA warp + four read-back warps
        cfg.gridDim = dim3((unsigned)grid);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = pair ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaFuncSetAttribute(bench_imma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(bench_imma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(bench_imma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        auto launch = [&]() -> int {
            cudaError_t e = f4 ? cudaLaunchKernelEx(&cfg, bench_imma_kernel<false, true>, iters, 0, (unsigned*)ctx->scratch, 0)
                          : pair ? cudaLaunchKernelEx(&cfg, bench_imma_kernel<true>, iters, ldtm, (unsigned*)ctx->scratch, bfull)
                                 : cudaLaunchKernelEx(&cfg, bench_imma_kernel<false>, iters, ldtm, (unsigned*)ctx->scratch, bfull);
            ctx->launches += 1;
            if (e != cudaSuccess) return fail(ctx, MMG_ECUDA, "bench_imma_kernel launch failed: %s", cudaGetErrorString(e));
            return MMG_OK;
        };
        // burst: the second of two launches, chip still cool
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(ctx->kev0, ctx->stream);
            MMG_TRY(launch());
            cudaEventRecord(ctx->kev1, ctx->stream);
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
        }
        // "..._sustained<ms>": launches queued back to back for about <ms> milliseconds, the rate of the LAST quarter of them --
        // the steady state under the board's power cap, which is what a kernel inside a seconds-long step can reach
        if (const char* sus = strstr(which, "sustained")) {
            const double sustain_ms = std::max(50.0, atof(sus + 9));
            const int reps = std::max(8, (int)(sustain_ms / std::max(ms, 1e-3f))), timed = std::max(1, reps / 4);
            for (int rep = 0; rep < reps; ++rep) {
                if (rep == reps - timed) cudaEventRecord(ctx->kev0, ctx->stream);
                MMG_TRY(launch());
            }
            cudaEventRecord(ctx->kev1, ctx->stream);
            MMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaEventElapsedTime(&ms, ctx->kev0, ctx->kev1);
            ms /= (float)timed;
        }
        *value = 2.0 * (double)grid * iters * TC_BM * TC_BN * TC_BK * (f4 ? 2 : 1) / (ms * 1e-3) / 1e12;   // 256 e2m1 values per 128-byte K block
        return MMG_OK;
    }
    return fail(ctx, MMG_EBADARG, "unknown microbench '%s'", which);
}

}  // extern "C"
