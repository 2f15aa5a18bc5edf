// Host side of the packed genotype upload (mmg_kinship_gram_i8_host): SNP-major int8 genotype codes 0..3 are packed to
// 2 bits each -- a quarter of the PCIe bytes -- by all host threads while the DMA engine moves other chunks unpacked, and
// unpacked again on the GPU (unpack2_kernel, api.cu).  The reference walks the same `snps` list chunk by chunk on the
// host too (kinship.py:29-32: `sp.array(snps[i:i+chunk], dtype='int8')`); this is that loop's data movement.
//
// Compiled by g++ (not nvcc): AVX2 through a target attribute with a runtime CPU check, scalar otherwise.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#define MMG_X86 1
#endif

namespace {

// one row: n codes -> ceil(n / 4) bytes (code j in bits 2 (j % 4) .. of byte j / 4), the rest of the dst_ld bytes zeroed.
// Returns the OR of all source bytes (anything outside 0..3 sets a bit above bit 1).
inline unsigned pack_row_scalar(const int8_t* src, int64_t n, uint8_t* dst, int64_t j0) {
    unsigned seen = 0;
    int64_t j = j0;
    for (; j + 4 <= n; j += 4) {
        const unsigned a = (uint8_t)src[j], b = (uint8_t)src[j + 1], c = (uint8_t)src[j + 2], d = (uint8_t)src[j + 3];
        seen |= a | b | c | d;
        dst[j >> 2] = (uint8_t)((a & 3u) | ((b & 3u) << 2) | ((c & 3u) << 4) | ((d & 3u) << 6));
    }
    if (j < n) {
        unsigned v = 0;
        for (int k = 0; j + k < n; ++k) {
            const unsigned a = (uint8_t)src[j + k];
            seen |= a;
            v |= (a & 3u) << (2 * k);
        }
        dst[j >> 2] = (uint8_t)v;
    }
    return seen;
}

#ifdef MMG_X86
__attribute__((target("avx2"))) unsigned pack_rows_avx2(const int8_t* src, int64_t rows, int64_t n, int64_t ld, uint8_t* dst, int64_t dst_ld) {
    const __m256i w1 = _mm256_set1_epi16(0x0401);          // bytes (1, 4): b0 + 4 b1 per 16-bit lane
    const __m256i w2 = _mm256_set1_epi32(0x00100001);      // words (1, 16): n0 + 16 n1 per 32-bit lane
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    __m256i seen = _mm256_setzero_si256();
    unsigned seen_tail = 0;
    const int64_t n4 = (n + 3) / 4;
    for (int64_t r = 0; r < rows; ++r) {
        const int8_t* s = src + r * ld;
        uint8_t* d = dst + r * dst_ld;
        int64_t j = 0;
        for (; j + 32 <= n; j += 32) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + j));
            seen = _mm256_or_si256(seen, v);
            const __m256i t16 = _mm256_maddubs_epi16(v, w1);
            const __m256i t32 = _mm256_madd_epi16(t16, w2);
            const __m256i g = _mm256_shuffle_epi8(t32, gather);
            const uint32_t lo = (uint32_t)_mm256_extract_epi32(g, 0), hi = (uint32_t)_mm256_extract_epi32(g, 4);
            const uint64_t out = (uint64_t)lo | ((uint64_t)hi << 32);
            std::memcpy(d + (j >> 2), &out, 8);
        }
        seen_tail |= pack_row_scalar(s, n, d, j);
        if (dst_ld > n4) std::memset(d + n4, 0, (size_t)(dst_ld - n4));
    }
    alignas(32) uint8_t tmp[32];
    _mm256_store_si256(reinterpret_cast<__m256i*>(tmp), seen);
    for (int k = 0; k < 32; ++k) seen_tail |= tmp[k];
    return seen_tail;
}
#endif

unsigned pack_rows_scalar(const int8_t* src, int64_t rows, int64_t n, int64_t ld, uint8_t* dst, int64_t dst_ld) {
    unsigned seen = 0;
    const int64_t n4 = (n + 3) / 4;
    for (int64_t r = 0; r < rows; ++r) {
        uint8_t* d = dst + r * dst_ld;
        seen |= pack_row_scalar(src + r * ld, n, d, 0);
        if (dst_ld > n4) std::memset(d + n4, 0, (size_t)(dst_ld - n4));
    }
    return seen;
}

unsigned pack_rows(const int8_t* src, int64_t rows, int64_t n, int64_t ld, uint8_t* dst, int64_t dst_ld) {
#ifdef MMG_X86
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return pack_rows_avx2(src, rows, n, ld, dst, dst_ld);
#endif
    return pack_rows_scalar(src, rows, n, ld, dst, dst_ld);
}

}  // namespace

// Packs rows [0, rows) of src (row stride ld) into dst (row stride dst_ld >= ceil(n / 4)) with `threads` host threads.
// Returns 0 when every code was in 0..3, 1 otherwise (dst is then not to be used).
extern "C" int mmg_host_pack2(const int8_t* src, int64_t rows, int64_t n, int64_t ld, uint8_t* dst, int64_t dst_ld, int threads) {
    if (rows <= 0) return 0;
    if (threads < 1) threads = 1;
    const int64_t per = (rows + threads - 1) / threads;
    std::vector<unsigned> seen((size_t)threads, 0u);
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) {
        const int64_t r0 = t * per, r1 = r0 + per < rows ? r0 + per : rows;
        if (r0 >= rows) break;
        pool.emplace_back([=, &seen] { seen[(size_t)t] = pack_rows(src + r0 * ld, r1 - r0, n, ld, dst + r0 * dst_ld, dst_ld); });
    }
    seen[0] = pack_rows(src, per < rows ? per : rows, n, ld, dst, dst_ld);
    for (auto& th : pool) th.join();
    unsigned all = 0;
    for (unsigned s : seen) all |= s;
    return (all & ~3u) ? 1 : 0;
}

// host threads one process should use: the cores of the box shared between the ranks torchrun started on it (LOCAL_WORLD_SIZE)
extern "C" int mmg_host_threads_default() {
    unsigned hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 4;
    const char* lws = std::getenv("LOCAL_WORLD_SIZE");
    const int ranks = lws ? std::atoi(lws) : 1;
    if (ranks > 1) hc = hc / (unsigned)ranks;
    if (hc < 1) hc = 1;
    return (int)(hc > 32 ? 32 : hc);
}
