// Host-side format conversion of the drop-in boundary (no device work, no physics): the reference keeps a lattice as
// N x N int32 +-1, column-major (definitions.hpp:16, Lattice::spins_) = 4 bytes per spin; the device wants 1 bit per
// spin.  Packing on the host before the PCIe copy cuts the transfer 32-fold (2.7 GB -> 84 MB for the headline batch).
// Transport format ("packed natural"): replica-major, then internal row y (= reference column j), then
// max(1, L/32) words per row; bit k of word w is 1 iff spin (i = 32w + k, j = y) is +1.
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/mcrg_b200.h"

namespace {

void pack_rows_scalar(const int32_t *src, uint32_t *dst, size_t n_words, int bits) {
    for (size_t q = 0; q < n_words; ++q) {
        uint32_t v = 0;
        for (int k = 0; k < bits; ++k) v |= (uint32_t)(src[q * bits + k] > 0) << k;
        dst[q] = v;
    }
}

#if defined(__x86_64__)
// The conversion streams 128 bytes of int32 per output word and is bound by how many cache-line fills one core keeps in
// flight, not by arithmetic: a software prefetch a few KB ahead (harmless past the end: prefetches never fault) and, where
// the CPU has it, AVX-512 compare-into-mask (two 64-byte loads per word, no movemask) raise the per-thread rate by ~40 %
// (measured single-threaded on a Xeon: 7.8 -> 11.3 GB/s).
__attribute__((target("avx512f"))) void pack_rows_avx512(const int32_t *src, uint32_t *dst, size_t n_words) {
    const __m512i zero = _mm512_setzero_si512();
    for (size_t q = 0; q < n_words; ++q) {
        const char *p = reinterpret_cast<const char *>(src + q * 32);
        _mm_prefetch(p + 4096, _MM_HINT_T0);
        _mm_prefetch(p + 4096 + 64, _MM_HINT_T0);
        const __mmask16 lo = _mm512_cmpgt_epi32_mask(_mm512_loadu_si512(p), zero);
        const __mmask16 hi = _mm512_cmpgt_epi32_mask(_mm512_loadu_si512(p + 64), zero);
        dst[q] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
}

__attribute__((target("avx2"))) void pack_rows_avx2(const int32_t *src, uint32_t *dst, size_t n_words) {
    const __m256i zero = _mm256_setzero_si256();
    for (size_t q = 0; q < n_words; ++q) {
        const __m256i *p = reinterpret_cast<const __m256i *>(src + q * 32);
        _mm_prefetch(reinterpret_cast<const char *>(p) + 2048, _MM_HINT_T0);
        _mm_prefetch(reinterpret_cast<const char *>(p) + 2048 + 64, _MM_HINT_T0);
        uint32_t v = 0;
        for (int k = 0; k < 4; ++k) {
            const __m256i x = _mm256_loadu_si256(p + k);
            v |= (uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_cmpgt_epi32(x, zero))) << (8 * k);
        }
        dst[q] = v;
    }
}
#endif

}  // namespace

extern "C" size_t mcrg_packed_words(int L, int count) {
    if (L < 1 || count < 0) return 0;
    return (size_t)count * L * (L >= 32 ? L / 32 : 1);
}

extern "C" int mcrg_host_pack_i32_colmajor(const int32_t *spins, int L, int count, uint32_t *packed, int n_threads) {
    if (!spins || !packed || L < 2 || (L & (L - 1)) || count < 0) return MCRG_ERR_ARG;
    const int bits = L >= 32 ? 32 : L;  // for L < 32 every row is one (partial) word
    const size_t n_words = mcrg_packed_words(L, count);
    if (n_threads < 1) n_threads = 1;
    n_threads = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(1, n_words / 4096));
    bool avx2 = false, avx512 = false;
#if defined(__x86_64__)
    avx2 = bits == 32 && __builtin_cpu_supports("avx2");
    avx512 = bits == 32 && __builtin_cpu_supports("avx512f");
    if (const char *e = getenv("MCRG_HOSTPACK_ISA")) {  // tests: force a narrower path ("avx2", "scalar")
        if (e[0] == 'a' && e[3] == '2') avx512 = false;
        if (e[0] == 's') avx512 = avx2 = false;
    }
#endif
    auto work = [&](size_t q0, size_t q1) {
#if defined(__x86_64__)
        if (avx512) {
            pack_rows_avx512(spins + q0 * 32, packed + q0, q1 - q0);
            return;
        }
        if (avx2) {
            pack_rows_avx2(spins + q0 * 32, packed + q0, q1 - q0);
            return;
        }
#endif
        pack_rows_scalar(spins + q0 * bits, packed + q0, q1 - q0, bits);
    };
    if (n_threads == 1) {
        work(0, n_words);
        return 0;
    }
    std::vector<std::thread> pool;
    const size_t per = (n_words + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const size_t q0 = std::min(n_words, (size_t)t * per), q1 = std::min(n_words, q0 + per);
        if (q0 < q1) pool.emplace_back(work, q0, q1);
    }
    for (auto &th : pool) th.join();
    return 0;
}

// Host read-bandwidth probe (bench.py's `e2e.host_roofline`): the packing above is a streaming read of the caller's int32
// lattices; this reads the same bytes with the same thread layout and no conversion, so that the packing rate can be quoted
// against what these host cores can read at all.  Returns the sum of the words (so the loads cannot be elided).
#if defined(__x86_64__)
namespace {
__attribute__((target("avx2"))) int64_t read_sum_avx2(const int32_t *src, size_t n_ints) {
    __m256i a = _mm256_setzero_si256(), b = a, c = a, d = a;
    size_t q = 0;
    for (; q + 32 <= n_ints; q += 32) {
        const __m256i *p = reinterpret_cast<const __m256i *>(src + q);
        _mm_prefetch(reinterpret_cast<const char *>(p) + 4096, _MM_HINT_T0);
        _mm_prefetch(reinterpret_cast<const char *>(p) + 4096 + 64, _MM_HINT_T0);
        a = _mm256_add_epi32(a, _mm256_loadu_si256(p));
        b = _mm256_add_epi32(b, _mm256_loadu_si256(p + 1));
        c = _mm256_add_epi32(c, _mm256_loadu_si256(p + 2));
        d = _mm256_add_epi32(d, _mm256_loadu_si256(p + 3));
    }
    a = _mm256_add_epi32(_mm256_add_epi32(a, b), _mm256_add_epi32(c, d));
    alignas(32) int32_t lanes[8];
    _mm256_store_si256(reinterpret_cast<__m256i *>(lanes), a);
    int64_t s = 0;
    for (int k = 0; k < 8; ++k) s += lanes[k];
    for (; q < n_ints; ++q) s += src[q];
    return s;
}
}  // namespace
#endif

extern "C" int mcrg_host_read_probe(const int32_t *buf, size_t n_ints, int n_threads, int64_t *sum_out) {
    if (!buf) return MCRG_ERR_ARG;
    if (n_threads < 1) n_threads = 1;
    std::vector<int64_t> part((size_t)n_threads, 0);
    auto work = [&](int t, size_t q0, size_t q1) {
#if defined(__x86_64__)
        if (__builtin_cpu_supports("avx2")) {
            part[t] = read_sum_avx2(buf + q0, q1 - q0);
            return;
        }
#endif
        int64_t s = 0;
        for (size_t q = q0; q < q1; ++q) s += buf[q];
        part[t] = s;
    };
    std::vector<std::thread> pool;
    const size_t per = ((n_ints + n_threads - 1) / n_threads + 31) & ~(size_t)31;
    for (int t = 0; t < n_threads; ++t) {
        const size_t q0 = std::min(n_ints, (size_t)t * per), q1 = std::min(n_ints, q0 + per);
        if (q0 < q1) pool.emplace_back(work, t, q0, q1);
    }
    for (auto &th : pool) th.join();
    int64_t s = 0;
    for (int64_t v : part) s += v;
    if (sum_out) *sum_out = s;
    return 0;
}
