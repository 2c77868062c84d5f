// host_pack.cpp — see host_pack.h.  Compiled by g++ (no CUDA): the AVX2 body is selected at run time.
#include "host_pack.h"

#include <immintrin.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

namespace skb {

namespace {

// code of one byte: A/a C/c G/g T/t -> 0..3, anything else 0 (= kmer_bits.cuh::pack4 on the device)
struct CodeTable {
    uint8_t t[256];
    CodeTable() {
        memset(t, 0, sizeof t);
        t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3;
    }
};
const CodeTable g_code;

inline uint32_t pack_word_scalar(const uint8_t* p, size_t n) {      // n <= 16 bases
    uint32_t w = 0;
    for (size_t i = 0; i < n; i++) w |= (uint32_t)g_code.t[p[i]] << (30 - 2 * i);
    return w;
}

void pack_scalar(const uint8_t* src, size_t n_bases, uint32_t* words) {
    size_t i = 0, w = 0;
    for (; i + 16 <= n_bases; i += 16, w++) {
        const uint8_t* p = src + i;
        uint32_t x = 0;
        for (int j = 0; j < 16; j++) x = (x << 2) | g_code.t[p[j]];
        words[w] = x;
    }
    if (i < n_bases) words[w] = pack_word_scalar(src + i, n_bases - i);
}

// The input is read once, front to back.  The hardware prefetchers stop at every 4 KB page boundary (and under a
// hypervisor the first touch of a page costs a nested page walk), which held one thread to ~6 GB/s; asking for the
// lines two pages ahead lifts that to ~10 GB/s.  Prefetches never fault, so running past the end of the buffer is safe.
#ifndef SKB_PACK_PREFETCH
#define SKB_PACK_PREFETCH 6144
#endif
inline void prefetch_ahead(const uint8_t* p) {
    _mm_prefetch((const char*)(p + SKB_PACK_PREFETCH), _MM_HINT_T0);
    _mm_prefetch((const char*)(p + SKB_PACK_PREFETCH + 64), _MM_HINT_T0);
}

// 32 ASCII bytes -> eight 32-bit lanes, lane i = the 8-bit code of bases 4i..4i+3 (first base in bits 7..6)
__attribute__((target("avx2"))) inline __m256i codes_of_32(__m256i v) {
    const __m256i nib = _mm256_set1_epi8(0x0F);
    // low nibble 1/3/7 (A C G) wants a high nibble of 4 or 6, low nibble 4 (T) wants 5 or 7
    const __m256i lut_lo = _mm256_setr_epi8(0, 1, 0, 1, 2, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 2, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i lut_hi = _mm256_setr_epi8(0, 0, 0, 0, 1, 2, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 1, 2, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i lut_code = _mm256_setr_epi8(0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i lo = _mm256_and_si256(v, nib);
    const __m256i hi = _mm256_and_si256(_mm256_srli_epi16(v, 4), nib);
    const __m256i ok = _mm256_and_si256(_mm256_shuffle_epi8(lut_lo, lo), _mm256_shuffle_epi8(lut_hi, hi));
    const __m256i bad = _mm256_cmpeq_epi8(ok, _mm256_setzero_si256());
    const __m256i code = _mm256_andnot_si256(bad, _mm256_shuffle_epi8(lut_code, lo));
    const __m256i pairs = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0104));      // c0 * 4 + c1
    return _mm256_madd_epi16(pairs, _mm256_set1_epi32(0x00010010));                   // p0 * 16 + p1
}

template <bool STREAM>
__attribute__((target("avx2"))) void pack_avx2_t(const uint8_t* src, size_t n_bases, uint32_t* words) {
    const __m256i rev4 = _mm256_setr_epi8(3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12, 3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12);
    const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    size_t i = 0;
    // scalar words until the output is 32-byte aligned, so that the main loop can use streaming stores (the staging
    // memory is written once and read by the copy engine: no reason to pull it through the caches)
    while (((uintptr_t)(words + i / 16) & 31) && i + 16 <= n_bases) {
        uint32_t x = 0;
        for (int j = 0; j < 16; j++) x = (x << 2) | g_code.t[src[i + j]];
        words[i / 16] = x;
        i += 16;
    }
    for (; i + 128 <= n_bases; i += 128) {
        prefetch_ahead(src + i);
        const __m256i a = codes_of_32(_mm256_loadu_si256((const __m256i*)(src + i)));
        const __m256i b = codes_of_32(_mm256_loadu_si256((const __m256i*)(src + i + 32)));
        const __m256i c = codes_of_32(_mm256_loadu_si256((const __m256i*)(src + i + 64)));
        const __m256i d = codes_of_32(_mm256_loadu_si256((const __m256i*)(src + i + 96)));
        // bytes of q: low lane a0..a3 b0..b3 c0..c3 d0..d3, high lane a4..a7 b4..b7 c4..c7 d4..d7
        __m256i q = _mm256_packus_epi16(_mm256_packus_epi32(a, b), _mm256_packus_epi32(c, d));
        q = _mm256_shuffle_epi8(q, rev4);                       // first base into the most significant byte of each word
        q = _mm256_permutevar8x32_epi32(q, order);              // words back into sequence order
        _mm256_stream_si256((__m256i*)(words + i / 16), q);
    }
    if (i < n_bases) pack_scalar(src + i, n_bases - i, words + i / 16);
    _mm_sfence();
}


// AVX-512 VBMI: one two-table byte permute is the whole 128-entry code lookup (bytes with bit 7 set are masked to 0),
// and vpmovdb gathers the 16 code bytes of 64 bases.
struct Tab128 {
    alignas(64) uint8_t t[128];
    Tab128() { memset(t, 0, sizeof t); t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3; }
};
const Tab128 g_tab128;

__attribute__((target("avx512f,avx512bw,avx512vbmi,avx512vl"))) inline __m128i words_of_64(__m512i v, __m512i lo, __m512i hi) {
    const __mmask64 ascii = ~_mm512_movepi8_mask(v);
    const __m512i code = _mm512_maskz_permutex2var_epi8(ascii, lo, v, hi);
    const __m512i pairs = _mm512_maddubs_epi16(code, _mm512_set1_epi16(0x0104));
    const __m512i quads = _mm512_madd_epi16(pairs, _mm512_set1_epi32(0x00010010));
    const __m128i bytes = _mm512_cvtepi32_epi8(quads);                    // byte i = bases 4i..4i+3
    return _mm_shuffle_epi8(bytes, _mm_setr_epi8(3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12));
}

template <bool STREAM>
__attribute__((target("avx512f,avx512bw,avx512vbmi,avx512vl"))) void pack_avx512_t(const uint8_t* src, size_t n_bases, uint32_t* words) {
    const __m512i lo = _mm512_load_si512((const void*)g_tab128.t), hi = _mm512_load_si512((const void*)(g_tab128.t + 64));
    size_t i = 0;
    while (((uintptr_t)(words + i / 16) & 31) && i + 16 <= n_bases) {
        uint32_t x = 0;
        for (int j = 0; j < 16; j++) x = (x << 2) | g_code.t[src[i + j]];
        words[i / 16] = x;
        i += 16;
    }
    for (; i + 128 <= n_bases; i += 128) {
        prefetch_ahead(src + i);
        const __m128i a = words_of_64(_mm512_loadu_si512((const void*)(src + i)), lo, hi);
        const __m128i b = words_of_64(_mm512_loadu_si512((const void*)(src + i + 64)), lo, hi);
        const __m256i q = _mm256_inserti128_si256(_mm256_castsi128_si256(a), b, 1);
        if (STREAM) _mm256_stream_si256((__m256i*)(words + i / 16), q);
        else _mm256_store_si256((__m256i*)(words + i / 16), q);
    }
    if (i < n_bases) pack_scalar(src + i, n_bases - i, words + i / 16);
    _mm_sfence();
}

bool have_avx512() {
    static const bool v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi") &&
                          __builtin_cpu_supports("avx512vl") && getenv("SKB_PACK_NO_AVX512") == nullptr;
    return v;
}
bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}

}  // namespace

void pack_avx2(const uint8_t* src, size_t n, uint32_t* words, bool stream = true) { stream ? pack_avx2_t<true>(src, n, words) : pack_avx2_t<false>(src, n, words); }
void pack_avx512(const uint8_t* src, size_t n, uint32_t* words, bool stream = true) { stream ? pack_avx512_t<true>(src, n, words) : pack_avx512_t<false>(src, n, words); }

void host_pack_bases(const uint8_t* src, size_t n_bases, uint32_t* words, bool stream) {
    if (have_avx512()) pack_avx512(src, n_bases, words, stream);
    else if (have_avx2()) pack_avx2(src, n_bases, words, stream);
    else pack_scalar(src, n_bases, words);
}
const char* host_pack_isa() { return have_avx512() ? "avx512vbmi" : have_avx2() ? "avx2" : "scalar"; }

// ------------------------------------------------------------------------------------------------ worker team
unsigned host_cpu_count() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return (unsigned)n;
    }
    const unsigned h = std::thread::hardware_concurrency();
    return h ? h : 1;
}

HostTeam::HostTeam(unsigned n_threads, std::function<void()> thread_init) : thread_init_(std::move(thread_init)) {
    if (n_threads == 0) n_threads = 1;
    workers_.reserve(n_threads);
    for (unsigned i = 0; i < n_threads; i++) workers_.emplace_back([this, i] { run(i); });
}

HostTeam::~HostTeam() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_work_.notify_all();
    for (auto& t : workers_) t.join();
}

void HostTeam::launch(std::function<void(unsigned)> fn) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = std::move(fn);
    pending_ = (unsigned)workers_.size();
    epoch_++;
    lk.unlock();
    cv_work_.notify_all();
}

void HostTeam::wait() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [this] { return pending_ == 0; });
}

void HostTeam::run(unsigned id) {
    if (thread_init_) thread_init_();
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    while (true) {
        cv_work_.wait(lk, [&] { return stop_ || epoch_ != seen; });
        if (stop_) return;
        seen = epoch_;
        lk.unlock();
        fn_(id);
        lk.lock();
        if (--pending_ == 0) cv_done_.notify_all();
    }
}

}  // namespace skb
