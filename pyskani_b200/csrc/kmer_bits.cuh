// kmer_bits.cuh — bit-level primitives of the seeding kernel (replaces the per-base loop of
// skani::seeding::fmh_seeds, call site reference lib.rs:165-171).
//
// Everything here is __host__ __device__ so that tests/test_kmer_bits_host.py can compile the very same
// code with g++ and check it against the oracle without a GPU.
//
// Layout: a "word" holds 16 bases, 2 bits each (A=0 C=1 G=2 T=3, anything else 0), FIRST base in the
// MOST significant pair.  Concatenating consecutive words (older word on the left) therefore gives the
// sequence as one big-endian bit string and a k-mer is a contiguous bit field of it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SKB_HD __host__ __device__ __forceinline__
#else
#define SKB_HD inline
#endif

namespace skb {

// ((hi:lo) >> s) & 0xffffffff for 0 <= s < 32
SKB_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s == 0 ? lo : (uint32_t)((((uint64_t)hi << 32) | lo) >> s);
#endif
}

// skani's mm_hash64.  NOTE the first step: the Rust source reads `!key.wrapping_add(key << 21)`, which
// parses as !(key + (key << 21)).  See oracle/skani_oracle.cpp.
SKB_HD uint64_t mm_hash64(uint64_t x) {
    x = ~(x + (x << 21));
    x ^= x >> 24;
    x *= 265;           // x + (x << 3) + (x << 8)
    x ^= x >> 14;
    x *= 21;            // x + (x << 2) + (x << 4)
    x ^= x >> 28;
    x += x << 31;
    return x;
}

// 4 ASCII bytes (little-endian in v: first base in the low byte) -> 8 bits, first base in bits 7..6.
// code = ((b >> 1) & 3) ^ ((b >> 2) & 1) maps A/a C/c G/g T/t to 0..3.  Validity is checked by rebuilding the
// upper-case letter each code stands for (0x41 + 2*b0 + 6*b1 + 11*b0*b1 per byte: A C G T) and comparing it with
// the input; only words that contain a foreign byte take the masking path.
SKB_HD uint32_t pack4(uint32_t v) {
    uint32_t code = ((v >> 1) & 0x03030303u) ^ ((v >> 2) & 0x01010101u);
    const uint32_t b0 = code & 0x01010101u, b1 = (code >> 1) & 0x01010101u;
    const uint32_t expect = 0x41414141u + b0 * 2u + b1 * 6u + (b0 & b1) * 11u;    // no carries between bytes
    const uint32_t x = (v & 0xDFDFDFDFu) ^ expect;                                  // zero byte <=> valid base
    if (x != 0u) {
        // high bit of each byte set iff that byte of x is 0
        const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
        const uint32_t ok = ~(t | x | 0x7F7F7F7Fu);
        code &= (ok >> 7) * 3u;                                                     // every other byte encodes as 0
    }
    return (code * 0x40100401u) >> 24;
}

// 16 ASCII bytes (x = bytes 0..3, ... , w = bytes 12..15) -> one word
SKB_HD uint32_t pack16(uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    return (pack4(x) << 24) | (pack4(y) << 16) | (pack4(z) << 8) | pack4(w);
}

SKB_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
#endif
}

// reverse complement of a word: base j of w -> complemented, at pair index j counted from the LSB
SKB_HD uint32_t revcomp_word(uint32_t w) {
    uint32_t x = brev32(~w);
    return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}

struct KmerPair {
    uint64_t f21, r21;   // forward / reverse-complement 21-mer ending at the position
};

// The 21-mers ending at base e (0..15) of word w0; w1, w2 are the two preceding words,
// r0/r1/r2 = revcomp_word(w0/w1/w2).
SKB_HD KmerPair kmers_at(uint32_t w2, uint32_t w1, uint32_t w0, uint32_t r2, uint32_t r1, uint32_t r0, int e) {
    KmerPair p;
    const uint32_t s = 30 - 2 * e;
    uint32_t flo = funnel_r(w0, w1, s);
    uint32_t fhi = funnel_r(w1, w2, s) & 0x3FFu;
    p.f21 = ((uint64_t)fhi << 32) | flo;
    const uint32_t sr = 24 + 2 * e;
    uint32_t rlo, rhi;
    if (sr < 32) {
        rlo = funnel_r(r2, r1, sr);
        rhi = funnel_r(r1, r0, sr) & 0x3FFu;
    } else {
        rlo = funnel_r(r1, r0, sr - 32);
        rhi = (r0 >> (sr - 32)) & 0x3FFu;
    }
    p.r21 = ((uint64_t)rhi << 32) | rlo;
    return p;
}

struct SeedEval {
    uint32_t kmer;       // canonical k-mer (k <= 16)
    bool canonical;      // forward < reverse (SeedPosition.canonical)
    bool is_seed;
    uint64_t marker;     // canonical 21-mer
    bool is_marker;
};

// One position of fmh_seeds.  kmask = 2k low bits set, kshift = 42 - 2k.
SKB_HD SeedEval eval_position(const KmerPair& p, uint32_t kmask, uint32_t kshift, uint64_t thr_seed, uint64_t thr_marker) {
    SeedEval o;
    uint32_t fk = (uint32_t)p.f21 & kmask;
    uint32_t rk = (uint32_t)(p.r21 >> kshift);
    o.canonical = fk < rk;
    o.kmer = o.canonical ? fk : rk;
    o.is_seed = mm_hash64((uint64_t)o.kmer) < thr_seed;
    o.marker = p.f21 < p.r21 ? p.f21 : p.r21;
    o.is_marker = mm_hash64(o.marker) < thr_marker;
    return o;
}

}  // namespace skb
