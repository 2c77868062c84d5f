// Host build of pyskani_b200/csrc/kmer_bits.cuh for CPU-side unit tests (no GPU needed).
// Walks one contig exactly the way the seeding kernel's threads do: 16-base words, two halo words,
// per-position extraction by funnel shifts.
#include "../../pyskani_b200/csrc/kmer_bits.cuh"
#include <cstring>
#include <vector>

extern "C" {

uint64_t shim_hash(uint64_t x) { return skb::mm_hash64(x); }
uint32_t shim_pack16(const uint8_t* p) {
    uint32_t v[4];
    std::memcpy(v, p, 16);
    return skb::pack16(v[0], v[1], v[2], v[3]);
}
uint32_t shim_revcomp_word(uint32_t w) { return skb::revcomp_word(w); }

// returns number of seeds; markers appended (with duplicates, in position order)
int64_t shim_scan(const uint8_t* seq, uint64_t len, int k, int c, int marker_c,
                  uint32_t* seed_kmer, uint32_t* seed_pos, uint8_t* seed_canon, int64_t cap,
                  uint64_t* markers, int64_t* n_markers) {
    const uint64_t nw = (len + 15) / 16;
    std::vector<uint8_t> padded(nw * 16 + 16, 0);
    std::memcpy(padded.data(), seq, len);
    std::vector<uint32_t> W(nw + 2, 0), R(nw + 2, 0);
    for (uint64_t w = 0; w < nw; w++) { W[w + 2] = shim_pack16(&padded[w * 16]); }
    for (uint64_t w = 0; w < nw + 2; w++) R[w] = skb::revcomp_word(W[w]);
    const uint32_t kmask = k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1);
    const uint32_t kshift = 42 - 2 * k;
    const uint64_t thr_seed = UINT64_MAX / (uint64_t)c, thr_marker = UINT64_MAX / (uint64_t)marker_c;
    int64_t ns = 0, nm = 0;
    for (uint64_t w = 0; w < nw; w++)
        for (int e = 0; e < 16; e++) {
            uint64_t pos = w * 16 + e;
            if (pos < 20 || pos >= len) continue;
            skb::KmerPair p = skb::kmers_at(W[w], W[w + 1], W[w + 2], R[w], R[w + 1], R[w + 2], e);
            skb::SeedEval ev = skb::eval_position(p, kmask, kshift, thr_seed, thr_marker);
            if (ev.is_seed) { if (ns < cap) { seed_kmer[ns] = ev.kmer; seed_pos[ns] = (uint32_t)pos; seed_canon[ns] = ev.canonical; } ns++; }
            if (ev.is_marker) { if (nm < cap) markers[nm] = ev.marker; nm++; }
        }
    *n_markers = nm;
    return ns;
}
}
