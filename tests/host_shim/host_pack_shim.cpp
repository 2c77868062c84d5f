// Host-side test shim: pyskani_b200/csrc/host_pack.cpp (the ingest compaction of skb_sketch_batch) against the device's
// own pack16 (kmer_bits.cuh compiled for the host).
#include <atomic>

#include "../../pyskani_b200/csrc/host_pack.cpp"
#include "../../pyskani_b200/csrc/kmer_bits.cuh"

extern "C" {

// packs n bases with the dispatching entry (isa = 0), the scalar body (1), the AVX2 body (2), the AVX-512 body (3) or the
// dispatching entry with ordinary stores (4);
// returns the words written, or 0 if this CPU lacks the instruction set
uint64_t shim_host_pack(const uint8_t* src, uint64_t n, uint32_t* words, int isa) {
    if (isa == 1) skb::pack_scalar(src, n, words);
    else if (isa == 2) { if (!skb::have_avx2()) return 0; skb::pack_avx2(src, n, words); }
    else if (isa == 3) { if (!skb::have_avx512()) return 0; skb::pack_avx512(src, n, words); }
    else if (isa == 4) skb::host_pack_bases(src, n, words, false);       // ordinary instead of streaming stores
    else skb::host_pack_bases(src, n, words);
    return (n + 15) / 16 + 1;
}
const char* shim_host_pack_isa() { return skb::host_pack_isa(); }

// the same words through the device routine: pack16 over 16-byte groups (the tail padded with a foreign byte)
void shim_device_pack(const uint8_t* src, uint64_t n, uint32_t* words) {
    for (uint64_t w = 0; 16 * w < n; w++) {
        uint8_t b[16];
        for (int j = 0; j < 16; j++) b[j] = 16 * w + j < n ? src[16 * w + j] : (uint8_t)'N';
        uint32_t v[4];
        for (int q = 0; q < 4; q++) v[q] = (uint32_t)b[4 * q] | ((uint32_t)b[4 * q + 1] << 8) | ((uint32_t)b[4 * q + 2] << 16) | ((uint32_t)b[4 * q + 3] << 24);
        words[w] = skb::pack16(v[0], v[1], v[2], v[3]);
    }
}

// a team of `threads` packs n bases in pieces of `piece` bases (a multiple of 16); returns the rounds that ran
int shim_team_pack(const uint8_t* src, uint64_t n, uint32_t* words, unsigned threads, uint64_t piece, int rounds) {
    skb::HostTeam team(threads);
    int ran = 0;
    for (int r = 0; r < rounds; r++) {
        std::atomic<uint64_t> next{0};
        team.launch([&](unsigned) {
            while (true) {
                const uint64_t o = next.fetch_add(piece);
                if (o >= n) break;
                skb::host_pack_bases(src + o, std::min<uint64_t>(piece, n - o), words + o / 16);
            }
        });
        team.wait();
        ran++;
    }
    return ran;
}
unsigned shim_cpu_count() { return skb::host_cpu_count(); }
}
