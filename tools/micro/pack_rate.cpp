// Host compaction rate (host_pack.cpp) by thread count and by where the output goes:
//   stream : every thread writes its part of ONE large output array with streaming stores (what the ingest pipeline does:
//            the staging mirrors the whole batch)
//   ring   : every thread rewrites its own small output block (stays in the cache: no DRAM write traffic at all)
// g++ -O3 -std=c++17 -pthread -I pyskani_b200/csrc tools/micro/pack_rate.cpp pyskani_b200/csrc/host_pack.cpp -o /tmp/pack_rate
#include "host_pack.h"
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
using namespace skb;
int main(int argc, char** argv) {
    const size_t n = (size_t)(argc > 1 ? atol(argv[1]) : 1250) << 20;
    uint8_t* src = (uint8_t*)aligned_alloc(2 << 20, n);
    if (getenv("PACK_NOHUGE")) madvise(src, n, MADV_NOHUGEPAGE); else madvise(src, n, MADV_HUGEPAGE);     // 4 KB or 2 MB pages for the input
    uint32_t* dst = (uint32_t*)aligned_alloc(4096, n / 4 + 4096);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < n; i += 8) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; uint64_t v = 0; for (int j = 0; j < 8; j++) v |= (uint64_t)"ACGT"[(s >> (2 * j)) & 3] << (8 * j); memcpy(src + i, &v, 8); }
    memset(dst, 0, n / 4);
    const size_t PIECE = 1 << 20, np = n / PIECE;
    printf("cpus %u, isa %s, %zu MB\n", host_cpu_count(), host_pack_isa(), n >> 20);
    for (unsigned threads : {1u, 4u, 8u, 12u, 15u, 16u, 24u, 31u}) {
        if (threads > host_cpu_count()) continue;
        HostTeam team(threads);
        for (int mode = 0; mode < 2; mode++) {
            double best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                std::atomic<size_t> next{0};
                auto t0 = std::chrono::steady_clock::now();
                team.launch([&](unsigned id) {
                    uint32_t* ring = dst + (size_t)id * (PIECE / 16);
                    while (true) { size_t p = next.fetch_add(1); if (p >= np) break; host_pack_bases(src + p * PIECE, PIECE, mode ? ring : dst + p * PIECE / 16, !(mode && getenv("SKB_PACK_NO_STREAM"))); }
                });
                team.wait();
                best = std::min(best, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
            }
            printf("threads %2u %-6s: %7.2f ms = %6.1f GB/s\n", threads, mode ? "ring" : "stream", best, n / best / 1e6);
        }
    }
}
