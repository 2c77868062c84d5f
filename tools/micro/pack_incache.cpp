// host_pack_bases on an input of the given size in KB, repeated until 2 GB have gone through (cache-resident inputs show the
// arithmetic limit of the packing loop, large ones the memory limit of one thread).
//   g++ -O3 -std=c++17 -pthread -I pyskani_b200/csrc tools/micro/pack_incache.cpp pyskani_b200/csrc/host_pack.cpp -o /tmp/pack_incache && /tmp/pack_incache 32
#include "host_pack.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
using namespace skb;
int main(int argc, char** argv) {
    const size_t n = (size_t)(argc > 1 ? atol(argv[1]) : 32) << 10;
    uint8_t* src = (uint8_t*)aligned_alloc(4096, n + 65536);
    uint32_t* dst = (uint32_t*)aligned_alloc(4096, n / 4 + 4096);
    for (size_t i = 0; i < n + 65536; i++) src[i] = "ACGT"[(i * 2654435761u >> 13) & 3];
    const int reps = (int)((size_t)(2u << 30) / n);
    for (int stream = 0; stream < 2; stream++) {
        auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < reps; r++) host_pack_bases(src, n, dst, stream);
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("%zu KB x %d, stream %d: %.1f GB/s\n", n >> 10, reps, stream, (double)n * reps / s / 1e9);
    }
}
