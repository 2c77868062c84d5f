// Streaming read bandwidth of the host per thread (AVX-512 loads over 256 MB per thread): the ceiling of host_pack.cpp, which
// reads every input byte once.   g++ -O3 -std=c++17 -pthread tools/micro/host_read_bw.cpp -o /tmp/host_read_bw && /tmp/host_read_bw 8
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
__attribute__((target("avx512f"))) long long rd(const unsigned char* p, size_t n) {
    __m512i a = _mm512_setzero_si512(), b = a, c = a, d = a;
    for (size_t i = 0; i + 256 <= n; i += 256) {
        a = _mm512_add_epi64(a, _mm512_loadu_si512(p + i)); b = _mm512_add_epi64(b, _mm512_loadu_si512(p + i + 64));
        c = _mm512_add_epi64(c, _mm512_loadu_si512(p + i + 128)); d = _mm512_add_epi64(d, _mm512_loadu_si512(p + i + 192));
    }
    a = _mm512_add_epi64(_mm512_add_epi64(a, b), _mm512_add_epi64(c, d));
    return _mm512_reduce_add_epi64(a);
}
int main(int argc, char** argv) {
    int T = argc > 1 ? atoi(argv[1]) : 1;
    size_t n = 256u << 20;
    std::vector<unsigned char*> bufs(T);
    for (auto& b : bufs) { b = (unsigned char*)aligned_alloc(4096, n); memset(b, 1, n); }
    long long sink = 0;
    double best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back([&, t] { sink += rd(bufs[t], n); });
        for (auto& x : th) x.join();
        best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    printf("%d threads: %.1f GB/s total, %.1f per thread (%lld)\n", T, T * (double)n / best / 1e9, (double)n / best / 1e9, sink);
}
