// Does a cudaMalloc of a fresh 1 GB slab on a BACKGROUND thread stall kernels and copies that another thread keeps issuing?
// nvcc -O2 -o /tmp/malloc_bg tools/micro/malloc_bg.cu && /tmp/malloc_bg
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void spin(unsigned long long cycles) { const unsigned long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
int main() {
    cudaFree(0);
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    void *h = nullptr, *d = nullptr;
    cudaHostAlloc(&h, 64 << 20, cudaHostAllocDefault); cudaMalloc(&d, 64 << 20);
    auto foreground = [&](const char* what, std::atomic<bool>* stop) {
        // batches of 10 x (20 us kernel) + one 16 MB H2D copy, synchronised: what a sketch call looks like to the driver
        std::vector<double> t;
        double t_all = now();
        int n = 0;
        while (stop ? !stop->load() : n < 300) {
            const double t0 = now();
            for (int k = 0; k < 10; k++) spin<<<148, 256, 0, st>>>(40000);
            cudaMemcpyAsync(d, h, 16 << 20, cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
            t.push_back(now() - t0); n++;
        }
        t_all = now() - t_all;
        double mx = 0, sum = 0; for (double x : t) { mx = x > mx ? x : mx; sum += x; }
        printf("%-34s %4d batches, mean %.3f ms, max %.3f ms\n", what, n, sum / n, mx);
    };
    foreground("foreground alone:", nullptr);
    std::atomic<bool> stop{false};
    std::thread bg([&] {
        cudaSetDevice(0);
        void* keep[8];
        for (int i = 0; i < 8; i++) { const double t0 = now(); cudaMalloc(&keep[i], (size_t)1 << 30); printf("  background cudaMalloc(1 GB) %d: %.1f ms\n", i, now() - t0); }
        stop.store(true);
    });
    foreground("foreground beside 8 x cudaMalloc:", &stop);
    bg.join();
    return 0;
}
