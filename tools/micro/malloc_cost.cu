// cost of fresh device memory: cudaMalloc vs cudaMallocAsync, by size (run on the GPU box)
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    cudaFree(0);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, 0);
    unsigned long long thr = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    for (int rep = 0; rep < 2; rep++)
    for (size_t mb : {64, 256, 1024, 4096}) {
        void* p = nullptr;
        double t0 = now(); cudaMalloc(&p, mb << 20); double t1 = now();
        cudaMemsetAsync(p, 1, mb << 20, st); cudaStreamSynchronize(st); double t2 = now();
        cudaFree(p); double t3 = now();
        void* q = nullptr;
        cudaMallocAsync(&q, mb << 20, st); cudaStreamSynchronize(st); double t4 = now();
        cudaFreeAsync(q, st); cudaStreamSynchronize(st); double t5 = now();
        printf("%5zu MB: cudaMalloc %.2f ms, first memset %.2f ms, cudaFree %.2f ms | cudaMallocAsync %.2f ms, cudaFreeAsync %.2f ms\n", mb, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4);
    }
    // a database growing by 1 GB slabs: is every new slab equally cheap?
    void* keep[16];
    for (int i = 0; i < 12; i++) {
        double t0 = now(); cudaMalloc(&keep[i], (size_t)1 << 30); double t1 = now();
        cudaMemsetAsync(keep[i], 1, (size_t)1 << 30, st); cudaStreamSynchronize(st);
        printf("slab %2d: cudaMalloc(1 GB) %.2f ms\n", i, t1 - t0);
    }
    return 0;
}
