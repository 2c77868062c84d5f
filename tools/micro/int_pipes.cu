// Issue rates of the integer instructions the seeding kernel is made of, alone and mixed (one B200, 24 warps per SM like
// seed_scan_kernel).  Prints warp instructions per cycle per SM sub-partition for every variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/int_pipes tools/micro/int_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ILP = 8, ITERS = 2048;

template <int V>
__global__ void __launch_bounds__(256, 3) k(uint32_t* out, uint32_t m, uint32_t two, uint32_t c, long long* cyc) {
    uint32_t x[ILP], y[ILP];
    uint64_t d[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x * 2654435761u + i * 40503u + m; y[i] = x[i] ^ (c + i); d[i] = ((uint64_t)y[i] << 32) | x[i]; }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (V == 0) { x[i] = (x[i] ^ y[i]) & (y[i] | c); }                                     // LOP3
            if (V == 1) { x[i] = __funnelshift_r(x[i], y[i], 7); }                                  // SHF
            if (V == 2) { asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m), "r"(y[i])); }     // IMAD
            if (V == 3) { asm("mul.wide.u32 %0, %1, %2;" : "=l"(d[i]) : "r"((uint32_t)d[i]), "r"(m)); }   // IMAD.WIDE
            if (V == 4) { asm("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m), "r"(y[i])); }     // IMAD.HI
            if (V == 5) { asm("{ .reg .u32 t; mad.hi.cc.u32 t, %0, %1, %2; madc.lo.u32 %0, %0, %3, 0; }"
                              : "+r"(y[i]) : "r"(m), "r"(c), "r"(two)); }                           // IMAD.HI -> P, IMAD.X
            if (V == 6) { x[i] = (x[i] ^ y[i]) & (y[i] | c); y[i] = __funnelshift_r(y[i], x[i], 9);
                          asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m), "r"(y[i])); }     // 2 ALU : 1 IMAD
            if (V == 7) { x[i] = (x[i] ^ y[i]) & (y[i] | c); y[i] = __funnelshift_r(y[i], x[i], 9);
                          asm("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m), "r"(y[i])); }     // 2 ALU : 1 IMAD.HI
            if (V == 8) { x[i] = (x[i] ^ y[i]) & (y[i] | c); y[i] = __funnelshift_r(y[i], x[i], 9);
                          asm("mul.wide.u32 %0, %1, %2;" : "=l"(d[i]) : "r"(x[i]), "r"(m));
                          x[i] = (uint32_t)d[i] ^ (uint32_t)(d[i] >> 32); }                                         // 3 ALU : 1 IMAD.WIDE
            if (V == 9) { x[i] = (x[i] ^ y[i]) & (y[i] | c); y[i] = __funnelshift_r(y[i], x[i], 9);
                          asm("{ .reg .u32 t; mad.hi.cc.u32 t, %1, %2, %3; madc.lo.u32 %0, %0, %4, 0; }"
                              : "+r"(y[i]) : "r"(x[i]), "r"(m), "r"(c), "r"(two)); }                // 2 ALU : IMAD.HI -> P + IMAD.X
            if (V == 10) { const bool p = x[i] > c; y[i] = y[i] + y[i] + (p ? 1u : 0u); x[i] ^= y[i]; }   // ISETP + SEL/IADD3 form
        }
    }
    const long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= x[i] ^ y[i] ^ (uint32_t)d[i] ^ (uint32_t)(d[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name, int instr_per_step, uint32_t* out, long long* cyc, int grid) {
    k<V><<<grid, 256>>>(out, 3u, 2u, 12345u, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<V><<<grid, 256>>>(out, 0xFFFFFFFFu, 2u, 12345u, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[4096]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < grid; i++) mean += (double)h[i]; mean /= grid;
    // per SM sub-partition: 6 warps (24 per SM), each ITERS * ILP * instr_per_step warp instructions
    const double per_smsp = 6.0 * ITERS * ILP * instr_per_step;
    printf("%-34s %2d instr/step  %8.0f cycles  %.3f warp instr / cycle / sub-partition  %.2f cycles / step  (%.3f ms)\n", name, instr_per_step, mean,
           per_smsp / mean, mean / (6.0 * ITERS * ILP), ms);
}

int main() {
    const int grid = 148 * 3;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4 * grid * 256); cudaMalloc(&cyc, 8 * grid);
    run<0>("LOP3", 1, out, cyc, grid);
    run<1>("SHF", 1, out, cyc, grid);
    run<2>("IMAD", 1, out, cyc, grid);
    run<3>("IMAD.WIDE", 1, out, cyc, grid);
    run<4>("IMAD.HI", 1, out, cyc, grid);
    run<5>("IMAD.HI->P + IMAD.X", 2, out, cyc, grid);
    run<6>("LOP3 + SHF + IMAD", 3, out, cyc, grid);
    run<7>("LOP3 + SHF + IMAD.HI + IMAD.MOV", 4, out, cyc, grid);
    run<8>("LOP3 + SHF + IMAD.WIDE + LOP3", 4, out, cyc, grid);
    run<9>("LOP3 + SHF + IMAD.HI->P + IMAD.X", 4, out, cyc, grid);
    run<10>("ISETP + SEL + 2 LOP3 + IMAD.IADD + IMAD.MOV", 6, out, cyc, grid);
    return 0;
}
