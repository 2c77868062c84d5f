// screen_kernels.cu — marker-containment screen.
//
// Replaces skani::screen::check_markers_quickly (reference lib.rs:623-628), which probes the smaller
// marker FxHashSet element by element into the larger one, with a warp-cooperative sorted-set
// intersection: one warp per (query, reference) pair, lanes stride over the smaller sorted list and
// binary-search the larger one.  The kernel produces the exact intersection size; the pass/fail
// decision (count > screen_val^21 * |smaller|, or the small-genome rescue) is a second tiny kernel so
// that the comparison is one IEEE multiply + compare, identical to the oracle's.
#include "skb_internal.cuh"

namespace skb {

namespace {

constexpr int SCREEN_WARPS = 8;

constexpr uint32_t SCREEN_SMEM_MARKERS = 24576;   // 192 KB of u64: query marker lists up to ~24 Mbp genomes

// Fast path: the CTA stages the query's sorted marker list in shared memory once, then walks references; all 256
// threads stride one reference's list (coalesced 8-byte loads) and binary-search the staged list on chip.
__global__ void __launch_bounds__(SCREEN_WARPS * 32)
marker_screen_smem_kernel(const GenomeView* __restrict__ queries, uint32_t n_queries,
                          const GenomeView* __restrict__ refs, uint32_t n_refs, uint32_t* __restrict__ count) {
    extern __shared__ uint64_t s_q[];
    __shared__ uint32_t s_part[SCREEN_WARPS];
    const uint32_t q = blockIdx.y;
    const uint64_t* qm = queries[q].markers;
    const uint32_t nq = queries[q].n_markers;
    if (nq > SCREEN_SMEM_MARKERS) return;           // handled by marker_screen_kernel
    for (uint32_t i = threadIdx.x; i < nq; i += blockDim.x) s_q[i] = __ldg(qm + i);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t r = blockIdx.x; r < n_refs; r += gridDim.x) {
        const uint64_t* b = refs[r].markers;
        const uint32_t nb = refs[r].n_markers;
        uint32_t c = 0;
        for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) {
            const uint64_t v = __ldg(b + i);
            uint32_t lo = 0, hi = nq;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_q[mid] < v) lo = mid + 1; else hi = mid;
            }
            c += (lo < nq && s_q[lo] == v) ? 1u : 0u;
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) s_part[warp] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < SCREEN_WARPS; w++) t += s_part[w];
            count[(size_t)q * n_refs + r] = t;
        }
        __syncthreads();
    }
}

// Fallback for marker lists that do not fit in shared memory: one warp per pair, both lists in global memory.
__global__ void __launch_bounds__(SCREEN_WARPS * 32)
marker_screen_kernel(const GenomeView* __restrict__ queries, uint32_t n_queries,
                     const GenomeView* __restrict__ refs, uint32_t n_refs, uint32_t* __restrict__ count) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.y;
    const uint32_t r = blockIdx.x * SCREEN_WARPS + warp;
    if (r >= n_refs) return;
    if (queries[q].n_markers <= SCREEN_SMEM_MARKERS) return;   // done by the shared-memory kernel
    const uint64_t* a = queries[q].markers; uint32_t na = queries[q].n_markers;
    const uint64_t* b = refs[r].markers;    uint32_t nb = refs[r].n_markers;
    if (na > nb) { const uint64_t* t = a; a = b; b = t; uint32_t tn = na; na = nb; nb = tn; }
    uint32_t c = 0;
    for (uint32_t i = lane; i < na; i += 32) {
        const uint64_t v = __ldg(a + i);
        uint32_t lo = 0, hi = nb;
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(b + mid) < v) lo = mid + 1; else hi = mid;
        }
        c += (lo < nb && __ldg(b + lo) == v) ? 1u : 0u;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) count[(size_t)q * n_refs + r] = c;
}

// pass = screen_val == 0 || (rescue_small && small < 20) || count > p21 * small
__global__ void screen_decide_kernel(const GenomeView* __restrict__ queries, uint32_t n_queries,
                                     const GenomeView* __restrict__ refs, uint32_t n_refs,
                                     const uint32_t* __restrict__ count, double p21, int always, int rescue_small,
                                     uint8_t* __restrict__ pass) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_queries * n_refs) return;
    uint32_t q = (uint32_t)(i / n_refs), r = (uint32_t)(i % n_refs);
    uint32_t small = queries[q].n_markers < refs[r].n_markers ? queries[q].n_markers : refs[r].n_markers;
    bool ok = always || (rescue_small && small < 20u) || ((double)count[i] > __dmul_rn(p21, (double)small));
    pass[i] = ok ? 1 : 0;
}

}  // namespace

void launch_marker_screen(const GenomeView* queries, uint32_t n_queries, const GenomeView* refs, uint32_t n_refs,
                          uint32_t* count, uint32_t max_query_markers, int n_sm, cudaStream_t st) {
    if (n_queries == 0 || n_refs == 0) return;
    const uint32_t staged = max_query_markers < SCREEN_SMEM_MARKERS ? max_query_markers : SCREEN_SMEM_MARKERS;
    const size_t smem = (size_t)staged * 8 + 16;
    cudaFuncSetAttribute(marker_screen_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((size_t)SCREEN_SMEM_MARKERS * 8 + 16));
    const uint32_t per_sm = smem > 100 * 1024 ? 1u : (smem > 48 * 1024 ? 2u : 4u);   // resident CTAs per SM at this footprint
    for (uint32_t q0 = 0; q0 < n_queries; q0 += 65535) {
        const uint32_t nq = n_queries - q0 < 65535 ? n_queries - q0 : 65535;
        // enough CTAs to fill the machine; each CTA walks references with stride gridDim.x
        uint32_t gx = ((uint32_t)n_sm * per_sm + nq - 1) / nq;
        if (gx > n_refs) gx = n_refs;
        if (gx == 0) gx = 1;
        marker_screen_smem_kernel<<<dim3(gx, nq), SCREEN_WARPS * 32, smem, st>>>(queries + q0, nq, refs, n_refs,
                                                                                 count + (size_t)q0 * n_refs);
        g_kernel_launches++;
        if (max_query_markers > SCREEN_SMEM_MARKERS) {
            dim3 grid((n_refs + SCREEN_WARPS - 1) / SCREEN_WARPS, nq);
            marker_screen_kernel<<<grid, SCREEN_WARPS * 32, 0, st>>>(queries + q0, nq, refs, n_refs,
                                                                      count + (size_t)q0 * n_refs);
            g_kernel_launches++;
        }
    }
}

void launch_screen_decide(const GenomeView* queries, uint32_t n_queries, const GenomeView* refs, uint32_t n_refs,
                          const uint32_t* count, double p21, int always, int rescue_small, uint8_t* pass,
                          cudaStream_t st) {
    size_t n = (size_t)n_queries * n_refs;
    if (n == 0) return;
    const int T = 256;
    screen_decide_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(queries, n_queries, refs, n_refs, count, p21,
                                                                    always, rescue_small, pass);
    g_kernel_launches++;
}

}  // namespace skb
