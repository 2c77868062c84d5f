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

__global__ void __launch_bounds__(SCREEN_WARPS * 32)
marker_screen_kernel(const GenomeView* __restrict__ queries, uint32_t n_queries,
                     const GenomeView* __restrict__ refs, uint32_t n_refs, uint32_t* __restrict__ count) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.y;
    const uint32_t r = blockIdx.x * SCREEN_WARPS + warp;
    if (r >= n_refs) return;
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
                          uint32_t* count, cudaStream_t st) {
    if (n_queries == 0 || n_refs == 0) return;
    for (uint32_t q0 = 0; q0 < n_queries; q0 += 65535) {
        uint32_t nq = n_queries - q0 < 65535 ? n_queries - q0 : 65535;
        dim3 grid((n_refs + SCREEN_WARPS - 1) / SCREEN_WARPS, nq);
        marker_screen_kernel<<<grid, SCREEN_WARPS * 32, 0, st>>>(queries + q0, nq, refs, n_refs,
                                                                  count + (size_t)q0 * n_refs);
        g_kernel_launches++;
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
