// screen_kernels.cu — marker-containment screen.
//
// Replaces skani::screen::check_markers_quickly (reference lib.rs:623-628), which probes the smaller marker FxHashSet
// element by element into the larger one.  Three ways to the same exact intersection sizes:
//   * marker_screen_smem_kernel — small pair matrices: up to four queries per CTA staged in shared memory (sorted list +
//     hashed bitmap), every reference list streams past them once;
//   * marker_screen_kernel      — fallback for marker sets too large for shared memory: a warp per pair, lanes stride over
//     the smaller sorted list and binary-search the larger one;
//   * marker_join_kernel        — large pair matrices: the database's markers as one sorted array of (marker, genome)
//     postings; a CTA looks a query's markers up once and counts per reference in shared memory.
// The pass/fail decision (count > screen_val^21 * |smaller|, or the small-genome rescue) is a second tiny kernel so that
// the comparison is one IEEE multiply + compare, identical to the oracle's.
#include <algorithm>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "skb_internal.cuh"

namespace skb {

namespace {

constexpr int SCREEN_WARPS = 8;

constexpr uint32_t SCREEN_SMEM_BYTES = 200 * 1024;   // shared memory the fast path may use per CTA
constexpr int SCREEN_QT = 4;                         // queries staged together per CTA (each streamed reference serves all)

// bytes of shared memory one staged query of n markers needs: sorted list + membership bitmap (>= 8 bits / marker)
__host__ __device__ inline uint32_t screen_bitmap_bits(uint32_t n) {
    uint32_t b = 1024;
    while (b < 8u * n && b < (1u << 22)) b <<= 1;
    return b;
}
__host__ __device__ inline uint32_t screen_query_bytes(uint32_t n) { return 8u * n + screen_bitmap_bits(n) / 8u; }

__device__ __forceinline__ uint32_t marker_slot(uint64_t v, uint32_t mask) {
    return (uint32_t)((v * 0x9E3779B97F4A7C15ull) >> 40) & mask;
}

// Fast path.  A CTA stages up to SCREEN_QT queries: each query's sorted marker list plus a bitmap over a hash of its
// markers.  References are then streamed once for all staged queries: 256 threads stride a reference's list with
// coalesced 8-byte loads; a marker is first tested against the bitmap (one shared-memory word; ~93 % of the probes of
// an unrelated pair end here) and only on a hit binary-searched in the staged list.  The count is exact.
constexpr int SCREEN_MAX_WARPS = 32;             // the fast path runs 8, 16 or 32 warps per CTA depending on its smem footprint

__global__ void __launch_bounds__(SCREEN_MAX_WARPS * 32)
marker_screen_smem_kernel(const GenomeView* __restrict__ queries, uint32_t n_queries,
                          const GenomeView* __restrict__ refs, uint32_t n_refs, uint32_t* __restrict__ count) {
    extern __shared__ uint64_t s_mem[];
    __shared__ uint32_t s_part[SCREEN_MAX_WARPS][SCREEN_QT];
    __shared__ uint64_t s_qv[SCREEN_MAX_WARPS][64];    // per-warp queue of bitmap hits: marker value ...
    __shared__ uint8_t s_qt[SCREEN_MAX_WARPS][64];     // ... and staged-query slot
    const uint32_t q0 = blockIdx.y * SCREEN_QT;
    const uint32_t nqt = min((uint32_t)SCREEN_QT, n_queries - q0);
    uint32_t nq[SCREEN_QT], bmask[SCREEN_QT];
    const uint64_t* list[SCREEN_QT];
    uint32_t* bits[SCREEN_QT];
    bool fits = true;
    {
        uint32_t off = 0;
#pragma unroll
        for (int t = 0; t < SCREEN_QT; t++) {
            nq[t] = (uint32_t)t < nqt ? queries[q0 + t].n_markers : 0u;
            const uint32_t bb = screen_bitmap_bits(nq[t]);
            bmask[t] = bb - 1u;
            list[t] = (const uint64_t*)((const char*)s_mem + off);
            bits[t] = (uint32_t*)((char*)s_mem + off + 8u * nq[t]);
            off += 8u * nq[t] + bb / 8u;
        }
        fits = off <= SCREEN_SMEM_BYTES;
    }
    if (!fits) return;                              // this group is handled by marker_screen_kernel
#pragma unroll
    for (int t = 0; t < SCREEN_QT; t++) {
        uint32_t* bt = bits[t];
        for (uint32_t i = threadIdx.x; i <= bmask[t] / 32u; i += blockDim.x) bt[i] = 0u;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < SCREEN_QT; t++) {
        if ((uint32_t)t >= nqt) break;
        const uint64_t* qm = queries[q0 + t].markers;
        uint64_t* lt = const_cast<uint64_t*>(list[t]);
        for (uint32_t i = threadIdx.x; i < nq[t]; i += blockDim.x) {
            const uint64_t v = __ldg(qm + i);
            lt[i] = v;
            const uint32_t sl = marker_slot(v, bmask[t]);
            atomicOr(&bits[t][sl >> 5], 1u << (sl & 31u));
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Bitmap hits (true matches and ~6 % false positives) are not searched on the spot — that would leave 1-2 lanes of a
    // warp running a 12-step search while 30 wait.  They are queued per warp and searched 32 at a time by full warps.
    uint64_t* qv = s_qv[warp];
    uint8_t* qt = s_qt[warp];
    for (uint32_t r = blockIdx.x; r < n_refs; r += gridDim.x) {
        const uint64_t* b = refs[r].markers;
        const uint32_t nb = refs[r].n_markers;
        uint32_t cnt[SCREEN_QT];                    // warp-uniform match counts
#pragma unroll
        for (int t = 0; t < SCREEN_QT; t++) cnt[t] = 0;
        uint32_t qn = 0;                            // queue length (warp-uniform)
        auto drain = [&](uint32_t take) {           // search the last `take` (<= 32) queued entries, one per lane
            const bool act = (uint32_t)lane < take;
            uint64_t v = 0; uint32_t tt = 0;
            if (act) { v = qv[qn - take + lane]; tt = qt[qn - take + lane]; }
            const uint64_t* lt = list[0]; uint32_t n = 0;
#pragma unroll
            for (int t = 0; t < SCREEN_QT; t++) if (tt == (uint32_t)t) { lt = list[t]; n = nq[t]; }
            uint32_t lo = 0, hi = act ? n : 0u;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (lt[mid] < v) lo = mid + 1; else hi = mid;
            }
            const bool found = act && lo < n && lt[lo] == v;
#pragma unroll
            for (int t = 0; t < SCREEN_QT; t++) cnt[t] += __popc(__ballot_sync(0xffffffffu, found && tt == (uint32_t)t));
            qn -= take;
            __syncwarp();                           // the drained slots are written again by other lanes of the warp
        };
        constexpr int U = 4;                        // independent 8-byte loads in flight per thread
        for (uint32_t i0 = threadIdx.x; i0 - threadIdx.x < nb; i0 += blockDim.x * U) {   // warp-uniform trip count
            uint64_t vv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t i = i0 + u * blockDim.x;
                vv[u] = i < nb ? __ldg(b + i) : ~0ull;       // ~0 is not a 42-bit marker: it never matches
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint64_t v = vv[u];
                const uint32_t h = (uint32_t)((v * 0x9E3779B97F4A7C15ull) >> 40);
#pragma unroll
                for (int t = 0; t < SCREEN_QT; t++) {
                    const uint32_t sl = h & bmask[t];
                    const bool hit = (bits[t][sl >> 5] >> (sl & 31u)) & 1u;
                    const uint32_t m = __ballot_sync(0xffffffffu, hit);
                    if (m) {
                        if (hit) { const uint32_t at = qn + __popc(m & ((1u << lane) - 1u)); qv[at] = v; qt[at] = (uint8_t)t; }
                        qn += __popc(m);
                        __syncwarp();
                        if (qn >= 32) drain(32);
                    }
                }
            }
        }
        __syncwarp();
        if (qn) drain(qn);                          // qn < 32 here
        if (lane == 0) {
#pragma unroll
            for (int t = 0; t < SCREEN_QT; t++) s_part[warp][t] = cnt[t];
        }
        __syncthreads();
        if (threadIdx.x < nqt) {
            uint32_t tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += s_part[w][threadIdx.x];
            count[(size_t)(q0 + threadIdx.x) * n_refs + r] = tot;
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
    {   // done by the shared-memory kernel when the query's group of SCREEN_QT fits
        const uint32_t g0 = q / SCREEN_QT * SCREEN_QT;
        uint32_t off = 0;
        for (uint32_t t = g0; t < g0 + SCREEN_QT; t++) off += screen_query_bytes(t < n_queries ? queries[t].n_markers : 0u);
        if (off <= SCREEN_SMEM_BYTES) return;
    }
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
                          uint32_t* count, const uint32_t* query_markers_host, int n_sm, cudaStream_t st) {
    if (n_queries == 0 || n_refs == 0) return;
    // shared memory of the largest group of SCREEN_QT consecutive queries that fits; groups that do not fit go to the fallback
    size_t smem = 0; bool any_big = false;
    for (uint32_t g0 = 0; g0 < n_queries; g0 += SCREEN_QT) {
        size_t off = 0;
        for (uint32_t t = g0; t < g0 + SCREEN_QT; t++) off += screen_query_bytes(t < n_queries ? query_markers_host[t] : 0u);   // empty slots still own a minimal bitmap
        if (off <= SCREEN_SMEM_BYTES) smem = off > smem ? off : smem; else any_big = true;
    }
    smem += 64;
    cudaFuncSetAttribute(marker_screen_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SCREEN_SMEM_BYTES + 64));
    // resident CTAs per SM at this footprint; fewer resident CTAs are compensated by wider CTAs (always 32 warps per SM)
    const uint32_t per_sm = smem > 100 * 1024 ? 1u : (smem > 48 * 1024 ? 2u : 4u);
    const uint32_t threads = per_sm == 1 ? 1024u : (per_sm == 2 ? 512u : 256u);
    const uint32_t n_groups = (n_queries + SCREEN_QT - 1) / SCREEN_QT;
    for (uint32_t g0 = 0; g0 < n_groups; g0 += 65535) {
        const uint32_t ng = n_groups - g0 < 65535 ? n_groups - g0 : 65535;
        const uint32_t qb = g0 * SCREEN_QT, nq = n_queries - qb < ng * SCREEN_QT ? n_queries - qb : ng * SCREEN_QT;
        // enough CTAs to fill the machine; each CTA walks references with stride gridDim.x
        // split the references over gx CTAs per group: measured on B200, ~8 CTAs per resident slot balance the
        // uneven per-pair cost best (related pairs search far more), as long as a CTA still streams >= 8 references
        // to amortise staging its queries
        const uint32_t slots = (uint32_t)n_sm * per_sm;
        const uint32_t gx_target = (8u * slots + ng - 1) / ng;      // ~8 CTAs per slot
        const uint32_t gx_fill = (slots + ng - 1) / ng;              // at least fill the machine once
        const uint32_t gx_amort = n_refs / 8u;                       // >= 8 references per CTA
        uint32_t gx = gx_target < (gx_amort > gx_fill ? gx_amort : gx_fill) ? gx_target : (gx_amort > gx_fill ? gx_amort : gx_fill);
        if (gx > n_refs) gx = n_refs;
        if (gx < 1) gx = 1;
        if (const char* e = getenv("SKB_SCREEN_GX")) { gx = (uint32_t)atoi(e); if (gx < 1) gx = 1; if (gx > n_refs) gx = n_refs; }
        marker_screen_smem_kernel<<<dim3(gx, ng), threads, smem, st>>>(queries + qb, nq, refs, n_refs,
                                                                                 count + (size_t)qb * n_refs);
        g_kernel_launches++;
    }
    if (any_big) {
        for (uint32_t q0 = 0; q0 < n_queries; q0 += 65535 / SCREEN_QT * SCREEN_QT) {
            const uint32_t step = 65535 / SCREEN_QT * SCREEN_QT;
            const uint32_t nq = n_queries - q0 < step ? n_queries - q0 : step;
            dim3 grid((n_refs + SCREEN_WARPS - 1) / SCREEN_WARPS, nq);
            marker_screen_kernel<<<grid, SCREEN_WARPS * 32, 0, st>>>(queries + q0, nq, refs, n_refs, count + (size_t)q0 * n_refs);
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

// ---------------------------------------------------------------- screen through a marker index (many queries x many refs)
// The pairwise kernels above stream every reference list past every query: |Q| x |R| x markers work even when the genomes
// share nothing.  For large pair matrices the database's markers are instead kept as ONE sorted array of (marker, genome)
// postings with a bucket table on the top bits (built once per database state).  A CTA owns one query and a tile of
// references: its counters sit in shared memory, each query marker is looked up once, and every posting found bumps a
// shared-memory counter.  Work = |Q| x markers lookups + the postings actually shared.  The counts are exactly the
// intersection sizes the pairwise kernels produce (marker sets are duplicate-free), so the decision kernel is the same.
namespace {

__global__ void marker_postings_kernel(const GenomeView* __restrict__ refs, const uint32_t* __restrict__ genome_off,
                                       uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t g = blockIdx.x;
    const GenomeView& R = refs[g];
    const uint32_t off = genome_off[g], n = R.n_markers;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { keys[off + i] = R.markers[i]; vals[off + i] = g; }
}

__global__ void marker_index_buckets_kernel(const uint64_t* __restrict__ keys, uint32_t n, uint32_t shift, uint32_t n_buckets,
                                            uint32_t* __restrict__ bucket) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_buckets) return;
    const uint64_t target = (uint64_t)b << shift;
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (keys[mid] < target) lo = mid + 1; else hi = mid; }
    bucket[b] = lo;
}

constexpr int JOIN_THREADS = 256;

__global__ void __launch_bounds__(JOIN_THREADS) marker_join_kernel(const GenomeView* __restrict__ queries,
                                                                    const uint64_t* __restrict__ keys,
                                                                    const uint32_t* __restrict__ vals,
                                                                    const uint32_t* __restrict__ bucket, uint32_t shift,
                                                                    uint32_t n_refs, uint32_t tile, uint32_t* __restrict__ count) {
    // gridDim.z > 1: the query's markers are divided among that many CTAs (few queries would otherwise leave most SMs
    // without a CTA, and a CTA's 5 000 dependent lookups take as long whether 125 or 1 000 of them run); the partial counts
    // are then ADDED to the zeroed matrix
    extern __shared__ uint32_t s_cnt[];
    const uint32_t t0 = blockIdx.x * tile, tn = min(tile, n_refs - t0);
    const GenomeView& Q = queries[blockIdx.y];
    for (uint32_t i = threadIdx.x; i < tn; i += JOIN_THREADS) s_cnt[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t m_all = Q.n_markers;
    const uint32_t per = ((m_all + gridDim.z - 1) / gridDim.z + 31u) & ~31u;         // markers of this CTA's part
    const uint32_t m0 = min(m_all, blockIdx.z * per), nm = min(m_all, m0 + per);
    // a warp takes 32 consecutive query markers: every lane finds its own first posting, then the warp walks the
    // posting run of each marker that has one with coalesced loads
    for (uint32_t base = m0 + (threadIdx.x >> 5) * 32; base < nm; base += JOIN_THREADS) {
        const uint32_t i = base + (uint32_t)lane;
        uint64_t m = ~0ull; uint32_t lo = 0, hi = 0;
        if (i < nm) {
            m = __ldg(Q.markers + i);
            const uint32_t b = (uint32_t)(m >> shift);
            lo = __ldg(bucket + b); hi = __ldg(bucket + b + 1);
            const uint32_t end = hi;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(keys + mid) < m) lo = mid + 1; else hi = mid; }
            hi = end;
        }
        const bool found = i < nm && lo < hi && __ldg(keys + lo) == m;
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, found);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint64_t ml = __shfl_sync(0xFFFFFFFFu, m, l);
            const uint32_t lol = __shfl_sync(0xFFFFFFFFu, lo, l), hil = __shfl_sync(0xFFFFFFFFu, hi, l);
            for (uint32_t j = lol;; j += 32) {
                const uint32_t idx = j + (uint32_t)lane;
                const bool ok = idx < hil && __ldg(keys + idx) == ml;
                if (ok) {
                    const uint32_t g = __ldg(vals + idx) - t0;
                    if (g < tn) atomicAdd(&s_cnt[g], 1u);
                }
                if (!__all_sync(0xFFFFFFFFu, ok)) break;
            }
        }
    }
    __syncthreads();
    uint32_t* row = count + (size_t)blockIdx.y * n_refs + t0;
    if (gridDim.z == 1) {
        for (uint32_t i = threadIdx.x; i < tn; i += JOIN_THREADS) row[i] = s_cnt[i];
    } else {
        for (uint32_t i = threadIdx.x; i < tn; i += JOIN_THREADS) if (s_cnt[i]) atomicAdd(&row[i], s_cnt[i]);
    }
}

}  // namespace

constexpr uint32_t MIDX_BUCKET_CAP = 96;      // largest bucket the partition path ranks by comparison

namespace {

// ---- index build by bucket partition (round 2): histogram of the top bits -> scan (= the bucket table) -> scatter with a
// cursor per bucket -> every posting ranks itself inside its bucket by (marker, genome).  The result is the array a stable
// radix sort by marker produces (postings are generated genome after genome), at a third of its cost: the sort makes six
// passes over 12 bytes per posting, and at 5 M postings each pass is latency-, not bandwidth-bound.
__global__ void marker_hist_kernel(const GenomeView* __restrict__ refs, uint32_t shift, uint32_t* __restrict__ cnt) {
    const GenomeView& R = refs[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < R.n_markers; i += blockDim.x) atomicAdd(&cnt[(uint32_t)(R.markers[i] >> shift)], 1u);
}
__global__ void marker_scatter_kernel(const GenomeView* __restrict__ refs, uint32_t shift, const uint32_t* __restrict__ bucket,
                                      uint32_t* __restrict__ cursor, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t g = blockIdx.x;
    const GenomeView& R = refs[g];
    for (uint32_t i = threadIdx.x; i < R.n_markers; i += blockDim.x) {
        const uint64_t m = R.markers[i];
        const uint32_t b = (uint32_t)(m >> shift);
        const uint32_t slot = bucket[b] + atomicAdd(&cursor[b], 1u);
        keys[slot] = m; vals[slot] = g;
    }
}
__global__ void marker_rank_kernel(uint32_t n, const uint64_t* __restrict__ k_in, const uint32_t* __restrict__ v_in,
                                   const uint32_t* __restrict__ bucket, uint32_t shift, uint64_t* __restrict__ keys,
                                   uint32_t* __restrict__ vals, uint32_t* __restrict__ overflow) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t k = k_in[j];
    const uint32_t v = v_in[j];
    const uint32_t b = (uint32_t)(k >> shift);
    const uint32_t lo = __ldg(bucket + b), hi = __ldg(bucket + b + 1);
    if (hi - lo > MIDX_BUCKET_CAP) { *overflow = 1u; return; }          // the host falls back to the radix sort
    uint32_t rank = 0;
    for (uint32_t t = lo; t < hi; t++) {
        const uint64_t kt = k_in[t];
        rank += (kt < k || (kt == k && v_in[t] < v)) ? 1u : 0u;
    }
    keys[lo + rank] = k; vals[lo + rank] = v;
}

}  // namespace

size_t marker_index_scratch_bytes(uint32_t n_postings, uint32_t n_buckets) {
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n_postings, 0, MARKER_BITS);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n_buckets + 1);
    return (((size_t)n_postings * 8 + 255) & ~(size_t)255) + (((size_t)n_postings * 4 + 255) & ~(size_t)255) +
           2 * (((size_t)n_buckets * 4 + 8 + 255) & ~(size_t)255) + 256 + std::max(sort_bytes, scan_bytes) + 256;
}

// overflow (device, 4 bytes): set when a bucket was too large for the partition path; the caller then calls again with
// use_sort = true.  Everything is asynchronous on `st`.
void build_marker_index(const GenomeView* refs, uint32_t n_refs, const uint32_t* genome_off, uint32_t n_postings,
                        uint64_t* keys, uint32_t* vals, uint32_t* bucket, uint32_t shift, uint32_t n_buckets,
                        void* scratch, size_t scratch_bytes, bool use_sort, uint32_t* overflow, cudaStream_t st) {
    if (n_refs == 0) return;
    char* p = (char*)scratch;
    uint64_t* k_in = (uint64_t*)p; p += ((size_t)n_postings * 8 + 255) & ~(size_t)255;
    uint32_t* v_in = (uint32_t*)p; p += ((size_t)n_postings * 4 + 255) & ~(size_t)255;
    uint32_t* cnt = (uint32_t*)p; p += ((size_t)n_buckets * 4 + 8 + 255) & ~(size_t)255;
    uint32_t* cursor = (uint32_t*)p; p += ((size_t)n_buckets * 4 + 8 + 255) & ~(size_t)255;
    p += 256;
    size_t tmp_bytes = scratch_bytes - (size_t)(p - (char*)scratch);
    if (use_sort) {
        marker_postings_kernel<<<n_refs, 256, 0, st>>>(refs, genome_off, k_in, v_in);
        g_kernel_launches++;
        if (n_postings) {
            cub::DeviceRadixSort::SortPairs(p, tmp_bytes, k_in, keys, v_in, vals, (int)n_postings, 0, MARKER_BITS, st);
            g_kernel_launches += 2 * ((MARKER_BITS + 7) / 8);
        }
        marker_index_buckets_kernel<<<(n_buckets + 1 + 255) / 256, 256, 0, st>>>(keys, n_postings, shift, n_buckets, bucket);
        g_kernel_launches++;
        return;
    }
    cudaMemsetAsync(cnt, 0, 4 * ((size_t)n_buckets + 1), st);
    cudaMemsetAsync(cursor, 0, 4 * ((size_t)n_buckets + 1), st);
    cudaMemsetAsync(overflow, 0, 4, st);
    marker_hist_kernel<<<n_refs, 256, 0, st>>>(refs, shift, cnt);
    cub::DeviceScan::ExclusiveSum(p, tmp_bytes, cnt, bucket, (int)n_buckets + 1, st);
    marker_scatter_kernel<<<n_refs, 256, 0, st>>>(refs, shift, bucket, cursor, k_in, v_in);
    if (n_postings) marker_rank_kernel<<<(n_postings + 255) / 256, 256, 0, st>>>(n_postings, k_in, v_in, bucket, shift, keys, vals, overflow);
    g_kernel_launches += 5;
}

void launch_marker_join(const GenomeView* queries, uint32_t n_queries, uint32_t n_refs, const uint64_t* keys,
                        const uint32_t* vals, const uint32_t* bucket, uint32_t shift, uint32_t* count, cudaStream_t st) {
    if (n_queries == 0 || n_refs == 0) return;
    const uint32_t tile = n_refs < 32768u ? n_refs : 32768u;
    const size_t smem = (size_t)tile * 4;
    cudaFuncSetAttribute(marker_join_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4);   // per launch: the attribute belongs to the current device
    const uint32_t tiles = (n_refs + tile - 1) / tile;
    // at least ~4 CTAs per SM: split the markers of each query when there are few queries
    uint32_t splits = 1;
    if ((uint64_t)tiles * n_queries < 592) splits = (uint32_t)std::min<uint64_t>(16, (592 + (uint64_t)tiles * n_queries - 1) / ((uint64_t)tiles * n_queries));
    if (splits > 1) cudaMemsetAsync(count, 0, sizeof(uint32_t) * (size_t)n_queries * n_refs, st);
    dim3 grid(tiles, n_queries, splits);
    marker_join_kernel<<<grid, JOIN_THREADS, smem, st>>>(queries, keys, vals, bucket, shift, n_refs, tile, count);
    g_kernel_launches++;
}

}  // namespace skb
