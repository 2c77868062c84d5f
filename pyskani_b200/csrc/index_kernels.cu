// index_kernels.cu — turns the seeding kernel's position-ordered output into the device-resident index.
//
// Replaces the hash containers skani fills inside fmh_seeds (reference lib.rs:165-171):
//   kmer_seeds_k : FxHashMap<kmer, SmallVec<SeedPosition>>  ->  seeds sorted by (genome, kmer, contig, pos)
//                                                                + a bucket table over the k-mer's top bits
//   marker_seeds : FxHashSet<u64>                            ->  sorted unique 21-mers per genome
// The k-mer order is built by a bucket partition (histogram -> per-genome scan, whose output is the bucket table ->
// scatter -> rank inside the <= 256-entry bucket); CUB's radix sort (library code, like cuBLAS for a GEMM) remains as
// the fallback for genomes whose buckets overflow, and sorts the marker sets (one segment per genome).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <thrust/iterator/transform_iterator.h>

#include "skb_internal.cuh"

namespace skb {

namespace {

__device__ __forceinline__ uint32_t upper_bound_u32(const uint32_t* a, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* a, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t* a, uint32_t n, uint64_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// key = genome << 32 | kmer ; value = index in position order
__global__ void make_seed_keys(uint32_t n, uint32_t n_genomes, const uint32_t* __restrict__ genome_seed_start,
                               const uint32_t* __restrict__ kmer_p, uint64_t* __restrict__ keys,
                               uint32_t* __restrict__ vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // genome g owns [start[g], start[g+1]); empty genomes are skipped by the upper bound
    uint32_t g = upper_bound_u32(genome_seed_start, n_genomes + 1, i) - 1;
    keys[i] = ((uint64_t)g << 32) | kmer_p[i];
    vals[i] = i;
}

__global__ void gather_kmer_order(uint32_t n, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                  const uint32_t* __restrict__ pos_p, const uint32_t* __restrict__ meta_p,
                                  const uint32_t* __restrict__ genome_seed_start,
                                  uint32_t* __restrict__ kmer_k, uint32_t* __restrict__ pos_k,
                                  uint32_t* __restrict__ meta_k, uint32_t* __restrict__ perm_k) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t src = vals[i];
    kmer_k[i] = (uint32_t)keys[i];
    pos_k[i] = pos_p[src];
    meta_k[i] = meta_p[src];
    if (perm_k) perm_k[i] = src - genome_seed_start[(uint32_t)(keys[i] >> 32)];      // genome-local position-order index
}

__global__ void strip_marker_keys(uint32_t n_in, const uint32_t* __restrict__ n_unique, const uint64_t* __restrict__ keys,
                                  uint64_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in || i >= *n_unique) return;
    out[i] = keys[i] & ((1ull << 42) - 1);
}

__global__ void marker_genome_offsets(uint32_t n_genomes, const uint32_t* __restrict__ n_unique,
                                      const uint64_t* __restrict__ keys, uint32_t* __restrict__ out) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_genomes) return;
    uint32_t n = *n_unique;
    out[g] = g == n_genomes ? n : lower_bound_u64(keys, n, (uint64_t)g << 42);
}

__global__ void build_buckets_kernel(const GenomeView* __restrict__ views, uint32_t n_genomes) {
    const GenomeView v = views[blockIdx.y];
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > v.n_buckets) return;
    uint32_t* out = const_cast<uint32_t*>(v.bucket);
    if (b == v.n_buckets) { out[b] = v.n_seeds; return; }
    out[b] = lower_bound_u32(v.kmer_k, v.n_seeds, b << v.bucket_shift);
}

__global__ void contig_starts_kernel(const GenomeView* __restrict__ views, uint32_t n_genomes) {
    const GenomeView v = views[blockIdx.y];
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > v.n_contigs) return;
    uint32_t* out = const_cast<uint32_t*>(v.contig_seed_start);
    // meta_p = contig << 1 | canonical, non-decreasing in contig
    uint32_t lo = 0, hi = v.n_seeds;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if ((v.meta_p[mid] >> 1) < c) lo = mid + 1; else hi = mid;
    }
    out[c] = lo;
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// small table "upload" done by the SMs: the source is pinned host memory mapped into the device address space, so the
// transfer does not queue behind bulk copies on the host->device copy engine
__global__ void pull_copy_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, size_t n_words) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace

size_t kmer_order_scratch_bytes(uint32_t n) {
    size_t cub_bytes = 0, seg_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 64);
    cub::DeviceSegmentedRadixSort::SortPairs(nullptr, seg_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                             (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 1 << 22,
                                             (const uint32_t*)nullptr, (const uint32_t*)nullptr, 0, 32);
    if (seg_bytes > cub_bytes) cub_bytes = seg_bytes;
    return align_up(cub_bytes) + 2 * align_up((size_t)n * 8) + 2 * align_up((size_t)n * 4) + 1024;
}

__global__ void iota_kernel(uint32_t* out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
__global__ void gather_pos_meta(uint32_t n, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ pos_p,
                                const uint32_t* __restrict__ meta_p, uint32_t n_genomes, const uint32_t* __restrict__ genome_seed_start,
                                uint32_t* __restrict__ pos_k, uint32_t* __restrict__ meta_k, uint32_t* __restrict__ perm_k) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t src = vals[i];
    pos_k[i] = pos_p[src];
    meta_k[i] = meta_p[src];
    // a segmented sort keeps every seed inside its genome: the genome of k-order slot i is the genome of its source
    if (perm_k) perm_k[i] = src - genome_seed_start[upper_bound_u32(genome_seed_start, n_genomes + 1, i) - 1];
}

void build_kmer_order(const IndexBuildArgs& a, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    const uint32_t n = a.n_seeds_total;
    if (n == 0) return;
    const int T = 256;
    char* p = (char*)scratch;
    if (a.max_genome_seeds <= SEGMENTED_SORT_MAX) {
        // Seeds are already grouped by genome (position order), so the (genome, kmer) order is a SEGMENTED sort by
        // k-mer: 32-bit keys straight from kmer_p, 2k bits, one CTA per genome working in L2.  Stable, so equal k-mers
        // keep their (contig, pos) order.
        uint32_t* vals_in = (uint32_t*)p; p += align_up((size_t)n * 4);
        uint32_t* vals_out = (uint32_t*)p; p += align_up((size_t)n * 4);
        size_t cub_bytes = scratch_bytes - (size_t)(p - (char*)scratch);
        iota_kernel<<<(n + T - 1) / T, T, 0, st>>>(vals_in, n);
        cub::DeviceSegmentedRadixSort::SortPairs(p, cub_bytes, a.kmer_p, a.kmer_k, vals_in, vals_out, (int)n, (int)a.n_genomes,
                                                 a.genome_seed_start, a.genome_seed_start + 1, 0, 2 * a.k, st);
        gather_pos_meta<<<(n + T - 1) / T, T, 0, st>>>(n, vals_out, a.pos_p, a.meta_p, a.n_genomes, a.genome_seed_start, a.pos_k, a.meta_k, a.perm_k);
        g_kernel_launches += 3;
        return;
    }
    uint64_t* keys_in = (uint64_t*)p; p += align_up((size_t)n * 8);
    uint64_t* keys_out = (uint64_t*)p; p += align_up((size_t)n * 8);
    uint32_t* vals_in = (uint32_t*)p; p += align_up((size_t)n * 4);
    uint32_t* vals_out = (uint32_t*)p; p += align_up((size_t)n * 4);
    size_t cub_bytes = scratch_bytes - (size_t)(p - (char*)scratch);
    make_seed_keys<<<(n + T - 1) / T, T, 0, st>>>(n, a.n_genomes, a.genome_seed_start, a.kmer_p, keys_in, vals_in);
    g_kernel_launches++;
    int gbits = 0;
    while ((1ull << gbits) < (uint64_t)a.n_genomes) gbits++;
    const int end_bit = 32 + gbits;   // k-mer bits above 2k are zero; sorting them is harmless for k < 16
    cub::DeviceRadixSort::SortPairs(p, cub_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit, st);
    g_kernel_launches += 1 + (end_bit + 7) / 8;
    gather_kmer_order<<<(n + T - 1) / T, T, 0, st>>>(n, keys_out, vals_out, a.pos_p, a.meta_p, a.genome_seed_start, a.kmer_k, a.pos_k, a.meta_k, a.perm_k);
    g_kernel_launches++;
}

// ------------------------------------------------------------------ k-mer order by bucket partition
// The k-mer order of a genome is (top-B-bit bucket, then k-mer, then position).  Buckets hold ~8 seeds, so instead of a
// 30-bit radix sort the seeds are (1) counted per bucket, (2) the counts scanned per genome — which IS the bucket table
// the anchor lookup needs —, (3) scattered to their bucket with an atomic cursor, and (4) each bucket's handful of
// entries is put in (k-mer, position-index) order by one thread.  The result is identical to the stable radix sort.
namespace {

__device__ __forceinline__ uint32_t genome_of(const BucketGenome* __restrict__ G, uint32_t n_genomes, uint32_t i) {
    uint32_t lo = 0, hi = n_genomes;           // last genome whose seed_start <= i and that is not empty
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (G[mid].seed_start <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void bucket_count_kernel(uint32_t n, uint32_t n_genomes, const BucketGenome* __restrict__ G,
                                    const uint32_t* __restrict__ kmer_p, uint32_t* __restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const BucketGenome g = G[genome_of(G, n_genomes, i)];
    atomicAdd(&counts[g.bucket_off + (__ldg(kmer_p + i) >> g.shift)], 1u);
}

// one CTA per genome: bucket[b] = number of seeds in buckets < b (b = 0..n_buckets), cursor[b] = bucket[b]
__global__ void __launch_bounds__(1024) bucket_scan_kernel(const BucketGenome* __restrict__ G, uint32_t* __restrict__ counts_to_cursor,
                                                           uint32_t* __restrict__ bucket) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const BucketGenome g = G[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 <= g.n_buckets; b0 += 1024) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t v = b < g.n_buckets ? counts_to_cursor[g.bucket_off + b] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        const uint32_t excl = s_carry + s_warp[warp] + inc - v;
        if (b <= g.n_buckets) { bucket[g.bucket_off + b] = excl; counts_to_cursor[g.bucket_off + b] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
}

// record of one seed inside its bucket: (k-mer, index in position order, position, meta)
__global__ void bucket_scatter_kernel(uint32_t n, uint32_t n_genomes, const BucketGenome* __restrict__ G,
                                      const uint32_t* __restrict__ kmer_p, const uint32_t* __restrict__ pos_p,
                                      const uint32_t* __restrict__ meta_p, uint32_t* __restrict__ cursor, uint4* __restrict__ tmp) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const BucketGenome g = G[genome_of(G, n_genomes, i)];
    const uint32_t km = __ldg(kmer_p + i);
    const uint32_t slot = atomicAdd(&cursor[g.bucket_off + (km >> g.shift)], 1u);
    tmp[g.seed_start + slot] = make_uint4(km, i - g.seed_start, __ldg(pos_p + i), __ldg(meta_p + i));
}

constexpr uint32_t BUCKET_RANK_MAX = 256;   // larger buckets (low-complexity genomes) send the batch to the radix-sort path

// one thread per scattered record: its rank inside the bucket = number of bucket entries with a smaller
// (k-mer, position index); the bucket's ~8 records sit in the same few cache lines for all of its threads
__global__ void bucket_rank_kernel(uint32_t n, uint32_t n_genomes, const BucketGenome* __restrict__ G,
                                   const uint32_t* __restrict__ bucket, const uint4* __restrict__ tmp,
                                   uint32_t* __restrict__ kmer_k, uint32_t* __restrict__ pos_k, uint32_t* __restrict__ meta_k,
                                   uint32_t* __restrict__ perm_k, uint32_t* __restrict__ overflow) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const BucketGenome g = G[genome_of(G, n_genomes, t)];
    const uint4 me = tmp[t];
    const uint32_t b = me.x >> g.shift;
    const uint32_t s = bucket[g.bucket_off + b], e = bucket[g.bucket_off + b + 1];
    if (e - s > BUCKET_RANK_MAX) { *overflow = 1u; return; }
    const uint64_t key = ((uint64_t)me.x << 32) | me.y;
    uint32_t rank = 0;
    const uint4* src = tmp + g.seed_start;
    for (uint32_t j = s; j < e; j++) {
        const uint4 o = src[j];
        rank += ((((uint64_t)o.x << 32) | o.y) < key) ? 1u : 0u;
    }
    const uint32_t dst = g.seed_start + s + rank;
    kmer_k[dst] = me.x; pos_k[dst] = me.z; meta_k[dst] = me.w;
    if (perm_k) perm_k[dst] = me.y;            // genome-local index of this seed in position order
}

}  // namespace

size_t bucket_order_scratch_bytes(uint32_t n_seeds, size_t bucket_total) {
    (void)bucket_total;
    return align_up((size_t)n_seeds * 16) + 1024;
}

// genomes_dev: [n_genomes] table; bucket: the store's bucket array [bucket_total]; overflow: device flag (pre-zeroed)
void build_kmer_order_buckets(uint32_t n_seeds, uint32_t n_genomes, const BucketGenome* genomes_dev, size_t bucket_total,
                              uint32_t* counts, int counts_ready, const uint32_t* kmer_p, const uint32_t* pos_p,
                              const uint32_t* meta_p, uint32_t* kmer_k, uint32_t* pos_k, uint32_t* meta_k, uint32_t* perm_k,
                              uint32_t* bucket, uint32_t* overflow, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (n_genomes == 0) return;
    (void)scratch_bytes;
    uint4* tmp = (uint4*)scratch;
    const int T = 256;
    if (counts_ready == 0) {
        cudaMemsetAsync(counts, 0, bucket_total * 4, st);
        if (n_seeds) bucket_count_kernel<<<(n_seeds + T - 1) / T, T, 0, st>>>(n_seeds, n_genomes, genomes_dev, kmer_p, counts);
        g_kernel_launches++;
    }
    bucket_scan_kernel<<<n_genomes, 1024, 0, st>>>(genomes_dev, counts, bucket);       // counts become the scatter cursors
    if (n_seeds) {
        bucket_scatter_kernel<<<(n_seeds + T - 1) / T, T, 0, st>>>(n_seeds, n_genomes, genomes_dev, kmer_p, pos_p, meta_p, counts, tmp);
        bucket_rank_kernel<<<(n_seeds + T - 1) / T, T, 0, st>>>(n_seeds, n_genomes, genomes_dev, bucket, tmp, kmer_k, pos_k, meta_k, perm_k, overflow);
    }
    g_kernel_launches += 3;
}

// ---- marker sets of a batch whose genomes have at most MARKER_SMEM_MAX markers each (16 Mbp at c = 1000): one CTA sorts a
// genome's markers in shared memory (bitonic), drops duplicates and leaves the result at the genome's pre-deduplication
// offset; a one-CTA scan turns the per-genome counts into the final offsets and a copy kernel closes the gaps.  Three
// launches instead of the dozen of the segmented radix sort + select path, which remains for larger genomes.
constexpr uint32_t MARKER_SMEM_MAX = 16384;

__global__ void __launch_bounds__(1024) marker_sort_smem_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ off_in,
                                                                uint64_t* __restrict__ sorted_unique, uint32_t* __restrict__ n_unique) {
    extern __shared__ uint64_t s_key[];
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_base;
    const uint32_t g = blockIdx.x, o = off_in[g], n = off_in[g + 1] - o;
    if (n == 0) { if (threadIdx.x == 0) n_unique[g] = 0; return; }
    uint32_t P = 2;
    while (P < n) P <<= 1;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) s_key[i] = i < n ? (keys_in[o + i] & ((1ull << 42) - 1)) : ~0ull;
    __syncthreads();
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const uint32_t i = 2 * t - (t & (j - 1)), l = i + j;        // i has bit j clear
                const uint64_t a = s_key[i], b = s_key[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { s_key[i] = b; s_key[l] = a; }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        const bool keep = i < n && (i == 0 || s_key[i] != s_key[i - 1]);
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { const uint32_t c = s_warp[w]; if (w < warp) before += c; total += c; }
        const uint32_t base = s_base;
        if (keep) sorted_unique[o + base + before + __popc(bal & ((1u << lane) - 1u))] = s_key[i];
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_unique[g] = s_base;
}

// out[g] = sum of counts[0..g) for g = 0..n (one CTA; n is the number of genomes of a batch)
__global__ void __launch_bounds__(1024) exclusive_offsets_kernel(const uint32_t* __restrict__ counts, uint32_t n, uint32_t* __restrict__ out) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t i0 = 0; i0 < n; i0 += 1024) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t v = i < n ? counts[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < 32; w++) { const uint32_t c = s_warp[w]; if (w < warp) before += c; total += c; }
        const uint32_t carry = s_carry;
        if (i < n) out[i] = carry + before + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_carry;
}

__global__ void marker_compact_kernel(const uint64_t* __restrict__ sorted_unique, const uint32_t* __restrict__ off_in,
                                      const uint32_t* __restrict__ off_out, uint64_t* __restrict__ markers_out) {
    const uint32_t g = blockIdx.x, src = off_in[g], dst = off_out[g], n = off_out[g + 1] - dst;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) markers_out[dst + i] = sorted_unique[src + i];
}

size_t marker_scratch_bytes(uint32_t n, uint32_t n_genomes) {
    size_t s1 = 0, s2 = 0, s3 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, s1, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)n, 0, 64);
    cub::DeviceSegmentedRadixSort::SortKeys(nullptr, s3, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)n, 1 << 22,
                                            (const uint32_t*)nullptr, (const uint32_t*)nullptr, 0, 64);
    if (s3 > s1) s1 = s3;
    cub::DeviceSelect::Unique(nullptr, s2, (const uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return align_up(s1 > s2 ? s1 : s2) + 2 * align_up((size_t)n * 8 + 8) + 256 + 1024 + align_up(4 * ((size_t)n_genomes + 1));
}

void build_marker_sets(uint32_t n_genomes, uint32_t n, uint64_t* marker_keys, uint64_t* markers_out,
                       uint32_t* genome_marker_out, const uint32_t* genome_marker_in, uint32_t max_genome_markers,
                       void* scratch, size_t scratch_bytes, cudaStream_t st) {
    const int T = 256;
    char* p = (char*)scratch;
    uint64_t* sorted = (uint64_t*)p; p += align_up((size_t)n * 8 + 8);
    uint64_t* uniq = (uint64_t*)p; p += align_up((size_t)n * 8 + 8);
    uint32_t* n_unique = (uint32_t*)p; p += 256;
    size_t cub_bytes = scratch_bytes - (size_t)(p - (char*)scratch);
    if (n == 0) {
        cudaMemsetAsync(genome_marker_out, 0, sizeof(uint32_t) * (n_genomes + 1), st);
        return;
    }
    if (genome_marker_in && max_genome_markers <= MARKER_SMEM_MAX && n_genomes <= 65535u * 16u) {
        uint32_t* per_genome = (uint32_t*)((char*)scratch + scratch_bytes - align_up(4 * ((size_t)n_genomes + 1)));
        uint32_t P = 2;
        while (P < max_genome_markers) P <<= 1;
        static_assert(MARKER_SMEM_MAX * 8 <= 200 * 1024, "marker sort tile must fit the opted-in shared memory");
        cudaFuncSetAttribute(marker_sort_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MARKER_SMEM_MAX * 8));
        marker_sort_smem_kernel<<<n_genomes, P >= 2048 ? 1024 : 256, (size_t)P * 8, st>>>(marker_keys, genome_marker_in, sorted, per_genome);
        exclusive_offsets_kernel<<<1, 1024, 0, st>>>(per_genome, n_genomes, genome_marker_out);
        marker_compact_kernel<<<n_genomes, 256, 0, st>>>(sorted, genome_marker_in, genome_marker_out, markers_out);
        g_kernel_launches += 3;
        return;
    }
    int gbits = 0;
    while ((1ull << gbits) < (uint64_t)n_genomes) gbits++;
    if (genome_marker_in && max_genome_markers <= SEGMENTED_SORT_MAX) {
        // marker keys arrive grouped by genome: sort the 42 marker bits inside each genome's segment
        cub::DeviceSegmentedRadixSort::SortKeys(p, cub_bytes, marker_keys, sorted, (int)n, (int)n_genomes, genome_marker_in,
                                                genome_marker_in + 1, 0, 42, st);
        g_kernel_launches += 1;
    } else {
        cub::DeviceRadixSort::SortKeys(p, cub_bytes, marker_keys, sorted, (int)n, 0, 42 + gbits, st);
        g_kernel_launches += 1 + (42 + gbits + 7) / 8;
    }
    cub::DeviceSelect::Unique(p, cub_bytes, sorted, uniq, n_unique, (int)n, st);
    g_kernel_launches += 2;
    strip_marker_keys<<<(n + T - 1) / T, T, 0, st>>>(n, n_unique, uniq, markers_out);
    marker_genome_offsets<<<(n_genomes + 1 + T - 1) / T, T, 0, st>>>(n_genomes, n_unique, uniq, genome_marker_out);
    g_kernel_launches += 2;
}

void launch_build_buckets(const GenomeView* views_dev, uint32_t n_genomes, uint32_t max_buckets, cudaStream_t st) {
    if (n_genomes == 0) return;
    const int T = 256;
    for (uint32_t g0 = 0; g0 < n_genomes; g0 += 65535) {
        uint32_t ng = n_genomes - g0 < 65535 ? n_genomes - g0 : 65535;
        dim3 grid((max_buckets + 1 + T - 1) / T, ng);
        build_buckets_kernel<<<grid, T, 0, st>>>(views_dev + g0, ng);
        g_kernel_launches++;
    }
}

void launch_contig_starts(const GenomeView* views_dev, uint32_t n_genomes, uint32_t max_contigs, cudaStream_t st) {
    if (n_genomes == 0) return;
    const int T = 128;
    for (uint32_t g0 = 0; g0 < n_genomes; g0 += 65535) {
        uint32_t ng = n_genomes - g0 < 65535 ? n_genomes - g0 : 65535;
        dim3 grid((max_contigs + 1 + T - 1) / T, ng);
        contig_starts_kernel<<<grid, T, 0, st>>>(views_dev + g0, ng);
        g_kernel_launches++;
    }
}

// indices of the set flags, ascending, plus their count
void select_passing(uint32_t n, const uint8_t* flags, uint32_t* out_idx, uint32_t* out_count, cudaStream_t st) {
    cub::CountingInputIterator<uint32_t> it(0);
    size_t bytes = 0;
    cub::DeviceSelect::Flagged(nullptr, bytes, it, flags, out_idx, out_count, (int)n, st);
    void* tmp = nullptr;
    cudaMallocAsync(&tmp, bytes + 16, st);
    cub::DeviceSelect::Flagged(tmp, bytes, it, flags, out_idx, out_count, (int)n, st);
    cudaFreeAsync(tmp, st);
    g_kernel_launches += 2;
}

// ---------------------------------------------------------------- device-to-device packing of sketches (multi-GPU exchange)
__global__ void segment_copy_kernel(const SegmentCopy* __restrict__ segs, uint32_t n_segs, char* __restrict__ dst_base) {
    // one CTA per 64 KB piece of a segment; segments are 4-byte granular (sources are only 4-byte aligned)
    const uint32_t s = blockIdx.y;
    if (s >= n_segs) return;
    const SegmentCopy sg = segs[s];
    const uint64_t words = sg.bytes >> 2;
    const uint32_t* src = (const uint32_t*)sg.src;
    uint32_t* dst = (uint32_t*)(dst_base + sg.dst_off);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

void launch_segment_copy(const SegmentCopy* d_segs, uint32_t n_segs, uint64_t max_bytes, void* dst_base, cudaStream_t st) {
    if (n_segs == 0) return;
    unsigned gx = (unsigned)((max_bytes / 4 + 256 * 16 - 1) / (256 * 16));
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    for (uint32_t s0 = 0; s0 < n_segs; s0 += 65535) {
        const uint32_t cnt = n_segs - s0 < 65535 ? n_segs - s0 : 65535;
        segment_copy_kernel<<<dim3(gx, cnt), 256, 0, st>>>(d_segs + s0, cnt, (char*)dst_base);
        g_kernel_launches++;
    }
}

void launch_pull_copy(void* dst, const void* src_pinned, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return;
    const size_t n = (bytes + 3) / 4;
    unsigned grid = (unsigned)((n + 255) / 256);
    if (grid > 1024) grid = 1024;
    pull_copy_kernel<<<grid, 256, 0, st>>>((uint32_t*)dst, (const uint32_t*)src_pinned, n);
    g_kernel_launches++;
}

size_t region_scan_scratch_bytes(uint32_t n) {
    size_t s = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, s, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)n + 1);
    return align_up(s) + 256;
}

void scan_region_counts(uint32_t n_regions, const uint64_t* region_cnt, uint64_t* region_start, void* scratch, size_t scratch_bytes,
                        cudaStream_t st) {
    // region_cnt[n_regions] is zero, so region_start[n_regions] is the batch total; 32-bit halves cannot carry into each
    // other because both totals are < 2^32
    cub::DeviceScan::ExclusiveSum(scratch, scratch_bytes, region_cnt, region_start, (int)n_regions + 1, st);
    g_kernel_launches += 2;
}

size_t scan_scratch_bytes(uint32_t n) {
    size_t s = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, s, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return align_up(s) + 256;
}

size_t sort_pairs_scratch_bytes(uint32_t n) {
    size_t s = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, s, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n, 0, 64);
    return align_up(s) + 256;
}

void sort_window_keys(uint32_t n, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in,
                      uint32_t* vals_out, int end_bit, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (n == 0) return;
    cub::DeviceRadixSort::SortPairs(scratch, scratch_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit, st);
    g_kernel_launches += 1 + (end_bit + 7) / 8;
}

struct MatchCount { __host__ __device__ __forceinline__ uint32_t operator()(const uint2& fc) const { return fc.y; } };
void scan_match_counts(const ChainBatch& b, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    // a_off[i] = sum of m_fc[0..i).y; the trailing element m_fc[n] is zero-initialised by the caller, so
    // a_off[n] is the total number of anchors
    auto counts = thrust::make_transform_iterator((const uint2*)b.m_fc, MatchCount());
    cub::DeviceScan::ExclusiveSum(scratch, scratch_bytes, counts, b.a_off, (int)b.n_qseeds_total + 1, st);
    g_kernel_launches += 2;
}

}  // namespace skb
