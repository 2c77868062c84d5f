// skb_internal.cuh — shared declarations of libskb's translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/skb.h"
#include "gbdt_model.h"

namespace skb {

// ---------------------------------------------------------------- device-side views
// One sketched genome as the kernels see it.  Two orders of the same seeds:
//   *_p : position order  (contig, pos)        — emitted directly by the seeding kernel
//   *_k : k-mer order     (kmer, contig, pos)  — the "device-sorted hash->position array" that replaces
//                                                 skani's FxHashMap seed index
// meta = contig_index << 1 | canonical.
struct GenomeView {
    const uint32_t* kmer_p;
    const uint32_t* pos_p;
    const uint32_t* meta_p;
    const uint32_t* kmer_k;
    const uint32_t* pos_k;
    const uint32_t* meta_k;
    const uint32_t* perm_k;        // k-order slot -> index of the same seed in position order (query side of the anchor join)
    const uint32_t* bucket;        // [n_buckets + 1] offsets into *_k by top bits of the k-mer
    const uint32_t* contig_seed_start;  // [n_contigs + 1] offsets into *_p
    const uint32_t* contig_len;    // [n_contigs]
    const uint32_t* contig_win_start;   // [n_contigs + 1] prefix sums of per-contig window capacity (len / 20000 + 1)
    const uint64_t* markers;       // sorted unique canonical 21-mers
    uint64_t total_len;
    uint32_t n_seeds;
    uint32_t n_markers;
    uint32_t n_contigs;
    uint32_t bucket_shift;         // bucket id = kmer >> bucket_shift
    uint32_t n_buckets;
    uint32_t win_cap;              // upper bound on 20 kb windows when this genome is the query
    uint32_t ctg_q90, ctg_q50, ctg_q10;   // contig-length quantiles (features of the learned-ANI model)
};

// ---------------------------------------------------------------- seeding
#ifndef SKB_SEED_THREADS
#define SKB_SEED_THREADS 256
#endif
constexpr int SEED_THREADS = SKB_SEED_THREADS;
constexpr int SEED_WARPS = SEED_THREADS / 32;
constexpr int WORDS_PER_LANE = 4;                                     // 16-base words each lane evaluates per tile
constexpr int TILE_WORDS = 32 * WORDS_PER_LANE;                       // 128 words per warp tile
constexpr int TILE_BASES = TILE_WORDS * 16;                           // 2048 bases per warp tile
constexpr int CHUNK_TILES = 8;                                        // most tiles per dynamically claimed region (16 kbp)

// One kept contig of the batch.  Tiles are numbered contig after contig; tile t of a contig covers its bases
// [t * TILE_BASES, min(len, (t + 1) * TILE_BASES)).
struct ContigDesc {
    uint64_t seq_off;    // byte offset of the contig's first base in the device sequence buffer (16-aligned)
    uint32_t len;
    uint32_t contig;     // index among the genome's kept contigs
    uint32_t genome;     // genome index inside the batch; bit 31 set on the genome's first kept contig
    uint32_t tile_start; // batch-wide id of the contig's first tile
};

// A seeding launch is cut into regions of CHUNK_TILES consecutive tiles; warps claim regions from an atomic counter and
// fill the region's private, ordered storage.  Regions are numbered launch after launch in tile order.
struct SeedScanArgs {
    const uint8_t* seq;
    const ContigDesc* contigs;   // descriptors of THIS launch (a contiguous slice of the batch's table)
    uint32_t n_contigs;
    uint32_t n_tiles;            // tiles of this launch
    uint32_t tile_base;          // batch-wide id of this launch's first tile
    uint32_t region_base;        // id of this launch's first region
    uint32_t n_chunks;           // regions of this launch = ceil(n_tiles / chunk_tiles)
    uint32_t chunk_tiles;        // tiles per region: CHUNK_TILES for large launches, fewer when that would leave warps idle
    uint32_t* chunk_counter;     // zero-initialised claim counter of this launch
    uint32_t n_warps;            // warps of this launch (grid * SEED_WARPS)
    uint32_t kmask, kshift;
    uint64_t thr_seed, thr_marker;
    uint64_t chk_seed, chk_marker;   // thresholds of the exact re-check at write-out (= thr_*; a test hook lowers them)
    // region storage: the region of a warp whose first tile is T starts at T * seed_tile_cap (resp. marker_tile_cap)
    // and may hold (tiles of the warp) * cap records, unless explicit tables are given (retry after an overflow)
    uint32_t seed_tile_cap, marker_tile_cap;
    const uint64_t* region_off;                                          // optional [n_regions + 1] exact layout (seeds | markers << 32)
    uint32_t* kmer_r; uint32_t* pos_r; uint32_t* meta_r;                 // region storage, seeds
    uint64_t* marker_r;                                                  // region storage, genome << 42 | 21-mer
    uint32_t* region_seed_src; uint32_t* region_marker_src;              // [n_regions] where each region's storage begins
    uint64_t* region_cnt;                                                // [n_regions + 1] exact counts, seeds | markers << 32 (also on overflow)
    uint32_t* genome_region;     // [n_genomes] region in which the genome's first tile lies (0xFFFFFFFF = no tile)
    uint32_t* genome_seed_local; uint32_t* genome_marker_local;          // [n_genomes] cursor of that region at that tile
    uint32_t* overflow;          // bit 0: a region was too small; bit 1: the high-word hash comparison let a position through
                                 // that the exact comparison rejects (the host repeats the batch with exact_compare = 1)
    uint32_t exact_compare;      // 0: compare the high words of hash and threshold, re-check hits when they are written
    uint32_t packed;             // 1: seq holds 2-bit words (contig at byte seq_off / 4) instead of ASCII (contig at byte seq_off)
    uint32_t fma_m1, fma_two;    // the constants 2^32 - 1 and 2 as launch parameters: multipliers the assembler cannot fold, which
                                 // keep the hit-mask arithmetic of the main loop on the FMA pipe (seed_kernels.cu, SKB_MASK_FMA)
};

// region_start[n_regions + 1] = exclusive scan of region_cnt (u64 lanes: seeds | markers << 32)
void scan_region_counts(uint32_t n_regions, const uint64_t* region_cnt, uint64_t* region_start, void* scratch, size_t scratch_bytes,
                        cudaStream_t st);
size_t region_scan_scratch_bytes(uint32_t n_regions);
void launch_genome_starts(uint32_t n_regions, const uint64_t* region_start, uint32_t n_genomes, const uint32_t* genome_region,
                          const uint32_t* genome_seed_local, const uint32_t* genome_marker_local, uint32_t* genome_seed_start,
                          uint32_t* genome_marker_start, cudaStream_t st);
// copies every region's records to their final, contiguous place (warp per region)
struct BucketGenome {
    uint32_t seed_start;   // first seed of the genome in the batch arrays
    uint32_t n_seeds;
    uint32_t shift;        // bucket id = kmer >> shift
    uint32_t n_buckets;
    uint32_t bucket_off;   // offset of the genome's (n_buckets + 1) entries in the batch's bucket array
};
struct RegionGatherArgs {
    uint32_t n_regions;
    const uint32_t* seed_src; const uint32_t* marker_src;       // [n_regions] region storage offsets
    const uint64_t* region_start;                               // [n_regions + 1] destinations (seeds | markers << 32)
    const uint32_t* kmer_r; const uint32_t* pos_r; const uint32_t* meta_r; const uint64_t* marker_r;
    uint32_t* kmer_p; uint32_t* pos_p; uint32_t* meta_p; uint64_t* marker_keys;
    // optional: while the seeds pass through, count them per (genome, k-mer bucket) for the index build
    const BucketGenome* genomes; uint32_t n_genomes; uint32_t* bucket_counts;     // bucket_counts == NULL: no histogram
};
void launch_region_gather(const RegionGatherArgs& a, cudaStream_t st);
void launch_counters_to_host(const uint32_t* seed_start, const uint32_t* marker_start, uint32_t n, const uint32_t* overflow,
                             uint32_t* host_out, cudaStream_t st);

void launch_seed_scan(const SeedScanArgs& a, int n_sm, cudaStream_t st);

// ---------------------------------------------------------------- index build (sort by k-mer, buckets, marker sets)
struct IndexBuildArgs {
    uint32_t n_genomes;
    uint32_t n_seeds_total, n_markers_total;
    uint32_t max_genome_seeds;            // largest per-genome seed count of the batch (chooses the sort strategy)
    const uint32_t* genome_seed_start;    // device [n_genomes+1]
    const uint32_t* kmer_p; const uint32_t* pos_p; const uint32_t* meta_p;
    uint32_t* kmer_k; uint32_t* pos_k; uint32_t* meta_k;
    uint32_t* perm_k;                     // optional: k-order slot -> genome-local position-order index
    int k;
};

constexpr uint32_t SEGMENTED_SORT_MAX = 262144;   // genomes up to ~32 Mbp are sorted by one CTA each

// sorts (genome,kmer) keys carrying the position-order index, then gathers the *_k arrays
void build_kmer_order(const IndexBuildArgs& a, void* scratch, size_t scratch_bytes, cudaStream_t st);
size_t kmer_order_scratch_bytes(uint32_t n_seeds_total);

// k-mer order through bucket partition (fast path of the index build); see index_kernels.cu
size_t bucket_order_scratch_bytes(uint32_t n_seeds, size_t bucket_total);
// counts (device, [bucket_total]): bucket histogram scratch; counts_ready != 0 means the caller already filled it,
// otherwise it is built here (measured: a separate histogram pass, 33 us per 4 M seeds, beats atomics inside the
// seeding kernel, +47 us).  It is consumed (turned into cursors).
void build_kmer_order_buckets(uint32_t n_seeds, uint32_t n_genomes, const BucketGenome* genomes_dev, size_t bucket_total,
                              uint32_t* counts, int counts_ready, const uint32_t* kmer_p, const uint32_t* pos_p,
                              const uint32_t* meta_p, uint32_t* kmer_k, uint32_t* pos_k, uint32_t* meta_k, uint32_t* perm_k,
                              uint32_t* bucket, uint32_t* overflow, void* scratch, size_t scratch_bytes, cudaStream_t st);

// sorts marker keys and removes duplicates per genome; writes marker values (42-bit) to markers_out and the
// per-genome offsets [n_genomes+1] to genome_marker_out (device)
// genome_marker_in (device, [n_genomes + 1], may be NULL): offsets of each genome's keys before de-duplication
void build_marker_sets(uint32_t n_genomes, uint32_t n_markers_total, uint64_t* marker_keys, uint64_t* markers_out,
                       uint32_t* genome_marker_out, const uint32_t* genome_marker_in, uint32_t max_genome_markers,
                       void* scratch, size_t scratch_bytes, cudaStream_t st);
size_t marker_scratch_bytes(uint32_t n_markers_total, uint32_t n_genomes);

// copies `bytes` (rounded up to 4) from pinned, device-mapped host memory to device memory with a kernel
struct SegmentCopy { const void* src; uint64_t dst_off; uint64_t bytes; };   // bytes: multiple of 4
void launch_segment_copy(const SegmentCopy* d_segs, uint32_t n_segs, uint64_t max_bytes, void* dst_base, cudaStream_t st);
void launch_pull_copy(void* dst, const void* src_pinned, size_t bytes, cudaStream_t st);

// bucket offsets + per-contig seed starts for a set of genomes described by views
void launch_build_buckets(const GenomeView* views_dev, uint32_t n_genomes, uint32_t max_buckets, cudaStream_t st);
void launch_contig_starts(const GenomeView* views_dev, uint32_t n_genomes, uint32_t max_contigs, cudaStream_t st);

// ---------------------------------------------------------------- screen
// count[q * n_refs + r] = | markers(q) ∩ markers(r) |
void launch_marker_screen(const GenomeView* queries, uint32_t n_queries, const GenomeView* refs, uint32_t n_refs,
                          uint32_t* count, const uint32_t* query_markers_host, int n_sm, cudaStream_t st);
void launch_screen_decide(const GenomeView* queries, uint32_t n_queries, const GenomeView* refs, uint32_t n_refs,
                          const uint32_t* count, double p21, int always, int rescue_small, uint8_t* pass,
                          cudaStream_t st);

constexpr int MARKER_BITS = 42;   // a marker is a canonical 21-mer
// marker index of a database: all (marker, genome) postings sorted by marker + bucket table on the top bits
size_t marker_index_scratch_bytes(uint32_t n_postings, uint32_t n_buckets);
// use_sort = false: bucket partition (may raise *overflow when a bucket is too large: call again with use_sort = true)
void build_marker_index(const GenomeView* refs, uint32_t n_refs, const uint32_t* genome_off, uint32_t n_postings,
                        uint64_t* keys, uint32_t* vals, uint32_t* bucket, uint32_t shift, uint32_t n_buckets,
                        void* scratch, size_t scratch_bytes, bool use_sort, uint32_t* overflow, cudaStream_t st);
void launch_marker_join(const GenomeView* queries, uint32_t n_queries, uint32_t n_refs, const uint64_t* keys,
                        const uint32_t* vals, const uint32_t* bucket, uint32_t shift, uint32_t* count, cudaStream_t st);

// ---------------------------------------------------------------- chain
// skani's chaining constants (frozen: pyskani exposes none of them).  The DP kernel uses them as compile-time values;
// skb_db_query fills ChainConsts from the same definitions.
constexpr int32_t DP_ANCHOR_SCORE = 20, DP_MAX_GAP = 300;
constexpr uint32_t DP_BP_BAND = 2500, DP_INDEX_BAND = 100;

struct ChainConsts {           // skani::chain::map_params_from_sketch (reference lib.rs:646-651), frozen per DESIGN.md
    uint32_t fragment_length;  // 20000
    int32_t anchor_score;      // 20
    int32_t min_anchors;       // 3
    int32_t min_score;         // 45
    int32_t max_gap;           // 300
    int32_t index_band;        // 100
    int32_t bp_band;           // 2500
    int32_t af_ext;            // 198
    double frac_cover_cutoff;  // 0.15
    int32_t robust, median;
    int32_t k;
    // learned-ANI correction (skani::regression, reference lib.rs:611-614): applied to the default (mean) estimate of pairs
    // with at least learned_min_cov aligned query bases when use_model != 0
    int32_t use_model;
    double learned_min_cov;    // 150000
    GbdtView model;            // device pointers
};

struct PairDesc {
    uint32_t q, r;             // indices into the query / reference view arrays
    uint32_t seed_off;         // offset of this pair's slice in the per-query-seed scratch arrays
    uint32_t win_off;          // offset of this pair's slice in the window arrays
    uint32_t bits_off;         // word offset of this pair's slice in the match bitmask (ceil(n_seeds / 32) words, 16-byte aligned)
};

struct WindowRec {             // one 20 kb query window of one pair
    uint32_t anchors;          // anchors in kept chains (component sizes)
    uint32_t seeds;            // query seeds between the first and the last kept chain coordinate
    uint32_t cov_q, cov_r;     // sum over kept chains of (span + af_ext)
    uint32_t n_chains;
};

struct PairResult {
    float ani, af_q, af_r;
    uint32_t n_windows, n_chains, n_anchors;
};

struct ChainBatch {            // device pointers of one batch of pairs
    const GenomeView* qviews; const GenomeView* rviews;
    const PairDesc* pairs; uint32_t n_pairs;
    uint32_t n_qseeds_total;   // sum of query seeds over pairs
    uint32_t n_win_total;      // sum of window capacities over pairs
    // per query seed of each pair
    uint2* m_fc;               // (first matching index in the reference's k-mer order, number of matches); zeroed per batch
    uint32_t* m_bits;          // bit i of a pair's slice = query seed i has at least one match
    const uint2* walk_groups;  // (first pair, count): consecutive pairs with the same query, walked by one CTA
    uint32_t n_walk_groups, walk_group_max;
    uint32_t* a_off;           // exclusive scan of m_cnt (+1 trailing element = total)
    unsigned long long* a_total64;   // the same total accumulated in 64 bits (guards the 32-bit scan against wrap-around)
    // anchors
    uint32_t anchor_cap;
    uint4* a_rec;              // one 16-byte record per anchor: (q_pos, r_pos, ref contig << 1 | reverse, query seed index)
    int32_t* a_f; uint32_t* a_root; uint32_t* a_aux;                    // DP score, component root, per-root size | flags
    unsigned long long* a_best;                                         // per-root best (score << 32 | ~index)
    uint32_t* big_list; uint32_t* big_count;   // window slots with more anchors than the thread-per-window DP takes
    uint32_t* win_bins; uint32_t* win_order;   // counting sort of the other windows by anchor count (descending)
    // windows
    uint32_t* win_start;       // [n_win_total] first query-seed index (pair-local) of each window
    uint32_t* win_end;         // [n_win_total] one past the last query-seed index
    uint32_t* win_contig;      // query contig
    uint32_t* pair_nwin;       // [n_pairs]
    WindowRec* win_rec;        // [n_win_total]
    uint64_t* sort_keys; uint32_t* sort_vals;   // [n_win_total]
    PairResult* results;       // [n_pairs]
};

void launch_match_count(const ChainBatch& b, cudaStream_t st);
void launch_anchor_fill(const ChainBatch& b, cudaStream_t st);
uint32_t walk_group_capacity(uint32_t max_query_seeds);
void launch_window_walk(const ChainBatch& b, const ChainConsts& c, uint32_t max_query_seeds, cudaStream_t st);
void launch_chain_dp(const ChainBatch& b, const ChainConsts& c, int n_sm, cudaStream_t st);
void launch_window_keys(const ChainBatch& b, cudaStream_t st);
void launch_ani_reduce(const ChainBatch& b, const ChainConsts& c, const uint64_t* sorted_keys,
                       const uint32_t* sorted_vals, cudaStream_t st);
void launch_gbdt_predict(const GbdtView& m, const float* rows, uint32_t n_rows, uint32_t stride, float* out, cudaStream_t st);
// exclusive scan of m_cnt into a_off (n+1 outputs); sort of window keys
void scan_match_counts(const ChainBatch& b, void* scratch, size_t scratch_bytes, cudaStream_t st);
size_t scan_scratch_bytes(uint32_t n);
void sort_window_keys(uint32_t n, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in,
                      uint32_t* vals_out, int end_bit, void* scratch, size_t scratch_bytes, cudaStream_t st);
size_t sort_pairs_scratch_bytes(uint32_t n);

extern unsigned long long g_kernel_launches;   // incremented by every launch_* wrapper

}  // namespace skb
