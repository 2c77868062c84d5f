// skb_api.cu — host side of libskb.so: the C ABI declared in include/skb.h.
//
// Orchestrates the kernels of seed_kernels.cu / index_kernels.cu / screen_kernels.cu / chain_kernels.cu on one
// CUDA stream per context.  All device memory comes from the stream-ordered allocator (cudaMallocAsync) with
// an unbounded release threshold, so scratch is recycled between calls instead of hitting cudaMalloc.
// There is deliberately no CPU implementation of any step in this file.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>

#include "host_pack.h"
#include "skb_internal.cuh"
#include "slab_pool.h"

namespace skb {

void launch_screen_decide(const GenomeView*, uint32_t, const GenomeView*, uint32_t, const uint32_t*, double, int, int,
                          uint8_t*, cudaStream_t);
void select_passing(uint32_t n, const uint8_t* flags, uint32_t* out_idx, uint32_t* out_count, cudaStream_t st);

// Storage of the sketches themselves: slabs from cudaMalloc with bump allocation (slab_pool.h)
struct CudaRawAlloc {
    void* operator()(size_t bytes) const {
        void* p = nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        const cudaError_t e = cudaMalloc(&p, bytes);
        if (std::getenv("SKB_TRACE"))
            std::fprintf(stderr, "[skb] slab: cudaMalloc of %zu MB took %.3f ms\n", bytes >> 20,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return p;
    }
};
struct CudaRawFree { void operator()(void* p) const { cudaFree(p); } };
using SlabPool = SlabPoolT<CudaRawAlloc, CudaRawFree>;

struct Core {
    int device = 0;
    SlabPool slabs;
    std::mutex slab_mu;       // sketches may be released by a thread that works on ANOTHER context of the same device (a
                              // database there holds the last reference): the pool has its own lock
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::string err;
    skb_stats_t stats{};
    cudaEvent_t ev[8]{};
    void* pinned = nullptr;        // staging for small contigs / tables
    size_t pinned_bytes = 0;
    uint32_t* pinned_cnt = nullptr; // small results written by kernels straight into host memory
    size_t pinned_cnt_words = 0;
    uint32_t* pinned_counts(size_t words);
    // grow-only scratch blocks reused by every call on this context (calls are serialised by `mu`): keeps the big
    // transient buffers out of the allocator so that repeated batches never re-map device memory
    struct Block { void* p = nullptr; size_t bytes = 0; };
    Block arena[20];
    void* scratch(int slot, size_t bytes);
    cudaStream_t copy_stream = nullptr;       // host->device copies of skb_sketch_batch run here, ahead of the kernels
    cudaStream_t aux_stream = nullptr;        // marker-set build of a batch, beside the k-mer order build on `stream`
    std::vector<cudaEvent_t> ev_pool;         // "chunk is on the device" events (timing disabled)
    // ingest pipeline of large host batches (host_pack.h): worker threads + pinned staging for the 2-bit chunks
    std::unique_ptr<HostTeam> team;
    uint64_t chain_batch_seeds = 160ull << 20;   // query seeds per chaining batch (set from the device's memory size)
    int host_threads = -1;                    // -1: default (SKB_HOST_THREADS, else min(32, cpus / LOCAL_WORLD_SIZE)); 0: off
    void* pack_stage = nullptr;               // pinned, mirrors the device layout of a batch at a quarter of its size
    size_t pack_stage_bytes = 0;
    void* raw_stage = nullptr;                // pinned staging of the pipeline's DMA thread for small contigs
    // which route pinned sources take is learned per context: best recent input GB/s of the mixed policy [0], of packing
    // every chunk [1] and of plain copies [2]; the fastest is used, the others re-measured in turn every 16th large call
    double ingest_rate[3] = {0.0, 0.0, 0.0};
    uint32_t ingest_calls = 0;
    int ingest_best = 0;
    cudaEvent_t pool_event(size_t i) {
        while (ev_pool.size() <= i) {
            cudaEvent_t e;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) return nullptr;   // waited for by the ingest feeder: sleep, do not spin
            ev_pool.push_back(e);
        }
        return ev_pool[i];
    }
    ~Core() {
        team.reset();
        cudaSetDevice(device);
        if (pack_stage) cudaFreeHost(pack_stage);
        if (raw_stage) cudaFreeHost(raw_stage);
        if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
        if (aux_stream) { cudaStreamSynchronize(aux_stream); cudaStreamDestroy(aux_stream); }
        if (stream) cudaStreamSynchronize(stream);
        for (auto& b : arena) if (b.p) cudaFree(b.p);
        slabs.destroy();
        for (auto& e : ev_pool) cudaEventDestroy(e);
        if (stream) cudaStreamSynchronize(stream);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (pinned) cudaFreeHost(pinned);
        if (pinned_cnt) cudaFreeHost(pinned_cnt);
        if (stream) cudaStreamDestroy(stream);
    }
};

// SKB_TRACE=1 prints host-side wall-clock marks (debug aid, stderr)
struct Trace {
    bool on; std::chrono::steady_clock::time_point t0; const char* what;
    explicit Trace(const char* w) : on(std::getenv("SKB_TRACE") != nullptr), t0(std::chrono::steady_clock::now()), what(w) {}
    void mark(const char* label) {
        if (!on) return;
        auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[skb] %s: %-28s %8.3f ms\n", what, label, std::chrono::duration<double, std::milli>(t - t0).count());
    }
};

struct Fail {
    int code;
    std::string msg;
};

#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            throw Fail{_e == cudaErrorMemoryAllocation ? SKB_ERR_NOMEM : SKB_ERR_CUDA,              \
                       std::string(#expr) + ": " + cudaGetErrorString(_e)};                        \
    } while (0)

uint32_t* Core::pinned_counts(size_t words) {
    if (pinned_cnt_words < words) {
        if (pinned_cnt) { CU(cudaStreamSynchronize(stream)); cudaFreeHost(pinned_cnt); pinned_cnt = nullptr; pinned_cnt_words = 0; }
        const size_t want = std::max<size_t>(words + words / 2, 4096);
        CU(cudaHostAlloc((void**)&pinned_cnt, 4 * want, cudaHostAllocDefault));
        pinned_cnt_words = want;
    }
    return pinned_cnt;
}

enum { SLOT_SEQ = 0, SLOT_KMER, SLOT_POS, SLOT_META, SLOT_MKEYS, SLOT_STATUS, SLOT_SORT, SLOT_MARK, SLOT_GS, SLOT_GM, SLOT_DESC, SLOT_MKEYS2, SLOT_RSCAN, SLOT_GM_IN, SLOT_BTAB, SLOT_BCOUNT, SLOT_CHAIN, SLOT_MIDX, SLOT_PACK, SLOT_SEQPK };

void* Core::scratch(int slot, size_t bytes) {
    Block& b = arena[slot];
    if (slot == SLOT_CHAIN && bytes > ((size_t)1 << 20)) {
        // test hook: pretend the device has no room for a chaining arena above this many MB (exercises the smaller-batch retry)
        if (const char* e = std::getenv("SKB_TEST_CHAIN_ARENA_MB"))
            if (bytes > ((size_t)std::strtoull(e, nullptr, 10) << 20)) throw Fail{SKB_ERR_NOMEM, "chaining arena above the test limit"};
    }
    if (b.bytes < bytes) {
        if (b.p) { CU(cudaStreamSynchronize(stream)); CU(cudaStreamSynchronize(copy_stream)); CU(cudaStreamSynchronize(aux_stream)); CU(cudaFree(b.p)); b.p = nullptr; b.bytes = 0; }
        const size_t want = bytes + bytes / 8 + 4096;
        CU(cudaMalloc(&b.p, want));
        b.bytes = want;
    }
    return b.p;
}

// stream-ordered device buffer
struct DevMem {
    void* p = nullptr;
    size_t bytes = 0;
    std::shared_ptr<Core> core;
    DevMem() = default;
    Slab* slab = nullptr;       // non-null: sketch storage from the context's slabs; null: transient, stream-ordered pool
    DevMem(const std::shared_ptr<Core>& c, size_t n) : bytes(n), core(c) {
        if (n) CU(cudaMallocAsync(&p, n, c->stream));
    }
    static DevMem persistent(const std::shared_ptr<Core>& c, size_t n) {
        DevMem m;
        m.bytes = n; m.core = c;
        if (n) {
            std::lock_guard<std::mutex> lk(c->slab_mu);
            m.p = c->slabs.alloc(n, &m.slab);
            if (!m.p) throw Fail{SKB_ERR_NOMEM, "out of device memory for sketch storage"};
        }
        return m;
    }
    DevMem(const DevMem&) = delete;
    DevMem& operator=(const DevMem&) = delete;
    DevMem(DevMem&& o) noexcept { *this = std::move(o); }
    DevMem& operator=(DevMem&& o) noexcept {
        if (this != &o) { release(); p = o.p; bytes = o.bytes; slab = o.slab; core = std::move(o.core); o.p = nullptr; o.bytes = 0; o.slab = nullptr; }
        return *this;
    }
    void release() {
        if (p && core) {
            if (slab) { std::lock_guard<std::mutex> lk(core->slab_mu); core->slabs.free(slab, p, bytes); }
            else { cudaSetDevice(core->device); cudaFreeAsync(p, core->stream); }
        }
        p = nullptr; bytes = 0; slab = nullptr;
    }
    ~DevMem() { release(); }
    template <typename T> T* as() const { return (T*)p; }
};

struct BatchStore {   // device arrays shared by the sketches that were produced together
    DevMem kmer_p, pos_p, meta_p, kmer_k, pos_k, meta_k, perm_k, bucket, contig_seed_start, contig_len, contig_win_start, markers;
    DevMem blob;      // sketches received through skb_sketch_unpack / an exchange block: every array is a slice of this one block
    // exchange blocks: the seed arrays ("bodies") may still be arriving over NVLink while the markers are already being
    // screened; whoever reads seed arrays first makes the context's stream wait for this event (ensure_seeds_ready)
    cudaEvent_t body_ev = nullptr;
    bool body_pending = false;
    ~BatchStore() { if (body_ev) cudaEventDestroy(body_ev); }
};

struct SketchImpl {
    std::shared_ptr<Core> core;
    std::shared_ptr<BatchStore> store;
    bool ref_only = false;     // transferred without the position-order arrays: usable as a reference, not as a query
    GenomeView view{};
    skb_sketch_info_t info{};
    std::vector<uint32_t> contig_len_host;
};

constexpr uint32_t FRAGMENT_LENGTH = 20000;   // skani::params::CHUNK_SIZE_DNA

inline void ensure_seeds_ready(const SketchImpl& I) {
    BatchStore& st = *I.store;
    if (st.body_pending) {
        CU(cudaStreamWaitEvent(I.core->stream, st.body_ev, 0));
        st.body_pending = false;
    }
}

}  // namespace skb

namespace skb {
struct ModelImpl {
    std::shared_ptr<Core> core;
    GbdtHost host;
    DevMem nodes, tree_off;
    GbdtView dev_view() const {
        GbdtView v = host.view();
        v.nodes = nodes.as<GbdtNode>(); v.tree_off = tree_off.as<uint32_t>();
        return v;
    }
};
// contig-length quantiles of a sketch (features of the learned-ANI model): element (n - 1) * q / 100 of the sorted lengths
inline void contig_quantiles(const std::vector<uint32_t>& lens, GenomeView& v) {
    v.ctg_q90 = v.ctg_q50 = v.ctg_q10 = 0;
    if (lens.empty()) return;
    std::vector<uint32_t> s(lens);
    std::sort(s.begin(), s.end());
    const size_t n = s.size();
    v.ctg_q90 = s[(n - 1) * 90 / 100]; v.ctg_q50 = s[(n - 1) * 50 / 100]; v.ctg_q10 = s[(n - 1) * 10 / 100];
}
}  // namespace skb

struct skb_ctx { std::shared_ptr<skb::Core> core; };
struct skb_model { std::shared_ptr<skb::ModelImpl> impl; };
struct skb_sketch { std::shared_ptr<skb::SketchImpl> impl; };
struct skb_db {
    std::shared_ptr<skb::Core> core;
    std::vector<std::shared_ptr<skb::SketchImpl>> items;
    skb::DevMem d_views;
    bool dirty = true;
    // marker index (postings sorted by marker), built on the first large screen after the database changed
    skb::DevMem idx_keys, idx_vals, idx_bucket;
    uint32_t idx_shift = 0, idx_postings = 0;
    bool idx_dirty = true;
    std::shared_ptr<skb::ModelImpl> model;      // learned-ANI ensemble, optional
};

namespace skb {

static void* ensure_pinned(Core& c, size_t bytes) {
    if (c.pinned_bytes < bytes) {
        if (c.pinned) cudaFreeHost(c.pinned);
        c.pinned = nullptr; c.pinned_bytes = 0;
        size_t want = std::max(bytes, (size_t)1 << 20);
        CU(cudaHostAlloc(&c.pinned, want, cudaHostAllocDefault));
        c.pinned_bytes = want;
    }
    return c.pinned;
}

template <typename T>
static void upload(Core& c, T* dst, const T* src, size_t n) {
    if (n) CU(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, c.stream));
}
template <typename T>
static void download(Core& c, T* dst, const T* src, size_t n) {
    if (n) CU(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, c.stream));
}

static inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Upload of a small host table on the compute stream while bulk copies may be saturating the host->device copy
// engine.  Small tables go as a pageable-memory memcpy, which the driver embeds in the command stream instead of
// queueing it on the copy engine; larger ones are staged in pinned memory and pulled by a kernel.
static void table_upload(Core& c, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return;
    if (bytes <= 32 * 1024) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c.stream));
    } else {
        void* h = ensure_pinned(c, bytes + 64);
        std::memcpy(h, src, bytes);
        launch_pull_copy(dst, h, bytes, c.stream);
        CU(cudaStreamSynchronize(c.stream));     // the staging block is shared
    }
}

// ------------------------------------------------------------------------------------------------ sketching
struct KeptContig { uint64_t off; uint32_t len; };

// Bucket table geometry of every genome of a batch: ~8 seeds per bucket, bucket id = kmer >> shift (shift <= 31).
struct BucketPlan { std::vector<BucketGenome> genomes; size_t total = 0; uint32_t max_buckets = 1; };
static BucketPlan plan_buckets(const skb_sketch_params_t& P, int seed, const std::vector<std::vector<uint32_t>>& contig_lens) {
    // sized from the EXPECTED seed count (total length / c), which is known before seeding: the seeding kernel can then
    // build the bucket histogram while it writes the seeds
    BucketPlan bp;
    const uint32_t n_genomes = (uint32_t)contig_lens.size();
    bp.genomes.resize(n_genomes);
    for (uint32_t g = 0; g < n_genomes; g++) {
        uint64_t tot = 0;
        for (uint32_t l : contig_lens[g]) tot += l;
        const uint64_t ns = seed ? tot / (uint64_t)P.c : 0;
        int B = 1;
        while (B < 2 * P.k - 1 && B < 20 && ((uint64_t)8 << B) < ns) B++;
        if (2 * P.k - B > 31) B = 2 * P.k - 31;
        BucketGenome& bg = bp.genomes[g];
        bg.seed_start = 0; bg.n_seeds = 0; bg.shift = (uint32_t)(2 * P.k - B); bg.n_buckets = 1u << B;
        bg.bucket_off = (uint32_t)bp.total;
        bp.total += bg.n_buckets + 1;
        bp.max_buckets = std::max(bp.max_buckets, bg.n_buckets);
    }
    return bp;
}

// Finishes a batch whose seeds/markers are described by per-genome contig tables; shared by the scan path
// (seq on device) and the import path (arrays from the host).
static void finish_batch(const std::shared_ptr<Core>& core, const skb_sketch_params_t& P, int seed,
                         uint32_t n_genomes, const std::vector<std::vector<uint32_t>>& contig_lens,
                         std::shared_ptr<BatchStore> store, const std::vector<uint32_t>& seed_start,
                         std::vector<uint32_t>& marker_start, const uint32_t* d_marker_start, skb_sketch_t** out,
                         const uint32_t* d_bucket_overflow = nullptr);

// A chunk of the batch whose bytes become available on the device when `ready` fires (host->device pipelining):
// flat contigs [previous contig_end, contig_end).
// host_wait (optional): blocks the calling thread until `ready` has been RECORDED by the ingest pipeline's threads (a
// stream wait on an event that was never recorded would be a no-op) and tells in which form the chunk arrived:
// 0 = ASCII in seq_dev, 1 = 2-bit words in seq_packed_dev.
struct ChunkPlan { uint32_t contig_end; cudaEvent_t ready; std::function<int()> host_wait; };

static void sketch_core(const std::shared_ptr<Core>& core, const skb_sketch_params_t& P, int seed, uint32_t n_genomes,
                        const uint32_t* gstart, const uint8_t* seq_dev, const uint64_t* offs, const uint64_t* lens,
                        skb_sketch_t** out, const std::vector<ChunkPlan>* plan = nullptr, const uint8_t* seq_packed_dev = nullptr) {
    Core& c = *core;
    cudaStream_t st = c.stream;
    Trace t2("sketch_core");
    if (P.k < 1 || P.k > 16) throw Fail{SKB_ERR_ARG, "k must be in 1..16 for DNA (skani panics above 16)"};
    if (P.c < 1 || P.marker_c < 1) throw Fail{SKB_ERR_ARG, "compression factors must be >= 1"};

    // ---- contig gate + tile table (reference lib.rs:155-174)
    std::vector<std::vector<uint32_t>> contig_lens(n_genomes);
    std::vector<ContigDesc> descs;
    std::vector<uint32_t> chunk_desc_end;      // descriptor index at which each planned chunk ends
    size_t plan_i = 0;
    uint64_t total_bases = 0, tile_count = 0;
    for (uint32_t g = 0; g < n_genomes; g++) {
        bool first = true;
        uint32_t kept = 0;
        for (uint32_t ci = gstart[g]; ci < gstart[g + 1]; ci++) {
            while (plan && plan_i < plan->size() && (*plan)[plan_i].contig_end <= ci) { chunk_desc_end.push_back((uint32_t)descs.size()); plan_i++; }
            if (lens[ci] < SKB_MIN_LENGTH_CONTIG) continue;
            if (lens[ci] > 0x7FFFFFFFull) throw Fail{SKB_ERR_ARG, "contigs of 2^31 bases or more are not supported"};
            if (offs[ci] & 15) throw Fail{SKB_ERR_ARG, "device contig offsets must be multiples of 16"};
            contig_lens[g].push_back((uint32_t)lens[ci]);
            total_bases += lens[ci];
            ContigDesc d;
            d.seq_off = offs[ci];
            d.len = (uint32_t)lens[ci];
            d.contig = kept;
            d.genome = g | (first ? 0x80000000u : 0u);
            d.tile_start = (uint32_t)tile_count;
            tile_count += (lens[ci] + TILE_BASES - 1) / TILE_BASES;
            first = false;
            descs.push_back(d);
            kept++;
        }
    }
    while (plan && plan_i < plan->size()) { chunk_desc_end.push_back((uint32_t)descs.size()); plan_i++; }
    BucketPlan bplan = plan_buckets(P, seed, contig_lens);
    if (total_bases >= 0x7FFFFFFFull) throw Fail{SKB_ERR_ARG, "a sketch batch is limited to 2^31 bases; split the call"};
    if (n_genomes >= (1u << 22)) throw Fail{SKB_ERR_ARG, "a sketch batch is limited to 2^22 genomes"};
    const uint32_t n_tiles = (uint32_t)tile_count;

    std::vector<uint32_t> seed_start(n_genomes + 1, 0), marker_start(n_genomes + 1, 0);
    auto store = std::make_shared<BatchStore>();

    if (n_tiles) {
        ContigDesc* d_descs = (ContigDesc*)c.scratch(SLOT_DESC, sizeof(ContigDesc) * descs.size());
        table_upload(c, d_descs, descs.data(), sizeof(ContigDesc) * descs.size());
        t2.mark("descs uploaded");
        // ---- launch plan: one launch per copy chunk (or one for everything); every warp of a launch owns a region
        struct Launch { uint32_t d0, d1, tile_base, n_tiles, n_warps, n_chunks, chunk_tiles, region_base; cudaEvent_t ready; const ChunkPlan* chunk; };
        std::vector<Launch> launches;
        uint32_t n_regions = 0;
        auto add_launch = [&](uint32_t d0, uint32_t d1, cudaEvent_t ready, const ChunkPlan* chunk = nullptr) {
            if (d1 <= d0) return;
            Launch L{};
            L.d0 = d0; L.d1 = d1; L.tile_base = descs[d0].tile_start;
            L.n_tiles = (d1 < descs.size() ? descs[d1].tile_start : n_tiles) - L.tile_base;
            // small calls (one genome: the query of Database.query) use smaller regions so that every warp of the GPU
            // gets work: 8-tile regions would leave a 5 Mbp genome to 305 of 4 736 warps
            const uint32_t all_warps = (uint32_t)c.n_sm * 4u * SEED_WARPS;
            L.chunk_tiles = std::max<uint32_t>(1, std::min<uint32_t>(CHUNK_TILES, n_tiles / (2 * all_warps)));   // by the size of the whole call
            L.n_chunks = (L.n_tiles + L.chunk_tiles - 1) / L.chunk_tiles;
            uint32_t grid = std::min<uint32_t>((uint32_t)c.n_sm * 4u, (L.n_chunks + SEED_WARPS - 1) / SEED_WARPS);
            L.n_warps = grid * SEED_WARPS; L.region_base = n_regions; L.ready = ready; L.chunk = chunk;
            n_regions += L.n_chunks;
            launches.push_back(L);
        };
        if (plan && plan->size() > 1) {
            uint32_t d0 = 0;
            for (size_t ch = 0; ch < plan->size(); ch++) { add_launch(d0, chunk_desc_end[ch], (*plan)[ch].ready, &(*plan)[ch]); d0 = chunk_desc_end[ch]; }
        } else {
            add_launch(0, (uint32_t)descs.size(), plan && plan->size() == 1 ? (*plan)[0].ready : nullptr, plan && plan->size() == 1 ? &(*plan)[0] : nullptr);
        }

        // ---- region bookkeeping (device): counts, storage offsets, scans, per-genome records, overflow flag
        const size_t g_bytes = sizeof(uint32_t) * (n_genomes + 1);
        // layout: u64 region_cnt[n+1] | u64 region_start[n+1] | u32 src offsets 2n | u32 genome records 3g | overflow | claims
        const size_t book_bytes = 16 * ((size_t)n_regions + 1) + 4 * ((size_t)n_regions * 2 + 3 * (size_t)n_genomes + 16 + launches.size());
        char* book = (char*)c.scratch(SLOT_STATUS, book_bytes);
        uint64_t* r_cnt = (uint64_t*)book; uint64_t* r_start = r_cnt + n_regions + 1;
        uint32_t* r_ssrc = (uint32_t*)(r_start + n_regions + 1); uint32_t* r_msrc = r_ssrc + n_regions;
        uint32_t* g_region = r_msrc + n_regions; uint32_t* g_slocal = g_region + n_genomes; uint32_t* g_mlocal = g_slocal + n_genomes;
        uint32_t* d_overflow = g_mlocal + n_genomes;
        uint32_t* d_claim = d_overflow + 1;     // one claim counter per launch
        const size_t rscan_bytes = region_scan_scratch_bytes(n_regions);
        void* rscan_scratch = c.scratch(SLOT_RSCAN, rscan_bytes);
        uint32_t* d_gs = (uint32_t*)c.scratch(SLOT_GS, g_bytes);
        uint32_t* d_gm = (uint32_t*)c.scratch(SLOT_GM, g_bytes);
        uint32_t* d_bcounts = (uint32_t*)c.scratch(SLOT_BCOUNT, 4 * bplan.total + 16);   // k-mer bucket histogram of the batch

        // per-tile capacities of the region storage: generous multiples of the expected hit counts; a region that
        // still overflows (low-complexity sequence) triggers one retry with the exact layout
        uint32_t seed_tile_cap = seed ? std::min<uint32_t>(TILE_BASES, TILE_BASES / P.c + TILE_BASES / P.c / 2 + 24) : 0;
        uint32_t marker_tile_cap = std::min<uint32_t>(TILE_BASES, 2 * (TILE_BASES / P.marker_c) + 16);
        uint32_t *t_kmer = nullptr, *t_pos = nullptr, *t_meta = nullptr;
        uint64_t* t_mreg = nullptr;
        uint32_t h_over = 0;
        // exact = 0: the seeding kernel compares only the high words of hash and threshold and re-checks every hit
        // exactly as it writes it; a position that slipped through (hash.hi == threshold.hi, lo above: one in ~6e9)
        // raises bit 1 of the flag and the batch is repeated with the exact 64-bit comparison (SKB_SEED_EXACT=1 forces it)
        // (also when the high word of a threshold is 0 - no seeds wanted, or a compression factor >= 2^32: the carry form of the
        // high-word test in the kernel needs a high word >= 1)
        const uint64_t thr_seed = seed ? UINT64_MAX / (uint64_t)P.c : 0, thr_marker = UINT64_MAX / (uint64_t)P.marker_c;
        int exact = (std::getenv("SKB_SEED_EXACT") || (thr_seed >> 32) == 0 || (thr_marker >> 32) == 0) ? 1 : 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            const size_t seed_store = attempt == 0 ? (size_t)n_tiles * seed_tile_cap : (size_t)seed_start[n_genomes];
            const size_t marker_store = attempt == 0 ? (size_t)n_tiles * marker_tile_cap : (size_t)marker_start[n_genomes];
            if (seed_store >= 0xFFFFFFFFull || marker_store >= 0xFFFFFFFFull) throw Fail{SKB_ERR_ARG, "sketch batch too large for 32-bit seed offsets; split the call"};
            t_kmer = (uint32_t*)c.scratch(SLOT_KMER, 4 * seed_store + 16); t_pos = (uint32_t*)c.scratch(SLOT_POS, 4 * seed_store + 16);
            t_meta = (uint32_t*)c.scratch(SLOT_META, 4 * seed_store + 16); t_mreg = (uint64_t*)c.scratch(SLOT_MKEYS, 8 * marker_store + 16);
            CU(cudaMemsetAsync(g_region, 0xFF, 4 * (size_t)n_genomes, st));
            CU(cudaMemsetAsync(d_overflow, 0, 4 * (1 + launches.size()), st));
            CU(cudaMemsetAsync(r_cnt + n_regions, 0, 8, st));
            SeedScanArgs a{};
            a.seq = seq_dev;
            a.kmask = P.k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * P.k)) - 1u);
            a.kshift = 42 - 2 * P.k;
            a.thr_seed = thr_seed;                                   // seed=False keeps markers only (A.4)
            a.thr_marker = thr_marker;
            a.chk_seed = a.thr_seed; a.chk_marker = a.thr_marker;
            if (std::getenv("SKB_SEED_TEST_INEXACT")) { a.chk_seed /= 2; a.chk_marker /= 2; }   // test hook: forces the exact repeat
            a.seed_tile_cap = seed_tile_cap; a.marker_tile_cap = marker_tile_cap;
            a.region_off = attempt == 0 ? nullptr : r_start;
            a.kmer_r = t_kmer; a.pos_r = t_pos; a.meta_r = t_meta; a.marker_r = t_mreg;
            a.region_seed_src = r_ssrc; a.region_marker_src = r_msrc; a.region_cnt = r_cnt;
            a.genome_region = g_region; a.genome_seed_local = g_slocal; a.genome_marker_local = g_mlocal;
            a.overflow = d_overflow;
            a.exact_compare = (uint32_t)exact;
            a.fma_m1 = 0xFFFFFFFFu; a.fma_two = 2;
            CU(cudaEventRecord(c.ev[1], st));
            for (size_t li = 0; li < launches.size(); li++) {
                const Launch& L = launches[li];
                const int packed = L.chunk && L.chunk->host_wait ? L.chunk->host_wait() : 0;
                if (L.ready) CU(cudaStreamWaitEvent(st, L.ready, 0));
                SeedScanArgs b2 = a;
                if (packed) { b2.seq = seq_packed_dev; b2.packed = 1; }
                b2.n_chunks = L.n_chunks; b2.chunk_tiles = L.chunk_tiles; b2.chunk_counter = d_claim + li;
                b2.contigs = d_descs + L.d0; b2.n_contigs = L.d1 - L.d0;
                b2.tile_base = L.tile_base; b2.n_tiles = L.n_tiles; b2.n_warps = L.n_warps; b2.region_base = L.region_base;
                launch_seed_scan(b2, c.n_sm, st);
            }
            CU(cudaEventRecord(c.ev[2], st));
            if (attempt == 0) {
                // the retry reads its layout from r_sstart / r_mstart, so the scan must not run again after it
                scan_region_counts(n_regions, r_cnt, r_start, rscan_scratch, rscan_bytes, st);
                launch_genome_starts(n_regions, r_start, n_genomes, g_region, g_slocal, g_mlocal, d_gs, d_gm, st);
            }
            // per-genome starts and the overflow flag reach the host through one small kernel that writes into pinned
            // host memory (three device->host copies would each cost a copy-engine round trip)
            uint32_t* h_counts = c.pinned_counts(2 * ((size_t)n_genomes + 1) + 1);
            launch_counters_to_host(attempt == 0 ? d_gs : nullptr, attempt == 0 ? d_gm : nullptr, n_genomes + 1, d_overflow, h_counts, st);
            t2.mark("enqueued seed+scan");
            CU(cudaStreamSynchronize(st));
            t2.mark("sync after seed");
            if (attempt == 0) {
                std::memcpy(seed_start.data(), h_counts, 4 * ((size_t)n_genomes + 1));
                std::memcpy(marker_start.data(), h_counts + n_genomes + 1, 4 * ((size_t)n_genomes + 1));
            }
            h_over = h_counts[2 * ((size_t)n_genomes + 1)];
            if (h_over & 2u) {
                if (exact) throw Fail{SKB_ERR_CUDA, "exact hash comparison flagged as inexact"};
                exact = 1; attempt = -1;          // start over: counts and layout of this pass include a false positive
                continue;
            }
            if (!h_over) break;
            if (attempt == 1) throw Fail{SKB_ERR_CUDA, "seed regions overflowed twice"};
        }
        // genomes without a tile never recorded a start: they are empty and begin where the next one begins
        for (uint32_t g = n_genomes; g-- > 0;) {
            if (seed_start[g] == 0xFFFFFFFFu) { seed_start[g] = seed_start[g + 1]; marker_start[g] = marker_start[g + 1]; }
        }
        const uint32_t ns = seed_start[n_genomes], nm = marker_start[n_genomes];

        t2.mark("host fixups");
        // ---- exact-size position-order arrays + contiguous marker keys: the gather is also the compaction copy
        store->kmer_p = DevMem::persistent(core, 4 * (size_t)ns); store->pos_p = DevMem::persistent(core, 4 * (size_t)ns);
        store->meta_p = DevMem::persistent(core, 4 * (size_t)ns);
        uint64_t* t_mkeys = (uint64_t*)c.scratch(SLOT_MKEYS2, 8 * (size_t)nm + 16);
        char* d_tab = nullptr;
        uint32_t* d_bover = nullptr;
        uint32_t* d_gm_in = nullptr;          // pre-deduplication marker offsets (d_gm is overwritten with the final ones)
        {
            RegionGatherArgs ga{};
            ga.n_regions = n_regions; ga.seed_src = r_ssrc; ga.marker_src = r_msrc; ga.region_start = r_start;
            ga.kmer_r = t_kmer; ga.pos_r = t_pos; ga.meta_r = t_meta; ga.marker_r = t_mreg;
            ga.kmer_p = store->kmer_p.as<uint32_t>(); ga.pos_p = store->pos_p.as<uint32_t>(); ga.meta_p = store->meta_p.as<uint32_t>();
            ga.marker_keys = t_mkeys;
            // the gather also builds the per-genome k-mer bucket histogram of the index build (saves a pass over kmer_p)
            for (uint32_t g = 0; g < n_genomes; g++) { bplan.genomes[g].seed_start = seed_start[g]; bplan.genomes[g].n_seeds = seed_start[g + 1] - seed_start[g]; }
            // one upload: [bucket plan per genome | bucket-overflow flag = 0 | pre-deduplication marker offsets]
            const size_t tab_bytes = sizeof(BucketGenome) * n_genomes;
            const size_t over_off = (tab_bytes + 15) & ~(size_t)15, gm_off = over_off + 16;
            std::vector<char> blob(gm_off + g_bytes, 0);
            std::memcpy(blob.data(), bplan.genomes.data(), tab_bytes);
            std::memcpy(blob.data() + gm_off, marker_start.data(), g_bytes);
            d_tab = (char*)c.scratch(SLOT_BTAB, blob.size());
            table_upload(c, d_tab, blob.data(), blob.size());
            d_bover = (uint32_t*)(d_tab + over_off);
            d_gm_in = (uint32_t*)(d_tab + gm_off);
            CU(cudaMemsetAsync(d_bcounts, 0, 4 * bplan.total, st));
            ga.genomes = (const BucketGenome*)d_tab; ga.n_genomes = n_genomes; ga.bucket_counts = d_bcounts;
            launch_region_gather(ga, st);
        }
        t2.mark("gather enqueued");
        // ---- marker sets, on the auxiliary stream: a chain of small latency-bound launches that runs beside the k-mer order
        store->markers = DevMem::persistent(core, 8 * (size_t)std::max<uint32_t>(nm, 1));
        {
            const size_t mark_bytes = marker_scratch_bytes(nm, n_genomes);
            void* mark_scratch = c.scratch(SLOT_MARK, mark_bytes);
            uint32_t max_gm = 0;
            for (uint32_t g = 0; g < n_genomes; g++) max_gm = std::max(max_gm, marker_start[g + 1] - marker_start[g]);
            CU(cudaEventRecord(c.ev[6], st));
            CU(cudaStreamWaitEvent(c.aux_stream, c.ev[6], 0));
            build_marker_sets(n_genomes, nm, t_mkeys, store->markers.as<uint64_t>(), d_gm, d_gm_in, max_gm, mark_scratch, mark_bytes, c.aux_stream);
            CU(cudaEventRecord(c.ev[7], c.aux_stream));
        }
        // ---- k-mer order
        store->kmer_k = DevMem::persistent(core, 4 * (size_t)ns); store->pos_k = DevMem::persistent(core, 4 * (size_t)ns);
        store->meta_k = DevMem::persistent(core, 4 * (size_t)ns); store->perm_k = DevMem::persistent(core, 4 * (size_t)ns);
        {
            // bucket partition: histogram (built by the gather) -> per-genome scan (= the bucket tables) -> scatter -> rank
            // inside the bucket
            store->bucket = DevMem::persistent(core, 4 * bplan.total);
            const size_t bscr_bytes = bucket_order_scratch_bytes(ns, bplan.total);
            void* bscr = c.scratch(SLOT_SORT, bscr_bytes);
            build_kmer_order_buckets(ns, n_genomes, (const BucketGenome*)d_tab, bplan.total, d_bcounts, 1, store->kmer_p.as<uint32_t>(),
                                     store->pos_p.as<uint32_t>(), store->meta_p.as<uint32_t>(), store->kmer_k.as<uint32_t>(),
                                     store->pos_k.as<uint32_t>(), store->meta_k.as<uint32_t>(), store->perm_k.as<uint32_t>(),
                                     store->bucket.as<uint32_t>(), d_bover, bscr, bscr_bytes, st);
        }
        CU(cudaStreamWaitEvent(st, c.ev[7], 0));
        CU(cudaGetLastError());                  // refused launches must not pass silently
        t2.mark("enqueued index");
        CU(cudaEventRecord(c.ev[3], st));
        finish_batch(core, P, seed, n_genomes, contig_lens, store, seed_start, marker_start, d_gm, out, d_bover);
    } else {
        CU(cudaEventRecord(c.ev[1], st)); CU(cudaEventRecord(c.ev[2], st)); CU(cudaEventRecord(c.ev[3], st));
        finish_batch(core, P, seed, n_genomes, contig_lens, store, seed_start, marker_start, nullptr, out);
    }
    t2.mark("finish_batch done");
}


// ------------------------------------------------------------------------------------------------ host ingest pipeline
// Large host batches reach the device through two concurrent routes (host_pack.h): the copy engine pulls chunks of
// plain ASCII out of the caller's memory, and a team of threads compacts other chunks to 2-bit words in pinned staging
// memory, from where a quarter of the bytes crosses the link.  Chunks (runs of whole contigs, ~16 MB) are claimed in
// order from one counter by whichever route is free, so the split adapts to the host: PCIe rate against pack rate.
// The calling thread meanwhile launches the seeding kernel of every chunk in order (sketch_core), variant by arrival form.

// DMA of ASCII contigs into the device layout: large contigs straight from the caller's memory (adjacent ones that keep
// the device displacement merged into one copy), small ones through a pinned staging block.
struct RawCopier {
    static constexpr uint64_t DIRECT = 1 << 18;
    static constexpr size_t STAGE = (size_t)8 << 20;
    Core& c; cudaStream_t cs; uint8_t* d_seq; bool merge;
    char* stage = nullptr; size_t used = 0; uint64_t stage_dev0 = 0;
    const uint8_t* run_src = nullptr; uint64_t run_dst = 0, run_bytes = 0;
    uint64_t bytes = 0;
    RawCopier(Core& c_, cudaStream_t cs_, uint8_t* d, char* stage_) : c(c_), cs(cs_), d_seq(d), merge(std::getenv("SKB_NO_COPY_MERGE") == nullptr), stage(stage_) {}
    void flush_run() {
        if (run_bytes) {
            CU(cudaMemcpyAsync((char*)d_seq + run_dst, run_src, run_bytes, cudaMemcpyHostToDevice, cs));
            run_bytes = 0;
        }
    }
    void flush_stage() {
        if (used) {
            CU(cudaMemcpyAsync((char*)d_seq + stage_dev0, stage, used, cudaMemcpyHostToDevice, cs));
            CU(cudaStreamSynchronize(cs));            // the staging block is reused
            used = 0;
        }
    }
    void add(const uint8_t* src, uint64_t len, uint64_t off) {
        bytes += len;
        if (len >= DIRECT) {
            // (the <= 31 padding bytes between merged contigs are copied along; they lie between two valid buffers on pages
            // that hold valid bytes, and the kernels never interpret bytes outside a contig)
            const bool adjacent = merge && run_bytes && off >= run_dst + run_bytes && off - (run_dst + run_bytes) < 32 &&
                                  (int64_t)(src - run_src) == (int64_t)(off - run_dst);
            if (adjacent) run_bytes = off - run_dst + len;
            else { flush_run(); run_src = src; run_dst = off; run_bytes = len; }
        } else {
            const uint64_t span = align16(len) + 16;
            if (used && (used + span > STAGE || stage_dev0 + used != off)) flush_stage();
            if (!used) stage_dev0 = off;
            std::memcpy(stage + used, src, len);
            used += span;
        }
    }
    void flush() { flush_run(); flush_stage(); }
};

struct Ingest {
    enum { MIX = 0, PACK_ONLY = 1 };
    size_t PIECE_BASES = (size_t)512 << 10;                    // unit of work of a packing thread (a multiple of 16)
    static constexpr unsigned RAW_DEPTH = 2;                     // ASCII chunks the copy engine may have queued
    struct Chunk { uint32_t c0, c1; cudaEvent_t ev; int mode; bool done; };
    struct Piece { const uint8_t* src; size_t n; uint32_t* dst; uint32_t chunk; };

    Core& c;
    const uint8_t* const* contigs; const uint64_t* lens; const uint64_t* offs;     // flat contig arrays; offs = ASCII device layout
    uint8_t* d_seq; uint8_t* d_pk; char* stage_pk; char* stage_raw;
    int policy;
    std::vector<Chunk> chunks;
    std::atomic<uint32_t> next_chunk{0};
    std::atomic<bool> abort{false};
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Piece> queue;
    std::vector<uint32_t> remaining;
    int err = 0; std::string err_msg;
    uint64_t raw_bytes = 0, packed_bytes = 0;      // bytes that crossed the link in either form (guarded by mu)

    Ingest(Core& c_, const uint8_t* const* ct, const uint64_t* ln, const uint64_t* of, uint8_t* dseq, uint8_t* dpk, char* spk, char* sraw, int pol)
        : c(c_), contigs(ct), lens(ln), offs(of), d_seq(dseq), d_pk(dpk), stage_pk(spk), stage_raw(sraw), policy(pol) {
        if (const char* e = std::getenv("SKB_PIECE_KB")) PIECE_BASES = std::max<size_t>(16, ((size_t)std::atoll(e) << 10) & ~(size_t)15);   // tuning hook
    }

    // calling thread: block until chunk ch is on its way (its event is recorded); returns the arrival form
    int wait(uint32_t ch) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return chunks[ch].done || err; });
        if (err) throw Fail{err, err_msg};
        return chunks[ch].mode;
    }
    void fail(int code, const std::string& msg) {
        std::lock_guard<std::mutex> lk(mu);
        if (!err) { err = code; err_msg = msg; }
        abort.store(true);
        cv.notify_all();
    }
    void mark_done(uint32_t ch, int mode, uint64_t link_bytes) {
        CU(cudaEventRecord(chunks[ch].ev, c.copy_stream));
        std::lock_guard<std::mutex> lk(mu);
        chunks[ch].mode = mode; chunks[ch].done = true;
        (mode ? packed_bytes : raw_bytes) += link_bytes;
        cv.notify_all();
    }
    bool kept(uint32_t i) const { return lens[i] >= SKB_MIN_LENGTH_CONTIG; }

    // ---- route 1: ASCII by DMA (one thread)
    void feed_raw() {
        RawCopier rc(c, c.copy_stream, d_seq, stage_raw);
        std::deque<uint32_t> outstanding;
        while (!abort.load()) {
            if (outstanding.size() >= RAW_DEPTH) { CU(cudaEventSynchronize(chunks[outstanding.front()].ev)); outstanding.pop_front(); }
            const uint32_t ch = next_chunk.fetch_add(1);
            if (ch >= chunks.size()) break;
            const uint64_t before = rc.bytes;
            for (uint32_t i = chunks[ch].c0; i < chunks[ch].c1; i++) if (kept(i)) rc.add(contigs[i], lens[i], offs[i]);
            rc.flush();
            mark_done(ch, 0, rc.bytes - before);
            outstanding.push_back(ch);
        }
    }
    // ---- route 2: 2-bit words through pinned staging (all other threads)
    std::chrono::steady_clock::time_point t_start = std::chrono::steady_clock::now();
    std::atomic<uint64_t> last_piece_ns{0};      // when the last piece was packed, from t_start
    std::atomic<uint64_t> finish_ns{0}, lock_ns{0}, pack_ns{0};      // SKB_TRACE: time the workers spent in CUDA calls / waiting for mu / packing
    void finish_packed(uint32_t ch) {
        const auto t_f0 = std::chrono::steady_clock::now();
        struct Acc { std::atomic<uint64_t>& a; std::chrono::steady_clock::time_point t0; ~Acc() { a += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); } } acc{finish_ns, t_f0};
        uint32_t first = chunks[ch].c0, last = chunks[ch].c1;
        while (first < last && !kept(first)) first++;
        while (last > first && !kept(last - 1)) last--;
        uint64_t bytes = 0;
        if (first < last) {
            const uint64_t b0 = offs[first] / 4, b1 = (offs[last - 1] + align16(lens[last - 1])) / 4;     // contiguous: the staging mirrors the device layout
            bytes = b1 - b0;
            CU(cudaMemcpyAsync(d_pk + b0, stage_pk + b0, bytes, cudaMemcpyHostToDevice, c.copy_stream));
        }
        mark_done(ch, 1, bytes);
    }
    void pack_loop() {
        while (!abort.load()) {
            Piece p{};
            bool have = false, empty_chunk = false;
            uint32_t claimed = 0;
            {
                // taking a piece and, when none is left, claiming the next chunk happen under one lock: at most one
                // chunk beyond the queue is ever claimed for packing, the rest stays available to the copy engine
                std::lock_guard<std::mutex> lk(mu);
                if (!queue.empty()) { p = queue.front(); queue.pop_front(); have = true; }
                else {
                    claimed = next_chunk.fetch_add(1);
                    if (claimed >= chunks.size()) break;
                    uint32_t n_pieces = 0;
                    for (uint32_t i = chunks[claimed].c0; i < chunks[claimed].c1; i++) {
                        if (!kept(i)) continue;
                        uint32_t* dst = reinterpret_cast<uint32_t*>(stage_pk + offs[i] / 4);
                        for (uint64_t o = 0; o < lens[i]; o += PIECE_BASES) {
                            queue.push_back(Piece{contigs[i] + o, (size_t)std::min<uint64_t>(PIECE_BASES, lens[i] - o), dst + o / 16, claimed});
                            n_pieces++;
                        }
                    }
                    remaining[claimed] = n_pieces;
                    if (n_pieces) { p = queue.front(); queue.pop_front(); have = true; }
                    else empty_chunk = true;
                }
            }
            if (empty_chunk) { finish_packed(claimed); continue; }
            if (!have) continue;
            const auto t_p0 = std::chrono::steady_clock::now();
            host_pack_bases(p.src, p.n, p.dst);
            const auto t_p1 = std::chrono::steady_clock::now();
            bool last;
            { std::lock_guard<std::mutex> lk(mu); last = --remaining[p.chunk] == 0; }
            const auto t_p2 = std::chrono::steady_clock::now();
            pack_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_p1 - t_p0).count();
            last_piece_ns.store((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_p1 - t_start).count());
            lock_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_p2 - t_p1).count();
            if (last) finish_packed(p.chunk);
        }
    }
    void worker(unsigned id) {
        try {
            if (id == 0 && policy == MIX) feed_raw();
            pack_loop();          // the feeding thread helps with what is left once every chunk is claimed
        } catch (const Fail& f) {
            fail(f.code, f.msg);
        } catch (const std::exception& e) {
            fail(SKB_ERR_CUDA, e.what());
        }
    }
};

// thread count of the ingest pipeline of this context (0 = pipeline off)
static unsigned resolve_host_threads(const Core& c) {
    if (c.host_threads >= 0) return (unsigned)c.host_threads;
    if (const char* e = std::getenv("SKB_HOST_THREADS")) return (unsigned)std::max(0, std::atoi(e));
    unsigned cpus = host_cpu_count();
    // one process per GPU sharing the host's cores (torchrun exports LOCAL_WORLD_SIZE)
    if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) {
        const int w = std::atoi(e);
        if (w > 1) {
            cpus = std::max(1u, cpus / (unsigned)w);
            // From four ranks per host on, the ranks' plain copies together already saturate the host's PCIe / memory fabric
            // (copy floor per GPU on the bench boxes: 55 GB/s alone, 29 GB/s at 4, 23 GB/s at 8) and a packing team on every
            // rank only adds load: 8 ranks 33.6 ms with the team against 31.7 ms without, 4 ranks 35.5-37 against 31.2 ms
            // (profiles/r2_bench_line_n8_ingest_*.json).  A rank cannot see that from its own timings -
            // its choice slows the OTHER ranks - so this is a rule, not something the context learns.
            if (w >= 4 || cpus < 8) return 0;
        }
    }
    // the calling thread spins in its stream synchronisations: it keeps one CPU, the team gets the others (with 16 CPUs,
    // 16 packing threads reach 80 GB/s, 15 reach 102 GB/s: profiles/r2_ingest_sweep.txt)
    return std::max(2u, std::min(32u, cpus) - 1u);
}

static void finish_batch(const std::shared_ptr<Core>& core, const skb_sketch_params_t& P, int seed,
                         uint32_t n_genomes, const std::vector<std::vector<uint32_t>>& contig_lens,
                         std::shared_ptr<BatchStore> store, const std::vector<uint32_t>& seed_start,
                         std::vector<uint32_t>& marker_start, const uint32_t* d_marker_start, skb_sketch_t** out,
                         const uint32_t* d_bucket_overflow) {
    // d_bucket_overflow != NULL: the k-mer order and the bucket tables were already produced by the bucket-partition
    // path (store->bucket is allocated and filled); the flag tells whether a bucket was too large for it
    // d_marker_start (device, may be NULL): per-genome offsets into the de-duplicated marker array, produced by the
    // marker kernels that are still in flight; they are fetched together with the end-of-batch synchronisation below
    Core& c = *core;
    cudaStream_t st = c.stream;
    Trace tf("finish_batch");
    // ---- per-genome tables: buckets, contig seed starts, contig lengths, window capacities
    std::vector<GenomeView> views(n_genomes);
    std::vector<uint32_t> h_clen, h_cwin;
    size_t cstart_total = 0;
    uint32_t max_contigs = 0;
    const BucketPlan bp = plan_buckets(P, seed, contig_lens);
    const size_t bucket_total = bp.total;
    const uint32_t max_buckets = bp.max_buckets;
    std::vector<size_t> bucket_off(n_genomes), cstart_off(n_genomes), clen_off(n_genomes);
    for (uint32_t g = 0; g < n_genomes; g++) {
        const uint32_t ns = seed_start[g + 1] - seed_start[g];
        GenomeView& v = views[g];
        v.n_buckets = bp.genomes[g].n_buckets;
        v.bucket_shift = bp.genomes[g].shift;
        bucket_off[g] = bp.genomes[g].bucket_off;
        const uint32_t nc = (uint32_t)contig_lens[g].size();
        max_contigs = std::max(max_contigs, nc);
        cstart_off[g] = cstart_total; cstart_total += nc + 1;
        clen_off[g] = h_clen.size();
        uint32_t wcap = 0; uint64_t tot = 0;
        for (uint32_t ci = 0; ci < nc; ci++) {
            h_clen.push_back(contig_lens[g][ci]);
            h_cwin.push_back(wcap);
            wcap += contig_lens[g][ci] / FRAGMENT_LENGTH + 1;
            tot += contig_lens[g][ci];
        }
        h_cwin.push_back(wcap);
        v.win_cap = wcap; v.total_len = tot; v.n_contigs = nc;
        v.n_seeds = ns; v.n_markers = 0;      // marker fields are filled in after the final synchronisation
        contig_quantiles(contig_lens[g], v);
    }
    if (!d_bucket_overflow) store->bucket = DevMem::persistent(core, 4 * bucket_total);
    store->contig_seed_start = DevMem::persistent(core, 4 * std::max<size_t>(cstart_total, 1));
    // contig lengths and window starts share one block (one upload): [lengths | window starts]
    const size_t clen_words = (h_clen.size() + 3) & ~(size_t)3;
    store->contig_len = DevMem::persistent(core, 4 * (clen_words + std::max<size_t>(cstart_total, 1)));
    uint32_t* cwin_base = store->contig_len.as<uint32_t>() + clen_words;
    for (uint32_t g = 0; g < n_genomes; g++) {
        GenomeView& v = views[g];
        const size_t so = seed_start[g];
        v.kmer_p = store->kmer_p.as<uint32_t>() + so; v.pos_p = store->pos_p.as<uint32_t>() + so;
        v.meta_p = store->meta_p.as<uint32_t>() + so; v.kmer_k = store->kmer_k.as<uint32_t>() + so;
        v.pos_k = store->pos_k.as<uint32_t>() + so; v.meta_k = store->meta_k.as<uint32_t>() + so;
        v.perm_k = store->perm_k.as<uint32_t>() + so;
        v.bucket = store->bucket.as<uint32_t>() + bucket_off[g];
        v.contig_seed_start = store->contig_seed_start.as<uint32_t>() + cstart_off[g];
        v.contig_win_start = cwin_base + cstart_off[g];
        v.contig_len = store->contig_len.as<uint32_t>() + clen_off[g];
        v.markers = nullptr;
    }
    {
        DevMem d_views(core, sizeof(GenomeView) * n_genomes);
        table_upload(c, d_views.p, views.data(), sizeof(GenomeView) * n_genomes);
        {
            std::vector<uint32_t> both(clen_words + h_cwin.size(), 0u);
            std::copy(h_clen.begin(), h_clen.end(), both.begin());
            std::copy(h_cwin.begin(), h_cwin.end(), both.begin() + clen_words);
            table_upload(c, store->contig_len.p, both.data(), 4 * both.size());
        }
        // per-genome layout of contig_win_start mirrors contig_seed_start (nc + 1 entries each)
        if (!d_bucket_overflow) launch_build_buckets(d_views.as<GenomeView>(), n_genomes, max_buckets, st);
        launch_contig_starts(d_views.as<GenomeView>(), n_genomes, max_contigs, st);
        if (d_marker_start) download(c, marker_start.data(), d_marker_start, n_genomes + 1);
        uint32_t h_bover = 0;
        if (d_bucket_overflow) download(c, &h_bover, d_bucket_overflow, 1);
        tf.mark("tables + downloads enqueued");
        CU(cudaStreamSynchronize(st));   // the one synchronisation of the index build
        tf.mark("synchronised");
        if (h_bover) {
            // a bucket was too large for the partition path (low-complexity genome): redo the k-mer order with the
            // radix sort and rebuild the bucket tables from it
            const uint32_t ns = seed_start[n_genomes];
            DevMem d_gs(core, 4 * (size_t)(n_genomes + 1));
            table_upload(c, d_gs.p, seed_start.data(), 4 * (size_t)(n_genomes + 1));
            const size_t sort_bytes = kmer_order_scratch_bytes(ns);
            void* sort_scratch = c.scratch(SLOT_SORT, sort_bytes);
            IndexBuildArgs ib{};
            ib.n_genomes = n_genomes; ib.n_seeds_total = ns; ib.genome_seed_start = d_gs.as<uint32_t>();
            for (uint32_t g = 0; g < n_genomes; g++) ib.max_genome_seeds = std::max(ib.max_genome_seeds, seed_start[g + 1] - seed_start[g]);
            ib.kmer_p = store->kmer_p.as<uint32_t>(); ib.pos_p = store->pos_p.as<uint32_t>(); ib.meta_p = store->meta_p.as<uint32_t>();
            ib.kmer_k = store->kmer_k.as<uint32_t>(); ib.pos_k = store->pos_k.as<uint32_t>(); ib.meta_k = store->meta_k.as<uint32_t>();
            ib.perm_k = store->perm_k.as<uint32_t>();
            ib.k = P.k;
            build_kmer_order(ib, sort_scratch, sort_bytes, st);
            launch_build_buckets(d_views.as<GenomeView>(), n_genomes, max_buckets, st);
            CU(cudaStreamSynchronize(st));
        }
    }
    for (uint32_t g = 0; g < n_genomes; g++) {
        views[g].n_markers = marker_start[g + 1] - marker_start[g];
        views[g].markers = store->markers.as<uint64_t>() + marker_start[g];
    }
    for (uint32_t g = 0; g < n_genomes; g++) {
        auto impl = std::make_shared<SketchImpl>();
        impl->core = core; impl->store = store; impl->view = views[g];
        impl->contig_len_host = contig_lens[g];
        impl->info.n_seeds = views[g].n_seeds; impl->info.n_markers = views[g].n_markers;
        impl->info.total_len = views[g].total_len; impl->info.n_contigs = views[g].n_contigs;
        impl->info.k = P.k; impl->info.c = P.c; impl->info.marker_c = P.marker_c; impl->info.has_seeds = seed ? 1 : 0;
        out[g] = new skb_sketch{impl};
    }
    tf.mark("handles built");
}

}  // namespace skb

// ================================================================================================ C ABI
using namespace skb;

namespace {

template <typename F>
int guarded(Core* core, F&& f) {
    try {
        if (core) { std::lock_guard<std::mutex> lk(core->mu); cudaSetDevice(core->device); return f(); }
        return f();
    } catch (const Fail& e) {
        if (core) core->err = e.msg;
        cudaGetLastError();
        return e.code;
    } catch (const std::bad_alloc&) {
        if (core) core->err = "host allocation failed";
        return SKB_ERR_NOMEM;
    } catch (const std::exception& e) {
        if (core) core->err = e.what();
        return SKB_ERR_CUDA;
    }
}

float elapsed(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

}  // namespace

extern "C" {

const char* skb_version(void) { return "pyskani_b200 libskb 0.1 (sm_100a)"; }

int skb_ctx_create(int device, skb_ctx_t** out) {
    if (!out) return SKB_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return SKB_ERR_CUDA; }
    auto core = std::make_shared<Core>();
    core->device = device;
    if (cudaSetDevice(device) != cudaSuccess) return SKB_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SKB_ERR_CUDA;
    core->n_sm = prop.multiProcessorCount;
    core->chain_batch_seeds = std::min<uint64_t>(384ull << 20, std::max<uint64_t>(16ull << 20, (uint64_t)((double)prop.totalGlobalMem * 0.15 / 72.0)));
    if (cudaStreamCreateWithFlags(&core->stream, cudaStreamNonBlocking) != cudaSuccess) return SKB_ERR_CUDA;
    if (cudaStreamCreateWithFlags(&core->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return SKB_ERR_CUDA;
    if (cudaStreamCreateWithFlags(&core->aux_stream, cudaStreamNonBlocking) != cudaSuccess) return SKB_ERR_CUDA;
    for (auto& e : core->ev) if (cudaEventCreate(&e) != cudaSuccess) return SKB_ERR_CUDA;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = new skb_ctx{core};
    return SKB_OK;
}

void skb_ctx_destroy(skb_ctx_t* ctx) { delete ctx; }
const char* skb_last_error(const skb_ctx_t* ctx) { return ctx ? ctx->core->err.c_str() : "null context"; }
int skb_ctx_stats(const skb_ctx_t* ctx, skb_stats_t* out) {
    if (!ctx || !out) return SKB_ERR_ARG;
    *out = ctx->core->stats;
    out->kernels_launched = g_kernel_launches;
    return SKB_OK;
}
int skb_ctx_sync(skb_ctx_t* ctx) {
    if (!ctx) return SKB_ERR_ARG;
    return guarded(ctx->core.get(), [&] { CU(cudaStreamSynchronize(ctx->core->stream)); return SKB_OK; });
}
void* skb_ctx_stream(skb_ctx_t* ctx) { return ctx ? (void*)ctx->core->stream : nullptr; }
int skb_ctx_set_priority(skb_ctx_t* ctx, int32_t high) {
    if (!ctx) return SKB_ERR_ARG;
    return guarded(ctx->core.get(), [&] {
        Core& c = *ctx->core;
        int least = 0, greatest = 0;              // numerically lower = higher priority; streams are created with `least`
        CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        for (cudaStream_t* sp : {&c.stream, &c.aux_stream}) {
            CU(cudaStreamSynchronize(*sp));
            cudaStream_t fresh = nullptr;
            CU(cudaStreamCreateWithPriority(&fresh, cudaStreamNonBlocking, high ? greatest : least));
            cudaStreamDestroy(*sp);
            *sp = fresh;
        }
        return SKB_OK;
    });
}
int skb_ctx_set_host_threads(skb_ctx_t* ctx, int32_t n) {
    if (!ctx) return SKB_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->core->mu);
    ctx->core->host_threads = n < 0 ? -1 : std::min<int32_t>(n, 256);
    return SKB_OK;
}

int skb_host_alloc(skb_ctx_t* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return SKB_ERR_ARG;
    return guarded(ctx->core.get(), [&] { CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault)); return SKB_OK; });
}
void skb_host_free(skb_ctx_t* ctx, void* p) { if (ctx && p) { cudaSetDevice(ctx->core->device); cudaFreeHost(p); } }
int skb_dev_alloc(skb_ctx_t* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return SKB_ERR_ARG;
    return guarded(ctx->core.get(), [&] { CU(cudaMalloc(out, bytes)); return SKB_OK; });
}
void skb_dev_free(skb_ctx_t* ctx, void* p) { if (ctx && p) { cudaSetDevice(ctx->core->device); cudaFree(p); } }
int skb_memcpy_h2d(skb_ctx_t* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return SKB_ERR_ARG;
    return guarded(ctx->core.get(), [&] {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->core->stream));
        CU(cudaStreamSynchronize(ctx->core->stream));
        return SKB_OK;
    });
}

int skb_sketch_batch_device(skb_ctx_t* ctx, const skb_sketch_params_t* params, int32_t seed, uint32_t n_genomes,
                            const uint32_t* genome_contig_start, const uint8_t* seq, const uint64_t* contig_offsets,
                            const uint64_t* contig_lens, skb_sketch_t** out) {
    if (!ctx || !params || !out || (n_genomes && !genome_contig_start)) return SKB_ERR_ARG;
    if (n_genomes && genome_contig_start[n_genomes] && (!contig_lens || !contig_offsets)) return SKB_ERR_ARG;
    return guarded(ctx->core.get(), [&] {
        Core& c = *ctx->core;
        CU(cudaEventRecord(c.ev[0], c.stream));
        sketch_core(ctx->core, *params, seed, n_genomes, genome_contig_start, seq, contig_offsets, contig_lens, out);
        CU(cudaEventRecord(c.ev[4], c.stream));
        CU(cudaStreamSynchronize(c.stream));
        c.stats.h2d_ms = 0; c.stats.seed_ms = elapsed(c.ev[1], c.ev[2]); c.stats.index_ms = elapsed(c.ev[2], c.ev[4]);
        c.stats.total_ms = elapsed(c.ev[0], c.ev[4]);
        return SKB_OK;
    });
}

int skb_sketch_batch(skb_ctx_t* ctx, const skb_sketch_params_t* params, int32_t seed, uint32_t n_genomes,
                     const uint32_t* genome_contig_start, const uint8_t* const* contigs, const uint64_t* contig_lens,
                     skb_sketch_t** out) {
    if (!ctx || !params || !out || (n_genomes && !genome_contig_start)) return SKB_ERR_ARG;
    // a genome without any contig is legal (the reference returns an empty sketch, lib.rs:155): the arrays may then be NULL
    if (n_genomes && genome_contig_start[n_genomes] && (!contig_lens || !contigs)) return SKB_ERR_ARG;
    return guarded(ctx->core.get(), [&] {
        Core& c = *ctx->core;
        cudaStream_t st = c.stream;
        const uint32_t n_contigs = n_genomes ? genome_contig_start[n_genomes] : 0;
        // device layout: 64 B guard | contig 0 (16-aligned) | contig 1 | ... | 64 B guard.  Short contigs are
        // dropped before the copy: they never reach the GPU (reference lib.rs:156).
        std::vector<uint64_t> offs(n_contigs, 0);
        uint64_t cur = 64;
        for (uint32_t i = 0; i < n_contigs; i++) {
            if (contig_lens[i] < SKB_MIN_LENGTH_CONTIG) continue;
            offs[i] = cur;
            cur += align16(contig_lens[i]) + 16;
        }
        cur += 64;
        Trace tr("sketch_batch");
        CU(cudaEventRecord(c.ev[0], st));
        uint8_t* d_seq = (uint8_t*)c.scratch(SLOT_SEQ, cur);
        tr.mark("alloc d_seq");
        // The copies run on their own stream in chunks of ~CHUNK bytes; each chunk records an event, and the seeding
        // launch of that chunk waits on it, so the kernels of chunk i overlap the PCIe transfer of chunk i+1.
        // Large contigs are copied straight from the caller's memory (true DMA when it is pinned); small ones are
        // packed through pinned staging first.
        // (alternating the chunks between two copy streams was measured: 11.8 instead of 10.8 ms per 505 MB step)
        cudaStream_t cs = c.copy_stream;
        CU(cudaEventRecord(c.ev[5], st));
        CU(cudaStreamWaitEvent(cs, c.ev[5], 0));          // d_seq's allocation is ordered on `st`
        cudaEvent_t ev_c0 = nullptr, ev_c1 = nullptr;
        if (tr.on) { cudaEventCreate(&ev_c0); cudaEventCreate(&ev_c1); cudaEventRecord(ev_c0, cs); }
        // granularity of the copy -> seeding hand-over: 16 MB when this thread enqueues plain copies; 32 MB for batches that
        // go through the ingest team, whose per-chunk costs (claim, copy + event by a worker, wake-up of this thread, launch)
        // are larger: 1.25 GB packed by 15 threads takes 12.3 ms in 16 MB chunks, 10.4 ms in 32 MB chunks (with 512 KB pieces),
        // 10.8 ms in 64 MB chunks (the last chunk's latency starts to show)
        uint64_t kept_bytes = 0;
        for (uint32_t i = 0; i < n_contigs; i++) if (contig_lens[i] >= SKB_MIN_LENGTH_CONTIG) kept_bytes += contig_lens[i];
        const char* ingest_env = std::getenv("SKB_INGEST");      // raw | pack | mix | auto (= unset, but also for small batches)
        const bool want_raw = ingest_env && std::strcmp(ingest_env, "raw") == 0;
        const unsigned n_threads = want_raw ? 0 : resolve_host_threads(c);
        constexpr uint64_t INGEST_MIN = (uint64_t)64 << 20; // smaller calls are not worth waking the thread team for
        uint64_t CHUNK = (uint64_t)(n_threads >= 2 && kept_bytes >= INGEST_MIN ? 32 : 16) << 20;
        if (const char* e = std::getenv("SKB_CHUNK_KB")) CHUNK = std::max<uint64_t>(1, (uint64_t)std::atoll(e)) << 10;   // test hook
        constexpr uint64_t SUB = (uint64_t)192 << 20;       // genomes are indexed in sub-batches of about this size
        constexpr uint64_t SUB_MAX = (uint64_t)1536 << 20;  // upper bound of a sub-batch (the batch limit is 2^31 bases)
        // The batch is processed in sub-batches (whole genomes): while the index of sub-batch i is built, the copies of
        // the later sub-batches keep the PCIe link busy, so only the last sub-batch's index build is exposed after the
        // final byte has arrived.
        struct SubBatch { uint32_t g0, g1; std::vector<ChunkPlan> plan; };
        std::vector<SubBatch> subs;
        struct ChunkRange { uint32_t c0, c1; };
        std::vector<ChunkRange> ranges;                     // flat contig range of every chunk, in order
        // cut points (cumulative bytes): equal parts of about SUB bytes (measured best on B200 among head/tail splits:
        // the index build runs ~2.5x slower while the copy engine is busy, so parts must stay small enough to keep up)
        std::vector<uint64_t> cuts;
        for (uint64_t cut = SUB; cut + SUB / 2 < kept_bytes; cut += SUB) cuts.push_back(cut);
        if (const char* e = std::getenv("SKB_SUB_CUTS")) {     // tuning hook: cumulative cut points in MB, comma separated
            cuts.clear();
            for (const char* p = e; *p;) { char* end; const double mb = std::strtod(p, &end); if (end == p) break; cuts.push_back((uint64_t)(mb * 1048576.0)); p = *end ? end + 1 : end; }
        }
        {
            uint64_t in_chunk = 0, in_sub = 0, done_bytes = 0;
            uint32_t chunk_c0 = 0;
            SubBatch cur_sub{0, 0, {}};
            auto close_chunk = [&](uint32_t contig_end) {
                cudaEvent_t e = c.pool_event(ranges.size());
                if (!e) throw Fail{SKB_ERR_CUDA, "cannot create an event"};
                ranges.push_back(ChunkRange{chunk_c0, contig_end});
                cur_sub.plan.push_back(ChunkPlan{contig_end, e, nullptr});
                chunk_c0 = contig_end;
                in_chunk = 0;
            };
            for (uint32_t g = 0; g < n_genomes; g++) {
                for (uint32_t i = genome_contig_start[g]; i < genome_contig_start[g + 1]; i++) {
                    const uint64_t len = contig_lens[i];
                    if (len < SKB_MIN_LENGTH_CONTIG) continue;
                    in_chunk += len; in_sub += len; done_bytes += len;
                    if (in_chunk >= CHUNK && i + 1 < genome_contig_start[g + 1]) close_chunk(i + 1);
                }
                const bool last = g + 1 == n_genomes;
                const bool end_sub = last || (subs.size() < cuts.size() && done_bytes >= cuts[subs.size()]) || in_sub >= SUB_MAX;
                if (in_chunk >= CHUNK || end_sub) close_chunk(genome_contig_start[g + 1]);
                if (end_sub) {
                    cur_sub.g1 = g + 1;
                    subs.push_back(std::move(cur_sub));
                    cur_sub = SubBatch{g + 1, g + 1, {}};
                    in_sub = 0;
                }
            }
        }
        // ---- how the bytes travel
        bool pipelined = n_threads >= 2 && ranges.size() >= 2 && (kept_bytes >= INGEST_MIN || ingest_env != nullptr);
        // Large batches without a forced policy: pageable sources (Python bytes) are left to the packing threads altogether -
        // the driver stages a pageable copy through its own buffers at a fraction of the link rate and holds the stream's lock
        // meanwhile, which stalls the copies the packing threads enqueue (through the extension: 23 GB/s with the DMA route,
        // 80 GB/s without).  For pinned sources the context LEARNS which of three policies is fastest on this host: both
        // routes at once [0], every chunk packed [1], or plain copies without the team [2].  With enough threads the host's
        // memory system, not PCIe, is the limit and the DMA route's reads only slow the packing threads down (16-CPU box,
        // 1.25 GB: 13.2 ms mixed, 12.2 ms packed, 24.5 ms copied); with 4 threads 16.3 / 23.7 / 24.5 ms; with several ranks on
        // a host whose PCIe or memory fabric is already saturated by the copies, the team only costs.  The first six large
        // calls try every policy twice, later calls stay with the best one (5 % hysteresis) and re-measure the others in turn
        // every 32nd call.  The sketches do not depend on the choice.
        int learned = -1;                   // policy under measurement in this call, -1: none
        bool force_pack = false;
        if (pipelined && (!ingest_env || std::strcmp(ingest_env, "auto") == 0)) {
            bool pageable = false;
            for (const ChunkRange& r : ranges) {
                uint32_t i = r.c0;
                while (i < r.c1 && contig_lens[i] < SKB_MIN_LENGTH_CONTIG) i++;
                if (i == r.c1) continue;
                cudaPointerAttributes at{};
                if (cudaPointerGetAttributes(&at, contigs[i]) != cudaSuccess) { cudaGetLastError(); pageable = true; break; }
                if (at.type == cudaMemoryTypeUnregistered) { pageable = true; break; }
            }
            if (pageable) force_pack = true;
            else if (!ingest_env) {
                // two rounds over the three policies first (a single sample is too easily spoilt by what other ranks of the
                // host happen to try at that moment), then the incumbent stays unless another policy is 5 % better
                int pick = -1;
                if (c.ingest_calls < 6) pick = (int)(c.ingest_calls % 3);
                else {
                    int best = c.ingest_best;
                    for (int k = 0; k < 3; k++) if (c.ingest_rate[k] > 1.05 * c.ingest_rate[best]) best = k;
                    c.ingest_best = best;
                    pick = best;
                    if (c.ingest_calls % 32 == 31) pick = (best + 1 + (int)((c.ingest_calls / 32) % 2)) % 3;
                }
                c.ingest_calls++;
                learned = pick;
                if (pick == 2) pipelined = false;
            }
        }
        auto ingest_t0 = std::chrono::steady_clock::now();      // restarted below, after one-time allocations
        float seed_ms = 0;
        uint64_t link_raw = 0, link_packed = 0;
        if (!pipelined) {
            // all copies are enqueued up front by this thread
            RawCopier rc(c, cs, d_seq, nullptr);
            for (uint32_t i = 0; i < n_contigs && !rc.stage; i++)
                if (contig_lens[i] >= SKB_MIN_LENGTH_CONTIG && contig_lens[i] < RawCopier::DIRECT) rc.stage = (char*)ensure_pinned(c, RawCopier::STAGE);
            size_t ch = 0;
            for (SubBatch& sb : subs)
                for (ChunkPlan& cp : sb.plan) {
                    for (uint32_t i = ranges[ch].c0; i < ranges[ch].c1; i++)
                        if (contig_lens[i] >= SKB_MIN_LENGTH_CONTIG) rc.add(contigs[i], contig_lens[i], offs[i]);
                    rc.flush();
                    CU(cudaEventRecord(cp.ready, cs));
                    ch++;
                }
            link_raw = rc.bytes;
            if (tr.on) cudaEventRecord(ev_c1, cs);
            tr.mark("copies enqueued");
            for (SubBatch& sb : subs) {
                sketch_core(ctx->core, *params, seed, sb.g1 - sb.g0, genome_contig_start + sb.g0, d_seq, offs.data(), contig_lens,
                            out + sb.g0, &sb.plan);
                seed_ms += elapsed(c.ev[1], c.ev[2]);
            }
        } else {
            if (!c.team || c.team->size() != n_threads) {
                c.team.reset();
                const int dev = c.device;
                c.team.reset(new HostTeam(n_threads, [dev] { cudaSetDevice(dev); }));
            }
            uint8_t* d_pk = (uint8_t*)c.scratch(SLOT_SEQPK, cur / 4 + 64);
            if (c.pack_stage_bytes < cur / 4 + 64) {
                if (c.pack_stage) { CU(cudaStreamSynchronize(cs)); cudaFreeHost(c.pack_stage); c.pack_stage = nullptr; c.pack_stage_bytes = 0; }
                const size_t want = cur / 4 + cur / 32 + 4096;
                CU(cudaHostAlloc(&c.pack_stage, want, cudaHostAllocDefault));
                c.pack_stage_bytes = want;
            }
            if (!c.raw_stage) CU(cudaHostAlloc(&c.raw_stage, RawCopier::STAGE, cudaHostAllocDefault));
            char* stage_raw = (char*)c.raw_stage;
            CU(cudaEventRecord(c.ev[5], st));
            CU(cudaStreamWaitEvent(cs, c.ev[5], 0));      // d_pk's allocation is ordered on `st` as well
            int policy = ingest_env && std::strcmp(ingest_env, "pack") == 0 ? Ingest::PACK_ONLY : Ingest::MIX;
            if (force_pack || learned == 1) policy = Ingest::PACK_ONLY;
            Ingest ing(c, contigs, contig_lens, offs.data(), d_seq, d_pk, (char*)c.pack_stage, stage_raw, policy);
            ing.chunks.resize(ranges.size());
            ing.remaining.assign(ranges.size(), 0);
            {
                size_t ch = 0;
                for (SubBatch& sb : subs)
                    for (ChunkPlan& cp : sb.plan) {
                        ing.chunks[ch] = Ingest::Chunk{ranges[ch].c0, ranges[ch].c1, cp.ready, 0, false};
                        const uint32_t id = (uint32_t)ch;
                        Ingest* pi = &ing;
                        cp.host_wait = [pi, id] { return pi->wait(id); };
                        ch++;
                    }
            }
            // the team must have left `ing` before this frame unwinds, whatever happens below
            struct TeamGuard {
                Ingest& ing; HostTeam& team; bool finished;
                ~TeamGuard() { if (!finished) ing.abort.store(true); team.wait(); }
            };
            ingest_t0 = std::chrono::steady_clock::now();
            c.team->launch([&ing](unsigned id) { ing.worker(id); });
            {
                TeamGuard guard{ing, *c.team, false};
                tr.mark("ingest team started");
                for (SubBatch& sb : subs) {
                    sketch_core(ctx->core, *params, seed, sb.g1 - sb.g0, genome_contig_start + sb.g0, d_seq, offs.data(), contig_lens,
                                out + sb.g0, &sb.plan, d_pk);
                    seed_ms += elapsed(c.ev[1], c.ev[2]);
                }
                guard.finished = true;                     // every chunk was consumed above: the team is leaving on its own
            }
            if (ing.err) throw Fail{ing.err, ing.err_msg};
            link_raw = ing.raw_bytes; link_packed = ing.packed_bytes;

            if (tr.on) cudaEventRecord(ev_c1, cs);
            if (tr.on) std::fprintf(stderr, "[skb] sketch_batch: ingest by %u threads (%s): %.1f MB as ASCII, %.1f MB as 2-bit words (= %.1f MB of bases); "
                                    "workers: %.2f ms packing, %.2f ms in copy/event calls, %.2f ms at the queue lock (summed over threads), last piece packed %.2f ms after the team started\n",
                                    n_threads, host_pack_isa(), link_raw / 1048576.0, link_packed / 1048576.0, link_packed * 4 / 1048576.0,
                                    ing.pack_ns.load() / 1e6, ing.finish_ns.load() / 1e6, ing.lock_ns.load() / 1e6, ing.last_piece_ns.load() / 1e6);
        }
        tr.mark("sketch_core done");
        CU(cudaEventRecord(c.ev[4], st));
        CU(cudaStreamSynchronize(st));
        tr.mark("final sync");
        if (learned >= 0) {
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - ingest_t0).count();
            const double rate = (double)kept_bytes / std::max(sec, 1e-6) / 1e9;
            // disturbances only ever make a call slower: remember the best recent rate of each policy (slowly forgotten)
            double& r = c.ingest_rate[learned];
            r = std::max(0.97 * r, rate);
        }
        if (tr.on) {
            float cms = 0, lead = 0;
            cudaEventElapsedTime(&cms, ev_c0, ev_c1); cudaEventElapsedTime(&lead, c.ev[0], ev_c0);
            std::fprintf(stderr, "[skb] sketch_batch: copy stream busy %.3f ms, started %.3f ms after the call's first event\n", cms, lead);
            cudaEventDestroy(ev_c0); cudaEventDestroy(ev_c1);
        }
        c.stats.h2d_raw_bytes = link_raw; c.stats.h2d_packed_bytes = link_packed;
        c.stats.h2d_ms = 0; c.stats.seed_ms = seed_ms;     // seeding launches wait on the copies: this includes PCIe time
        c.stats.total_ms = elapsed(c.ev[0], c.ev[4]); c.stats.index_ms = c.stats.total_ms - seed_ms;
        return SKB_OK;
    });
}

void skb_sketch_free(skb_sketch_t* s) {
    if (!s) return;
    std::shared_ptr<Core> core = s->impl ? s->impl->core : nullptr;   // outlives the lock below
    if (core) {
        std::lock_guard<std::mutex> lk(core->mu);
        cudaSetDevice(core->device);
        delete s;
    } else {
        delete s;
    }
}

void skb_sketch_free_many(uint32_t n, skb_sketch_t* const* s) {
    // one lock for the whole set: a database's worth of handles is released in microseconds instead of n calls
    for (uint32_t i = 0; i < n;) {
        if (!s[i]) { i++; continue; }
        std::shared_ptr<Core> core = s[i]->impl ? s[i]->impl->core : nullptr;
        if (!core) { delete s[i]; i++; continue; }
        std::lock_guard<std::mutex> lk(core->mu);
        cudaSetDevice(core->device);
        for (; i < n && (!s[i] || !s[i]->impl || s[i]->impl->core == core); i++) delete s[i];
    }
}

int skb_sketch_info(const skb_sketch_t* s, skb_sketch_info_t* out) {
    if (!s || !out) return SKB_ERR_ARG;
    *out = s->impl->info;
    return SKB_OK;
}

int skb_sketch_export(const skb_sketch_t* s, uint64_t* kmer, uint32_t* pos, uint32_t* contig, uint8_t* canonical,
                      uint64_t* markers, uint32_t* contig_lengths) {
    if (!s) return SKB_ERR_ARG;
    SketchImpl& I = *s->impl;
    return guarded(I.core.get(), [&] {
        Core& c = *I.core;
        // markers-only exports (Database.flush / save_markers) must not pull the seed arrays off the device
        const bool want_seeds = kmer || pos || contig || canonical;
        if (want_seeds) ensure_seeds_ready(I);
        const uint32_t n = want_seeds ? I.view.n_seeds : 0;
        std::vector<uint32_t> k32(kmer ? n : 0), p32(pos ? n : 0), m32(contig || canonical ? n : 0);
        if (kmer) download(c, k32.data(), I.view.kmer_k, n);
        if (pos) download(c, p32.data(), I.view.pos_k, n);
        if (contig || canonical) download(c, m32.data(), I.view.meta_k, n);
        if (markers) download(c, markers, I.view.markers, I.view.n_markers);
        CU(cudaStreamSynchronize(c.stream));
        for (uint32_t i = 0; i < n; i++) {
            if (kmer) kmer[i] = k32[i];
            if (pos) pos[i] = p32[i];
            if (contig) contig[i] = m32[i] >> 1;
            if (canonical) canonical[i] = (uint8_t)(m32[i] & 1u);
        }
        if (contig_lengths) std::memcpy(contig_lengths, I.contig_len_host.data(), 4 * I.contig_len_host.size());
        return SKB_OK;
    });
}

int skb_sketch_import(skb_ctx_t* ctx, const skb_sketch_params_t* params, int32_t has_seeds, uint64_t n_seeds,
                      const uint64_t* kmer, const uint32_t* pos, const uint32_t* contig, const uint8_t* canonical,
                      uint64_t n_markers, const uint64_t* markers, uint32_t n_contigs, const uint32_t* contig_lengths,
                      skb_sketch_t** out) {
    if (!ctx || !params || !out) return SKB_ERR_ARG;
    if (n_seeds && (!kmer || !pos || !contig || !canonical)) return SKB_ERR_ARG;
    if (n_markers && !markers) return SKB_ERR_ARG;
    if (n_contigs && !contig_lengths) return SKB_ERR_ARG;
    return guarded(ctx->core.get(), [&] {
        Core& c = *ctx->core;
        cudaStream_t st = c.stream;
        if (params->k < 1 || params->k > 16) throw Fail{SKB_ERR_ARG, "k must be in 1..16"};
        if (n_seeds >= 0x7FFFFFFFull || n_markers >= 0x7FFFFFFFull) throw Fail{SKB_ERR_ARG, "sketch too large"};
        // position order on the host (import is the Database.load path, not a hot path), then the same
        // device index build as a fresh sketch
        const uint32_t n = (uint32_t)n_seeds;
        std::vector<uint32_t> ord(n);
        for (uint32_t i = 0; i < n; i++) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) {
            if (contig[a] != contig[b]) return contig[a] < contig[b];
            return pos[a] < pos[b];
        });
        std::vector<uint32_t> hk(n), hp(n), hm(n);
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t s = ord[i];
            if (contig[s] >= n_contigs) throw Fail{SKB_ERR_ARG, "seed refers to a contig that does not exist"};
            if (kmer[s] >> (2 * params->k)) throw Fail{SKB_ERR_ARG, "k-mer wider than 2k bits"};
            hk[i] = (uint32_t)kmer[s]; hp[i] = pos[s]; hm[i] = (contig[s] << 1) | (canonical[s] ? 1u : 0u);
        }
        std::vector<uint64_t> hmark(markers, markers + n_markers);
        std::sort(hmark.begin(), hmark.end());
        hmark.erase(std::unique(hmark.begin(), hmark.end()), hmark.end());
        auto store = std::make_shared<BatchStore>();
        store->kmer_p = DevMem::persistent(ctx->core, 4 * (size_t)n); store->pos_p = DevMem::persistent(ctx->core, 4 * (size_t)n);
        store->meta_p = DevMem::persistent(ctx->core, 4 * (size_t)n); store->kmer_k = DevMem::persistent(ctx->core, 4 * (size_t)n);
        store->pos_k = DevMem::persistent(ctx->core, 4 * (size_t)n); store->meta_k = DevMem::persistent(ctx->core, 4 * (size_t)n);
        store->perm_k = DevMem::persistent(ctx->core, 4 * (size_t)n);
        store->markers = DevMem::persistent(ctx->core, 8 * std::max<size_t>(hmark.size(), 1));
        CU(cudaMemcpyAsync(store->kmer_p.p, hk.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(store->pos_p.p, hp.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(store->meta_p.p, hm.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(store->markers.p, hmark.data(), 8 * hmark.size(), cudaMemcpyHostToDevice, st));
        std::vector<uint32_t> seed_start{0, n}, marker_start{0, (uint32_t)hmark.size()};
        if (n) {
            DevMem d_gs(ctx->core, 8);
            CU(cudaMemcpyAsync(d_gs.p, seed_start.data(), 8, cudaMemcpyHostToDevice, st));
            DevMem scratch(ctx->core, kmer_order_scratch_bytes(n));
            IndexBuildArgs ib{};
            ib.n_genomes = 1; ib.n_seeds_total = n; ib.genome_seed_start = d_gs.as<uint32_t>(); ib.max_genome_seeds = n;
            ib.kmer_p = store->kmer_p.as<uint32_t>(); ib.pos_p = store->pos_p.as<uint32_t>(); ib.meta_p = store->meta_p.as<uint32_t>();
            ib.kmer_k = store->kmer_k.as<uint32_t>(); ib.pos_k = store->pos_k.as<uint32_t>(); ib.meta_k = store->meta_k.as<uint32_t>();
            ib.perm_k = store->perm_k.as<uint32_t>();
            ib.k = params->k;
            build_kmer_order(ib, scratch.p, scratch.bytes, st);
            CU(cudaStreamSynchronize(st));   // host vectors above go out of scope after this call
        }
        CU(cudaStreamSynchronize(st));
        std::vector<std::vector<uint32_t>> cl(1);
        cl[0].assign(contig_lengths, contig_lengths + n_contigs);
        finish_batch(ctx->core, *params, has_seeds, 1, cl, store, seed_start, marker_start, nullptr, out);
        return SKB_OK;
    });
}

int skb_db_create(skb_ctx_t* ctx, skb_db_t** out) {
    if (!ctx || !out) return SKB_ERR_ARG;
    *out = new skb_db{ctx->core};
    return SKB_OK;
}
void skb_db_destroy(skb_db_t* db) {
    if (!db) return;
    auto core = db->core;
    std::lock_guard<std::mutex> lk(core->mu);
    cudaSetDevice(core->device);
    delete db;
}
int skb_db_add(skb_db_t* db, skb_sketch_t* s, uint32_t* index_out) {
    if (!db || !s) return SKB_ERR_ARG;
    return guarded(db->core.get(), [&] {
        if (s->impl->core->device != db->core->device) throw Fail{SKB_ERR_ARG, "sketch belongs to a context of another device"};
        if (!db->items.empty()) {
            const auto& a = db->items[0]->info; const auto& b = s->impl->info;
            if (a.k != b.k || a.c != b.c || a.marker_c != b.marker_c) throw Fail{SKB_ERR_ARG, "sketch parameters differ from the database's"};
        }
        if (index_out) *index_out = (uint32_t)db->items.size();
        db->items.push_back(s->impl);
        db->dirty = true; db->idx_dirty = true;
        return SKB_OK;
    });
}
int skb_db_add_many(skb_db_t* db, uint32_t n, skb_sketch_t* const* sketches, uint32_t* index_out) {
    if (!db || (n && !sketches)) return SKB_ERR_ARG;
    return guarded(db->core.get(), [&] {
        for (uint32_t i = 0; i < n; i++) {
            if (!sketches[i]) throw Fail{SKB_ERR_ARG, "null sketch"};
            if (sketches[i]->impl->core->device != db->core->device) throw Fail{SKB_ERR_ARG, "sketch belongs to a context of another device"};
            const auto& a0 = db->items.empty() ? sketches[0]->impl->info : db->items[0]->info;
            const auto& b0 = sketches[i]->impl->info;
            if (a0.k != b0.k || a0.c != b0.c || a0.marker_c != b0.marker_c) throw Fail{SKB_ERR_ARG, "sketch parameters differ from the database's"};
        }
        if (index_out) *index_out = (uint32_t)db->items.size();
        for (uint32_t i = 0; i < n; i++) db->items.push_back(sketches[i]->impl);
        db->dirty = true; db->idx_dirty = true;
        return SKB_OK;
    });
}
int skb_db_replace(skb_db_t* db, uint32_t index, skb_sketch_t* s) {
    if (!db || !s) return SKB_ERR_ARG;
    return guarded(db->core.get(), [&] {
        if (index >= db->items.size()) throw Fail{SKB_ERR_KEY, "no sketch at this index"};
        if (s->impl->core->device != db->core->device) throw Fail{SKB_ERR_ARG, "sketch belongs to a context of another device"};
        const auto& a = db->items[0]->info; const auto& b = s->impl->info;
        if (db->items.size() > 1 && (a.k != b.k || a.c != b.c || a.marker_c != b.marker_c)) throw Fail{SKB_ERR_ARG, "sketch parameters differ from the database's"};
        CU(cudaStreamSynchronize(db->core->stream));     // nothing in flight may still read the old sketch's arrays
        db->items[index] = s->impl;
        db->dirty = true; db->idx_dirty = true;
        return SKB_OK;
    });
}
uint64_t skb_db_size(const skb_db_t* db) { return db ? db->items.size() : 0; }

int skb_model_load_json(skb_ctx_t* ctx, const char* json, size_t len, skb_model_t** out) {
    if (!ctx || !json || !out) return SKB_ERR_ARG;
    *out = nullptr;
    return guarded(ctx->core.get(), [&] {
        auto m = std::make_shared<ModelImpl>();
        m->core = ctx->core;
        try { m->host = gbdt_parse(json, len); }
        catch (const std::runtime_error& e) { throw Fail{SKB_ERR_ARG, e.what()}; }
        Core& c = *ctx->core;
        m->nodes = DevMem::persistent(ctx->core, sizeof(GbdtNode) * std::max<size_t>(m->host.nodes.size(), 1));
        m->tree_off = DevMem::persistent(ctx->core, 4 * m->host.tree_off.size());
        CU(cudaMemcpyAsync(m->nodes.p, m->host.nodes.data(), sizeof(GbdtNode) * m->host.nodes.size(), cudaMemcpyHostToDevice, c.stream));
        CU(cudaMemcpyAsync(m->tree_off.p, m->host.tree_off.data(), 4 * m->host.tree_off.size(), cudaMemcpyHostToDevice, c.stream));
        CU(cudaStreamSynchronize(c.stream));
        *out = new skb_model{m};
        return SKB_OK;
    });
}
void skb_model_free(skb_model_t* m) {
    if (!m) return;
    std::shared_ptr<Core> core = m->impl ? m->impl->core : nullptr;
    if (core) { std::lock_guard<std::mutex> lk(core->mu); cudaSetDevice(core->device); delete m; }
    else delete m;
}
int skb_model_info(const skb_model_t* m, uint32_t* n_trees, uint32_t* n_nodes, uint32_t* n_features) {
    if (!m) return SKB_ERR_ARG;
    if (n_trees) *n_trees = (uint32_t)m->impl->host.tree_off.size() - 1;
    if (n_nodes) *n_nodes = (uint32_t)m->impl->host.nodes.size();
    if (n_features) *n_features = m->impl->host.n_features;
    return SKB_OK;
}
int skb_model_predict(skb_model_t* m, const float* rows, uint32_t n_rows, uint32_t n_features, float* out) {
    if (!m || (n_rows && (!rows || !out))) return SKB_ERR_ARG;
    return guarded(m->impl->core.get(), [&] {
        if (n_features < m->impl->host.n_features || n_features > GBDT_FEATURES) throw Fail{SKB_ERR_ARG, "feature rows are narrower than the model's feature_size (or wider than 10)"};
        if (n_rows == 0) return SKB_OK;
        Core& c = *m->impl->core;
        DevMem d_rows(m->impl->core, 4 * (size_t)n_rows * n_features), d_out(m->impl->core, 4 * (size_t)n_rows);
        CU(cudaMemcpyAsync(d_rows.p, rows, 4 * (size_t)n_rows * n_features, cudaMemcpyHostToDevice, c.stream));
        launch_gbdt_predict(m->impl->dev_view(), d_rows.as<float>(), n_rows, n_features, d_out.as<float>(), c.stream);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, d_out.p, 4 * (size_t)n_rows, cudaMemcpyDeviceToHost, c.stream));
        CU(cudaStreamSynchronize(c.stream));
        return SKB_OK;
    });
}
int skb_db_set_model(skb_db_t* db, skb_model_t* m) {
    if (!db) return SKB_ERR_ARG;
    return guarded(db->core.get(), [&] {
        if (m && m->impl->core != db->core) throw Fail{SKB_ERR_ARG, "model belongs to another context"};
        CU(cudaStreamSynchronize(db->core->stream));
        db->model = m ? m->impl : nullptr;
        return SKB_OK;
    });
}

void skb_hits_free(skb_hit_t* hits) { delete[] hits; }

}  // extern "C"


// ---------------------------------------------------------------- device-to-device transfer of sketches
namespace {
constexpr uint32_t PACK_MAGIC = 0x534B4251u;   // "SKBQ"
constexpr int PACK_ARRAYS = 12;
// A packed set of sketches has a HEAD (descriptor, marker sets, per-contig tables: all the screen needs) and a BODY (seed
// arrays and bucket tables: what chaining needs).  off[] of head arrays is relative to the head payload, of body arrays
// to the body payload; skb_sketch_pack puts the body right behind the head, an exchange block keeps them in two regions
// so that they can travel as two collectives.
struct PackHeader { uint32_t magic, n; uint64_t head_bytes, body_bytes; uint32_t ref_only, pad; };
struct PackSketch {
    uint64_t total_len;
    uint32_t n_seeds, n_markers, n_contigs, bucket_shift, n_buckets, win_cap;
    int32_t k, c, marker_c, has_seeds;
    uint64_t off[PACK_ARRAYS];                 // kmer_p pos_p meta_p perm_k kmer_k pos_k meta_k bucket | cstart clen cwin markers
};
constexpr bool pack_is_head(int a) { return a >= 8; }
inline uint64_t pad16(uint64_t x) { return (x + 15) & ~(uint64_t)15; }
inline uint64_t pad256(uint64_t x) { return (x + 255) & ~(uint64_t)255; }
// byte sizes of the eleven device arrays of one sketch, in PackSketch::off order.  ref_only leaves out what only a QUERY
// needs (position-order seeds, per-contig seed starts): half of the bytes of a sketch.
void pack_sizes(const GenomeView& v, bool ref_only, uint64_t* sz) {
    const uint64_t n = v.n_seeds, nc = v.n_contigs;
    for (int i = 0; i < 4; i++) sz[i] = ref_only || !v.kmer_p ? 0 : 4 * n;        // what only a query needs
    for (int i = 4; i < 7; i++) sz[i] = 4 * n;
    sz[7] = v.bucket ? 4 * ((uint64_t)v.n_buckets + 1) : 0;
    sz[8] = ref_only || !v.contig_seed_start ? 0 : 4 * (nc + 1); sz[9] = 4 * nc; sz[10] = 4 * (nc + 1);
    sz[11] = 8 * (uint64_t)v.n_markers;
}
void pack_totals(uint32_t n, skb_sketch_t* const* sketches, bool ref_only, uint64_t& head, uint64_t& body, uint64_t& meta) {
    head = body = 0;
    meta = sizeof(PackHeader) + (uint64_t)n * sizeof(PackSketch);
    for (uint32_t i = 0; i < n; i++) {
        if (!sketches[i]) throw Fail{SKB_ERR_ARG, "null sketch"};
        uint64_t sz[PACK_ARRAYS];
        pack_sizes(sketches[i]->impl->view, ref_only || sketches[i]->impl->ref_only, sz);
        for (int a = 0; a < PACK_ARRAYS; a++) (pack_is_head(a) ? head : body) += pad16(sz[a]);
        meta += 4 * (uint64_t)sketches[i]->impl->view.n_contigs;
    }
}
}  // namespace

extern "C" {

int skb_sketch_pack_size(uint32_t n, skb_sketch_t* const* sketches, uint64_t* payload_bytes, uint64_t* meta_bytes) {
    if ((n && !sketches) || !payload_bytes || !meta_bytes) return SKB_ERR_ARG;
    try {
        uint64_t head = 0, body = 0, meta = 0;
        pack_totals(n, sketches, false, head, body, meta);
        *payload_bytes = pad256(head) + body; *meta_bytes = meta;
    } catch (const Fail&) { return SKB_ERR_ARG; }
    return SKB_OK;
}

}  // extern "C"

namespace {

// Fills the host descriptor of n sketches and lists the device arrays to gather: head arrays go to head_dst + off, body
// arrays to body_dst + off.  Shared by skb_sketch_pack and skb_exchange_pack.
void build_pack_meta(Core& c, uint32_t n, skb_sketch_t* const* sketches, bool ref_only, char* m, uint64_t head_dst, uint64_t body_dst,
                     std::vector<SegmentCopy>& segs, uint64_t& max_bytes) {
    uint64_t head = 0, body = 0, meta = 0;
    pack_totals(n, sketches, ref_only, head, body, meta);
    bool any_ref_only = ref_only;
    PackSketch* ps = (PackSketch*)(m + sizeof(PackHeader));
    uint32_t* clens = (uint32_t*)(m + sizeof(PackHeader) + (uint64_t)n * sizeof(PackSketch));
    uint64_t off_h = 0, off_b = 0;
    max_bytes = 0;
    for (uint32_t i = 0; i < n; i++) {
        const SketchImpl& I = *sketches[i]->impl;
        if (I.core.get() != &c) throw Fail{SKB_ERR_ARG, "sketch belongs to another context"};
        ensure_seeds_ready(I);
        any_ref_only = any_ref_only || I.ref_only;
        const GenomeView& v = I.view;
        PackSketch p{};
        p.total_len = v.total_len; p.n_seeds = v.n_seeds; p.n_markers = v.n_markers; p.n_contigs = v.n_contigs;
        p.bucket_shift = v.bucket_shift; p.n_buckets = v.n_buckets; p.win_cap = v.win_cap;
        p.k = I.info.k; p.c = I.info.c; p.marker_c = I.info.marker_c; p.has_seeds = I.info.has_seeds;
        uint64_t sz[PACK_ARRAYS];
        pack_sizes(v, ref_only || I.ref_only, sz);
        const void* src[PACK_ARRAYS] = {v.kmer_p, v.pos_p, v.meta_p, v.perm_k, v.kmer_k, v.pos_k, v.meta_k, v.bucket,
                                        v.contig_seed_start, v.contig_len, v.contig_win_start, v.markers};
        for (int a = 0; a < PACK_ARRAYS; a++) {
            uint64_t& off = pack_is_head(a) ? off_h : off_b;
            p.off[a] = off;
            if (sz[a]) {
                segs.push_back(SegmentCopy{src[a], (pack_is_head(a) ? head_dst : body_dst) + off, sz[a]});
                max_bytes = std::max(max_bytes, sz[a]);
            }
            off += pad16(sz[a]);
        }
        std::memcpy(&ps[i], &p, sizeof(p));
        if (v.n_contigs) std::memcpy(clens, I.contig_len_host.data(), 4 * (size_t)v.n_contigs);
        clens += v.n_contigs;
    }
    PackHeader hd{PACK_MAGIC, n, head, body, any_ref_only ? 1u : 0u, 0u};
    std::memcpy(m, &hd, sizeof(hd));
}

// Sketch handles whose arrays are views into the head / body payloads the descriptor `m` describes; `store` keeps the
// memory alive.  Returns the number of sketches.
uint32_t views_from_meta(const std::shared_ptr<Core>& core, const std::shared_ptr<BatchStore>& store, const char* m,
                         uint64_t meta_bytes, const char* head_base, const char* body_base, skb_sketch_t** out) {
    PackHeader hd;
    std::memcpy(&hd, m, sizeof(hd));
    if (meta_bytes < sizeof(PackHeader) + (uint64_t)hd.n * sizeof(PackSketch)) throw Fail{SKB_ERR_ARG, "truncated pack descriptor"};
    const PackSketch* ps = (const PackSketch*)(m + sizeof(PackHeader));
    const uint32_t* clens = (const uint32_t*)(m + sizeof(PackHeader) + (uint64_t)hd.n * sizeof(PackSketch));
    uint64_t nc_total = 0;
    for (uint32_t i = 0; i < hd.n; i++) { PackSketch p; std::memcpy(&p, &ps[i], sizeof(p)); nc_total += p.n_contigs; }
    if (meta_bytes < sizeof(PackHeader) + (uint64_t)hd.n * sizeof(PackSketch) + 4 * nc_total) throw Fail{SKB_ERR_ARG, "truncated pack descriptor"};
    const bool ref_only = hd.ref_only != 0;
    for (uint32_t i = 0; i < hd.n; i++) {
        PackSketch p;
        std::memcpy(&p, &ps[i], sizeof(p));
        GenomeView v{};
        if (!ref_only) {
            v.kmer_p = (const uint32_t*)(body_base + p.off[0]); v.pos_p = (const uint32_t*)(body_base + p.off[1]);
            v.meta_p = (const uint32_t*)(body_base + p.off[2]); v.perm_k = (const uint32_t*)(body_base + p.off[3]);
            v.contig_seed_start = (const uint32_t*)(head_base + p.off[8]);
        }
        v.kmer_k = (const uint32_t*)(body_base + p.off[4]);
        v.pos_k = (const uint32_t*)(body_base + p.off[5]); v.meta_k = (const uint32_t*)(body_base + p.off[6]);
        v.bucket = (const uint32_t*)(body_base + p.off[7]);
        v.contig_len = (const uint32_t*)(head_base + p.off[9]); v.contig_win_start = (const uint32_t*)(head_base + p.off[10]);
        v.markers = (const uint64_t*)(head_base + p.off[11]);
        v.total_len = p.total_len; v.n_seeds = p.n_seeds; v.n_markers = p.n_markers; v.n_contigs = p.n_contigs;
        v.bucket_shift = p.bucket_shift; v.n_buckets = p.n_buckets; v.win_cap = p.win_cap;
        auto impl = std::make_shared<SketchImpl>();
        impl->core = core; impl->store = store; impl->view = v; impl->ref_only = ref_only;
        impl->contig_len_host.assign(clens, clens + p.n_contigs);
        contig_quantiles(impl->contig_len_host, impl->view);
        clens += p.n_contigs;
        impl->info.n_seeds = p.n_seeds; impl->info.n_markers = p.n_markers; impl->info.total_len = p.total_len;
        impl->info.n_contigs = p.n_contigs; impl->info.k = p.k; impl->info.c = p.c; impl->info.marker_c = p.marker_c;
        impl->info.has_seeds = p.has_seeds; impl->info.reference_only = ref_only ? 1 : 0;
        out[i] = new skb_sketch{impl};
    }
    return hd.n;
}

}  // namespace

extern "C" {

int skb_sketch_pack(skb_ctx_t* ctx, uint32_t n, skb_sketch_t* const* sketches, void* payload_dev, uint64_t payload_bytes,
                    void* meta_host, uint64_t meta_bytes) {
    if (!ctx || (n && !sketches) || !meta_host || (payload_bytes && !payload_dev)) return SKB_ERR_ARG;
    return guarded(ctx->core.get(), [&] {
        Core& c = *ctx->core;
        uint64_t head = 0, body = 0, need_m = 0;
        pack_totals(n, sketches, false, head, body, need_m);
        if (payload_bytes < pad256(head) + body || meta_bytes < need_m) throw Fail{SKB_ERR_ARG, "pack buffers too small (see skb_sketch_pack_size)"};
        std::vector<SegmentCopy> segs;
        uint64_t max_bytes = 0;
        build_pack_meta(c, n, sketches, false, (char*)meta_host, 0, pad256(head), segs, max_bytes);
        if (!segs.empty()) {
            const size_t tb = sizeof(SegmentCopy) * segs.size();
            void* d_segs = c.scratch(SLOT_PACK, tb);
            table_upload(c, d_segs, segs.data(), tb);
            launch_segment_copy((const SegmentCopy*)d_segs, (uint32_t)segs.size(), max_bytes, payload_dev, c.stream);
            CU(cudaGetLastError());
        }
        CU(cudaStreamSynchronize(c.stream));      // the payload is complete (and `segs` no longer needed) on return
        return SKB_OK;
    });
}

int skb_sketch_unpack(skb_ctx_t* ctx, const void* meta_host, uint64_t meta_bytes, const void* payload_dev, uint64_t payload_bytes,
                      skb_sketch_t** out, uint32_t out_cap, uint32_t* n_out) {
    if (!ctx || !meta_host || meta_bytes < sizeof(PackHeader) || !n_out) return SKB_ERR_ARG;
    return guarded(ctx->core.get(), [&] {
        Core& c = *ctx->core;
        PackHeader hd;
        std::memcpy(&hd, meta_host, sizeof(hd));
        if (hd.magic != PACK_MAGIC) throw Fail{SKB_ERR_ARG, "not a sketch pack descriptor"};
        const uint64_t total = pad256(hd.head_bytes) + hd.body_bytes;
        if (payload_bytes < total || (total && !payload_dev)) throw Fail{SKB_ERR_ARG, "truncated pack payload"};
        *n_out = hd.n;
        if (hd.n == 0) return SKB_OK;
        if (!out || out_cap < hd.n) throw Fail{SKB_ERR_ARG, "output array too small for the packed sketches"};
        auto store = std::make_shared<BatchStore>();
        store->blob = DevMem::persistent(ctx->core, std::max<uint64_t>(total, 16));
        if (total) CU(cudaMemcpyAsync(store->blob.p, payload_dev, total, cudaMemcpyDeviceToDevice, c.stream));
        const char* base = (const char*)store->blob.p;
        views_from_meta(ctx->core, store, (const char*)meta_host, meta_bytes, base, base + pad256(hd.head_bytes), out);
        CU(cudaStreamSynchronize(c.stream));      // the caller may reuse the payload buffer on return
        return SKB_OK;
    });
}

}  // extern "C"

// ---------------------------------------------------------------- zero-copy exchange region
struct skb_exchange {
    std::shared_ptr<skb::Core> core;
    std::shared_ptr<skb::BatchStore> store;     // store->blob is the block; adopted sketches share it
    uint64_t bytes = 0;
};

extern "C" {

int skb_exchange_segment_size(uint32_t n, skb_sketch_t* const* sketches, int32_t reference_only, uint64_t* head_bytes,
                              uint64_t* body_bytes, uint64_t* meta_bytes) {
    if ((n && !sketches) || !head_bytes || !body_bytes || !meta_bytes) return SKB_ERR_ARG;
    try {
        uint64_t head = 0, body = 0, meta = 0;
        pack_totals(n, sketches, reference_only != 0, head, body, meta);
        *meta_bytes = meta;
        *head_bytes = pad256(meta) + pad256(head);
        *body_bytes = pad256(body);
    } catch (const Fail&) { return SKB_ERR_ARG; }
    return SKB_OK;
}

int skb_exchange_create(skb_ctx_t* ctx, uint64_t bytes, skb_exchange_t** out) {
    if (!ctx || !out) return SKB_ERR_ARG;
    *out = nullptr;
    return guarded(ctx->core.get(), [&] {
        auto ex = std::make_unique<skb_exchange>();
        ex->core = ctx->core;
        ex->store = std::make_shared<BatchStore>();
        ex->store->blob = DevMem::persistent(ctx->core, std::max<uint64_t>(bytes, 256));
        ex->bytes = bytes;
        *out = ex.release();
        return SKB_OK;
    });
}

void* skb_exchange_ptr(skb_exchange_t* ex) { return ex ? ex->store->blob.p : nullptr; }

void skb_exchange_free(skb_exchange_t* ex) {
    if (!ex) return;
    auto core = ex->core;
    std::lock_guard<std::mutex> lk(core->mu);
    cudaSetDevice(core->device);
    delete ex;
}

int skb_exchange_pack(skb_exchange_t* ex, uint64_t head_offset, uint64_t body_offset, uint32_t n, skb_sketch_t* const* sketches,
                      int32_t reference_only) {
    if (!ex || (n && !sketches) || (head_offset & 255) || (body_offset & 255)) return SKB_ERR_ARG;
    return guarded(ex->core.get(), [&] {
        Core& c = *ex->core;
        uint64_t head = 0, body = 0, need_m = 0;
        pack_totals(n, sketches, reference_only != 0, head, body, need_m);
        const uint64_t head_payload = head_offset + pad256(need_m);
        if (head_payload + pad256(head) > ex->bytes || body_offset + pad256(body) > ex->bytes)
            throw Fail{SKB_ERR_ARG, "segment does not fit into the exchange block"};
        // descriptor: built in pinned memory, copied to the front of the head segment; arrays: one gather kernel
        const size_t meta_pad = pad256(need_m);
        std::vector<SegmentCopy> segs;
        uint64_t max_bytes = 0;
        std::vector<char> meta(need_m);
        build_pack_meta(c, n, sketches, reference_only != 0, meta.data(), head_payload, body_offset, segs, max_bytes);
        const size_t tb = sizeof(SegmentCopy) * segs.size();
        char* h = (char*)ensure_pinned(c, meta_pad + tb + 64);
        std::memcpy(h, meta.data(), need_m);
        if (tb) std::memcpy(h + meta_pad, segs.data(), tb);
        char* blk = (char*)ex->store->blob.p;
        CU(cudaMemcpyAsync(blk + head_offset, h, need_m, cudaMemcpyHostToDevice, c.stream));
        if (!segs.empty()) {
            void* d_segs = c.scratch(SLOT_PACK, tb);
            CU(cudaMemcpyAsync(d_segs, h + meta_pad, tb, cudaMemcpyHostToDevice, c.stream));
            launch_segment_copy((const SegmentCopy*)d_segs, (uint32_t)segs.size(), max_bytes, blk, c.stream);
            CU(cudaGetLastError());
        }
        // the pinned staging block is shared by later calls: wait for the two small uploads (the gather kernel may
        // still be running; callers order their collective behind it on this stream)
        CU(cudaEventRecord(c.ev[5], c.stream));
        CU(cudaEventSynchronize(c.ev[5]));
        return SKB_OK;
    });
}

int skb_exchange_order_after(skb_exchange_t* ex, void* foreign_stream, int32_t bodies) {
    if (!ex) return SKB_ERR_ARG;
    return guarded(ex->core.get(), [&] {
        Core& c = *ex->core;
        BatchStore& st = *ex->store;
        if (!bodies) {
            // heads (descriptors + marker sets) have been enqueued on the caller's stream: this context waits for them now
            CU(cudaEventRecord(c.ev[5], (cudaStream_t)foreign_stream));
            CU(cudaStreamWaitEvent(c.stream, c.ev[5], 0));
        } else {
            // bodies (seed arrays): recorded now, waited for by whoever reads seed arrays of an adopted sketch first
            if (!st.body_ev) CU(cudaEventCreateWithFlags(&st.body_ev, cudaEventDisableTiming));
            CU(cudaEventRecord(st.body_ev, (cudaStream_t)foreign_stream));
            st.body_pending = true;
        }
        return SKB_OK;
    });
}

int skb_exchange_adopt(skb_exchange_t* ex, uint32_t n_segments, const uint64_t* head_offsets, const uint64_t* body_offsets,
                       const uint64_t* meta_bytes, skb_sketch_t** out, uint32_t out_cap, uint32_t* counts) {
    if (!ex || (n_segments && (!head_offsets || !body_offsets || !meta_bytes || !counts))) return SKB_ERR_ARG;
    return guarded(ex->core.get(), [&] {
        Core& c = *ex->core;
        uint64_t total_meta = 0;
        for (uint32_t i = 0; i < n_segments; i++) {
            if ((head_offsets[i] & 255) || (body_offsets[i] & 255) || meta_bytes[i] < sizeof(PackHeader) ||
                head_offsets[i] + meta_bytes[i] > ex->bytes || body_offsets[i] > ex->bytes)
                throw Fail{SKB_ERR_ARG, "bad exchange segment"};
            total_meta += pad256(meta_bytes[i]);
        }
        // all descriptors come to the host with one synchronisation
        char* h = (char*)ensure_pinned(c, total_meta + 64);
        const char* blk = (const char*)ex->store->blob.p;
        uint64_t ho = 0;
        for (uint32_t i = 0; i < n_segments; i++) {
            CU(cudaMemcpyAsync(h + ho, blk + head_offsets[i], meta_bytes[i], cudaMemcpyDeviceToHost, c.stream));
            ho += pad256(meta_bytes[i]);
        }
        CU(cudaStreamSynchronize(c.stream));
        uint32_t n_total = 0;
        ho = 0;
        for (uint32_t i = 0; i < n_segments; i++) {
            PackHeader hd;
            std::memcpy(&hd, h + ho, sizeof(hd));
            if (hd.magic != PACK_MAGIC) throw Fail{SKB_ERR_ARG, "exchange segment does not start with a sketch descriptor (collective not finished?)"};
            const uint64_t head_payload = head_offsets[i] + pad256(meta_bytes[i]);
            if (head_payload + hd.head_bytes > ex->bytes || body_offsets[i] + hd.body_bytes > ex->bytes)
                throw Fail{SKB_ERR_ARG, "exchange segment exceeds the block"};
            if (n_total + hd.n > out_cap || (hd.n && !out)) throw Fail{SKB_ERR_ARG, "output array too small for the adopted sketches"};
            views_from_meta(ex->core, ex->store, h + ho, meta_bytes[i], blk + head_payload, blk + body_offsets[i], out + n_total);
            counts[i] = hd.n;
            n_total += hd.n;
            ho += pad256(meta_bytes[i]);
        }
        return SKB_OK;
    });
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ query
namespace {

const GenomeView* db_views(skb_db& db) {
    Core& c = *db.core;
    if (db.dirty) {
        const size_t n = db.items.size();
        db.d_views = DevMem(db.core, sizeof(GenomeView) * std::max<size_t>(n, 1));
        std::vector<GenomeView> h(n);
        for (size_t i = 0; i < n; i++) h[i] = db.items[i]->view;
        // pageable source: the call returns once the bytes are staged, so `h` may go out of scope; consumers are
        // ordered behind the copy on the same stream
        CU(cudaMemcpyAsync(db.d_views.p, h.data(), sizeof(GenomeView) * n, cudaMemcpyHostToDevice, c.stream));
        db.dirty = false;
    }
    return db.d_views.as<GenomeView>();
}

// (Re)builds the marker index of the database if sketches were added since the last build.  Returns false if the
// database cannot be indexed (no markers, or more than 2^31 postings).
bool db_marker_index(skb_db& db, const GenomeView* d_r) {
    Core& c = *db.core;
    const uint32_t nr = (uint32_t)db.items.size();
    if (!db.idx_dirty) return db.idx_postings != 0;
    std::vector<uint32_t> off(nr + 1, 0);
    uint64_t total = 0;
    for (uint32_t i = 0; i < nr; i++) { off[i] = (uint32_t)total; total += db.items[i]->view.n_markers; }
    db.idx_dirty = false; db.idx_postings = 0;
    if (total == 0 || total >= 0x7FFFFFFFull) return false;
    off[nr] = (uint32_t)total;
    int B = 1;
    while (B < 24 && ((uint64_t)16 << B) < total) B++;
    const uint32_t nb = 1u << B;
    db.idx_shift = (uint32_t)(MARKER_BITS - B);
    db.idx_keys = DevMem::persistent(db.core, 8 * total);
    db.idx_vals = DevMem::persistent(db.core, 4 * total);
    db.idx_bucket = DevMem::persistent(db.core, 4 * ((size_t)nb + 1));
    const size_t off_bytes = (4 * ((size_t)nr + 1) + 255) & ~(size_t)255;
    const size_t scr_bytes = marker_index_scratch_bytes((uint32_t)total, nb);
    char* scr = (char*)c.scratch(SLOT_MIDX, off_bytes + 256 + scr_bytes);
    uint32_t* d_over = (uint32_t*)(scr + off_bytes);
    CU(cudaMemcpyAsync(scr, off.data(), 4 * ((size_t)nr + 1), cudaMemcpyHostToDevice, c.stream));
    // bucket partition first; a bucket beyond its capacity (many genomes sharing the same markers) falls back to the sort
    bool use_sort = std::getenv("SKB_MIDX_SORT") != nullptr;       // test hook
    for (int attempt = 0; attempt < 2; attempt++) {
        build_marker_index(d_r, nr, (const uint32_t*)scr, (uint32_t)total, db.idx_keys.as<uint64_t>(), db.idx_vals.as<uint32_t>(),
                           db.idx_bucket.as<uint32_t>(), db.idx_shift, nb, scr + off_bytes + 256, scr_bytes, use_sort, d_over, c.stream);
        uint32_t over = 0;
        if (!use_sort) download(c, &over, d_over, 1);
        CU(cudaStreamSynchronize(c.stream));      // `off` is pageable host memory
        if (!over) break;
        use_sort = true;
    }
    db.idx_postings = (uint32_t)total;
    return true;
}

struct ScreenOut {
    std::vector<uint32_t> pass_idx;   // q * n_refs + r, ascending
};

// x^21 in the order of Rust's f64::powi (LLVM's powi expansion and compiler-rt's __powidf2: square and multiply from the
// lowest exponent bit, one IEEE rounding per product), so that the cut-off is the double skani forms if it calls powi
double pow21(double x) {
    double r = 1.0;
    for (int b = SKB_MARKER_K;;) {
        if (b & 1) r *= x;
        b >>= 1;
        if (b == 0) break;
        x *= x;
    }
    return r;
}

// runs the screen for all (query, ref) pairs; optionally returns the dense arrays
void run_screen(skb_db& db, const std::vector<std::shared_ptr<SketchImpl>>& queries, const GenomeView* d_q,
                double screen_val, int rescue_small, ScreenOut* out, uint8_t* pass_host, uint32_t* shared_host) {
    Core& c = *db.core;
    cudaStream_t st = c.stream;
    const uint32_t nq = (uint32_t)queries.size(), nr = (uint32_t)db.items.size();
    const size_t n = (size_t)nq * nr;
    if (n == 0) return;
    if (n >= 0x7FFFFFFFull) throw Fail{SKB_ERR_ARG, "more than 2^31 pairs in one call; split the queries"};
    Trace ts("screen");
    const GenomeView* d_r = db_views(db);
    DevMem d_count(db.core, 4 * n), d_pass(db.core, n);
    ts.mark("views + buffers");
    std::vector<uint32_t> qm(nq);
    for (uint32_t i = 0; i < nq; i++) qm[i] = queries[i]->view.n_markers;
    // large pair matrices go through the database's marker index, small ones through the pairwise kernels
    // (SKB_SCREEN_MODE=index|pairwise forces one of them: used by the parity tests)
    bool use_index = n >= 16384;
    if (const char* e = std::getenv("SKB_SCREEN_MODE")) use_index = std::strcmp(e, "index") == 0;
    if (use_index && db_marker_index(db, d_r)) {
        ts.mark("marker index ready");
        launch_marker_join(d_q, nq, nr, db.idx_keys.as<uint64_t>(), db.idx_vals.as<uint32_t>(), db.idx_bucket.as<uint32_t>(),
                           db.idx_shift, d_count.as<uint32_t>(), st);
    } else
        launch_marker_screen(d_q, nq, d_r, nr, d_count.as<uint32_t>(), qm.data(), c.n_sm, st);
    launch_screen_decide(d_q, nq, d_r, nr, d_count.as<uint32_t>(), pow21(screen_val), screen_val == 0.0, rescue_small,
                         d_pass.as<uint8_t>(), st);
    CU(cudaGetLastError());
    ts.mark("join + decide enqueued");
    if (pass_host) download(c, pass_host, d_pass.as<uint8_t>(), n);
    if (shared_host) download(c, shared_host, d_count.as<uint32_t>(), n);
    if (out && n <= (1u << 16)) {
        // few pairs: fetch the flags themselves and list the survivors on the host (one synchronisation, no select pass)
        std::vector<uint8_t> flags(n);
        download(c, flags.data(), d_pass.as<uint8_t>(), n);
        CU(cudaStreamSynchronize(st));
        for (uint32_t i = 0; i < (uint32_t)n; i++) if (flags[i]) out->pass_idx.push_back(i);
    } else if (out) {
        DevMem d_idx(db.core, 4 * n + 4);
        uint32_t* d_cnt = d_idx.as<uint32_t>() + n;
        select_passing((uint32_t)n, d_pass.as<uint8_t>(), d_idx.as<uint32_t>(), d_cnt, st);
        uint32_t cnt = 0;
        download(c, &cnt, d_cnt, 1);
        CU(cudaStreamSynchronize(st));
        ts.mark("survivor count on the host");
        out->pass_idx.resize(cnt);
        download(c, out->pass_idx.data(), d_idx.as<uint32_t>(), cnt);
    }
    CU(cudaStreamSynchronize(st));
    ts.mark("done");
}

}  // namespace

extern "C" {

int skb_db_screen(skb_db_t* db, uint32_t n_queries, skb_sketch_t* const* queries, double cutoff, int32_t rescue_small,
                  uint8_t* pass, uint32_t* shared) {
    if (!db || (n_queries && !queries)) return SKB_ERR_ARG;
    return guarded(db->core.get(), [&] {
        std::vector<std::shared_ptr<SketchImpl>> qs;
        std::vector<GenomeView> hv;
        for (uint32_t i = 0; i < n_queries; i++) {
            if (!queries[i]) throw Fail{SKB_ERR_ARG, "null query sketch"};
            if (queries[i]->impl->core->device != db->core->device) throw Fail{SKB_ERR_ARG, "query sketch belongs to a context of another device"};
            qs.push_back(queries[i]->impl); hv.push_back(queries[i]->impl->view);
        }
        DevMem d_q(db->core, sizeof(GenomeView) * std::max<uint32_t>(n_queries, 1));
        CU(cudaMemcpyAsync(d_q.p, hv.data(), sizeof(GenomeView) * n_queries, cudaMemcpyHostToDevice, db->core->stream));
        run_screen(*db, qs, d_q.as<GenomeView>(), cutoff, rescue_small, nullptr, pass, shared);
        return SKB_OK;
    });
}

int skb_db_query(skb_db_t* db, uint32_t n_queries, skb_sketch_t* const* queries, const skb_query_opts_t* opts,
                 skb_hit_t** hits, uint64_t* n_hits, uint64_t* n_screened_in) {
    if (!db || !opts || !hits || !n_hits || (n_queries && !queries)) return SKB_ERR_ARG;
    *hits = nullptr; *n_hits = 0;
    if (n_screened_in) *n_screened_in = 0;
    return guarded(db->core.get(), [&] {
        Core& c = *db->core;
        cudaStream_t st = c.stream;
        if (opts->learned_ani == 1 && !db->model)
            throw Fail{SKB_ERR_UNSUPPORTED,
                       "learned_ani=True needs skani's GBDT model, whose weights are not part of pyskani's sources: attach "
                       "one with skb_db_set_model (Database.set_model / PYSKANI_B200_MODEL), or pass learned_ani=False"};
        const uint32_t nr = (uint32_t)db->items.size();
        if (nr == 0 || n_queries == 0) return SKB_OK;
        const skb_sketch_info_t& dbi = db->items[0]->info;
        std::vector<std::shared_ptr<SketchImpl>> qs;
        std::vector<GenomeView> hv;
        for (uint32_t i = 0; i < n_queries; i++) {
            if (!queries[i]) throw Fail{SKB_ERR_ARG, "null query sketch"};
            if (queries[i]->impl->core->device != db->core->device) throw Fail{SKB_ERR_ARG, "query sketch belongs to a context of another device"};
            const auto& qi = queries[i]->impl->info;
            if (qi.k != dbi.k || qi.c != dbi.c || qi.marker_c != dbi.marker_c) throw Fail{SKB_ERR_ARG, "query sketch parameters differ from the database's"};
            if (queries[i]->impl->ref_only) throw Fail{SKB_ERR_ARG, "this sketch was transferred as a reference only (no position-order seeds): it cannot be a query"};
            qs.push_back(queries[i]->impl); hv.push_back(queries[i]->impl->view);
        }
        Trace tq("db_query");
        CU(cudaEventRecord(c.ev[0], st));
        DevMem d_q(db->core, sizeof(GenomeView) * n_queries);
        CU(cudaMemcpyAsync(d_q.p, hv.data(), sizeof(GenomeView) * n_queries, cudaMemcpyHostToDevice, st));
        const GenomeView* d_r = db_views(*db);

        // ---- screen (reference lib.rs:603-637)
        const double screen_val = opts->cutoff != 0.0 ? opts->cutoff : 0.80;   // SEARCH_ANI_CUTOFF_DEFAULT
        ScreenOut so;
        run_screen(*db, qs, d_q.as<GenomeView>(), screen_val, opts->faster_small ? 0 : 1, &so, nullptr, nullptr);
        CU(cudaEventRecord(c.ev[1], st));
        tq.mark("screen done");
        if (n_screened_in) *n_screened_in = so.pass_idx.size();

        // ---- chain survivors in batches (reference lib.rs:640-657)
        ChainConsts C{};
        C.fragment_length = FRAGMENT_LENGTH; C.anchor_score = DP_ANCHOR_SCORE; C.min_anchors = 3; C.min_score = 45; C.max_gap = DP_MAX_GAP;
        C.index_band = (int32_t)DP_INDEX_BAND; C.bp_band = (int32_t)DP_BP_BAND; C.af_ext = 198; C.frac_cover_cutoff = 0.15;   // D_FRAC_COVER_CUTOFF / 100 (lib.rs:589)
        C.robust = opts->robust; C.median = opts->median; C.k = dbi.k;
        // lib.rs:611-614: learned = learned_ani.unwrap_or_else(|| use_learned_ani(c, false, false, median)); the model is
        // only consulted for the default (mean) estimate
        const bool learned = opts->learned_ani == 1 || (opts->learned_ani < 0 && dbi.c >= 70 && !opts->median);
        C.use_model = learned && db->model && !opts->robust && !opts->median ? 1 : 0;
        C.learned_min_cov = 150000.0;
        if (C.use_model) C.model = db->model->dev_view();

        std::vector<skb_hit_t> all_hits;
        const size_t n_pass = so.pass_idx.size();
        if (n_pass) {
            // chaining reads seed arrays: sketches adopted from an exchange block may still be waiting for theirs
            for (const auto& it : db->items) ensure_seeds_ready(*it);
            for (const auto& q : qs) ensure_seeds_ready(*q);
        }
        // Query seeds per chaining batch: large batches amortise the ~25 launches, the kernel tails and the host round trip
        // of a batch (10^6-pair all-vs-all, 354 M query seeds: 10.9 ms in three batches of <= 160 M, 10.4 ms in one).  A batch
        // needs ~12 B per query seed + ~54 B per seed for the anchor arrays (1.5 anchors per seed reserved): up to 384 M seeds
        // = 25 GB, bounded by 15 % of the device's memory (read once, when the context is created: cudaMemGetInfo costs up to
        // 15 ms per call on some boxes) and halved for good whenever the arena cannot be allocated.
        uint64_t MAX_BATCH_SEEDS = c.chain_batch_seeds;
        if (const char* e = std::getenv("SKB_CHAIN_BATCH_MSEEDS")) MAX_BATCH_SEEDS = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10)) << 20;   // tuning / test hook
        size_t p0 = 0;
        while (p0 < n_pass) {
            std::vector<PairDesc> pairs;
            uint64_t seeds = 0, wins = 0, bit_words = 0;
            uint32_t max_qseeds = 0;
            size_t p1 = p0;
            while (p1 < n_pass) {
                const uint32_t q = so.pass_idx[p1] / nr, r = so.pass_idx[p1] % nr;
                const GenomeView& qv = qs[q]->view;
                if (!pairs.empty() && (seeds + qv.n_seeds > MAX_BATCH_SEEDS || pairs.size() >= 65535)) break;
                pairs.push_back(PairDesc{q, r, (uint32_t)seeds, (uint32_t)wins, (uint32_t)bit_words});
                bit_words += ((qv.n_seeds + 31) / 32 + 3) / 4 * 4;     // 16-byte aligned slices (bulk-copy source)
                seeds += qv.n_seeds; wins += qv.win_cap;
                max_qseeds = std::max(max_qseeds, qv.n_seeds);
                p1++;
            }
            if (seeds >= 0x7FFFFFFFull) throw Fail{SKB_ERR_ARG, "query sketch too large for one chaining batch"};
            const uint32_t np = (uint32_t)pairs.size();
            // The number of anchors is only known on the device.  Size the anchor arrays from an estimate (1.5 x the
            // smaller seed count of every pair, times 1.5), run the whole batch, and read the true total back together with the
            // results: one synchronisation per batch; a batch whose estimate was too small is simply run again.
            uint64_t est = 1024;
            for (const PairDesc& pd : pairs) { const uint64_t m = std::min(qs[pd.q]->view.n_seeds, db->items[pd.r]->view.n_seeds); est += m + m / 2; }
            if (est > 0x7FFFFFFFull) est = 0x7FFFFFFFull;
            if (const char* e = std::getenv("SKB_FORCE_ANCHOR_EST")) est = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10));   // test hook: exercises the rerun
            std::vector<PairResult> res(np);
            // walk groups: runs of pairs with the same query, cut so that the groups about fill the SMs once
            std::vector<uint2> groups;
            uint32_t group_max = std::min<uint32_t>(walk_group_capacity(max_qseeds), std::max<uint32_t>(1, (np + c.n_sm - 1) / c.n_sm));
            if (const char* e = std::getenv("SKB_WALK_GROUP"))    // test hook: group size independent of the batch size
                group_max = std::min<uint32_t>(walk_group_capacity(max_qseeds), std::max<uint32_t>(1, (uint32_t)std::atoi(e)));
            for (uint32_t i = 0; i < np;) {
                uint32_t j = i + 1;
                while (j < np && j - i < group_max && pairs[j].q == pairs[i].q) j++;
                groups.push_back(make_uint2(i, j - i));
                i = j;
            }
            const size_t nw = std::max<uint64_t>(wins, 1);
            const size_t scan_bytes = scan_scratch_bytes((uint32_t)seeds + 1), sort_bytes = sort_pairs_scratch_bytes((uint32_t)wins);
            bool retry_smaller = false;
            for (int attempt = 0; attempt < 2; attempt++) {
                const size_t na = (size_t)est;
                // every transient array of the batch is carved out of ONE grow-only block of the context: a cold
                // stream-ordered pool needed 20-150 ms to grow through the fifteen allocations this used to make
                size_t total = 0;
                auto plan = [&](size_t bytes) { const size_t o = total; total += (bytes + 255) & ~(size_t)255; return o; };
                const size_t o_pairs = plan(sizeof(PairDesc) * np), o_fc = plan(8 * (seeds + 1)),
                             o_aoff = plan(4 * (seeds + 2)), o_bits = plan(4 * (bit_words + 4)), o_scan = plan(scan_bytes),
                             o_a = plan(na * 16), o_fra = plan(na * 4 * 3), o_best = plan(na * 8), o_big = plan(4 * (nw + 4)), o_bins = plan(4 * 192), o_order = plan(4 * nw),
                             o_w = plan(nw * 4 * 3 + 4 * (size_t)np),
                             o_rec = plan(nw * sizeof(WindowRec)), o_keys = plan(nw * 8 * 2), o_vals = plan(nw * 4 * 2),
                             o_res = plan(sizeof(PairResult) * np), o_sort = plan(sort_bytes),
                             o_groups = plan(sizeof(uint2) * groups.size()), o_total = plan(8);
                char* base = nullptr;
                try {
                    base = (char*)c.scratch(SLOT_CHAIN, total);
                } catch (const Fail& f) {
                    // no room for a batch of this size (other tenants of the device, a huge database): halve the batch size
                    // of this context for good and plan the remaining pairs again
                    if (f.code != SKB_ERR_NOMEM || np <= 1) throw;
                    cudaGetLastError();
                    c.chain_batch_seeds = std::max<uint64_t>(1ull << 20, std::min(c.chain_batch_seeds, seeds) / 2);
                    MAX_BATCH_SEEDS = c.chain_batch_seeds;
                    retry_smaller = true;
                    break;
                }
                ChainBatch B{};
                B.qviews = d_q.as<GenomeView>(); B.rviews = d_r; B.n_pairs = np;
                B.n_qseeds_total = (uint32_t)seeds; B.n_win_total = (uint32_t)wins;
                B.pairs = (PairDesc*)(base + o_pairs);
                CU(cudaMemcpyAsync(base + o_pairs, pairs.data(), sizeof(PairDesc) * np, cudaMemcpyHostToDevice, st));
                B.m_fc = (uint2*)(base + o_fc); B.a_off = (uint32_t*)(base + o_aoff);
                B.m_bits = (uint32_t*)(base + o_bits);
                B.walk_groups = (const uint2*)(base + o_groups); B.n_walk_groups = (uint32_t)groups.size(); B.walk_group_max = group_max;
                CU(cudaMemcpyAsync(base + o_groups, groups.data(), sizeof(uint2) * groups.size(), cudaMemcpyHostToDevice, st));
                CU(cudaMemsetAsync(B.m_fc, 0, 8 * (seeds + 1), st));      // only matched seeds are written
                B.a_total64 = (unsigned long long*)(base + o_total);
                CU(cudaMemsetAsync(B.a_total64, 0, 8, st));
                launch_match_count(B, st);
                scan_match_counts(B, base + o_scan, scan_bytes, st);
                B.anchor_cap = (uint32_t)est;
                B.a_rec = (uint4*)(base + o_a);
                B.a_f = (int32_t*)(base + o_fra); B.a_root = (uint32_t*)(B.a_f + na); B.a_aux = B.a_root + na;
                B.a_best = (unsigned long long*)(base + o_best);
                B.big_count = (uint32_t*)(base + o_big); B.big_list = B.big_count + 4;
                CU(cudaMemsetAsync(B.big_count, 0, 16, st));
                B.win_bins = (uint32_t*)(base + o_bins); B.win_order = (uint32_t*)(base + o_order);
                CU(cudaMemsetAsync(B.win_bins, 0, 4 * 192, st));
                B.win_start = (uint32_t*)(base + o_w); B.win_end = B.win_start + nw; B.win_contig = B.win_end + nw; B.pair_nwin = B.win_contig + nw;
                B.win_rec = (WindowRec*)(base + o_rec);
                CU(cudaMemsetAsync(B.win_start, 0, 8 * nw, st));   // start == end == 0 marks an unused slot
                B.sort_keys = (uint64_t*)(base + o_keys); B.sort_vals = (uint32_t*)(base + o_vals);
                uint64_t* keys_sorted = B.sort_keys + nw; uint32_t* vals_sorted = B.sort_vals + nw;
                B.results = (PairResult*)(base + o_res);
                tq.mark("batch allocated");
                launch_anchor_fill(B, st);
                launch_window_walk(B, C, max_qseeds, st);
                launch_chain_dp(B, C, c.n_sm, st);
                if (C.robust || C.median) {
                    // only the trimmed mean and the median need the windows ordered by their anchors / seeds ratio
                    launch_window_keys(B, st);
                    int pbits = 1;
                    while ((1ull << pbits) <= (uint64_t)np) pbits++;
                    sort_window_keys((uint32_t)wins, B.sort_keys, keys_sorted, B.sort_vals, vals_sorted, 32 + pbits, base + o_sort,
                                     sort_bytes, st);
                    launch_ani_reduce(B, C, keys_sorted, vals_sorted, st);
                } else {
                    launch_ani_reduce(B, C, nullptr, nullptr, st);
                }
                CU(cudaGetLastError());          // a launch that was refused (configuration) must not pass silently
                uint32_t n_anchors = 0;
                unsigned long long n_anchors64 = 0;
                download(c, &n_anchors, B.a_off + seeds, 1);
                download(c, &n_anchors64, B.a_total64, 1);
                download(c, res.data(), B.results, np);
                tq.mark("batch enqueued");
                CU(cudaStreamSynchronize(st));   // also keeps `pairs` alive until its upload has finished
                tq.mark("batch synced");
                // the anchor offsets are a 32-bit scan: repeat-rich genome pairs (multi-copy k-mers on both sides multiply)
                // could wrap it, after which the capacity check below would pass on garbage
                if (n_anchors64 >= 0x7FFFFFFFull)
                    throw Fail{SKB_ERR_ARG, "more than 2^31 k-mer matches in one chaining batch (highly repetitive sketches); query fewer genomes per call"};
                if (n_anchors <= B.anchor_cap) break;
                if (attempt == 1) throw Fail{SKB_ERR_CUDA, "anchor arrays overflowed twice"};
                est = n_anchors;
            }
            if (retry_smaller) continue;       // plan again from p0 with the smaller batch size
            for (uint32_t i = 0; i < np; i++) {
                if (res[i].ani > 0.1f) {   // reference lib.rs:654
                    skb_hit_t h{};
                    h.query_index = pairs[i].q; h.ref_index = pairs[i].r;
                    h.ani = res[i].ani; h.af_query = res[i].af_q; h.af_ref = res[i].af_r;
                    h.n_windows = res[i].n_windows; h.n_chains = res[i].n_chains; h.n_anchors = res[i].n_anchors;
                    all_hits.push_back(h);
                }
            }
            p0 = p1;
        }
        CU(cudaEventRecord(c.ev[2], st));
        CU(cudaStreamSynchronize(st));
        c.stats.screen_ms = elapsed(c.ev[0], c.ev[1]); c.stats.chain_ms = elapsed(c.ev[1], c.ev[2]);
        c.stats.total_ms = elapsed(c.ev[0], c.ev[2]);
        if (!all_hits.empty()) {
            *hits = new skb_hit_t[all_hits.size()];
            std::memcpy(*hits, all_hits.data(), sizeof(skb_hit_t) * all_hits.size());
        }
        *n_hits = all_hits.size();
        return SKB_OK;
    });
}

}  // extern "C"
