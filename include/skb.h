/* skb.h — C ABI of libskb.so: the B200-native sketch -> screen -> chain -> ANI path behind pyskani.
 *
 * Every entry point replaces one Rust-to-Rust call that the reference (althonos/pyskani,
 * src/pyskani/_skani/lib.rs) makes into the crate `skani`; the call site is cited on each
 * declaration.  The ABI is plain C: opaque handles, pointers and sizes, integer status codes.
 * There is NO CPU fallback behind it: every function that computes launches sm_100a kernels and
 * fails with SKB_ERR_CUDA if no device is usable.
 *
 * Ownership: input byte ranges are borrowed for the duration of the call only (the reference
 * borrows them the same way, utils.rs:81-101); handles returned through an out-parameter are owned
 * by the caller and released with the matching *_free / *_destroy; hit arrays are library-allocated
 * and released with skb_hits_free.
 * Threading: a context serialises its own calls with an internal mutex (the reference guards its
 * state with RwLocks, lib.rs:135-136).  One context per GPU is the normal arrangement.  Two contexts on the SAME device run
 * concurrently (own stream, own scratch): a database of one may hold, and be queried with, finished sketches of the other,
 * which is how an all-vs-all hides its queries under the host->device ingest of the next sketch batch
 * (pyskani_b200/parallel.py: all_vs_all_pipelined).
 */
#ifndef SKB_H
#define SKB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKB_OK            0
#define SKB_ERR_ARG       1   /* -> ValueError   */
#define SKB_ERR_CUDA      2   /* -> RuntimeError */
#define SKB_ERR_NOMEM     3   /* -> MemoryError  */
#define SKB_ERR_KEY       4   /* -> KeyError     */
#define SKB_ERR_UNSUPPORTED 5 /* -> NotImplementedError / RuntimeError */

#define SKB_MIN_LENGTH_CONTIG 500   /* skani::params::MIN_LENGTH_CONTIG, gate at lib.rs:156 */
#define SKB_MARKER_K 21             /* skani::params::K_MARKER_DNA */

typedef struct skb_ctx    skb_ctx_t;     /* one GPU: device id, streams, scratch arenas */
typedef struct skb_sketch skb_sketch_t;  /* one sketched genome, resident in HBM (skani::types::Sketch) */
typedef struct skb_db     skb_db_t;      /* ordered set of sketches queried together (Database.markers + Database.sketches) */

/* SketchParams::new(marker_c, c, k, false, false) — lib.rs:416 */
typedef struct {
    int32_t k;          /* seed k-mer length, <= 16 (skani panics above 16)      */
    int32_t c;          /* FracMinHash compression of seeds   (default 125)      */
    int32_t marker_c;   /* FracMinHash compression of markers (default 1000)     */
} skb_sketch_params_t;

/* The part of CommandParams that pyskani lets the caller set — lib.rs:573-614 */
typedef struct {
    double  cutoff;        /* screen_val; 0.0 means "use SEARCH_ANI_CUTOFF_DEFAULT = 0.80" (lib.rs:603-609) */
    int32_t learned_ani;   /* -1 = None, 0 = False, 1 = True (lib.rs:593,611-614)                          */
    int32_t median;        /* lib.rs:582 */
    int32_t robust;        /* lib.rs:581 */
    int32_t faster_small;  /* rescue_small = !faster_small (lib.rs:597) */
} skb_query_opts_t;

/* AniEstResult fields that Hit exposes — hit.rs:77-104 */
typedef struct {
    uint32_t query_index;  /* index of the query in the call's query array  */
    uint32_t ref_index;    /* index of the reference inside the database    */
    float    ani;          /* Hit.identity            */
    float    af_query;     /* Hit.query_fraction      */
    float    af_ref;       /* Hit.reference_fraction  */
    uint32_t n_windows;    /* diagnostics: 20 kb query windows that contributed */
    uint32_t n_chains;     /* diagnostics: kept chains                          */
    uint32_t n_anchors;    /* diagnostics: k-mer matches joined                 */
} skb_hit_t;

typedef struct {
    uint64_t n_seeds, n_markers, total_len;
    uint32_t n_contigs;    /* contigs that passed the MIN_LENGTH_CONTIG gate */
    int32_t  k, c, marker_c;
    int32_t  has_seeds;    /* sketched with seed=True */
    int32_t  reference_only; /* received through an exchange block without the query-side arrays */
} skb_sketch_info_t;

/* Timings of the last call on this context, measured with CUDA events on the context's stream. */
typedef struct {
    float h2d_ms, seed_ms, index_ms, screen_ms, chain_ms, total_ms;
    uint64_t kernels_launched;   /* cumulative count of this library's kernel launches */
    /* skb_sketch_batch: bytes that crossed the host->device link as ASCII and as 2-bit words (4 bases per byte; large
     * batches are partly compacted by host threads before the copy, see skb_ctx_set_host_threads) */
    uint64_t h2d_raw_bytes, h2d_packed_bytes;
} skb_stats_t;

/* ---- context ---- */
int  skb_ctx_create(int device, skb_ctx_t** out);
void skb_ctx_destroy(skb_ctx_t* ctx);
const char* skb_last_error(const skb_ctx_t* ctx);     /* message of the last failing call */
int  skb_ctx_stats(const skb_ctx_t* ctx, skb_stats_t* out);
int  skb_ctx_sync(skb_ctx_t* ctx);
/* CUDA stream all work of this context is enqueued on (a cudaStream_t), for event timing by callers */
void* skb_ctx_stream(skb_ctx_t* ctx);
/* Host threads of the ingest pipeline of skb_sketch_batch (host buffers).  Batches of 64 MB and more travel two ways at
 * once: the copy engine pulls chunks of ASCII out of the caller's memory while n - 1 threads compact other chunks to
 * 2-bit words (4:1) in pinned staging memory and one thread feeds the copy engine; the seeding kernel reads either form
 * and produces bit-identical sketches.  n = 0 or 1: every byte travels as ASCII; n < 0: default = SKB_HOST_THREADS, else
 * min(32, usable CPUs / LOCAL_WORLD_SIZE) - 1, and 0 from four ranks per host on or below 8 CPUs per rank; which of the two
 * policies (both routes, or every chunk compacted) a context uses for pinned sources is measured on its first large calls.  SKB_INGEST=raw|pack|mix
 * overrides the policy (test hook). */
int  skb_ctx_set_host_threads(skb_ctx_t* ctx, int32_t n);
/* Two contexts sharing one device (a pipelined all-vs-all sketches on one and queries on the other, each fed by its own
 * host thread): high != 0 gives the context's compute streams the HIGHEST scheduling priority of the device, so its kernels
 * do not queue behind the other context's (0: back to the default, the lowest).  Call it between, not during, other calls;
 * skb_ctx_stream() changes.  (Sleeping instead of spinning waits were measured for either side and cost more than the CPU
 * they free: DESIGN.md section 7.) */
int  skb_ctx_set_priority(skb_ctx_t* ctx, int32_t high);

/* pinned host staging memory (so that callers can keep inputs where an async H2D copy can reach them) */
int  skb_host_alloc(skb_ctx_t* ctx, size_t bytes, void** out);
void skb_host_free(skb_ctx_t* ctx, void* p);
/* device memory for device-resident inputs of skb_sketch_batch_device */
int  skb_dev_alloc(skb_ctx_t* ctx, size_t bytes, void** out);
void skb_dev_free(skb_ctx_t* ctx, void* p);
int  skb_memcpy_h2d(skb_ctx_t* ctx, void* dst, const void* src, size_t bytes);

/* ---- sketching: Database::_sketch (lib.rs:140-185) = Sketch::new + per-contig fmh_seeds (lib.rs:165-171) ----
 * n_genomes genomes are sketched in one pass.  Genome g owns contigs
 * [genome_contig_start[g], genome_contig_start[g+1]) of the flat arrays contigs[] / contig_lens[].
 * Contigs shorter than SKB_MIN_LENGTH_CONTIG are skipped and do not consume a contig index.
 * out[g] receives a new sketch handle.  Host version: contigs[i] are host pointers (pinned or pageable),
 * the H2D copy happens inside the call. */
int skb_sketch_batch(skb_ctx_t* ctx, const skb_sketch_params_t* params, int32_t seed,
                     uint32_t n_genomes, const uint32_t* genome_contig_start,
                     const uint8_t* const* contigs, const uint64_t* contig_lens,
                     skb_sketch_t** out);
/* Device version: all contigs already sit in one device buffer `seq`; contig i occupies
 * seq[contig_offsets[i] .. contig_offsets[i] + contig_lens[i]).  Offsets must be multiples of 16 and the
 * buffer must extend 64 readable bytes before the first and after the last contig. */
int skb_sketch_batch_device(skb_ctx_t* ctx, const skb_sketch_params_t* params, int32_t seed,
                            uint32_t n_genomes, const uint32_t* genome_contig_start,
                            const uint8_t* seq, const uint64_t* contig_offsets, const uint64_t* contig_lens,
                            skb_sketch_t** out);
void skb_sketch_free(skb_sketch_t* s);
void skb_sketch_free_many(uint32_t n, skb_sketch_t* const* s);   /* the same for n handles with one lock acquisition */
int  skb_sketch_info(const skb_sketch_t* s, skb_sketch_info_t* out);
/* Copy a sketch to host arrays (any pointer may be NULL).  Seeds come sorted by (kmer, contig, pos);
 * canonical[i] = SeedPosition.canonical; markers sorted ascending and unique (the reference keeps both in
 * hash containers whose order is arbitrary).  Used by save() and by the parity tests. */
int  skb_sketch_export(const skb_sketch_t* s, uint64_t* kmer, uint32_t* pos, uint32_t* contig,
                       uint8_t* canonical, uint64_t* markers, uint32_t* contig_lengths);
/* Rebuild a device sketch from host arrays (Database.load / Database.open path, lib.rs:93-122).
 * Seeds may come in any order. */
int  skb_sketch_import(skb_ctx_t* ctx, const skb_sketch_params_t* params, int32_t has_seeds,
                       uint64_t n_seeds, const uint64_t* kmer, const uint32_t* pos, const uint32_t* contig,
                       const uint8_t* canonical, uint64_t n_markers, const uint64_t* markers,
                       uint32_t n_contigs, const uint32_t* contig_lengths, skb_sketch_t** out);

/* ---- device-to-device transfer of sketches (the one exchange step of a multi-GPU all-vs-all: SURVEY.md section 8e;
 * the reference has no counterpart, its Database lives in one process, lib.rs:132-137) ----
 * pack: the device arrays of n sketches are concatenated into a caller-provided DEVICE buffer (which a collective can
 * then send over NVLink) plus a small HOST descriptor; unpack: rebuilds the sketches on another context from such a pair
 * with one device-to-device copy — no host staging and no re-sorting, unlike skb_sketch_export / skb_sketch_import.
 * The payload is position independent. */
int  skb_sketch_pack_size(uint32_t n, skb_sketch_t* const* sketches, uint64_t* payload_bytes, uint64_t* meta_bytes);
int  skb_sketch_pack(skb_ctx_t* ctx, uint32_t n, skb_sketch_t* const* sketches, void* payload_dev, uint64_t payload_bytes,
                     void* meta_host, uint64_t meta_bytes);
/* out receives *n_out handles (the count is part of the descriptor); out_cap is the capacity of out[]. */
int  skb_sketch_unpack(skb_ctx_t* ctx, const void* meta_host, uint64_t meta_bytes, const void* payload_dev,
                       uint64_t payload_bytes, skb_sketch_t** out, uint32_t out_cap, uint32_t* n_out);

/* ---- zero-copy exchange region (multi-GPU all-vs-all) ----
 * One block of sketch storage that holds, per rank, a HEAD segment [descriptor | marker sets | per-contig tables] and a
 * BODY segment [seed arrays | bucket tables].  Every rank packs its own two segments in place, the caller's collectives
 * (all-gather with per-rank sizes, straight into this block) fill the others, and the peers' sketches are then adopted
 * as views into the block: no staging buffer, no padding, no second device-to-device copy.  Heads and bodies travel as
 * two collectives so that the marker screen can start while the (25x larger) bodies are still on the wire.
 * reference_only != 0 leaves out what only a QUERY needs (position-order seeds): half of the bytes; such sketches can be
 * database members but not queries.  The block lives until the last adopted sketch is freed. */
typedef struct skb_exchange skb_exchange_t;
/* bytes of the two segments these sketches need (multiples of 256) and of the descriptor at the front of the head */
int  skb_exchange_segment_size(uint32_t n, skb_sketch_t* const* sketches, int32_t reference_only, uint64_t* head_bytes,
                               uint64_t* body_bytes, uint64_t* meta_bytes);
int  skb_exchange_create(skb_ctx_t* ctx, uint64_t bytes, skb_exchange_t** out);
void* skb_exchange_ptr(skb_exchange_t* ex);                    /* device address of the block */
/* writes the two segments of these sketches at the given offsets (multiples of 256); asynchronous on the context's stream */
int  skb_exchange_pack(skb_exchange_t* ex, uint64_t head_offset, uint64_t body_offset, uint32_t n,
                       skb_sketch_t* const* sketches, int32_t reference_only);
/* Orders this context behind work the caller enqueued on ANOTHER CUDA stream (the collective's): bodies == 0: the heads;
 * the context's stream waits now (call before skb_exchange_adopt).  bodies != 0: the bodies; the wait is deferred until
 * seed arrays of an adopted sketch are first read (chaining, export, re-packing), so screening overlaps the transfer. */
int  skb_exchange_order_after(skb_exchange_t* ex, void* foreign_cuda_stream, int32_t bodies);
/* adopts the sketches of n_segments peers (segment offsets as packed there, descriptor sizes meta_bytes[i]) once the
 * heads have arrived.  Handles are appended to out[] peer after peer; counts[i] receives the number of sketches of peer i. */
int  skb_exchange_adopt(skb_exchange_t* ex, uint32_t n_segments, const uint64_t* head_offsets, const uint64_t* body_offsets,
                        const uint64_t* meta_bytes, skb_sketch_t** out, uint32_t out_cap, uint32_t* counts);
void skb_exchange_free(skb_exchange_t* ex);                    /* drops the caller's reference to the block */

/* ---- database: the (markers, sketches) pair a Database owns (lib.rs:132-137) ---- */
int  skb_db_create(skb_ctx_t* ctx, skb_db_t** out);
void skb_db_destroy(skb_db_t* db);
/* Database.sketch's two pushes (lib.rs:501-508).  The database shares ownership of the sketch. */
int  skb_db_add(skb_db_t* db, skb_sketch_t* s, uint32_t* index_out);
/* the same for n sketches in one call; index_out (may be NULL) receives the index of the first one */
int  skb_db_add_many(skb_db_t* db, uint32_t n, skb_sketch_t* const* sketches, uint32_t* index_out);
/* Puts `s` in the place of the sketch at `index`: what a second Database.sketch() under an existing name amounts to
 * in the reference, whose sketch store is a HashMap keyed by name (insert replaces, lib.rs:45,64) */
int  skb_db_replace(skb_db_t* db, uint32_t index, skb_sketch_t* s);
uint64_t skb_db_size(const skb_db_t* db);

/* ---- learned-ANI model: skani::regression::get_model / use_learned_ani (lib.rs:611-614) ----
 * skani embeds gradient-boosted regression ensembles (crate gbdt 0.1.3) as serde_json text; their weights are not part
 * of pyskani's sources, so the model is loaded from text in the same format that the caller supplies.  A database with
 * a model corrects the default (mean) ANI estimate on the device exactly when the reference would:
 *   learned_ani == 1, or learned_ani == -1 (None) with c >= 70 and median == 0          (lib.rs:611-613);
 * robust / median estimates are returned uncorrected.  learned_ani == 1 without a model is SKB_ERR_UNSUPPORTED. */
typedef struct skb_model skb_model_t;
int  skb_model_load_json(skb_ctx_t* ctx, const char* json, size_t len, skb_model_t** out);
void skb_model_free(skb_model_t* m);
int  skb_model_info(const skb_model_t* m, uint32_t* n_trees, uint32_t* n_nodes, uint32_t* n_features);
/* evaluates the ensemble on the device for n_rows feature rows of n_features f32 each (parity / diagnostics entry) */
int  skb_model_predict(skb_model_t* m, const float* rows, uint32_t n_rows, uint32_t n_features, float* out);
/* attaches a model to the database (shared ownership); NULL detaches */
int  skb_db_set_model(skb_db_t* db, skb_model_t* m);

/* ---- query: lib.rs:616-657 for n_queries queries at once ----
 * For every (query, reference) pair: check_markers_quickly (lib.rs:623-628); for survivors
 * map_params_from_sketch + chain_seeds (lib.rs:646-653); keep results with ani > 0.1 (lib.rs:654).
 * Hits come back grouped by query, references in database order (the reference iterates a HashSet,
 * lib.rs:640, so its order is arbitrary).  n_screened_in (may be NULL) = pairs that passed the screen. */
int  skb_db_query(skb_db_t* db, uint32_t n_queries, skb_sketch_t* const* queries,
                  const skb_query_opts_t* opts, skb_hit_t** hits, uint64_t* n_hits, uint64_t* n_screened_in);
void skb_hits_free(skb_hit_t* hits);

/* check_markers_quickly alone (lib.rs:623-628) for every (query, reference) pair:
 * pass[q * db_size + r] = 1/0, shared[...] = marker intersection size (either may be NULL). */
int  skb_db_screen(skb_db_t* db, uint32_t n_queries, skb_sketch_t* const* queries,
                   double cutoff, int32_t rescue_small, uint8_t* pass, uint32_t* shared);

/* Library build info */
const char* skb_version(void);

#ifdef __cplusplus
}
#endif
#endif
