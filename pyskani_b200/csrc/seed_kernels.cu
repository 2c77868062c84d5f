// seed_kernels.cu — FracMinHash seeding on sm_100a.
//
// Replaces skani::seeding::fmh_seeds (one call per contig at reference lib.rs:165-171) for a whole batch
// of genomes in ONE launch.  Work unit: a tile of TILE_BASES (16 384) consecutive bases of one contig.
//   * each thread loads 4 x 16 ASCII bytes with 128-bit read-only loads (warp-contiguous 512 B each) and packs
//     them to 2-bit words in shared memory
//   * k-mers are cut out of three consecutive words with funnel shifts (no rolling dependency chain)
//   * both hashes (k-mer and 21-mer marker) are evaluated for all 16 positions of a word in 32-bit halves:
//     multiplies go to the FMA pipe (IMAD.WIDE / IMAD), xor-shifts to the ALU pipe, so both pipes issue
//   * popcounts are scanned over the CTA, and the CTA obtains its global output offset with a
//     single-pass decoupled look-back over tile status words, so seeds leave the kernel already ordered
//     by (genome, contig, position) — no atomics on the data path, no second pass over the sequence
//   * hit positions are re-extracted and written to their exact slot.
// The kernel is persistent: CTAs draw tile ids from an atomic counter, which also gives the look-back
// its forward-progress guarantee (a CTA only ever waits on tiles that were claimed before its own).
#include "kmer_bits.cuh"
#include "skb_internal.cuh"

namespace skb {

unsigned long long g_kernel_launches = 0;

namespace {

constexpr uint64_t ST_AGG = 1ull << 62;     // tile aggregate published
constexpr uint64_t ST_INC = 2ull << 62;     // inclusive prefix published
constexpr uint64_t ST_MASK = 3ull << 62;
constexpr uint64_t CNT_MASK = 0x7FFFFFFFull;   // seeds in bits 0..30, markers in bits 31..61

__device__ __forceinline__ uint64_t ld_relaxed(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- mm_hash64 in 32-bit halves (see kmer_bits.cuh::mm_hash64 for the reference form) -------------------
struct U64 { uint32_t lo, hi; };

__device__ __forceinline__ U64 mul_c(U64 x, uint32_t c) {            // x * c mod 2^64: IMAD.WIDE + IMAD (FMA pipe)
    const uint64_t p = (uint64_t)x.lo * c;
    U64 r;
    r.lo = (uint32_t)p;
    r.hi = x.hi * c + (uint32_t)(p >> 32);
    return r;
}
#ifndef SKB_XS_FMA
#define SKB_XS_FMA 0
#endif
// x >> S for 0 < S < 32.  With SKB_XS_FMA the shifts are expressed as multiplies (IMAD.HI / IMAD) so that they
// issue on the FMA pipe and leave the ALU pipe to the xors and compares.
template <int S>
__device__ __forceinline__ uint32_t shr_hi(uint32_t hi) {
#if SKB_XS_FMA
    return __umulhi(hi, 1u << (32 - S));
#else
    return hi >> S;
#endif
}
template <int S>
__device__ __forceinline__ uint32_t shr_lo(uint32_t lo, uint32_t hi) {
#if SKB_XS_FMA
    return hi * (1u << (32 - S)) + __umulhi(lo, 1u << (32 - S));     // disjoint bit ranges: + == |
#else
    return __funnelshift_r(lo, hi, S);
#endif
}
template <int S>
__device__ __forceinline__ U64 xorshr(U64 x) {                        // x ^ (x >> S), 0 < S < 32
    U64 r;
    r.lo = x.lo ^ shr_lo<S>(x.lo, x.hi);
    r.hi = x.hi ^ shr_hi<S>(x.hi);
    return r;
}
// hash(x) < thr, x given in halves.  Step 1 is ~(x * (2^21 + 1)); the complement is folded into the first
// xor-shift:  ~a ^ (~a >> 24) == a ^ (a >> 24) ^ 0xFFFFFF00_00000000.
__device__ __forceinline__ bool hash_below(U64 x, uint32_t thr_lo, uint32_t thr_hi) {
    U64 a = mul_c(x, (1u << 21) + 1u);
    U64 b;
    b.lo = a.lo ^ shr_lo<24>(a.lo, a.hi);
    b.hi = a.hi ^ shr_hi<24>(a.hi) ^ 0xFFFFFF00u;
    b = mul_c(b, 265u);
    b = xorshr<14>(b);
    b = mul_c(b, 21u);
    b = xorshr<28>(b);
    b = mul_c(b, 0x80000001u);                                        // x + (x << 31)
    return b.hi < thr_hi || (b.hi == thr_hi && b.lo < thr_lo);
}

struct WordCtx { uint32_t w0, w1, w2, r0, r1, r2; };

// masks of the 16 positions of one word: bit e of *smask / *mmask = position e is a seed / marker
__device__ __forceinline__ void eval_word(const WordCtx& c, uint32_t kmask, uint32_t kshift, uint32_t ts_lo, uint32_t ts_hi,
                                          uint32_t tm_lo, uint32_t tm_hi, uint32_t& smask, uint32_t& mmask) {
    smask = 0; mmask = 0;
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const uint32_t s = 30 - 2 * e;
        const uint32_t flo = __funnelshift_r(c.w0, c.w1, s);
        const uint32_t fhi = __funnelshift_r(c.w1, c.w2, s) & 0x3FFu;
        constexpr int dummy = 0; (void)dummy;
        const uint32_t sr = 24 + 2 * e;
        uint32_t rlo, rhi;
        if (sr < 32) { rlo = __funnelshift_r(c.r2, c.r1, sr); rhi = __funnelshift_r(c.r1, c.r0, sr) & 0x3FFu; }
        else         { rlo = __funnelshift_r(c.r1, c.r0, sr - 32); rhi = (c.r0 >> (sr - 32)) & 0x3FFu; }
        // seed k-mer
        const uint32_t fk = flo & kmask;
        const uint32_t rk = __funnelshift_r(rlo, rhi, kshift);
        const uint32_t km = min(fk, rk);
        if (hash_below(U64{km, 0u}, ts_lo, ts_hi)) smask |= 1u << e;
        // marker 21-mer: canonical = min of the two 42-bit values
        const bool fsmall = fhi < rhi || (fhi == rhi && flo < rlo);
        const U64 mk{fsmall ? flo : rlo, fsmall ? fhi : rhi};
        if (hash_below(mk, tm_lo, tm_hi)) mmask |= 1u << e;
    }
}

__global__ void __launch_bounds__(SEED_THREADS, 3) seed_scan_kernel(const SeedScanArgs a) {
    __shared__ uint32_t s_pk[TILE_WORDS + 2];
    __shared__ uint32_t s_masks[TILE_WORDS];      // smask | mmask << 16 per word
    __shared__ uint64_t s_wseed[SEED_THREADS / 32], s_wmark[SEED_THREADS / 32];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t ts_lo = (uint32_t)a.thr_seed, ts_hi = (uint32_t)(a.thr_seed >> 32);
    const uint32_t tm_lo = (uint32_t)a.thr_marker, tm_hi = (uint32_t)(a.thr_marker >> 32);

    while (true) {
        if (t == 0) s_tile = atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        const uint32_t tile_id = s_tile;                 // local to this launch
        if (tile_id >= a.n_tiles) break;
        const uint32_t gtile = a.tile_base + tile_id;    // id within the batch (descriptors hold batch-wide tile ids)
        // contig of this tile: last descriptor whose tile_start <= gtile (uniform search, L1-resident table)
        uint32_t lo = 0, hi = a.n_contigs;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&a.contigs[mid].tile_start) <= gtile) lo = mid; else hi = mid;
        }
        const ContigDesc cd = a.contigs[lo];
        const uint32_t pos0 = (gtile - cd.tile_start) * (uint32_t)TILE_BASES;   // contig position of the tile's first base
        const uint32_t n = min((uint32_t)TILE_BASES, cd.len - pos0);             // bases in the tile
        const uint8_t* base = a.seq + cd.seq_off + pos0;
        const bool first_of_genome = (cd.genome & 0x80000000u) && pos0 == 0;

        // ---- load + pack: word j*256 + t for j = 0..3 (each warp load covers 512 contiguous bytes)
        uint4 v[WORDS_PER_THREAD];
#pragma unroll
        for (int j = 0; j < WORDS_PER_THREAD; j++) {
            const uint32_t w = j * SEED_THREADS + t;
            v[j] = make_uint4(0, 0, 0, 0);
            if (16u * w < n) v[j] = ld_stream16(base + 16u * w);
        }
#pragma unroll
        for (int j = 0; j < WORDS_PER_THREAD; j++) s_pk[2 + j * SEED_THREADS + t] = pack16(v[j].x, v[j].y, v[j].z, v[j].w);
        if (t < 2) {
            uint32_t h = 0;
            if (pos0 > 0) {   // the two words before the tile belong to the same contig
                uint4 q = ld_stream16(base - 32 + 16 * t);
                h = pack16(q.x, q.y, q.z, q.w);
            }
            s_pk[t] = h;
        }
        __syncthreads();

        // ---- evaluate
        uint64_t cs = 0, cm = 0;                   // per sub-tile counts, 16 bits each
#pragma unroll 1
        for (int j = 0; j < WORDS_PER_THREAD; j++) {
            const uint32_t w = j * SEED_THREADS + t;
            uint32_t mk = 0;
            if (16u * w < n) {
                WordCtx c;
                c.w2 = s_pk[w]; c.w1 = s_pk[w + 1]; c.w0 = s_pk[w + 2];
                c.r0 = revcomp_word(c.w0); c.r1 = revcomp_word(c.w1); c.r2 = revcomp_word(c.w2);
                uint32_t sm, mm;
                eval_word(c, a.kmask, a.kshift, ts_lo, ts_hi, tm_lo, tm_hi, sm, mm);
                const uint32_t left = n - 16u * w;
                uint32_t valid = left >= 16 ? 0xFFFFu : ((1u << left) - 1u);
                const uint32_t p0 = pos0 + 16u * w;
                if (p0 < SKB_MARKER_K - 1) {                  // the first window ends at position 20
                    const uint32_t skip = SKB_MARKER_K - 1 - p0;
                    valid &= skip >= 16 ? 0u : ~((1u << skip) - 1u);
                }
                sm &= valid; mm &= valid;
                mk = sm | (mm << 16);
                cs |= (uint64_t)__popc(sm) << (16 * j);
                cm |= (uint64_t)__popc(mm) << (16 * j);
            }
            s_masks[w] = mk;
        }

        // ---- CTA scan: four sub-tiles at once, 16-bit lanes inside a u64 (a sub-tile holds <= 4096 hits)
        uint64_t is = cs, im = cm;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t ns = __shfl_up_sync(0xffffffffu, is, o), nm = __shfl_up_sync(0xffffffffu, im, o);
            if (lane >= o) { is += ns; im += nm; }
        }
        if (lane == 31) { s_wseed[warp] = is; s_wmark[warp] = im; }
        __syncthreads();

        if (warp == 0) {
            constexpr int NW = SEED_THREADS / 32;
            const uint64_t ws = lane < NW ? s_wseed[lane] : 0ull, wm = lane < NW ? s_wmark[lane] : 0ull;
            uint64_t wis = ws, wim = wm;
#pragma unroll
            for (int o = 1; o < NW; o <<= 1) {
                const uint64_t ns = __shfl_up_sync(0xffffffffu, wis, o), nm = __shfl_up_sync(0xffffffffu, wim, o);
                if (lane >= o) { wis += ns; wim += nm; }
            }
            const uint64_t tot_s = __shfl_sync(0xffffffffu, wis, NW - 1), tot_m = __shfl_sync(0xffffffffu, wim, NW - 1);
            // sub-tile bases: base_j = sum of totals of sub-tiles < j; fold them into the per-warp exclusive offsets
            uint64_t sub_s = 0, sub_m = 0;
            uint32_t acc_s = 0, acc_m = 0;
#pragma unroll
            for (int j = 0; j < WORDS_PER_THREAD; j++) {
                sub_s |= (uint64_t)acc_s << (16 * j); sub_m |= (uint64_t)acc_m << (16 * j);
                acc_s += (uint32_t)(tot_s >> (16 * j)) & 0xFFFFu; acc_m += (uint32_t)(tot_m >> (16 * j)) & 0xFFFFu;
            }
            // acc_* can reach 16384 (< 2^16): every 16-bit lane stays in range
            if (lane < NW) { s_wseed[lane] = (wis - ws) + sub_s; s_wmark[lane] = (wim - wm) + sub_m; }

            // ---- decoupled look-back
            const uint64_t agg = (uint64_t)acc_s | ((uint64_t)acc_m << 31);
            uint64_t excl = 0;
            if (tile_id == 0) {
                // the first tile of a launch continues from the running total of the previous launch of the batch
                // (chunked host->device pipelining); it publishes an inclusive prefix directly, never an aggregate
                excl = a.base_in ? ld_relaxed(a.base_in) : 0ull;
                if (lane == 0) st_relaxed(&a.tile_status[0], ST_INC | (excl + agg));
            } else {
                if (lane == 0) st_relaxed(&a.tile_status[tile_id], ST_AGG | agg);
                int64_t j0 = (int64_t)tile_id - 1;
                while (true) {
                    const int64_t j = j0 - lane;
                    uint64_t s;
                    do {
                        s = j >= 0 ? ld_relaxed(&a.tile_status[j]) : ST_INC;
                    } while (__any_sync(0xffffffffu, (s & ST_MASK) == 0));
                    const uint32_t inc_mask = __ballot_sync(0xffffffffu, (s & ST_MASK) == ST_INC);
                    uint64_t val = s & ~ST_MASK;
                    if (inc_mask) {
                        const int first = __ffs(inc_mask) - 1;   // nearest predecessor holding an inclusive prefix
                        if (lane > first) val = 0;
                        excl += warp_sum_u64(val);
                        break;
                    }
                    excl += warp_sum_u64(val);
                    j0 -= 32;
                }
                if (lane == 0) st_relaxed(&a.tile_status[tile_id], ST_INC | (excl + agg));
            }
            if (lane == 0) {
                s_base = excl;
                if (first_of_genome) {
                    const uint32_t g = cd.genome & 0x7FFFFFFFu;
                    a.genome_seed_start[g] = (uint32_t)(excl & CNT_MASK);
                    a.genome_marker_start[g] = (uint32_t)((excl >> 31) & CNT_MASK);
                }
                if (tile_id == a.n_tiles - 1) {
                    const uint64_t inc = excl + agg;
                    if (a.base_out) st_relaxed(a.base_out, inc);
                    if (a.is_last) {
                        a.genome_seed_start[a.n_genomes] = (uint32_t)(inc & CNT_MASK);
                        a.genome_marker_start[a.n_genomes] = (uint32_t)((inc >> 31) & CNT_MASK);
                    }
                }
            }
        }
        __syncthreads();

        // ---- ordered write-out
        if (cs | cm) {
            const uint64_t ex_s = (is - cs) + s_wseed[warp], ex_m = (im - cm) + s_wmark[warp];
            const uint64_t b = s_base;
            const uint32_t base_s = (uint32_t)(b & CNT_MASK), base_m = (uint32_t)((b >> 31) & CNT_MASK);
            const uint64_t gkey = (uint64_t)(cd.genome & 0x7FFFFFFFu) << 42;
#pragma unroll 1
            for (int j = 0; j < WORDS_PER_THREAD; j++) {
                const uint32_t w = j * SEED_THREADS + t;
                const uint32_t mk = s_masks[w];
                uint32_t both = (mk | (mk >> 16)) & 0xFFFFu;
                if (!both) continue;
                const uint32_t w2 = s_pk[w], w1 = s_pk[w + 1], w0 = s_pk[w + 2];
                const uint32_t r0 = revcomp_word(w0), r1 = revcomp_word(w1), r2 = revcomp_word(w2);
                uint32_t so = base_s + ((uint32_t)(ex_s >> (16 * j)) & 0xFFFFu);
                uint32_t mo = base_m + ((uint32_t)(ex_m >> (16 * j)) & 0xFFFFu);
                const uint32_t p0 = pos0 + 16u * w;
                while (both) {
                    const int e = __ffs(both) - 1;
                    both &= both - 1;
                    const KmerPair kp = kmers_at(w2, w1, w0, r2, r1, r0, e);
                    if ((mk >> e) & 1u) {
                        const uint32_t fk = (uint32_t)kp.f21 & a.kmask;
                        const uint32_t rk = (uint32_t)(kp.r21 >> a.kshift);
                        const bool canon = fk < rk;
                        if (so < a.seed_cap) {
                            a.kmer_p[so] = canon ? fk : rk;
                            a.pos_p[so] = p0 + e;
                            a.meta_p[so] = (cd.contig << 1) | (uint32_t)canon;
                        } else {
                            *a.overflow = 1u;
                        }
                        so++;
                    }
                    if ((mk >> (16 + e)) & 1u) {
                        if (mo < a.marker_cap) a.marker_keys[mo] = gkey | (kp.f21 < kp.r21 ? kp.f21 : kp.r21);
                        else *a.overflow = 1u;
                        mo++;
                    }
                }
            }
        }
        // the next iteration's first __syncthreads separates these reads of shared state from its writes
    }
}

}  // namespace

void launch_seed_scan(const SeedScanArgs& a, int n_sm, cudaStream_t st) {
    if (a.n_tiles == 0) return;
    uint32_t grid = (uint32_t)n_sm * 3u;
    if (grid > a.n_tiles) grid = a.n_tiles;
    seed_scan_kernel<<<grid, SEED_THREADS, 0, st>>>(a);
    g_kernel_launches++;
}

}  // namespace skb
