// seed_kernels.cu — FracMinHash seeding on sm_100a.
//
// Replaces skani::seeding::fmh_seeds (one call per contig at reference lib.rs:165-171) for a whole batch
// of genomes in ONE launch.  Work unit: a tile of TILE_BASES (2 048) consecutive bases of one contig, owned by
// one WARP; warps claim regions of CHUNK_TILES consecutive tiles from one atomic counter and own a private, ordered
// output region for each.
//   * each lane loads 4 x 16 ASCII bytes with 128-bit read-only loads (warp-contiguous 512 B each) and packs
//     them to 2-bit words in the warp's slice of shared memory
//   * k-mers are cut out of three consecutive words with funnel shifts (no rolling dependency chain)
//   * both hashes (k-mer and 21-mer marker) are evaluated for all 16 positions of a word in 32-bit halves:
//     multiplies go to the FMA pipe (IMAD.WIDE / IMAD), xor-shifts and compares to the ALU pipe
//   * popcounts are scanned over the warp with shuffles; hit positions are re-extracted and written to their
//     exact slot of the warp's region, so a region is ordered by (genome, contig, position)
//   * apart from the claim counter there is no block barrier, no atomic on data and no inter-warp dependency: warps
//     never wait on each other.
// A one-CTA scan over the region counts and a gather (which doubles as the copy into exact-size arrays) stitch
// the regions together in tile order, so seeds come out ordered without a sort and without a second pass over
// the sequence.
#include "kmer_bits.cuh"
#include "skb_internal.cuh"

namespace skb {

unsigned long long g_kernel_launches = 0;

namespace {

__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// ---- mm_hash64 in 32-bit halves (see kmer_bits.cuh::mm_hash64 for the reference form) -------------------
struct U64 { uint32_t lo, hi; };

#ifndef SKB_MUL_SPLIT
#define SKB_MUL_SPLIT 3
#endif
// x * c mod 2^64.  Plain: IMAD.WIDE + IMAD.  SPLIT: the low and high word of x.lo * c by separate multiplies (IMAD +
// IMAD.HI + IMAD) - one more issue slot of the FMA pipe, but no 64-bit result: tools/micro/int_pipes.cu shows that an
// IMAD.WIDE in a stream of ALU-pipe instructions costs that pipe about one slot as well (LOP3 + SHF + IMAD.WIDE + LOP3 takes
// 3.9 ALU slots, not 3).  SKB_MUL_SPLIT: 1 = both hashes, 2 = the seed hash, 3 = the marker hash (the measured optimum: with
// both split the FMA pipe becomes the busier one).
// (c_again == c from constant memory, which the compiler cannot see through: given the same operand twice it fuses the two
// multiplies back into a wide one.)
__constant__ uint32_t c_hash_mul[3] = {(1u << 21) + 1u, 265u, 21u};
template <bool SPLIT>
__device__ __forceinline__ U64 mul_c(U64 x, uint32_t c, uint32_t c_again = 0) {
    U64 r;
    if (SPLIT) {
        r.lo = x.lo * c;
        r.hi = x.hi * c + __umulhi(x.lo, c_again);
    } else {
        const uint64_t p = (uint64_t)x.lo * c;
        r.lo = (uint32_t)p;
        r.hi = x.hi * c + (uint32_t)(p >> 32);
    }
    return r;
}
#ifndef SKB_XS_FMA
#define SKB_XS_FMA 0
#endif
// x >> S for 0 < S < 32.  With SKB_XS_FMA shifts are expressed as multiplies (IMAD.HI / IMAD) so that they issue on the FMA
// pipe and leave the ALU pipe to the xors and compares: 1 = every shift of the xor-shift steps, 2 = the high-word shifts,
// 3 = the high-word shifts of the seed hash only, 4 = those of the marker hash only.
template <int S, bool FMA>
__device__ __forceinline__ uint32_t shr_hi(uint32_t hi) {
    if (FMA) return __umulhi(hi, 1u << (32 - S));
    return hi >> S;
}
template <int S>
__device__ __forceinline__ uint32_t shr_lo(uint32_t lo, uint32_t hi) {
#if SKB_XS_FMA == 1
    return hi * (1u << (32 - S)) + __umulhi(lo, 1u << (32 - S));     // disjoint bit ranges: + == |
#else
    return __funnelshift_r(lo, hi, S);
#endif
}
template <int S, bool FMA>
__device__ __forceinline__ U64 xorshr(U64 x) {                        // x ^ (x >> S), 0 < S < 32
    U64 r;
    r.lo = x.lo ^ shr_lo<S>(x.lo, x.hi);
    r.hi = x.hi ^ shr_hi<S, FMA>(x.hi);
    return r;
}
// hash(x) < thr, x given in halves.  Step 1 is ~(x * (2^21 + 1)); the complement is folded into the first
// xor-shift:  ~a ^ (~a >> 24) == a ^ (a >> 24) ^ 0xFFFFFF00_00000000.
template <bool MARKER>
__device__ __forceinline__ U64 hash_halves(U64 x) {
    constexpr bool FMA = SKB_XS_FMA == 1 || SKB_XS_FMA == 2 || (SKB_XS_FMA == 3 && !MARKER) || (SKB_XS_FMA == 4 && MARKER);
    constexpr bool SPLIT = SKB_MUL_SPLIT == 1 || (SKB_MUL_SPLIT == 2 && !MARKER) || (SKB_MUL_SPLIT == 3 && MARKER);
    U64 a = mul_c<SPLIT>(x, (1u << 21) + 1u, c_hash_mul[0]);
    U64 b;
    b.lo = a.lo ^ shr_lo<24>(a.lo, a.hi);
    b.hi = a.hi ^ shr_hi<24, FMA>(a.hi) ^ 0xFFFFFF00u;
    b = mul_c<SPLIT>(b, 265u, c_hash_mul[1]);
    b = xorshr<14, FMA>(b);
    b = mul_c<SPLIT>(b, 21u, c_hash_mul[2]);
    b = xorshr<28, FMA>(b);
    return mul_c<false>(b, 0x80000001u);                                     // x + (x << 31)
}
// EXACT: the 64-bit comparison hash < thr.  Otherwise the comparison of the high words only, hash.hi <= thr.hi: one
// ISETP instead of two per hash (4.5 % of the ALU-pipe work of the loop).  It accepts a superset: the extra elements are
// the keys whose hash has hi == thr.hi and lo >= thr.lo, one position in ~6e9.  Every accepted position is re-checked
// exactly when it is written out (1 % of the positions); a false positive raises the launch's `inexact` flag and the host
// repeats the batch with EXACT = true, so results never depend on the shortcut.
template <bool EXACT, bool MARKER>
__device__ __forceinline__ bool hash_below(U64 x, uint32_t thr_lo, uint32_t thr_hi) {
    const U64 b = hash_halves<MARKER>(x);
    if (EXACT) return (((uint64_t)b.hi << 32) | b.lo) < (((uint64_t)thr_hi << 32) | thr_lo);
    return b.hi <= thr_hi;
}

struct WordCtx { uint32_t w0, w1, w2, r0, r1, r2; };

#ifndef SKB_MASK_FMA
#define SKB_MASK_FMA 1
#endif
// The high-word comparison and the hit mask on the FMA pipe (the ALU pipe is the one that limits this kernel; ISETP + SEL
// + IADD3 for the mask were 5 of its 42 instructions per position).  With M = 2^32 - 1 the high word of h * M is h - 1
// (0 for h = 0), so the carry of  hi(h * M) + (2^32 - thr)  is exactly [h > thr] for thr >= 1: ONE multiply-add whose only
// output is the carry predicate (IMAD.HI RZ, P, ...), and  miss = miss * 2 + carry  (IMAD.X) builds the mask of a word's 16
// positions from the last position to the first.  M and 2 arrive as launch parameters so that the assembler cannot fold
// the multiplies back into ALU-pipe instructions.  thr = 0 (no seeds wanted) has no such form: the host then asks for
// the exact comparison.
__device__ __forceinline__ uint32_t push_miss(uint32_t miss, uint32_t h, uint32_t m1, uint32_t comp, uint32_t two) {
    uint32_t r;
    asm("{ .reg .u32 t; mad.hi.cc.u32 t, %1, %2, %3; madc.lo.u32 %0, %4, %5, 0; }"
        : "=r"(r) : "r"(h), "r"(m1), "r"(comp), "r"(miss), "r"(two));
    return r;
}

// masks of the 16 positions of one word: bit e of *smask / *mmask = position e is a seed / marker
template <bool EXACT>
__device__ __forceinline__ void eval_word(const WordCtx& c, uint32_t kmask, uint32_t kshift, uint32_t ts_lo, uint32_t ts_hi,
                                          uint32_t tm_lo, uint32_t tm_hi, uint32_t m1, uint32_t two, uint32_t& smask, uint32_t& mmask) {
    smask = 0; mmask = 0;
    constexpr bool FMA_MASK = !EXACT && SKB_MASK_FMA;
    uint32_t miss_s = 0, miss_m = 0;                       // FMA_MASK: bit e = position e is NOT a seed / marker
    const uint32_t comp_s = 0u - ts_hi, comp_m = 0u - tm_hi;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int e = FMA_MASK ? 15 - i : i;
        const uint32_t s = 30 - 2 * e;
        const uint32_t flo = __funnelshift_r(c.w0, c.w1, s);
        const uint32_t fhi = __funnelshift_r(c.w1, c.w2, s) & 0x3FFu;
        const uint32_t sr = 24 + 2 * e;
        uint32_t rlo, rhi;
        if (sr < 32) { rlo = __funnelshift_r(c.r2, c.r1, sr); rhi = __funnelshift_r(c.r1, c.r0, sr) & 0x3FFu; }
        else         { rlo = __funnelshift_r(c.r1, c.r0, sr - 32); rhi = (c.r0 >> (sr - 32)) & 0x3FFu; }
        // seed k-mer
        const uint32_t fk = flo & kmask;
        const uint32_t rk = __funnelshift_r(rlo, rhi, kshift);
        const uint32_t km = min(fk, rk);
        // marker 21-mer: canonical = min of the two 42-bit values
        const bool fsmall = (((uint64_t)fhi << 32) | flo) < (((uint64_t)rhi << 32) | rlo);
        const U64 mk{fsmall ? flo : rlo, fsmall ? fhi : rhi};
        if (FMA_MASK) {
            miss_s = push_miss(miss_s, hash_halves<false>(U64{km, 0u}).hi, m1, comp_s, two);
            miss_m = push_miss(miss_m, hash_halves<true>(mk).hi, m1, comp_m, two);
        } else {
            if (hash_below<EXACT, false>(U64{km, 0u}, ts_lo, ts_hi)) smask |= 1u << e;
            if (hash_below<EXACT, true>(mk, tm_lo, tm_hi)) mmask |= 1u << e;
        }
    }
    if (FMA_MASK) { smask = ~miss_s & 0xFFFFu; mmask = ~miss_m & 0xFFFFu; }
}

#ifndef SKB_SEED_MINBLOCKS
#define SKB_SEED_MINBLOCKS 3
#endif
constexpr uint32_t HIT_LIST_CAP = 96;      // hits of one tile that take the compact write-out (2 048 bases hold ~16 seeds, ~2 markers)

__device__ __forceinline__ uint32_t ld_stream4(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// PACKED: the launch's input already consists of 2-bit words (host_pack.h: chunks that the host compacted before the
// PCIe link); word i of a contig sits at byte offset seq_off / 4 + 4 i of a.seq.  Everything after the load is shared.
template <bool EXACT, bool PACKED>
__global__ void __launch_bounds__(SEED_THREADS, SKB_SEED_MINBLOCKS) seed_scan_kernel(const SeedScanArgs a) {
    // Warps are independent: private packed-word and mask buffers, private output region, no block barriers.
    __shared__ uint32_t s_pk[SEED_WARPS][TILE_WORDS + 2];     // 2-bit packed words (two words of the previous tile in front)
    __shared__ uint32_t s_rc[SEED_WARPS][TILE_WORDS + 2];     // their reverse complements: computed once, read three times
    __shared__ uint32_t s_masks[SEED_WARPS][TILE_WORDS];      // smask | mmask << 16 per word
    __shared__ uint16_t s_hits[SEED_WARPS][2][HIT_LIST_CAP];  // (word << 4 | position) of the tile's seeds / markers, in order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ts_lo = (uint32_t)a.thr_seed, ts_hi = (uint32_t)(a.thr_seed >> 32);
    const uint32_t tm_lo = (uint32_t)a.thr_marker, tm_hi = (uint32_t)(a.thr_marker >> 32);
    uint32_t* pk = s_pk[warp];
    uint32_t* rc = s_rc[warp];
    uint32_t* masks = s_masks[warp];
    uint16_t* hits_s = s_hits[warp][0];
    uint16_t* hits_m = s_hits[warp][1];

    // Regions are claimed dynamically (one atomic per chunk_tiles <= CHUNK_TILES tiles) so that no warp idles while another still
    // has a long static range in front of it; a region's place in the output depends only on its id.
    while (true) {
    uint32_t chunk = 0;
    if (lane == 0) chunk = atomicAdd(a.chunk_counter, 1u);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk >= a.n_chunks) break;
    const uint32_t t0 = chunk * a.chunk_tiles;
    const uint32_t t1 = min(t0 + a.chunk_tiles, a.n_tiles);
    const uint32_t region = a.region_base + chunk;
    uint32_t seed_off, seed_cap, marker_off, marker_cap;
    if (a.region_off) {       // exact layout (retry): exclusive scan of the packed counts
        const uint64_t o0 = a.region_off[region], o1 = a.region_off[region + 1];
        seed_off = (uint32_t)o0; seed_cap = (uint32_t)o1 - seed_off;
        marker_off = (uint32_t)(o0 >> 32); marker_cap = (uint32_t)(o1 >> 32) - marker_off;
    } else {
        seed_off = (a.tile_base + t0) * a.seed_tile_cap; seed_cap = (t1 - t0) * a.seed_tile_cap;
        marker_off = (a.tile_base + t0) * a.marker_tile_cap; marker_cap = (t1 - t0) * a.marker_tile_cap;
    }
    uint32_t cur_s = 0, cur_m = 0;                             // records written so far (warp-uniform)

    if (t0 < t1) {
        // descriptor of the first tile: last one whose tile_start <= tile_base + t0
        uint32_t ci = 0;
        {
            uint32_t lo = 0, hi = a.n_contigs;
            const uint32_t g0 = a.tile_base + t0;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(&a.contigs[mid].tile_start) <= g0) lo = mid; else hi = mid;
            }
            ci = lo;
        }
        ContigDesc cd = a.contigs[ci];
        uint32_t next_start = ci + 1 < a.n_contigs ? __ldg(&a.contigs[ci + 1].tile_start) : 0xFFFFFFFFu;

        for (uint32_t tile = t0; tile < t1; tile++) {
            const uint32_t gtile = a.tile_base + tile;
            while (gtile >= next_start) {                      // uniform: move on to the next contig
                ci++;
                cd = a.contigs[ci];
                next_start = ci + 1 < a.n_contigs ? __ldg(&a.contigs[ci + 1].tile_start) : 0xFFFFFFFFu;
            }
            const uint32_t pos0 = (gtile - cd.tile_start) * (uint32_t)TILE_BASES;   // contig position of the tile's first base
            const uint32_t n = min((uint32_t)TILE_BASES, cd.len - pos0);             // bases in the tile
            const uint8_t* base = a.seq + cd.seq_off + pos0;
            const uint32_t genome = cd.genome & 0x7FFFFFFFu;
            if ((cd.genome & 0x80000000u) && pos0 == 0 && lane == 0) {               // first tile of a genome
                a.genome_region[genome] = region;
                a.genome_seed_local[genome] = cur_s;
                a.genome_marker_local[genome] = cur_m;
            }

            if (PACKED) {
                // ---- load: word j*32 + lane for j = 0..3 (each warp load covers 128 contiguous bytes)
                const uint32_t* wbase = reinterpret_cast<const uint32_t*>(a.seq + (cd.seq_off >> 2)) + (pos0 >> 4);
                uint32_t v[WORDS_PER_LANE];
#pragma unroll
                for (int j = 0; j < WORDS_PER_LANE; j++) {
                    const uint32_t w = j * 32 + lane;
                    v[j] = 16u * w < n ? ld_stream4(wbase + w) : 0u;
                }
                const uint32_t hv = (lane < 2 && pos0 > 0) ? ld_stream4(wbase - 2 + lane) : 0u;   // the two words before the tile
#pragma unroll
                for (int j = 0; j < WORDS_PER_LANE; j++) {
                    pk[2 + j * 32 + lane] = v[j];
                    rc[2 + j * 32 + lane] = revcomp_word(v[j]);
                }
                if (lane < 2) { pk[lane] = hv; rc[lane] = revcomp_word(hv); }
            } else {
            // ---- load + pack: word j*32 + lane for j = 0..3 (each warp load covers 512 contiguous bytes)
            uint4 v[WORDS_PER_LANE];
#pragma unroll
            for (int j = 0; j < WORDS_PER_LANE; j++) {
                const uint32_t w = j * 32 + lane;
                v[j] = make_uint4(0, 0, 0, 0);
                if (16u * w < n) v[j] = ld_stream16(base + 16u * w);
            }
            uint4 hv = make_uint4(0, 0, 0, 0);
            if (lane < 2 && pos0 > 0) hv = ld_stream16(base - 32 + 16 * lane);      // the two words before the tile
#pragma unroll
            for (int j = 0; j < WORDS_PER_LANE; j++) {
                const uint32_t word = pack16(v[j].x, v[j].y, v[j].z, v[j].w);
                pk[2 + j * 32 + lane] = word;
                rc[2 + j * 32 + lane] = revcomp_word(word);
            }
            if (lane < 2) {
                const uint32_t word = pos0 > 0 ? pack16(hv.x, hv.y, hv.z, hv.w) : 0u;
                pk[lane] = word;
                rc[lane] = revcomp_word(word);
            }
            }
            __syncwarp();

            // ---- evaluate
            uint64_t cs = 0, cm = 0;                   // per sub-tile counts, 16 bits each
#pragma unroll 1
            for (int j = 0; j < WORDS_PER_LANE; j++) {
                const uint32_t w = j * 32 + lane;
                uint32_t mk = 0;
                if (16u * w < n) {
                    WordCtx c;
                    c.w2 = pk[w]; c.w1 = pk[w + 1]; c.w0 = pk[w + 2];
                    c.r2 = rc[w]; c.r1 = rc[w + 1]; c.r0 = rc[w + 2];
                    uint32_t sm, mm;
                    eval_word<EXACT>(c, a.kmask, a.kshift, ts_lo, ts_hi, tm_lo, tm_hi, a.fma_m1, a.fma_two, sm, mm);
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
                masks[w] = mk;
            }

            // ---- warp scan: four sub-tiles at once, 16-bit lanes inside a u64 (a sub-tile holds <= 512 hits)
            uint64_t is = cs, im = cm;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t ns = __shfl_up_sync(0xffffffffu, is, o), nm = __shfl_up_sync(0xffffffffu, im, o);
                if (lane >= o) { is += ns; im += nm; }
            }
            const uint64_t tot_s = __shfl_sync(0xffffffffu, is, 31), tot_m = __shfl_sync(0xffffffffu, im, 31);
            if (tot_s | tot_m) {
                // sub-tile bases: base_j = sum of the totals of sub-tiles < j
                uint64_t sub_s = 0, sub_m = 0;
                uint32_t acc_s = 0, acc_m = 0;
#pragma unroll
                for (int j = 0; j < WORDS_PER_LANE; j++) {
                    sub_s |= (uint64_t)acc_s << (16 * j); sub_m |= (uint64_t)acc_m << (16 * j);
                    acc_s += (uint32_t)(tot_s >> (16 * j)) & 0xFFFFu; acc_m += (uint32_t)(tot_m >> (16 * j)) & 0xFFFFu;
                }
                const uint64_t ex_s = (is - cs) + sub_s, ex_m = (im - cm) + sub_m;   // <= 2048 per 16-bit lane
                const uint64_t gkey = (uint64_t)genome << 42;
                // ---- ordered write-out into the warp's region.
                // Usual case: every lane drops (word, position) of its hits into a tile-wide list at its scanned offset
                // (a few instructions per hit), then the list is worked off with ALL lanes busy and coalesced stores -
                // in the per-lane form below 7 of 8 lanes idle while one rebuilds its k-mers.
                const bool compact = acc_s <= HIT_LIST_CAP && acc_m <= HIT_LIST_CAP;
                if (compact) {
                    if (cs | cm) {
#pragma unroll 1
                        for (int j = 0; j < WORDS_PER_LANE; j++) {
                            const uint32_t w = j * 32 + lane;
                            const uint32_t mk = masks[w];
                            uint32_t sm = mk & 0xFFFFu, mm = mk >> 16;
                            uint32_t so = (uint32_t)(ex_s >> (16 * j)) & 0xFFFFu, mo = (uint32_t)(ex_m >> (16 * j)) & 0xFFFFu;
                            while (sm) { hits_s[so++] = (uint16_t)((w << 4) | (uint32_t)(__ffs(sm) - 1)); sm &= sm - 1; }
                            while (mm) { hits_m[mo++] = (uint16_t)((w << 4) | (uint32_t)(__ffs(mm) - 1)); mm &= mm - 1; }
                        }
                    }
                    __syncwarp();
                    for (uint32_t h = lane; h < acc_s; h += 32) {
                        const uint32_t ent = hits_s[h], w = ent >> 4, e = ent & 15u;
                        const KmerPair kp = kmers_at(pk[w], pk[w + 1], pk[w + 2], rc[w], rc[w + 1], rc[w + 2], (int)e);
                        const uint32_t fk = (uint32_t)kp.f21 & a.kmask;
                        const uint32_t rk = (uint32_t)(kp.r21 >> a.kshift);
                        const bool canon = fk < rk;
                        const uint32_t km = canon ? fk : rk;
                        if (!EXACT && !(mm_hash64((uint64_t)km) < a.chk_seed)) atomicOr(a.overflow, 2u);
                        const uint32_t so = cur_s + h;
                        if (so < seed_cap) {
                            a.kmer_r[seed_off + so] = km;
                            a.pos_r[seed_off + so] = pos0 + 16u * w + e;
                            a.meta_r[seed_off + so] = (cd.contig << 1) | (uint32_t)canon;
                        } else atomicOr(a.overflow, 1u);
                    }
                    for (uint32_t h = lane; h < acc_m; h += 32) {
                        const uint32_t ent = hits_m[h], w = ent >> 4, e = ent & 15u;
                        const KmerPair kp = kmers_at(pk[w], pk[w + 1], pk[w + 2], rc[w], rc[w + 1], rc[w + 2], (int)e);
                        const uint64_t mkr = kp.f21 < kp.r21 ? kp.f21 : kp.r21;
                        if (!EXACT && !(mm_hash64(mkr) < a.chk_marker)) atomicOr(a.overflow, 2u);
                        const uint32_t mo = cur_m + h;
                        if (mo < marker_cap) a.marker_r[marker_off + mo] = gkey | mkr;
                        else atomicOr(a.overflow, 1u);
                    }
                } else if (cs | cm) {
                    // dense tiles (tiny compression factors): every lane writes its own hits
#pragma unroll 1
                    for (int j = 0; j < WORDS_PER_LANE; j++) {
                        const uint32_t w = j * 32 + lane;
                        const uint32_t mk = masks[w];
                        uint32_t both = (mk | (mk >> 16)) & 0xFFFFu;
                        if (!both) continue;
                        const uint32_t w2 = pk[w], w1 = pk[w + 1], w0 = pk[w + 2];
                        const uint32_t r2 = rc[w], r1 = rc[w + 1], r0 = rc[w + 2];
                        uint32_t so = cur_s + ((uint32_t)(ex_s >> (16 * j)) & 0xFFFFu);
                        uint32_t mo = cur_m + ((uint32_t)(ex_m >> (16 * j)) & 0xFFFFu);
                        const uint32_t p0 = pos0 + 16u * w;
                        while (both) {
                            const int e = __ffs(both) - 1;
                            both &= both - 1;
                            const KmerPair kp = kmers_at(w2, w1, w0, r2, r1, r0, e);
                            if ((mk >> e) & 1u) {
                                const uint32_t fk = (uint32_t)kp.f21 & a.kmask;
                                const uint32_t rk = (uint32_t)(kp.r21 >> a.kshift);
                                const bool canon = fk < rk;
                                const uint32_t km = canon ? fk : rk;
                                if (!EXACT && !(mm_hash64((uint64_t)km) < a.chk_seed)) atomicOr(a.overflow, 2u);
                                if (so < seed_cap) {
                                    a.kmer_r[seed_off + so] = km;
                                    a.pos_r[seed_off + so] = p0 + e;
                                    a.meta_r[seed_off + so] = (cd.contig << 1) | (uint32_t)canon;
                                } else atomicOr(a.overflow, 1u);
                                so++;
                            }
                            if ((mk >> (16 + e)) & 1u) {
                                const uint64_t mkr = kp.f21 < kp.r21 ? kp.f21 : kp.r21;
                                if (!EXACT && !(mm_hash64(mkr) < a.chk_marker)) atomicOr(a.overflow, 2u);
                                if (mo < marker_cap) a.marker_r[marker_off + mo] = gkey | mkr;
                                else atomicOr(a.overflow, 1u);
                                mo++;
                            }
                        }
                    }
                }
                cur_s += acc_s; cur_m += acc_m;
            }
            __syncwarp();     // the next tile overwrites pk / rc / masks / the hit lists
        }
    }
    if (lane == 0) {
        a.region_cnt[region] = (uint64_t)cur_s | ((uint64_t)cur_m << 32);
        a.region_seed_src[region] = seed_off; a.region_marker_src[region] = marker_off;
    }
    }   // next region
}

// per-genome starts from the scanned region offsets (region_start[r] = seeds | markers << 32 before region r)
__global__ void genome_starts_kernel(uint32_t n_regions, const uint64_t* __restrict__ region_start, uint32_t n_genomes,
                                     const uint32_t* __restrict__ genome_region, const uint32_t* __restrict__ genome_seed_local,
                                     const uint32_t* __restrict__ genome_marker_local, uint32_t* __restrict__ genome_seed_start,
                                     uint32_t* __restrict__ genome_marker_start) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_genomes) return;
    if (g == n_genomes) {
        const uint64_t tot = region_start[n_regions];
        genome_seed_start[g] = (uint32_t)tot; genome_marker_start[g] = (uint32_t)(tot >> 32);
        return;
    }
    const uint32_t r = genome_region[g];
    if (r == 0xFFFFFFFFu) { genome_seed_start[g] = 0xFFFFFFFFu; genome_marker_start[g] = 0xFFFFFFFFu; return; }
    const uint64_t st = region_start[r];
    genome_seed_start[g] = (uint32_t)st + genome_seed_local[g];
    genome_marker_start[g] = (uint32_t)(st >> 32) + genome_marker_local[g];
}

// [seed_start[n] | marker_start[n] | overflow] -> pinned host memory (seed_start == NULL: only the flag)
__global__ void counters_to_host_kernel(const uint32_t* __restrict__ seed_start, const uint32_t* __restrict__ marker_start, uint32_t n,
                                        const uint32_t* __restrict__ overflow, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (seed_start && i < n) { out[i] = seed_start[i]; out[n + i] = marker_start[i]; }
    if (i == 0) out[2 * (size_t)n] = *overflow;
}

// last genome whose seed_start <= i (empty genomes share their successor's start and are skipped by the search order)
__device__ __forceinline__ uint32_t gather_genome_of(const BucketGenome* __restrict__ G, uint32_t n_genomes, uint32_t i) {
    uint32_t lo = 0, hi = n_genomes;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&G[mid].seed_start) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) region_gather_kernel(const RegionGatherArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= a.n_regions) return;
    const uint64_t st0 = a.region_start[r], st1 = a.region_start[r + 1];
    {
        const uint32_t dst = (uint32_t)st0, n = (uint32_t)st1 - dst, src = a.seed_src[r];
        // genome of the region's first and last seed (uniform); regions that straddle genomes look every seed up
        uint32_t g_first = 0, g_last = 0;
        if (a.bucket_counts && n) { g_first = gather_genome_of(a.genomes, a.n_genomes, dst); g_last = gather_genome_of(a.genomes, a.n_genomes, dst + n - 1); }
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t km = a.kmer_r[src + i];
            a.kmer_p[dst + i] = km; a.pos_p[dst + i] = a.pos_r[src + i]; a.meta_p[dst + i] = a.meta_r[src + i];
            if (a.bucket_counts) {
                const uint32_t g = g_first == g_last ? g_first : gather_genome_of(a.genomes, a.n_genomes, dst + i);
                atomicAdd(&a.bucket_counts[__ldg(&a.genomes[g].bucket_off) + (km >> __ldg(&a.genomes[g].shift))], 1u);
            }
        }
    }
    {
        const uint32_t dst = (uint32_t)(st0 >> 32), n = (uint32_t)(st1 >> 32) - dst, src = a.marker_src[r];
        for (uint32_t i = lane; i < n; i += 32) a.marker_keys[dst + i] = a.marker_r[src + i];
    }
}

}  // namespace

void launch_seed_scan(const SeedScanArgs& a, int n_sm, cudaStream_t st) {
    if (a.n_warps == 0) return;
    const dim3 grid(a.n_warps / SEED_WARPS), block(SEED_THREADS);
    if (a.packed) {
        if (a.exact_compare) seed_scan_kernel<true, true><<<grid, block, 0, st>>>(a);
        else seed_scan_kernel<false, true><<<grid, block, 0, st>>>(a);
    } else {
        if (a.exact_compare) seed_scan_kernel<true, false><<<grid, block, 0, st>>>(a);
        else seed_scan_kernel<false, false><<<grid, block, 0, st>>>(a);
    }
    g_kernel_launches++;
}

void launch_genome_starts(uint32_t n_regions, const uint64_t* region_start, uint32_t n_genomes, const uint32_t* genome_region,
                          const uint32_t* genome_seed_local, const uint32_t* genome_marker_local, uint32_t* genome_seed_start,
                          uint32_t* genome_marker_start, cudaStream_t st) {
    genome_starts_kernel<<<(n_genomes + 1 + 255) / 256, 256, 0, st>>>(n_regions, region_start, n_genomes, genome_region,
                                                                       genome_seed_local, genome_marker_local, genome_seed_start,
                                                                       genome_marker_start);
    g_kernel_launches++;
}

void launch_counters_to_host(const uint32_t* seed_start, const uint32_t* marker_start, uint32_t n, const uint32_t* overflow,
                             uint32_t* host_out, cudaStream_t st) {
    counters_to_host_kernel<<<seed_start ? (n + 255) / 256 : 1, 256, 0, st>>>(seed_start, marker_start, n, overflow, host_out);
    g_kernel_launches++;
}

void launch_region_gather(const RegionGatherArgs& a, cudaStream_t st) {
    if (a.n_regions == 0) return;
    region_gather_kernel<<<(a.n_regions + 7) / 8, 256, 0, st>>>(a);
    g_kernel_launches++;
}

}  // namespace skb
