// seed_kernels.cu — FracMinHash seeding on sm_100a.
//
// Replaces skani::seeding::fmh_seeds (one call per contig at reference lib.rs:165-171) for a whole batch
// of genomes in ONE launch.  Work unit: a tile of TILE_BASES consecutive bases of one contig.
//   * each thread loads 16 ASCII bytes with one 128-bit read-only load and packs them to a 2-bit word
//   * k-mers are cut out of three consecutive words with funnel shifts (no rolling dependency chain)
//   * both hashes (k-mer and 21-mer marker) are evaluated for all 16 positions, giving two 16-bit masks
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

__global__ void __launch_bounds__(SEED_THREADS, 4) seed_scan_kernel(const SeedScanArgs a) {
    __shared__ uint32_t s_pk[SEED_THREADS + 2];
    __shared__ uint32_t s_warp[SEED_THREADS / 32];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_base;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    while (true) {
        if (t == 0) s_tile = atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        const uint32_t tile_id = s_tile;
        if (tile_id >= a.n_tiles) break;
        const Tile tl = a.tiles[tile_id];
        const uint8_t* base = a.seq + tl.seq_off;
        const bool active = 16u * t < tl.n;

        uint32_t w0 = 0;
        if (active) {
            uint4 v = ld_stream16(base + 16 * t);
            w0 = pack16(v.x, v.y, v.z, v.w);
        }
        s_pk[t + 2] = w0;
        if (t < 2) {
            uint32_t h = 0;
            if (tl.pos0 > 0) {   // the two words before the tile belong to the same contig
                uint4 v = ld_stream16(base - 32 + 16 * t);
                h = pack16(v.x, v.y, v.z, v.w);
            }
            s_pk[t] = h;
        }
        __syncthreads();
        const uint32_t w2 = s_pk[t], w1 = s_pk[t + 1];
        const uint32_t r0 = revcomp_word(w0), r1 = revcomp_word(w1), r2 = revcomp_word(w2);

        uint32_t smask = 0, mmask = 0;
        if (active) {
#pragma unroll
            for (int e = 0; e < 16; e++) {
                KmerPair kp = kmers_at(w2, w1, w0, r2, r1, r0, e);
                SeedEval ev = eval_position(kp, a.kmask, a.kshift, a.thr_seed, a.thr_marker);
                smask |= (uint32_t)ev.is_seed << e;
                mmask |= (uint32_t)ev.is_marker << e;
            }
            const uint32_t left = tl.n - 16u * t;                 // bases of the contig from this word on
            uint32_t valid = left >= 16 ? 0xFFFFu : ((1u << left) - 1u);
            const uint32_t p0 = tl.pos0 + 16u * t;                // contig position of base 0 of the word
            if (p0 < SKB_MARKER_K - 1) {                          // first window ends at position 20
                uint32_t skip = SKB_MARKER_K - 1 - p0;
                valid &= skip >= 16 ? 0u : ~((1u << skip) - 1u);
            }
            smask &= valid;
            mmask &= valid;
        }

        // ---- CTA scan of (seeds | markers << 16)
        const uint32_t cnt = __popc(smask) | (__popc(mmask) << 16);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();

        if (warp == 0) {
            constexpr int NW = SEED_THREADS / 32;
            const uint32_t wt = lane < NW ? s_warp[lane] : 0u;
            uint32_t winc = wt;
#pragma unroll
            for (int o = 1; o < NW; o <<= 1) {
                uint32_t n = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += n;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, winc, NW - 1);
            if (lane < NW) s_warp[lane] = winc - wt;              // exclusive prefix of each warp

            // ---- decoupled look-back
            const uint64_t agg = (uint64_t)(total & 0xFFFFu) | ((uint64_t)(total >> 16) << 31);
            uint64_t excl = 0;
            if (tile_id == 0) {
                if (lane == 0) st_relaxed(&a.tile_status[0], ST_INC | agg);
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
                if (tl.genome & 0x80000000u) {
                    const uint32_t g = tl.genome & 0x7FFFFFFFu;
                    a.genome_seed_start[g] = (uint32_t)(excl & CNT_MASK);
                    a.genome_marker_start[g] = (uint32_t)((excl >> 31) & CNT_MASK);
                }
                if (tile_id == a.n_tiles - 1) {
                    const uint64_t inc = excl + agg;
                    a.genome_seed_start[a.n_genomes] = (uint32_t)(inc & CNT_MASK);
                    a.genome_marker_start[a.n_genomes] = (uint32_t)((inc >> 31) & CNT_MASK);
                }
            }
        }
        __syncthreads();

        // ---- ordered write-out
        if (smask | mmask) {
            const uint32_t ex = (incl - cnt) + s_warp[warp];
            const uint64_t b = s_base;
            uint32_t so = (uint32_t)(b & CNT_MASK) + (ex & 0xFFFFu);
            uint32_t mo = (uint32_t)((b >> 31) & CNT_MASK) + (ex >> 16);
            const uint32_t p0 = tl.pos0 + 16u * t;
            const uint64_t gkey = (uint64_t)(tl.genome & 0x7FFFFFFFu) << 42;
            uint32_t both = smask | mmask;
            while (both) {
                const int e = __ffs(both) - 1;
                both &= both - 1;
                KmerPair kp = kmers_at(w2, w1, w0, r2, r1, r0, e);
                if ((smask >> e) & 1u) {
                    const uint32_t fk = (uint32_t)kp.f21 & a.kmask;
                    const uint32_t rk = (uint32_t)(kp.r21 >> a.kshift);
                    const bool canon = fk < rk;
                    if (so < a.seed_cap) {
                        a.kmer_p[so] = canon ? fk : rk;
                        a.pos_p[so] = p0 + e;
                        a.meta_p[so] = (tl.contig << 1) | (uint32_t)canon;
                    } else {
                        *a.overflow = 1u;
                    }
                    so++;
                }
                if ((mmask >> e) & 1u) {
                    if (mo < a.marker_cap) a.marker_keys[mo] = gkey | (kp.f21 < kp.r21 ? kp.f21 : kp.r21);
                    else *a.overflow = 1u;
                    mo++;
                }
            }
        }
        // the next iteration's first __syncthreads separates these reads of shared state from its writes
    }
}

}  // namespace

void launch_seed_scan(const SeedScanArgs& a, int n_sm, cudaStream_t st) {
    if (a.n_tiles == 0) return;
    uint32_t grid = (uint32_t)n_sm * 4u;
    if (grid > a.n_tiles) grid = a.n_tiles;
    seed_scan_kernel<<<grid, SEED_THREADS, 0, st>>>(a);
    g_kernel_launches++;
}

}  // namespace skb
