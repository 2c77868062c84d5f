// chain_kernels.cu — anchor lookup, banded sparse chaining and ANI/AF on sm_100a.
//
// Replaces skani::chain::chain_seeds (reference lib.rs:652-653) for a whole batch of screened-in pairs.
// The algorithm is the one frozen in oracle/skani_oracle.cpp (orc_chain_params_default):
//   1. anchors   = every (query seed, reference seed) pair with equal k-mer, in
//                  (q_contig, q_pos, r_contig, r_pos) order.  The query side is walked in position order and
//                  each seed is looked up in the reference's k-mer-sorted array through its bucket table,
//                  so the anchors come out already sorted: no per-pair sort.
//   2. windows   : a window opens at the first anchor of a contig / the first anchor >= start + 20000 bp.
//   3. DP        : inside a window, f[i] = max(20, max_j f[j] + 20 - |dr - dq|) over the <= 100 previous
//                  anchors within 2500 bp on the query, same reference contig and strand, dq > 0, dr > 0,
//                  |dr - dq| <= 300; ties go to the nearest predecessor.  One warp per window: lane l keeps
//                  the newest anchor whose index is congruent to l (mod 32) in registers, so the 32 most recent
//                  anchors are scored against the current one on chip; older ones (rarely in band) are read back
//                  from the anchor arrays (dp_window below).
//   4. chains    = components of the back-pointer forest; score = best f in the component, extent = root ..
//                  best anchor, weight = component size; keep size >= 3 and score >= 45; greedy by score
//                  without query overlap inside the window.
//   5. window ANI = min(1, anchors / query seeds spanned)^(1/k); genome ANI = seed-weighted mean (or
//                  10-90 % trimmed, or median); AF = sum of (chain span + 198) / genome length.
// Scores are integers (20 per anchor minus integer gaps), so the int32 DP is bit-identical to skani's f64.
#include <cstdio>
#include <cstdlib>
#include "skb_internal.cuh"

namespace skb {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t AUX_PROCESSED = 0x80000000u, AUX_ACCEPTED = 0x40000000u, AUX_SIZE = 0x3FFFFFFFu;

// first index in [lo, hi) whose key >= target, keys non-decreasing; all lanes of the warp cooperate
template <typename KeyFn>
__device__ __forceinline__ uint32_t warp_lower_bound(uint32_t lo, uint32_t hi, uint32_t target, int lane, KeyFn key) {
    while (true) {
        const uint32_t n = hi - lo;
        if (n == 0) return lo;
        if (n <= 32) {
            const uint32_t idx = lo + lane;
            const bool ge = idx >= hi || key(idx) >= target;
            const uint32_t b = __ballot_sync(FULL, ge);
            return b ? lo + (uint32_t)(__ffs(b) - 1) : hi;   // n == 32 with no hit leaves b == 0
        }
        const uint32_t stride = (n + 31) / 32;
        const uint32_t idx = lo + lane * stride;
        const bool ge = idx >= hi || key(idx) >= target;
        const uint32_t b = __ballot_sync(FULL, ge);
        if (b == 0) {                       // every probe below target and inside the range
            lo = lo + 31 * stride + 1;
            continue;
        }
        const int f = __ffs(b) - 1;
        if (f == 0) return lo;
        const uint32_t new_lo = lo + (uint32_t)(f - 1) * stride + 1;
        const uint32_t new_hi = min(lo + (uint32_t)f * stride, hi);
        lo = new_lo; hi = new_hi;
    }
}

// ------------------------------------------------------------------ 1a. match counts
// The query is walked in K-MER order: the 32 k-mers of a warp are neighbours in the sorted k-mer space, so their bucket
// entries and the reference k-mers they are compared with share a few sectors (in position order every lane touched its
// own: ~20 L1 sector lookups per query seed, the L1 tag stage was the limit).  The result goes to the seed's slot in
// POSITION order through perm_k - one 8-byte store per MATCHED seed into the zeroed (first, count) array; the anchors
// built from it therefore still come out sorted by (q_contig, q_pos, r_contig, r_pos): no per-pair sort.
__global__ void match_count_kernel(const ChainBatch b) {
    const PairDesc pd = b.pairs[blockIdx.y];
    const GenomeView& Q = b.qviews[pd.q];
    const GenomeView& R = b.rviews[pd.r];
    const uint32_t nq = Q.n_seeds;
    if (R.n_seeds == 0) return;
    const int lane = threadIdx.x & 31;
    unsigned long long total = 0;       // 64-bit anchor count of this thread's seeds (a_off is a 32-bit scan and may wrap)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) {
        const uint32_t km = __ldg(Q.kmer_k + i);
        const uint32_t bk = km >> R.bucket_shift;
        uint32_t lo = __ldg(R.bucket + bk), hi = __ldg(R.bucket + bk + 1);
        const uint32_t end = hi;
        while (lo < hi) {               // buckets hold ~8 seeds: a short search
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(R.kmer_k + mid) < km) lo = mid + 1; else hi = mid;
        }
        const uint32_t first = lo;
        while (lo < end && __ldg(R.kmer_k + lo) == km) lo++;
        const uint32_t cnt = lo - first;
        if (cnt) {
            b.m_fc[pd.seed_off + __ldg(Q.perm_k + i)] = make_uint2(first, cnt);
            total += cnt;
        }
    }
    for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
    if (lane == 0 && total) atomicAdd(b.a_total64, total);
}

// ------------------------------------------------------------------ 1b. anchors
__global__ void anchor_fill_kernel(const ChainBatch b) {
    const PairDesc pd = b.pairs[blockIdx.y];
    const GenomeView& Q = b.qviews[pd.q];
    const GenomeView& R = b.rviews[pd.r];
    const uint32_t nq = Q.n_seeds;
    const int lane = threadIdx.x & 31;
    // warp-aligned strips of 32 consecutive query seeds: the strip's "has a match" bits are one word of the bitmask the
    // window walk reads
    for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; i0 < nq; i0 += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + lane;
        uint2 fc = make_uint2(0u, 0u);
        if (i < nq) fc = b.m_fc[pd.seed_off + i];
        const uint32_t bal = __ballot_sync(FULL, fc.y != 0u);
        if (lane == 0) b.m_bits[pd.bits_off + (i0 >> 5)] = bal;
        if (fc.y == 0u) continue;
        const uint32_t first = fc.x, cnt = fc.y;
        const uint32_t off = b.a_off[pd.seed_off + i];
        const uint32_t qp = __ldg(Q.pos_p + i);
        const uint32_t qm = __ldg(Q.meta_p + i);
        for (uint32_t m = 0; m < cnt; m++) {
            const uint32_t o = off + m;
            if (o >= b.anchor_cap) break;
            const uint32_t rm = __ldg(R.meta_k + first + m);
            // (q_pos, r_pos, ref contig << 1 | reverse_match, query seed index): one 16-byte store
            b.a_rec[o] = make_uint4(qp, __ldg(R.pos_k + first + m), (rm & ~1u) | ((qm ^ rm) & 1u), i);
        }
    }
}

// ------------------------------------------------------------------ 2. windows
// A window opens at the first matched query seed of a contig and then at the first matched seed whose position is
// >= the previous opening position + F.  That is a serial chain per contig (~230 links for a 5 Mbp contig), so the
// cost is the latency of one link.  Window j of contig c lands in slot win_off + contig_win_start[c] + j.
//
// Fast path (window_walk_smem_kernel): one CTA per pair stages the query's seed positions and the "has a match" bitmask
// (written by match_count_kernel) in shared memory with TMA bulk copies (4 B + 1 bit per seed; <= WALK_SMEM_SEEDS
// seeds), then one warp per contig follows the chain with every link resolved on chip.  Fallback (window_walk_kernel): same walk with 32-ary searches in global memory.
constexpr uint32_t WALK_SMEM_SEEDS = 49152;     // 192 KB of positions + 6 KB of bits

__global__ void window_walk_kernel(const ChainBatch b, const uint32_t F, const int only_large) {
    const PairDesc pd = b.pairs[blockIdx.x];
    const GenomeView& Q = b.qviews[pd.q];
    if (only_large && Q.n_seeds <= WALK_SMEM_SEEDS) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t* aoff = b.a_off + pd.seed_off;
    for (uint32_t c = warp; c < Q.n_contigs; c += nwarps) {
        const uint32_t cs = Q.contig_seed_start[c], ce = Q.contig_seed_start[c + 1];
        uint32_t slot = pd.win_off + Q.contig_win_start[c];
        uint32_t s = cs;
        while (s < ce) {
            // first matched seed i >= s  <=>  first j in [s+1, ce] with aoff[j] > aoff[s]; i = j - 1
            const uint32_t base = aoff[s];
            const uint32_t j = warp_lower_bound(s + 1, ce + 1, base + 1, lane, [&](uint32_t x) { return aoff[x]; });
            if (j > ce) break;
            const uint32_t i = j - 1;
            const uint32_t p0 = __ldg(Q.pos_p + i);
            const uint32_t e = warp_lower_bound(i + 1, ce, p0 + F, lane, [&](uint32_t x) { return __ldg(Q.pos_p + x); });
            if (lane == 0) {
                b.win_start[slot] = i;
                b.win_end[slot] = e;
                b.win_contig[slot] = blockIdx.x;   // pair index, read back by the DP kernel
            }
            slot++;
            s = e;
        }
    }
}

// first set bit at index >= from in bits[0..n), or n
__device__ __forceinline__ uint32_t next_set_bit(const uint32_t* bits, uint32_t from, uint32_t n) {
    uint32_t w = from >> 5;
    const uint32_t nw = (n + 31) >> 5;
    if (w >= nw) return n;
    uint32_t cur = bits[w] & (0xFFFFFFFFu << (from & 31));
    while (cur == 0) {
        if (++w >= nw) return n;
        cur = bits[w];
    }
    const uint32_t r = (w << 5) + (uint32_t)__ffs(cur) - 1u;
    return r < n ? r : n;
}

// ---- TMA 1-D bulk copy (cp.async.bulk, SASS: UBLKCP) + mbarrier, used to stage a pair's arrays in shared memory
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// One CTA per *walk group*: up to walk_group_max consecutive pairs of the batch that share their query.  The query's
// seed positions are staged once, the match bitmask of every pair of the group next to them, and the (pair, contig) chains
// of the group are spread over the warps: with many pairs per query (all-vs-all) a group keeps 8+ chains in flight per SM
// instead of one.
__global__ void __launch_bounds__(1024) window_walk_smem_kernel(const ChainBatch b, const uint32_t F) {
    extern __shared__ __align__(128) uint32_t sm[];
    __shared__ uint64_t s_bar;
    const uint2 grp = b.walk_groups[blockIdx.x];      // first pair, number of pairs
    const PairDesc pd0 = b.pairs[grp.x];
    const GenomeView& Q = b.qviews[pd0.q];
    const uint32_t n = Q.n_seeds;
    if (n > WALK_SMEM_SEEDS) return;                  // handled by the global-memory kernel
    // The query's seed positions keep their 16-byte phase in shared memory, so the 16-byte aligned middle of the
    // array can go through one TMA bulk copy per 32 KB; the (<= 3 + 3) unaligned head/tail words are copied by threads.
    const uint32_t* src = Q.pos_p;
    const uint32_t phase = (uint32_t)(((uintptr_t)src >> 2) & 3u);
    uint32_t* s_pos = sm + phase;                     // [n]
    const uint32_t head = min(n, (4u - phase) & 3u);
    const uint32_t mid = ((n - head) >> 2) << 2;
    const uint32_t n_bit_words = (((n + 31) >> 5) + 3u) & ~3u;                 // padded to 16 bytes (so is the source slice)
    uint32_t* s_bits0 = sm + ((phase + n + 3u) & ~3u);                         // 16-byte aligned; pair k at + k * n_bit_words
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&s_bar, mid * 4u + grp.y * n_bit_words * 4u);
        for (uint32_t off = 0; off < mid; off += 8192) {
            const uint32_t cnt = min(8192u, mid - off);
            bulk_g2s(s_pos + head + off, src + head + off, cnt * 4u, &s_bar);
        }
        if (n_bit_words)
            for (uint32_t k = 0; k < grp.y; k++)
                bulk_g2s(s_bits0 + k * n_bit_words, b.m_bits + b.pairs[grp.x + k].bits_off, n_bit_words * 4u, &s_bar);
    }
    for (uint32_t i = threadIdx.x; i < n - mid; i += blockDim.x) {             // head and tail words
        const uint32_t j = i < head ? i : mid + i;
        s_pos[j] = __ldg(src + j);
    }
    mbar_wait(&s_bar, 0);
    __syncthreads();
    const uint32_t n_contigs = Q.n_contigs;
    for (uint32_t item = warp; item < grp.y * n_contigs; item += nwarps) {
        const uint32_t k = item / n_contigs, c = item - k * n_contigs;
        const uint32_t pair = grp.x + k;
        const PairDesc pd = b.pairs[pair];
        const uint32_t* s_bits = s_bits0 + k * n_bit_words;
        const uint32_t cs = Q.contig_seed_start[c], ce = Q.contig_seed_start[c + 1];
        uint32_t slot = pd.win_off + Q.contig_win_start[c];
        uint32_t s = cs;
        // expected number of seeds per window of this contig
        const uint32_t clen = Q.contig_len[c];
        const uint32_t est = clen ? (uint32_t)(((uint64_t)F * (ce - cs)) / clen) : 0u;
        while (s < ce) {
            // first matched seed at or after s (uniform across the warp: every lane runs the same scan)
            const uint32_t i = next_set_bit(s_bits, s, ce);
            if (i >= ce) break;
            const uint32_t target = s_pos[i] + F;
            uint32_t lo = i + 1, e = 0;
            bool done = false;
            // one-shot guess: seeds are roughly evenly spaced, so the answer is close to i + est; probe 32 consecutive
            // seeds around it and accept if the bracket is inside (the seed before the probes is still below target)
            if (est > 16) {
                const uint32_t g0 = min(i + 1 + (est - 16), ce);
                const uint32_t idx = g0 + (uint32_t)lane;
                const bool ge = idx >= ce || s_pos[idx] >= target;
                const uint32_t bal = __ballot_sync(FULL, ge);
                const bool before_ok = g0 == i + 1 || s_pos[g0 - 1] < target;     // uniform: same address for every lane
                if (before_ok) {
                    if (bal) { e = g0 + (uint32_t)(__ffs(bal) - 1); done = true; }
                    else lo = g0 + 32;                                             // every probe below target: continue behind them
                }
            }
            // gallop: 32 probes, 8 seeds apart, then resolve inside the 8-seed bracket
            while (!done) {
                const uint32_t idx = lo + (uint32_t)lane * 8u;
                const bool ge = idx >= ce || s_pos[idx] >= target;
                const uint32_t bal = __ballot_sync(FULL, ge);
                if (bal == 0) { lo += 32u * 8u - 7u; continue; }      // all 32 probes below target: restart after the last one
                const int f = __ffs(bal) - 1;
                const uint32_t blo = f == 0 ? lo : lo + (uint32_t)(f - 1) * 8u + 1u;
                const uint32_t bhi = min(lo + (uint32_t)f * 8u, ce);  // answer in [blo, bhi]
                const uint32_t j = blo + (uint32_t)lane;
                const bool ge2 = j >= bhi || s_pos[j] >= target;       // lanes >= 8 are past bhi: vote true
                e = blo + (uint32_t)(__ffs(__ballot_sync(FULL, ge2)) - 1);
                done = true;
            }
            if (lane == 0) {
                b.win_start[slot] = i;
                b.win_end[slot] = e;
                b.win_contig[slot] = pair;
            }
            slot++;
            s = e;
        }
    }
}

// ------------------------------------------------------------------ 3+4. DP, chains, per-window record

__device__ __forceinline__ void tuple_min(int32_t& sc, uint32_t& qs, uint32_t& rs, uint32_t& idx,
                                          int32_t sc2, uint32_t qs2, uint32_t rs2, uint32_t idx2) {
    // order: score desc, qs asc, rs asc, idx asc
    bool take = sc2 > sc || (sc2 == sc && (qs2 < qs || (qs2 == qs && (rs2 < rs || (rs2 == rs && idx2 < idx)))));
    if (take) { sc = sc2; qs = qs2; rs = rs2; idx = idx2; }
}

// The chaining DP of one window, run by one warp.
// Lane l keeps the newest anchor with index == l (mod 32) in registers: the 32 nearest predecessors of the current anchor
// are always on chip.  Anchors are ordered by query position and ~20 of them fall inside the 2 500 bp band, so older
// predecessors (distance 33..index_band) are only needed in repeat-rich windows; they are read back from the anchor
// arrays (f of a finished strip of 32 is already stored).  Which anchors of a strip need that is known when the strip
// is loaded (anchor i needs it iff anchor i - 32 is still within the band), so it costs one ballot per strip.
// The strip itself sits in shared memory as (qp, rp, meta) records: one broadcast LDS.128 per anchor.
// Score and distance travel through ONE warp reduction as (score + bias) << 8 | (255 - distance): maximal score first,
// nearest predecessor on ties.  WIDE: windows with so many anchors that the packed score could overflow (>= 2^22 /
// anchor_score anchors) use one reduction for the score and one for the distance.
// The loop only records each anchor's predecessor; the component roots follow afterwards by pointer jumping
// (~log2(chain length) rounds over the window instead of a dependent lookup per anchor).
// The common path is straight-line code (selects instead of branches) so that the warp stays provably converged.

template <bool WIDE>
__device__ __forceinline__ void dp_window(const uint4* __restrict__ rec_a, int32_t* f_a, uint32_t* root_a,
                                          const uint32_t n, const int lane, uint4* s_strip) {
    // skani's chaining constants are frozen (pyskani exposes none of them): compile-time values
    // here, shared with the host through skb_internal.cuh (ChainConsts is filled from the same constants)
    constexpr uint32_t bp_band = DP_BP_BAND, index_band = DP_INDEX_BAND;
    constexpr int32_t max_gap = DP_MAX_GAP, anchor_score = DP_ANCHOR_SCORE;
    constexpr int32_t link_bias = anchor_score + max_gap + 1;      // score of a valid link + max_gap + 1 - gap  >=  1
    uint32_t g_qp = 0, g_rp = 0, g_meta = 0xFFFFFFFFu, g_par = 0;   // meta 0xFFFFFFFF never matches
    int32_t g_f = 0;
    for (uint32_t sb = 0; sb < n; sb += 32) {
        const uint32_t mine = sb + lane;
        uint32_t sq = 0, sr = 0, sm = 0;
        if (mine < n) { const uint4 r4 = rec_a[mine]; sq = r4.x; sr = r4.y; sm = r4.z; }
        s_strip[lane] = make_uint4(sq, sr, sm, 0u);
        // g_* still holds anchor mine - 32 here: bit u of need_old = anchor sb + u must also look at distances > 32
        const uint32_t need_old = __ballot_sync(FULL, mine < n && mine >= 32u && (sq - g_qp) <= bp_band);
        __syncwarp();
        const uint32_t lim = min(32u, n - sb);
        uint32_t kd = 255u - (uint32_t)(32 - lane);                 // 255 - distance of this lane's newest anchor from sb + u
        for (uint32_t u = 0; u < lim; u++) {
            const uint32_t i = sb + u;                              // window-local anchor index; owner lane = u
            const uint4 cur = s_strip[u];
            const uint32_t cq = cur.x, cr = cur.y, cm = cur.z;
            const uint32_t revmask = 0u - (cm & 1u);
            // ---- the 32 nearest predecessors (registers)
            const int32_t dq = (int32_t)(cq - g_qp);
            const int32_t dr = (int32_t)(((cr - g_rp) ^ revmask) - revmask);      // reverse strand: g.rp - cr
            const int32_t gap = abs(dr - dq);
            int32_t best; uint32_t bestd;
            if (!WIDE) {
                // key = ok ? packed : 0, with the four conditions chained through one predicate (the compiler otherwise
                // emits one select per condition)
                const uint32_t packed = ((uint32_t)(g_f + link_bias - gap) << 8) | kd;
                uint32_t key;
                asm("{\n\t.reg .pred p;\n\t"
                    "setp.lt.u32 p, %1, %2;\n\t"
                    "setp.eq.and.u32 p, %3, %4, p;\n\t"
                    "setp.gt.and.s32 p, %5, 0, p;\n\t"
                    "setp.le.and.s32 p, %6, %7, p;\n\t"
                    "selp.u32 %0, %8, 0, p;\n\t}"
                    : "=r"(key)
                    : "r"((uint32_t)(dq - 1)), "r"(bp_band), "r"(g_meta), "r"(cm), "r"(dr), "r"(gap), "r"(max_gap), "r"(packed));
                const uint32_t mk = __reduce_max_sync(FULL, key);
                const int32_t m = (int32_t)(mk >> 8) - (max_gap + 1);
                const bool take = m > anchor_score;                 // mk == 0 gives m < 0
                best = take ? m : anchor_score;
                bestd = take ? 255u - (mk & 255u) : 0u;
            } else {
                const bool ok = ((uint32_t)(dq - 1) < bp_band) & (g_meta == cm) & (dr > 0) & (gap <= max_gap);
                const int32_t sc = ok ? g_f + anchor_score - gap : INT32_MIN;
                const int32_t m = __reduce_max_sync(FULL, sc);
                const uint32_t dm = __reduce_min_sync(FULL, sc == m ? 255u - kd : 0x7FFFFFFFu);
                const bool take = m > anchor_score;
                best = take ? m : anchor_score;
                bestd = take ? dm : 0u;
            }
            // ---- older predecessors: only if the anchor 32 back is still in band (then 64 back, 96 back)
            if ((need_old >> u) & 1u) {
                const uint32_t d0 = 255u - kd;
                for (uint32_t g = 1; g < 4; g++) {
                    const uint32_t d = d0 + 32u * g;
                    bool inband = d <= i && d <= index_band;
                    int32_t sc = INT32_MIN;
                    if (inband) {
                        const uint32_t j = i - d;
                        const uint4 pj = rec_a[j];
                        const uint32_t pq = pj.x;
                        inband = (cq - pq) <= bp_band;
                        if (inband && pj.z == cm) {
                            const uint32_t pr = pj.y;
                            const int32_t dq2 = (int32_t)(cq - pq);
                            const int32_t dr2 = (int32_t)(((cr - pr) ^ revmask) - revmask);
                            const int32_t gap2 = abs(dr2 - dq2);
                            if (dq2 > 0 && dr2 > 0 && gap2 <= max_gap) sc = f_a[j] + anchor_score - gap2;
                        }
                    }
                    const int32_t m = __reduce_max_sync(FULL, sc);
                    if (m > best) {                               // strictly better than anything nearer
                        best = m;
                        bestd = __reduce_min_sync(FULL, sc == m ? d : 0x7FFFFFFFu);
                    }
                    if (!((__ballot_sync(FULL, inband) >> u) & 1u)) break;
                }
            }
            const bool own = lane == (int)u;
            g_qp = own ? cq : g_qp; g_rp = own ? cr : g_rp; g_meta = own ? cm : g_meta;
            g_f = own ? best : g_f; g_par = own ? bestd : g_par;
            kd = kd == 223u ? 254u : kd - 1u;
        }
        if (mine < n) { f_a[mine] = g_f; root_a[mine] = mine - g_par; }      // predecessor (itself if none)
        __syncwarp();                              // the strip's f is read by other lanes from here on; s_strip is rewritten
    }
    // ---- component roots: pointer jumping over the predecessor links until nothing moves
    for (;;) {
        bool moved = false;
        for (uint32_t sb = 0; sb < n; sb += 32) {
            const uint32_t i = sb + lane;
            if (i < n) {
                const uint32_t p = *(volatile const uint32_t*)(root_a + i);      // generic: shared or global
                const uint32_t pp = *(volatile const uint32_t*)(root_a + p);
                if (pp != p) { root_a[i] = pp; moved = true; }
            }
        }
        __syncwarp();
        if (!__any_sync(FULL, moved)) break;
    }
}

constexpr int DP_WARPS = 4;
constexpr uint32_t DP_SMEM_ANCHORS = 256;
__device__ __forceinline__ uint32_t ldv(const uint32_t* p) { return *(volatile const uint32_t*)p; }
__device__ __forceinline__ unsigned long long ldv(const unsigned long long* p) { return *(volatile const unsigned long long*)p; }

// Anchor range of a window slot: false for unused slots and when the anchor arrays were sized too small (the host then
// reruns the batch with the exact size).
__device__ __forceinline__ bool window_anchors(const ChainBatch& b, uint32_t slot, uint32_t& A0, uint32_t& n) {
    const uint32_t ws = b.win_start[slot], we = b.win_end[slot];
    if (we <= ws) return false;
    const PairDesc pd = b.pairs[b.win_contig[slot]];
    A0 = b.a_off[pd.seed_off + ws];
    const uint32_t A1 = b.a_off[pd.seed_off + we];
    if (A1 > b.anchor_cap) return false;
    n = A1 - A0;
    return true;
}

// ---- warp per window: the windows the thread-per-window kernel below passes on (more than DPT_MAX_ANCHORS anchors:
// repeat-rich regions), taken from big_list by a persistent grid
__global__ void __launch_bounds__(DP_WARPS * 32) chain_dp_kernel(const ChainBatch b, const ChainConsts C) {
    __shared__ uint4 s_strip[DP_WARPS][32];          // the strip of 32 anchors a warp is working on; later its candidate list
    __shared__ __align__(16) uint32_t s_work[DP_WARPS][4 * DP_SMEM_ANCHORS];   // root | size+flags | best end (64 bit)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_big = *b.big_count;
    for (uint32_t w = blockIdx.x * DP_WARPS + warp; w < n_big; w += gridDim.x * DP_WARPS) {
    const uint32_t slot = b.big_list[w];
    uint32_t A0 = 0, n = 0;
    if (!window_anchors(b, slot, A0, n)) continue;
    const uint4* rec_a = b.a_rec + A0;
    int32_t* f_a = b.a_f + A0;
    // Per-anchor working state of the chain phases (component root, size | flags, best end): shared memory for windows
    // of up to DP_SMEM_ANCHORS anchors, the global arrays beyond
    const bool small = n <= DP_SMEM_ANCHORS;
    uint32_t* root_a = small ? s_work[warp] : b.a_root + A0;
    uint32_t* aux_a = small ? s_work[warp] + DP_SMEM_ANCHORS : b.a_aux + A0;
    unsigned long long* best_a = small ? (unsigned long long*)(s_work[warp] + 2 * DP_SMEM_ANCHORS) : b.a_best + A0;
    for (uint32_t i = lane; i < n; i += 32) { aux_a[i] = 0u; best_a[i] = 0ull; }
    __threadfence_block();
    __syncwarp();
    // candidate list: at most n / min_anchors entries
    int32_t* cand_a = (small && n <= 128u * (uint32_t)max(C.min_anchors, 1)) ? (int32_t*)s_strip[warp] : f_a;

    // ---------------- DP
    if ((uint64_t)n * (uint32_t)DP_ANCHOR_SCORE < (1u << 22)) dp_window<false>(rec_a, f_a, root_a, n, lane, s_strip[warp]);
    else dp_window<true>(rec_a, f_a, root_a, n, lane, s_strip[warp]);

    // ---------------- per-component size and best end
    for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t r = root_a[i];
        atomicAdd(&aux_a[r], 1u);
        atomicMax(&best_a[r], ((unsigned long long)(uint32_t)f_a[i] << 32) | (0xFFFFFFFFu - i));
    }
    __threadfence_block();
    __syncwarp();

    // ---------------- candidate chains -> compact list of roots in cand_a[0..ncand)
    uint32_t ncand = 0;
    for (uint32_t sb = 0; sb < n; sb += 32) {
        const uint32_t i = sb + lane;
        bool cand = false;
        if (i < n && root_a[i] == i) {
            const uint32_t size = ldv(&aux_a[i]) & AUX_SIZE;
            const int32_t score = (int32_t)(ldv(&best_a[i]) >> 32);
            cand = size >= (uint32_t)C.min_anchors && score >= C.min_score;
        }
        const uint32_t bal = __ballot_sync(FULL, cand);
        // when the list shares f_a: f_a[0..ncand) is overwritten only at indices < i's strip start or by earlier lanes of
        // the strip, whose f values are no longer needed (f lives on in best_a)
        __syncwarp();
        if (cand) cand_a[ncand + __popc(bal & ((1u << lane) - 1u))] = (int32_t)i;
        ncand += __popc(bal);
        __syncwarp();
    }

    // ---------------- greedy selection without query overlap
    uint32_t w_anchors = 0, w_lo = 0xFFFFFFFFu, w_hi = 0, w_lo_qi = 0, w_hi_qi = 0, w_covq = 0, w_covr = 0, w_chains = 0;
    for (uint32_t round = 0; round < ncand; round++) {
        int32_t sc = INT32_MIN; uint32_t qs = 0xFFFFFFFFu, rs = 0xFFFFFFFFu, idx = 0xFFFFFFFFu;
        for (uint32_t t = lane; t < ncand; t += 32) {
            const uint32_t r = (uint32_t)cand_a[t];
            if (ldv(&aux_a[r]) & AUX_PROCESSED) continue;
            const unsigned long long bb = ldv(&best_a[r]);
            const uint32_t bi = 0xFFFFFFFFu - (uint32_t)bb;
            const uint4 rr = rec_a[r];
            const uint32_t rs2 = (rr.z & 1u) ? rec_a[bi].y : rr.y;
            tuple_min(sc, qs, rs, idx, (int32_t)(bb >> 32), rr.x, rs2, r);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int32_t sc2 = __shfl_xor_sync(FULL, sc, o);
            const uint32_t qs2 = __shfl_xor_sync(FULL, qs, o), rs2 = __shfl_xor_sync(FULL, rs, o), idx2 = __shfl_xor_sync(FULL, idx, o);
            tuple_min(sc, qs, rs, idx, sc2, qs2, rs2, idx2);
        }
        const uint32_t r = idx;                                 // uniform: the best unprocessed candidate
        const unsigned long long bb = ldv(&best_a[r]);
        const uint32_t bi = 0xFFFFFFFFu - (uint32_t)bb;
        const uint4 rec_r = rec_a[r], rec_b = rec_a[bi];
        const uint32_t cqs = rec_r.x, cqe = rec_b.x;
        bool ov = false;
        for (uint32_t t = lane; t < ncand; t += 32) {
            const uint32_t r2 = (uint32_t)cand_a[t];
            if (!(ldv(&aux_a[r2]) & AUX_ACCEPTED)) continue;
            const uint32_t bi2 = 0xFFFFFFFFu - (uint32_t)ldv(&best_a[r2]);
            const uint32_t lo = max(cqs, rec_a[r2].x), hi = min(cqe, rec_a[bi2].x);
            ov |= hi >= lo;
        }
        ov = __any_sync(FULL, ov);
        const uint32_t size = ldv(&aux_a[r]) & AUX_SIZE;
        if (lane == 0) aux_a[r] = size | AUX_PROCESSED | (ov ? 0u : AUX_ACCEPTED);
        __threadfence_block();
        __syncwarp();
        if (!ov) {
            const bool rv = rec_r.z & 1u;
            const uint32_t crs = rv ? rec_b.y : rec_r.y, cre = rv ? rec_r.y : rec_b.y;
            w_anchors += size;
            if (cqs < w_lo) { w_lo = cqs; w_lo_qi = rec_r.w; }
            if (cqe >= w_hi) { w_hi = cqe; w_hi_qi = rec_b.w; }
            w_covq += (cqe - cqs) + (uint32_t)C.af_ext;
            w_covr += (cre - crs) + (uint32_t)C.af_ext;
            w_chains++;
        }
    }
    if (lane == 0) {
        WindowRec rec;
        rec.anchors = w_anchors;
        rec.seeds = w_chains ? (w_hi_qi - w_lo_qi + 1u) : 0u;
        rec.cov_q = w_covq; rec.cov_r = w_covr; rec.n_chains = w_chains;
        b.win_rec[slot] = rec;
    }
    __syncwarp();
    }
}

// ---- thread per window: the common case.
// A 20 kb window of a bacterial genome pair holds 10-200 anchors of which ~20 lie inside the 2 500 bp band of the current
// one.  One THREAD walks one window: per anchor it scores the predecessors still in band, nearest first, and stops at the
// first one outside - ~15 instructions per predecessor actually examined, against 42 warp-wide instructions per anchor
// (32 lanes, most of them idle or out of band) in the warp-per-window formulation above.
//   * the last DPT_RING anchors of every thread sit in a shared-memory ring laid out [slot][thread]; all threads of a
//     warp advance through their windows in lock step (same anchor index i), so ring slot (i - d) & 31 is the same row
//     for every lane and each LDS.128 is conflict free.  Ring entry: (q_pos, diagonal, meta, f | root << 16), where
//     diagonal = r_pos - q_pos on the forward strand and -(r_pos + q_pos) on the reverse strand, so that
//     dr - dq = diagonal(current) - diagonal(predecessor) on both strands: gap and dr cost three instructions;
//   * predecessors farther back than the ring (only while the anchor 32 back is still in band: repeat-dense stretches)
//     are read from the anchor records / f / root arrays;
//   * component roots need no pointer jumping: the predecessor precedes the anchor, so root[i] = root[pred] is final;
//   * per-component size and best end are kept in registers for the component of the previous anchor and spilled to the
//     per-anchor arrays when the component changes (chains rarely interleave).
// Windows with more than DPT_MAX_ANCHORS anchors (their f and root would not fit the 16-bit halves of the ring word, and
// one long window would stall the other 31 lanes) are appended to big_list for the warp-per-window kernel.
constexpr int DPT_THREADS = 64;
constexpr uint32_t DPT_RING = 32;
constexpr uint32_t DPT_MAX_ANCHORS = 512;
constexpr uint32_t DPT_BINS = DPT_MAX_ANCHORS / 4;        // windows are ordered by anchor count, 4 anchors per bin
constexpr uint32_t NO_ROOT = 0xFFFFFFFFu;

// ---- windows in descending order of their anchor count (counting sort over DPT_BINS bins): the 32 windows a warp
// walks in lock step then have nearly the same length (lanes idle 36 % of the time in slot order), and the longest
// windows start first.  bins[0..DPT_BINS) = histogram, then cursors; bins[DPT_BINS] = number of listed windows.
__global__ void window_bins_kernel(const ChainBatch b) {
    __shared__ uint32_t s_bins[DPT_BINS];
    for (uint32_t t = threadIdx.x; t < DPT_BINS; t += blockDim.x) s_bins[t] = 0u;
    __syncthreads();
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t A0 = 0, n = 0;
    if (slot < b.n_win_total && window_anchors(b, slot, A0, n) && n) {
        if (n > DPT_MAX_ANCHORS) b.big_list[atomicAdd(b.big_count, 1u)] = slot;
        else atomicAdd(&s_bins[(n - 1u) >> 2], 1u);
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < DPT_BINS; t += blockDim.x)
        if (s_bins[t]) atomicAdd(&b.win_bins[t], s_bins[t]);
}
__global__ void window_bins_scan_kernel(const ChainBatch b) {      // one warp: descending exclusive scan -> cursors
    const int lane = threadIdx.x;
    uint32_t run = 0;
    for (int base = (int)DPT_BINS - 32; base >= 0; base -= 32) {
        const uint32_t idx = (uint32_t)base + 31u - (uint32_t)lane;           // lane 0 takes the largest bin of the chunk
        const uint32_t c = b.win_bins[idx];
        uint32_t incl = c;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
        b.win_bins[idx] = run + incl - c;
        run += __shfl_sync(FULL, incl, 31);
    }
    if (lane == 0) b.win_bins[DPT_BINS] = run;
}
__global__ void __launch_bounds__(1024) window_order_kernel(const ChainBatch b) {
    // one global atomic per (CTA, bin) instead of one per window: the windows of a pair fall into few bins
    __shared__ uint32_t s_cnt[DPT_BINS], s_base[DPT_BINS];
    for (uint32_t t = threadIdx.x; t < DPT_BINS; t += blockDim.x) s_cnt[t] = 0u;
    __syncthreads();
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t A0 = 0, n = 0, local = 0, bin = 0;
    const bool listed = slot < b.n_win_total && window_anchors(b, slot, A0, n) && n && n <= DPT_MAX_ANCHORS;
    if (listed) { bin = (n - 1u) >> 2; local = atomicAdd(&s_cnt[bin], 1u); }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < DPT_BINS; t += blockDim.x)
        if (s_cnt[t]) s_base[t] = atomicAdd(&b.win_bins[t], s_cnt[t]);
    __syncthreads();
    if (listed) b.win_order[s_base[bin] + local] = slot;
}

// one predecessor from the ring against the current anchor; tq = q_pos - 1, best = max of f[j] - gap so far (> 0 to link)
__device__ __forceinline__ void dpt_pred(const uint4 p, const uint32_t tq, const int32_t cD, const uint32_t cm, const uint32_t d,
                                         int32_t& best, uint32_t& bestd) {
    const uint32_t t = tq - p.x;                           // dq - 1: in band and dq >= 1  <=>  t < band (unsigned)
    const int32_t delta = cD - (int32_t)p.y;               // dr - dq on either strand
    const int32_t gap = abs(delta);
    const int32_t sc = (int32_t)(p.w & 0xFFFFu) - gap;
    // dr = dq + delta > 0  <=>  t + delta >= 0
    if (t < DP_BP_BAND && p.z == cm && (int32_t)t + delta >= 0 && gap <= DP_MAX_GAP && sc > best) { best = sc; bestd = d; }
}

struct RootStat { uint32_t root, size, bidx; int32_t bf; };
__device__ __forceinline__ void stat_store(const RootStat& s, uint32_t* aux_a, unsigned long long* best_a) {
    if (s.root != NO_ROOT) {
        aux_a[s.root] = s.size;
        best_a[s.root] = ((unsigned long long)(uint32_t)s.bf << 32) | (0xFFFFFFFFu - s.bidx);
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr uint32_t DPT_STAGE = 4;          // anchor records in flight per thread (cp.async, global -> shared)

__global__ void __launch_bounds__(DPT_THREADS) chain_dp_thread_kernel(const ChainBatch b, const ChainConsts C) {
    __shared__ uint4 s_ring[DPT_RING][DPT_THREADS];
    __shared__ uint4 s_stage[DPT_STAGE][DPT_THREADS];     // the next anchor records of every thread, filled asynchronously
    constexpr uint32_t bp_band = DP_BP_BAND, index_band = DP_INDEX_BAND;
    constexpr int32_t max_gap = DP_MAX_GAP, anchor_score = DP_ANCHOR_SCORE;
    const int tid = threadIdx.x;
    const uint32_t rank = blockIdx.x * DPT_THREADS + tid;
    const uint32_t n_listed = b.win_bins[DPT_BINS];
    uint32_t slot = 0, A0 = 0, n = 0;
    if (rank < n_listed) { slot = b.win_order[rank]; window_anchors(b, slot, A0, n); }
    const uint32_t nmax = __reduce_max_sync(FULL, n);
    if (nmax == 0) return;
    const uint4* rec_a = b.a_rec + A0;
    // per anchor: f | root << 16 (read back only for predecessors beyond the ring; later the candidate list);
    // root_a: the list of component roots in order of creation
    uint32_t* fr_a = (uint32_t*)(b.a_f + A0); uint32_t* root_a = b.a_root + A0; uint32_t* aux_a = b.a_aux + A0;
    unsigned long long* best_a = b.a_best + A0;
    // per-component size and best end: the components of the last two distinct roots stay in registers (a chain and the
    // stray anchor that interrupts it), older ones are spilled to aux_a / best_a
    RootStat sa{NO_ROOT, 0u, 0u, 0}, sb{NO_ROOT, 0u, 0u, 0};
    uint32_t n_roots = 0;
    // anchor records arrive through a DPT_STAGE-deep cp.async pipeline: the data comes from DRAM (~1 us away) and a
    // register prefetch of one step is not far enough ahead; one commit group per step keeps the group count uniform
#pragma unroll
    for (uint32_t k = 0; k < DPT_STAGE; k++) {
        if (k < n) cp_async16(&s_stage[k][tid], rec_a + k);
        cp_async_commit();
    }
    cp_async_wait<DPT_STAGE - 1>();
    // the ring starts out full of entries that are out of band for every anchor of the window (q_pos 2^30 bases ahead):
    // the first anchors then run the same branch-free code as all others
    {
        const uint32_t q0 = n ? s_stage[0][tid].x : 0u;
        const uint4 far = make_uint4(q0 + 0x40000000u, 0u, 0xFFFFFFFFu, 0u);
#pragma unroll
        for (uint32_t k = 0; k < DPT_RING; k++) s_ring[k][tid] = far;
    }
    // ---------------- DP, lock step over the anchor index
    for (uint32_t i = 0; i < nmax; i++) {
        const bool act = i < n;
        cp_async_wait<DPT_STAGE - 1>();                    // record i has landed (groups are committed once per step)
        const uint4 r = s_stage[i & (DPT_STAGE - 1)][tid];
        if (i + DPT_STAGE < n) cp_async16(&s_stage[i & (DPT_STAGE - 1)][tid], rec_a + i + DPT_STAGE);
        cp_async_commit();
        const uint32_t cq = r.x, cm = r.z, tq = cq - 1u;
        const int32_t cD = (cm & 1u) ? -(int32_t)(r.y + r.x) : (int32_t)(r.y - r.x);
        int32_t best = 0;                 // max over valid predecessors of f[j] - gap; a link needs f[j] + 20 - gap > 20
        uint32_t bestd = 0;
        bool live = act;                  // predecessors are ordered by q_pos: once one is out of band, all older ones are
#pragma unroll 1
        for (uint32_t d0 = 1; d0 <= DPT_RING; d0 += 4) {
            if (!__any_sync(FULL, live)) break;
            const uint4 p0 = s_ring[(i - d0) & (DPT_RING - 1)][tid], p1 = s_ring[(i - d0 - 1) & (DPT_RING - 1)][tid];
            const uint4 p2 = s_ring[(i - d0 - 2) & (DPT_RING - 1)][tid], p3 = s_ring[(i - d0 - 3) & (DPT_RING - 1)][tid];
            dpt_pred(p0, tq, cD, cm, d0, best, bestd);
            dpt_pred(p1, tq, cD, cm, d0 + 1, best, bestd);
            dpt_pred(p2, tq, cD, cm, d0 + 2, best, bestd);
            dpt_pred(p3, tq, cD, cm, d0 + 3, best, bestd);
            live = live && (cq - p3.x) <= bp_band;
        }
        if (live && i > DPT_RING) {
            // the whole ring is in band (repeat-dense stretch): older predecessors from the arrays, up to the index band
            const uint32_t dend = min(i, index_band);
            for (uint32_t d = DPT_RING + 1; d <= dend; d++) {
                const uint4 p = rec_a[i - d];
                const uint32_t dq = cq - p.x;
                if (dq > bp_band) break;
                if (dq == 0u || p.z != cm) continue;
                const int32_t pD = (cm & 1u) ? -(int32_t)(p.y + p.x) : (int32_t)(p.y - p.x);
                const int32_t delta = cD - pD;
                const int32_t gap = abs(delta);
                const int32_t sc = (int32_t)(fr_a[i - d] & 0xFFFFu) - gap;
                if ((int32_t)dq + delta > 0 && gap <= max_gap && sc > best) { best = sc; bestd = d; }
            }
        }
        if (act) {
            const int32_t f = anchor_score + best;
            uint32_t root = i;
            if (bestd) root = (bestd <= DPT_RING ? s_ring[(i - bestd) & (DPT_RING - 1)][tid].w : fr_a[i - bestd]) >> 16;
            const uint32_t fr = (uint32_t)f | (root << 16);
            s_ring[i & (DPT_RING - 1)][tid] = make_uint4(cq, (uint32_t)cD, cm, fr);
            fr_a[i] = fr;
            if (root == i) root_a[n_roots++] = i;
            if (root != sa.root) {
                // the other cached component becomes the current one; a third one evicts the older entry
                const RootStat t = sa; sa = sb; sb = t;
                if (root != sa.root) {
                    stat_store(sa, aux_a, best_a);
                    if (root == i) sa = RootStat{i, 0u, i, INT32_MIN};
                    else {
                        const unsigned long long bb = best_a[root];
                        sa = RootStat{root, aux_a[root], 0xFFFFFFFFu - (uint32_t)bb, (int32_t)(bb >> 32)};
                    }
                }
            }
            sa.size++;
            if (f > sa.bf) { sa.bf = f; sa.bidx = i; }      // strict: the first maximal end in DP order
        }
    }
    cp_async_wait<0>();
    if (n == 0) return;
    stat_store(sa, aux_a, best_a);
    stat_store(sb, aux_a, best_a);

    // ---------------- candidate chains: roots with enough anchors and score; the list overwrites f (no longer needed)
    int32_t* f_a = (int32_t*)fr_a;
    uint32_t ncand = 0;
    for (uint32_t t = 0; t < n_roots; t++) {
        const uint32_t i = root_a[t];
        uint32_t size; int32_t score;
        if (i == sa.root) { size = sa.size; score = sa.bf; }            // the usual case: still in registers
        else if (i == sb.root) { size = sb.size; score = sb.bf; }
        else { size = aux_a[i]; score = (int32_t)(best_a[i] >> 32); }
        if (size >= (uint32_t)C.min_anchors && score >= C.min_score) f_a[ncand++] = (int32_t)i;
    }
    // ---------------- greedy selection without query overlap: (score desc, q start asc, r start asc, index asc)
    uint32_t w_anchors = 0, w_lo = 0xFFFFFFFFu, w_hi = 0, w_lo_qi = 0, w_hi_qi = 0, w_covq = 0, w_covr = 0, w_chains = 0;
    for (uint32_t round = 0; round < ncand; round++) {
        int32_t sc = INT32_MIN; uint32_t qs = 0xFFFFFFFFu, rs = 0xFFFFFFFFu, idx = 0xFFFFFFFFu;
        for (uint32_t t = 0; t < ncand; t++) {
            const uint32_t r2 = (uint32_t)f_a[t];
            if (aux_a[r2] & AUX_PROCESSED) continue;
            const unsigned long long bb = best_a[r2];
            const uint4 rr = rec_a[r2];
            const uint32_t rs2 = (rr.z & 1u) ? rec_a[0xFFFFFFFFu - (uint32_t)bb].y : rr.y;
            tuple_min(sc, qs, rs, idx, (int32_t)(bb >> 32), rr.x, rs2, r2);
        }
        const uint32_t rt = idx;
        const uint32_t bi = 0xFFFFFFFFu - (uint32_t)best_a[rt];
        const uint4 rec_r = rec_a[rt], rec_b = rec_a[bi];
        const uint32_t cqs = rec_r.x, cqe = rec_b.x;
        bool ov = false;
        for (uint32_t t = 0; t < ncand && !ov; t++) {
            const uint32_t r2 = (uint32_t)f_a[t];
            if (!(aux_a[r2] & AUX_ACCEPTED)) continue;
            const uint32_t bi2 = 0xFFFFFFFFu - (uint32_t)best_a[r2];
            ov = min(cqe, rec_a[bi2].x) >= max(cqs, rec_a[r2].x);
        }
        const uint32_t size = aux_a[rt] & AUX_SIZE;
        aux_a[rt] = size | AUX_PROCESSED | (ov ? 0u : AUX_ACCEPTED);
        if (!ov) {
            const bool rv = rec_r.z & 1u;
            const uint32_t crs = rv ? rec_b.y : rec_r.y, cre = rv ? rec_r.y : rec_b.y;
            w_anchors += size;
            if (cqs < w_lo) { w_lo = cqs; w_lo_qi = rec_r.w; }
            if (cqe >= w_hi) { w_hi = cqe; w_hi_qi = rec_b.w; }
            w_covq += (cqe - cqs) + (uint32_t)C.af_ext;
            w_covr += (cre - crs) + (uint32_t)C.af_ext;
            w_chains++;
        }
    }
    WindowRec rec;
    rec.anchors = w_anchors;
    rec.seeds = w_chains ? (w_hi_qi - w_lo_qi + 1u) : 0u;
    rec.cov_q = w_covq; rec.cov_r = w_covr; rec.n_chains = w_chains;
    b.win_rec[slot] = rec;
}

// ------------------------------------------------------------------ 5a. sort keys: pair << 32 | floor(ratio * 2^32)
__global__ void window_keys_kernel(const ChainBatch b) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= b.n_win_total) return;
    uint64_t key = (uint64_t)b.n_pairs << 32;                // unused / chainless windows sort behind every pair
    if (b.win_end[slot] > b.win_start[slot]) {
        const WindowRec rec = b.win_rec[slot];
        if (rec.n_chains && rec.seeds) {
            uint64_t rk = ((uint64_t)rec.anchors << 32) / rec.seeds;     // exact order of the rationals (seeds < 2^15)
            if (rk > 0xFFFFFFFFull) rk = 0xFFFFFFFFull;
            key = ((uint64_t)b.win_contig[slot] << 32) | rk;
        }
    }
    b.sort_keys[slot] = key;
    b.sort_vals[slot] = slot;
}

// ------------------------------------------------------------------ 5b. per-pair ANI / AF
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__global__ void ani_reduce_kernel(const ChainBatch b, const ChainConsts C, const uint64_t* __restrict__ keys,
                                  const uint32_t* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= b.n_pairs) return;
    const PairDesc pd = b.pairs[p];
    const GenomeView& Q = b.qviews[pd.q];
    const GenomeView& R = b.rviews[pd.r];
    PairResult res{-1.f, 0.f, 0.f, 0, 0, b.a_off[pd.seed_off + Q.n_seeds] - b.a_off[pd.seed_off]};
    const double inv_k = 1.0 / (double)C.k;
    double wsum = 0, ssum = 0, covq = 0, covr = 0, chains = 0, usum = 0;
    uint32_t n = 0;
    double median_ani = 0;
    if (keys == nullptr) {
        // seed-weighted mean (the default): the order of the windows does not matter, so the pair's window slots are read
        // in place and the key sort is skipped
        const uint32_t w0 = pd.win_off, w1 = pd.win_off + Q.win_cap;
        for (uint32_t t = w0 + lane; t < w1; t += 32) {
            if (b.win_end[t] <= b.win_start[t]) continue;
            const WindowRec rec = b.win_rec[t];
            if (!rec.n_chains || !rec.seeds) continue;
            n++;
            covq += rec.cov_q; covr += rec.cov_r; chains += rec.n_chains;
            double ratio = (double)rec.anchors / (double)rec.seeds;
            if (ratio > 1.0) ratio = 1.0;
            const double a = pow(ratio, inv_k);
            wsum += a * (double)rec.seeds;
            ssum += (double)rec.seeds;
            usum += a;
        }
        n = __reduce_add_sync(FULL, n);
    } else {
        // robust / median: windows in ascending order of anchors / seeds; segment of this pair in the sorted key array
        uint32_t lo = 0, hi = b.n_win_total;
        {
            uint32_t l = 0, h = b.n_win_total;
            const uint64_t t0 = (uint64_t)p << 32, t1 = (uint64_t)(p + 1) << 32;
            while (l < h) { uint32_t m = (l + h) >> 1; if (keys[m] < t0) l = m + 1; else h = m; }
            lo = l; h = b.n_win_total;
            while (l < h) { uint32_t m = (l + h) >> 1; if (keys[m] < t1) l = m + 1; else h = m; }
            hi = l;
        }
        n = hi - lo;
        if (n) {
            uint32_t s_lo = 0, s_hi = n;
            if (C.robust) { s_lo = n / 10; s_hi = n * 9 / 10; if (s_hi <= s_lo) { s_lo = 0; s_hi = n; } }
            for (uint32_t t = lane; t < n; t += 32) {
                const WindowRec rec = b.win_rec[vals[lo + t]];
                covq += rec.cov_q; covr += rec.cov_r; chains += rec.n_chains;
                if (t >= s_lo && t < s_hi) {
                    double ratio = (double)rec.anchors / (double)rec.seeds;
                    if (ratio > 1.0) ratio = 1.0;
                    wsum += pow(ratio, inv_k) * (double)rec.seeds;
                    ssum += (double)rec.seeds;
                }
            }
            if (C.median) {
                const WindowRec rec = b.win_rec[vals[lo + n / 2]];
                double ratio = (double)rec.anchors / (double)rec.seeds;
                if (ratio > 1.0) ratio = 1.0;
                median_ani = pow(ratio, inv_k);
            }
        }
    }
    res.n_windows = n;
    if (n) {
        wsum = warp_sum_f64(wsum); ssum = warp_sum_f64(ssum);
        covq = warp_sum_f64(covq); covr = warp_sum_f64(covr); chains = warp_sum_f64(chains);
        double ani = C.median ? median_ani : wsum / ssum;
        double afq = covq / (double)Q.total_len, afr = covr / (double)R.total_len;
        if (afq > 1.0) afq = 1.0;
        if (afr > 1.0) afr = 1.0;
        if (afq < C.frac_cover_cutoff && afr < C.frac_cover_cutoff) ani = -1.0;
        if (C.use_model && keys == nullptr && ani > 0.0 && covq >= C.learned_min_cov) {
            // learned-ANI correction: features of the pair -> gradient-boosted trees -> ANI in percent.
            // Feature order (DESIGN.md "learned ANI"): ANI %, standard deviation of the window ANIs %, reference contig-length
            // quantiles 90/50/10, query contig-length quantiles 90/50/10, mean aligned length per chain, aligned bases.
            const double mean_u = warp_sum_f64(usum) / (double)n;
            double dev = 0;
            for (uint32_t t = pd.win_off + lane; t < pd.win_off + Q.win_cap; t += 32) {
                if (b.win_end[t] <= b.win_start[t]) continue;
                const WindowRec rec = b.win_rec[t];
                if (!rec.n_chains || !rec.seeds) continue;
                double ratio = (double)rec.anchors / (double)rec.seeds;
                if (ratio > 1.0) ratio = 1.0;
                const double d = pow(ratio, inv_k) - mean_u;
                dev += d * d;
            }
            dev = warp_sum_f64(dev);
            float x[GBDT_FEATURES];
            x[0] = (float)(ani * 100.0); x[1] = (float)(sqrt(dev / (double)n) * 100.0);
            x[2] = (float)R.ctg_q90; x[3] = (float)R.ctg_q50; x[4] = (float)R.ctg_q10;
            x[5] = (float)Q.ctg_q90; x[6] = (float)Q.ctg_q50; x[7] = (float)Q.ctg_q10;
            x[8] = (float)(covq / chains); x[9] = (float)covq;
            // trees are evaluated 32 at a time (one per lane) and added in order, one f32 rounding per product and per sum,
            // exactly as gbdt-rs accumulates them
            float pred = C.model.bias;
            for (uint32_t t0 = 0; t0 < C.model.n_trees; t0 += 32) {
                const float v = t0 + lane < C.model.n_trees ? gbdt_tree(C.model, t0 + lane, x) : 0.f;
                const uint32_t cnt = min(32u, C.model.n_trees - t0);
                for (uint32_t j = 0; j < cnt; j++)
                    pred = __fadd_rn(pred, __fmul_rn(C.model.shrinkage, __shfl_sync(FULL, v, j)));
            }
            ani = (double)pred / 100.0;
            if (ani > 1.0) ani = 1.0;
            if (ani < 0.0) ani = 0.0;
        }
        res.ani = (float)ani; res.af_q = (float)afq; res.af_r = (float)afr;
        res.n_chains = (uint32_t)chains;
    }
    if (lane == 0) b.results[p] = res;
}

// rows of n_features f32 -> ensemble prediction, one thread per row (the same gbdt_tree / rounding as ani_reduce_kernel)
__global__ void gbdt_predict_kernel(const GbdtView m, const float* __restrict__ rows, uint32_t n_rows, uint32_t stride, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    float x[GBDT_FEATURES];
    for (uint32_t f = 0; f < GBDT_FEATURES; f++) x[f] = f < stride ? rows[(size_t)i * stride + f] : 0.f;
    out[i] = gbdt_predict(m, x);
}

}  // namespace

void launch_gbdt_predict(const GbdtView& m, const float* rows, uint32_t n_rows, uint32_t stride, float* out, cudaStream_t st) {
    if (n_rows == 0) return;
    gbdt_predict_kernel<<<(n_rows + 127) / 128, 128, 0, st>>>(m, rows, n_rows, stride, out);
    g_kernel_launches++;
}

void launch_match_count(const ChainBatch& b, cudaStream_t st) {
    if (b.n_pairs == 0) return;
    dim3 grid(32, b.n_pairs);
    match_count_kernel<<<grid, 256, 0, st>>>(b);
    g_kernel_launches++;
}
void launch_anchor_fill(const ChainBatch& b, cudaStream_t st) {
    if (b.n_pairs == 0) return;
    dim3 grid(32, b.n_pairs);
    anchor_fill_kernel<<<grid, 256, 0, st>>>(b);
    g_kernel_launches++;
}
constexpr size_t WALK_SMEM_MAX = 226 * 1024;     // dynamic part; the kernel also has a few bytes of static shared memory
size_t walk_smem_bytes(uint32_t n_seeds, uint32_t group) {
    return ((size_t)n_seeds + 8) * 4 + (size_t)group * ((((size_t)n_seeds + 31) / 32 + 3) / 4 * 4) * 4 + 64;
}
// most pairs of one query that fit beside its positions in the 227 KB of one CTA (at most one chain per warp)
uint32_t walk_group_capacity(uint32_t max_query_seeds) {
    const uint32_t n = max_query_seeds < WALK_SMEM_SEEDS ? max_query_seeds : WALK_SMEM_SEEDS;
    uint32_t g = 1;
    while (g < 32 && walk_smem_bytes(n, g + 1) <= WALK_SMEM_MAX) g++;
    return g;
}
void launch_window_walk(const ChainBatch& b, const ChainConsts& c, uint32_t max_query_seeds, cudaStream_t st) {
    if (b.n_pairs == 0) return;
    // shared-memory walk for every pair whose query fits, global-memory walk for the rest
    // per launch, not once per process: the attribute belongs to the current device
    cudaFuncSetAttribute(window_walk_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WALK_SMEM_MAX);
    const uint32_t n = max_query_seeds < WALK_SMEM_SEEDS ? max_query_seeds : WALK_SMEM_SEEDS;
    const size_t bytes = walk_smem_bytes(n, b.walk_group_max);
    window_walk_smem_kernel<<<b.n_walk_groups, 1024, bytes, st>>>(b, c.fragment_length);
    g_kernel_launches++;
    if (max_query_seeds > WALK_SMEM_SEEDS) {
        window_walk_kernel<<<b.n_pairs, 128, 0, st>>>(b, c.fragment_length, 1);
        g_kernel_launches++;
    }
}
void launch_chain_dp(const ChainBatch& b, const ChainConsts& c, int n_sm, cudaStream_t st) {
    if (b.n_win_total == 0) return;
    // thread per window for everything up to DPT_MAX_ANCHORS anchors; it lists the larger windows, which a persistent grid
    // of warp-per-window CTAs then takes (the list length is only known on the device)
    window_bins_kernel<<<(b.n_win_total + 255) / 256, 256, 0, st>>>(b);
    window_bins_scan_kernel<<<1, 32, 0, st>>>(b);
    window_order_kernel<<<(b.n_win_total + 1023) / 1024, 1024, 0, st>>>(b);
    chain_dp_thread_kernel<<<(b.n_win_total + DPT_THREADS - 1) / DPT_THREADS, DPT_THREADS, 0, st>>>(b, c);
    g_kernel_launches += 4;
    chain_dp_kernel<<<n_sm * 4, DP_WARPS * 32, 0, st>>>(b, c);
    g_kernel_launches++;
}
void launch_window_keys(const ChainBatch& b, cudaStream_t st) {
    if (b.n_win_total == 0) return;
    window_keys_kernel<<<(b.n_win_total + 255) / 256, 256, 0, st>>>(b);
    g_kernel_launches++;
}
void launch_ani_reduce(const ChainBatch& b, const ChainConsts& c, const uint64_t* sorted_keys,
                       const uint32_t* sorted_vals, cudaStream_t st) {
    if (b.n_pairs == 0) return;
    ani_reduce_kernel<<<(b.n_pairs + 3) / 4, 128, 0, st>>>(b, c, sorted_keys, sorted_vals);
    g_kernel_launches++;
}

}  // namespace skb
