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
__global__ void match_count_kernel(const ChainBatch b) {
    const PairDesc pd = b.pairs[blockIdx.y];
    const GenomeView& Q = b.qviews[pd.q];
    const GenomeView& R = b.rviews[pd.r];
    const uint32_t nq = Q.n_seeds;
    const int lane = threadIdx.x & 31;
    unsigned long long total = 0;       // 64-bit anchor count of this thread's seeds (a_off is a 32-bit scan and may wrap)
    // warp-aligned strips of 32 consecutive query seeds: the strip's "has a match" bits are one word of the bitmask
    for (uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; i0 < nq; i0 += gridDim.x * blockDim.x) {
        const uint32_t i = i0 + lane;
        uint32_t first = 0, cnt = 0;
        if (i < nq && R.n_seeds) {
            const uint32_t km = __ldg(Q.kmer_p + i);
            const uint32_t bk = km >> R.bucket_shift;
            uint32_t lo = __ldg(R.bucket + bk), hi = __ldg(R.bucket + bk + 1);
            const uint32_t end = hi;
            while (lo < hi) {               // buckets hold ~8 seeds: a short search
                uint32_t mid = (lo + hi) >> 1;
                if (__ldg(R.kmer_k + mid) < km) lo = mid + 1; else hi = mid;
            }
            first = lo;
            while (lo < end && __ldg(R.kmer_k + lo) == km) lo++;
            cnt = lo - first;
        }
        if (i < nq) {
            b.m_first[pd.seed_off + i] = first;
            b.m_cnt[pd.seed_off + i] = cnt;
        }
        total += cnt;
        const uint32_t bal = __ballot_sync(FULL, cnt != 0);
        if (lane == 0) b.m_bits[pd.bits_off + (i0 >> 5)] = bal;
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
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) {
        const uint32_t cnt = b.m_cnt[pd.seed_off + i];
        if (cnt == 0) continue;
        const uint32_t first = b.m_first[pd.seed_off + i];
        const uint32_t off = b.a_off[pd.seed_off + i];
        const uint32_t qp = __ldg(Q.pos_p + i);
        const uint32_t qm = __ldg(Q.meta_p + i);
        for (uint32_t m = 0; m < cnt; m++) {
            const uint32_t o = off + m;
            if (o >= b.anchor_cap) break;
            const uint32_t rm = __ldg(R.meta_k + first + m);
            b.a_qi[o] = i;
            b.a_qp[o] = qp;
            b.a_rp[o] = __ldg(R.pos_k + first + m);
            b.a_meta[o] = (rm & ~1u) | ((qm ^ rm) & 1u);    // ref contig << 1 | reverse_match
            b.a_aux[o] = 0u;                                // component size / flags and best end, accumulated by the DP kernel
            b.a_best[o] = 0ull;
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
__device__ __forceinline__ void dp_window(const uint32_t* __restrict__ qp_a, const uint32_t* __restrict__ rp_a,
                                          const uint32_t* __restrict__ meta_a, int32_t* f_a, uint32_t* root_a,
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
        if (mine < n) { sq = qp_a[mine]; sr = rp_a[mine]; sm = meta_a[mine]; }
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
                        const uint32_t pq = qp_a[j];
                        inband = (cq - pq) <= bp_band;
                        if (inband && meta_a[j] == cm) {
                            const uint32_t pr = rp_a[j];
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

__global__ void __launch_bounds__(DP_WARPS * 32) chain_dp_kernel(const ChainBatch b, const ChainConsts C) {
    __shared__ uint4 s_strip[DP_WARPS][32];          // the strip of 32 anchors a warp is working on; later its candidate list
    __shared__ __align__(16) uint32_t s_work[DP_WARPS][4 * DP_SMEM_ANCHORS];   // root | size+flags | best end (64 bit)
    const int lane = threadIdx.x & 31;
    const uint32_t slot = blockIdx.x * DP_WARPS + (threadIdx.x >> 5);
    if (slot >= b.n_win_total) return;
    const uint32_t ws = b.win_start[slot], we = b.win_end[slot];
    if (we <= ws) return;                                   // unused slot
    const PairDesc pd = b.pairs[b.win_contig[slot]];
    const uint32_t A0 = b.a_off[pd.seed_off + ws], A1 = b.a_off[pd.seed_off + we];
    if (A1 > b.anchor_cap) return;                          // anchor arrays were sized too small: the host reruns the batch
    const uint32_t n = A1 - A0;
    const uint32_t* qp_a = b.a_qp + A0; const uint32_t* rp_a = b.a_rp + A0; const uint32_t* meta_a = b.a_meta + A0;
    int32_t* f_a = b.a_f + A0;
    // Per-anchor working state of the chain phases (component root, size | flags, best end).  Windows of up to
    // DP_SMEM_ANCHORS anchors - nearly all of them - keep it in shared memory: the phases after the DP are chains of
    // dependent scattered accesses, which cost ~30 cycles there against an L2 round trip each in the global arrays.
    const int warp = threadIdx.x >> 5;
    const bool small = n <= DP_SMEM_ANCHORS;
    uint32_t* root_a = small ? s_work[warp] : b.a_root + A0;
    uint32_t* aux_a = small ? s_work[warp] + DP_SMEM_ANCHORS : b.a_aux + A0;
    unsigned long long* best_a = small ? (unsigned long long*)(s_work[warp] + 2 * DP_SMEM_ANCHORS) : b.a_best + A0;
    if (small) {
        for (uint32_t i = lane; i < n; i += 32) { aux_a[i] = 0u; best_a[i] = 0ull; }      // the global arrays come zeroed
        __syncwarp();
    }
    // candidate list: at most n / min_anchors entries
    int32_t* cand_a = (small && n <= 128u * (uint32_t)max(C.min_anchors, 1)) ? (int32_t*)s_strip[warp] : f_a;

    // ---------------- DP
    if ((uint64_t)n * (uint32_t)DP_ANCHOR_SCORE < (1u << 22)) dp_window<false>(qp_a, rp_a, meta_a, f_a, root_a, n, lane, s_strip[warp]);
    else dp_window<true>(qp_a, rp_a, meta_a, f_a, root_a, n, lane, s_strip[warp]);

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
            const bool rv = meta_a[r] & 1u;
            const uint32_t rs2 = rv ? rp_a[bi] : rp_a[r];
            tuple_min(sc, qs, rs, idx, (int32_t)(bb >> 32), qp_a[r], rs2, r);
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
        const uint32_t cqs = qp_a[r], cqe = qp_a[bi];
        bool ov = false;
        for (uint32_t t = lane; t < ncand; t += 32) {
            const uint32_t r2 = (uint32_t)cand_a[t];
            if (!(ldv(&aux_a[r2]) & AUX_ACCEPTED)) continue;
            const uint32_t bi2 = 0xFFFFFFFFu - (uint32_t)ldv(&best_a[r2]);
            const uint32_t lo = max(cqs, qp_a[r2]), hi = min(cqe, qp_a[bi2]);
            ov |= hi >= lo;
        }
        ov = __any_sync(FULL, ov);
        const uint32_t size = ldv(&aux_a[r]) & AUX_SIZE;
        if (lane == 0) aux_a[r] = size | AUX_PROCESSED | (ov ? 0u : AUX_ACCEPTED);
        __threadfence_block();
        __syncwarp();
        if (!ov) {
            const bool rv = meta_a[r] & 1u;
            const uint32_t crs = rv ? rp_a[bi] : rp_a[r], cre = rv ? rp_a[r] : rp_a[bi];
            w_anchors += size;
            if (cqs < w_lo) { w_lo = cqs; w_lo_qi = b.a_qi[A0 + r]; }
            if (cqe >= w_hi) { w_hi = cqe; w_hi_qi = b.a_qi[A0 + bi]; }
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
void launch_chain_dp(const ChainBatch& b, const ChainConsts& c, cudaStream_t st) {
    if (b.n_win_total == 0) return;
    chain_dp_kernel<<<(b.n_win_total + DP_WARPS - 1) / DP_WARPS, DP_WARPS * 32, 0, st>>>(b, c);
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
