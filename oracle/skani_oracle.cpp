// skani_oracle.cpp — CPU ORACLE (test infrastructure, NOT the product path).
// See skani_oracle.h for provenance, scope and the "who may load this" rule.
//
// Restates, in plain scalar C++:
//   fmh_seeds              (skani v0.3.0 seeding.rs; call site reference lib.rs:165-171)
//   check_markers_quickly  (skani v0.3.0 screen.rs;  call site reference lib.rs:623-628)
//   chain_seeds            (skani v0.3.0 chain.rs;   call site reference lib.rs:652-653)
// following SURVEY.md Appendix A (the crate's sources are not on this machine).
#include "skani_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int MARKER_K = 21;            // K_MARKER_DNA
constexpr uint64_t MIN_LENGTH_CONTIG = 500;  // skani::params::MIN_LENGTH_CONTIG (reference lib.rs:156)

// skani's mm_hash64 is written in Rust as `key = !key.wrapping_add(key << 21);`.  A method call binds
// tighter than unary `!`, so the first step evaluates to ~(key + (key << 21)) — NOT Thomas Wang's
// (~key) + (key << 21).  This form reproduces the seed/marker counts probed in SURVEY.md Appendix B
// (37,237 / 4,539 for EC590 and 37,384 / 4,551 for K-12); the textbook form gives 37,111 / 4,437.
inline uint64_t mm_hash64(uint64_t x) {
    x = ~(x + (x << 21));
    x ^= x >> 24;
    x = x + (x << 3) + (x << 8);
    x ^= x >> 14;
    x = x + (x << 2) + (x << 4);
    x ^= x >> 28;
    x = x + (x << 31);
    return x;
}

struct Seed {
    uint64_t kmer;
    uint32_t pos;      // index of the LAST base of the k-mer
    uint32_t contig;   // index among kept contigs
    uint8_t canonical; // forward k-mer < reverse-complement k-mer
};

struct Lut {
    uint8_t t[256];
    Lut() {
        std::memset(t, 0, sizeof t);  // everything that is not ACGT/acgt encodes as 0
        t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3;
    }
};
const Lut LUT;

}  // namespace

struct orc_sketch {
    int k, c, marker_c;
    bool has_seeds;
    uint64_t total_len = 0;
    std::vector<uint32_t> contig_lengths;
    std::vector<Seed> seeds;        // sorted (kmer, contig, pos)
    std::vector<uint64_t> markers;  // sorted unique
    std::vector<std::vector<uint32_t>> pos_by_contig;  // sorted seed positions per contig
};

namespace {

// fmh_seeds for one contig
void fmh_seeds(const uint8_t* s, uint64_t len, int k, int c, int marker_c, uint32_t contig_index,
               bool seed, std::vector<Seed>& seeds, std::vector<uint64_t>& markers) {
    if (len < (uint64_t)MARKER_K) return;
    const uint64_t mask_m = (~0ULL) >> (64 - 2 * MARKER_K);
    const uint64_t mask_k = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    const int rshift = 2 * (MARKER_K - k);
    const uint64_t thr_seed = UINT64_MAX / (uint64_t)c;
    const uint64_t thr_marker = UINT64_MAX / (uint64_t)marker_c;
    uint64_t f = 0, r = 0;
    for (uint64_t i = 0; i < (uint64_t)MARKER_K - 1; i++) {
        uint64_t b = LUT.t[s[i]];
        f = ((f << 2) | b) & mask_m;
        r = (r >> 2) | ((3 - b) << (2 * MARKER_K - 2));
    }
    for (uint64_t i = MARKER_K - 1; i < len; i++) {
        uint64_t b = LUT.t[s[i]];
        f = ((f << 2) | b) & mask_m;
        r = (r >> 2) | ((3 - b) << (2 * MARKER_K - 2));
        uint64_t fk = f & mask_k, rk = r >> rshift;
        bool canon = fk < rk;
        uint64_t km = canon ? fk : rk;
        if (seed && mm_hash64(km) < thr_seed)
            seeds.push_back(Seed{km, (uint32_t)i, contig_index, (uint8_t)canon});
        uint64_t mk = f < r ? f : r;
        if (mm_hash64(mk) < thr_marker) markers.push_back(mk);
    }
}

struct Anchor {
    uint32_t qc, qp, rc, rp;
    uint8_t rev;
};

struct Chain {
    double score;
    int n_anchors;
    uint32_t qs, qe, rs, re;
    uint32_t qc, rc;
    int chunk;
    std::vector<int> path;  // anchor indices (global), root..best
};

inline uint64_t overlap_len(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    uint32_t lo = std::max(a0, b0), hi = std::min(a1, b1);
    return hi >= lo ? (uint64_t)(hi - lo) + 1 : 0;
}

inline uint64_t count_in(const std::vector<uint32_t>& v, uint32_t lo, uint32_t hi) {
    if (hi < lo) return 0;
    auto a = std::lower_bound(v.begin(), v.end(), lo);
    auto b = std::upper_bound(v.begin(), v.end(), hi);
    return (uint64_t)(b - a);
}

struct WindowAcc {
    uint64_t anchors = 0;
    uint32_t lo = UINT32_MAX, hi = 0;
    uint64_t seeds_in_chains = 0;
    uint32_t a_lo = UINT32_MAX, a_hi = 0;  // first/last anchor (all anchors) of the window
    uint32_t qc = 0;
    bool any = false;
};

thread_local std::vector<double>* g_dump = nullptr;  // debug sink used by orc_chain_dump (fit tooling only)

void chain_impl(const orc_sketch* R0, const orc_sketch* Q0, const orc_chain_params_t& P, orc_result_t* out) {
    std::memset(out, 0, sizeof *out);
    out->ani = -1.f; out->ani_f64 = -1.0;
    bool sw = false;
    if (P.switch_mode == 1) sw = true;
    else if (P.switch_mode == 2) sw = R0->total_len < Q0->total_len;
    const orc_sketch* R = sw ? Q0 : R0;
    const orc_sketch* Q = sw ? R0 : Q0;
    out->switched = sw;
    const int k = R->k;

    // ---- anchors: join on equal k-mer, all position pairs (A.6 step 2) ----
    std::vector<Anchor> A;
    {
        size_t i = 0, j = 0;
        const auto& qs = Q->seeds; const auto& rs = R->seeds;
        while (i < qs.size() && j < rs.size()) {
            if (qs[i].kmer < rs[j].kmer) i++;
            else if (qs[i].kmer > rs[j].kmer) j++;
            else {
                uint64_t km = qs[i].kmer;
                size_t i1 = i, j1 = j;
                while (i1 < qs.size() && qs[i1].kmer == km) i1++;
                while (j1 < rs.size() && rs[j1].kmer == km) j1++;
                for (size_t a = i; a < i1; a++)
                    for (size_t b = j; b < j1; b++)
                        A.push_back(Anchor{qs[a].contig, qs[a].pos, rs[b].contig, rs[b].pos,
                                           (uint8_t)(qs[a].canonical != rs[b].canonical)});
                i = i1; j = j1;
            }
        }
    }
    out->n_anchors = (int64_t)A.size();
    if (A.empty()) return;
    std::sort(A.begin(), A.end(), [](const Anchor& a, const Anchor& b) {
        if (a.qc != b.qc) return a.qc < b.qc;
        if (a.qp != b.qp) return a.qp < b.qp;
        if (a.rc != b.rc) return a.rc < b.rc;
        if (a.rp != b.rp) return a.rp < b.rp;
        return a.rev < b.rev;
    });

    // ---- chunks ----
    const uint32_t F = (uint32_t)P.fragment_length;
    std::vector<int> chunk_begin;  // index into A
    {
        uint32_t cur_c = UINT32_MAX, cur_w = UINT32_MAX, start = 0;
        for (size_t i = 0; i < A.size(); i++) {
            bool open = false;
            if (A[i].qc != cur_c) open = true;
            else if (P.chunk_mode == 0) open = (A[i].qp / F) != cur_w;
            else if (P.chunk_mode == 1) open = A[i].qp >= start + F;
            if (open) {
                chunk_begin.push_back((int)i);
                cur_c = A[i].qc; cur_w = A[i].qp / F; start = A[i].qp;
            }
        }
        chunk_begin.push_back((int)A.size());
    }
    const int n_chunks = (int)chunk_begin.size() - 1;

    // ---- DP + chains per chunk ----
    std::vector<Chain> kept;
    std::vector<double> f;
    std::vector<int> ptr, comp, ord;
    for (int ch = 0; ch < n_chunks; ch++) {
        const int b = chunk_begin[ch], e = chunk_begin[ch + 1], n = e - b;
        ord.resize(n);
        std::iota(ord.begin(), ord.end(), b);
        if (P.order_by_ref)
            std::sort(ord.begin(), ord.end(), [&](int x, int y) {
                const Anchor &a = A[x], &c2 = A[y];
                if (a.rc != c2.rc) return a.rc < c2.rc;
                if (a.rp != c2.rp) return a.rp < c2.rp;
                if (a.qc != c2.qc) return a.qc < c2.qc;
                if (a.qp != c2.qp) return a.qp < c2.qp;
                return a.rev < c2.rev;
            });
        f.assign(n, 0.0); ptr.assign(n, 0);
        for (int i = 0; i < n; i++) {
            const Anchor& ci = A[ord[i]];
            double best = P.anchor_score; int bp = i;
            int jlo = std::max(0, i - P.index_band);
            for (int j = i - 1; j >= jlo; j--) {
                const Anchor& cj = A[ord[j]];
                int64_t dprim = P.order_by_ref ? (cj.rc == ci.rc ? (int64_t)ci.rp - cj.rp : INT64_MAX)
                                               : (int64_t)ci.qp - cj.qp;
                if (dprim > P.bp_band) break;
                if (cj.rc != ci.rc || cj.qc != ci.qc || cj.rev != ci.rev) continue;
                int64_t dq, dr;
                if (P.order_by_ref) {
                    dr = (int64_t)ci.rp - cj.rp;
                    dq = ci.rev ? (int64_t)cj.qp - ci.qp : (int64_t)ci.qp - cj.qp;
                } else {
                    dq = (int64_t)ci.qp - cj.qp;
                    dr = ci.rev ? (int64_t)cj.rp - ci.rp : (int64_t)ci.rp - cj.rp;
                }
                if (P.strict_dr ? (dq <= 0 || dr <= 0) : (dq < 0 || dr < 0)) continue;
                double gap = (double)std::llabs(dr - dq);
                if (gap > P.max_gap) continue;
                double sc = f[j] + P.anchor_score - gap;
                if (sc > best) { best = sc; bp = j; }
            }
            f[i] = best; ptr[i] = bp;
        }
        // components of the pointer forest: root id
        comp.assign(n, 0);
        for (int i = 0; i < n; i++) comp[i] = (ptr[i] == i) ? i : comp[ptr[i]];
        // best end per component (strict > : first maximal in DP order)
        std::vector<int> best_of(n, -1), comp_size(n, 0);
        for (int i = 0; i < n; i++) {
            int r = comp[i];
            comp_size[r]++;
            if (best_of[r] < 0 || f[i] > f[best_of[r]]) best_of[r] = i;
        }
        std::vector<Chain> cand;
        for (int r = 0; r < n; r++) {
            if (best_of[r] < 0) continue;
            int bi = best_of[r];
            Chain c; c.score = f[bi]; c.chunk = ch;
            int cur = bi;
            while (true) { c.path.push_back(ord[cur]); if (ptr[cur] == cur) break; cur = ptr[cur]; }
            std::reverse(c.path.begin(), c.path.end());
            c.n_anchors = P.count_mode == 1 ? comp_size[r] : (int)c.path.size();
            if ((int)c.path.size() < P.min_anchors && c.n_anchors < P.min_anchors) continue;
            if (c.n_anchors < P.min_anchors || c.score < P.min_score) continue;
            uint32_t qs = UINT32_MAX, qe = 0, rs = UINT32_MAX, re = 0;
            if (P.count_mode == 2) {  // extents of the whole component (rejected by the golden fit)
                for (int i2 = 0; i2 < n; i2++) if (comp[i2] == r) {
                    const Anchor& a2 = A[ord[i2]];
                    qs = std::min(qs, a2.qp); qe = std::max(qe, a2.qp);
                    rs = std::min(rs, a2.rp); re = std::max(re, a2.rp);
                }
            } else
            for (int ai : c.path) {
                qs = std::min(qs, A[ai].qp); qe = std::max(qe, A[ai].qp);
                rs = std::min(rs, A[ai].rp); re = std::max(re, A[ai].rp);
            }
            c.qs = qs; c.qe = qe; c.rs = rs; c.re = re;
            c.qc = A[c.path[0]].qc; c.rc = A[c.path[0]].rc;
            cand.push_back(std::move(c));
        }
        std::stable_sort(cand.begin(), cand.end(), [](const Chain& a, const Chain& b) {
            if (a.score != b.score) return a.score > b.score;
            if (a.qs != b.qs) return a.qs < b.qs;
            return a.rs < b.rs;
        });
        size_t first_kept = kept.size();
        for (auto& c : cand) {
            bool ok = true;
            for (size_t t = first_kept; t < kept.size() && ok; t++) {
                const Chain& o = kept[t];
                if (P.overlap_side == 0 || P.overlap_side == 2) {
                    if (o.qc == c.qc) {
                        uint64_t ov = overlap_len(o.qs, o.qe, c.qs, c.qe);
                        uint64_t sh = std::min<uint64_t>(o.qe - o.qs, c.qe - c.qs) + 1;
                        if ((double)ov > P.overlap_tol * (double)sh) ok = false;
                    }
                }
                if (ok && (P.overlap_side == 1 || P.overlap_side == 2)) {
                    if (o.rc == c.rc) {
                        uint64_t ov = overlap_len(o.rs, o.re, c.rs, c.re);
                        uint64_t sh = std::min<uint64_t>(o.re - o.rs, c.re - c.rs) + 1;
                        if ((double)ov > P.overlap_tol * (double)sh) ok = false;
                    }
                }
            }
            if (ok) kept.push_back(std::move(c));
        }
    }
    out->n_chains = (int64_t)kept.size();
    if (g_dump)
        for (const Chain& c : kept) {
            double row[10] = {(double)c.chunk, (double)c.qs, (double)c.qe, (double)c.rs, (double)c.re, c.score,
                              (double)c.n_anchors, (double)c.path.size(), (double)A[c.path[0]].rev,
                              (double)A[chunk_begin[c.chunk]].qp};
            g_dump->insert(g_dump->end(), row, row + 10);
        }
    if (kept.empty()) return;

    // ---- per-window accumulation ----
    // window key: chunk id for modes 0/1; (qc, qp / F) for mode 2
    std::vector<WindowAcc> W;
    std::vector<uint64_t> wkey;
    auto win_of = [&](uint64_t key) -> WindowAcc& {
        // keys arrive grouped but not necessarily sorted: linear probe from the back is enough
        for (size_t t = wkey.size(); t-- > 0;) if (wkey[t] == key) return W[t];
        wkey.push_back(key); W.emplace_back();
        return W.back();
    };
    if (P.chunk_mode == 2) {
        std::sort(kept.begin(), kept.end(), [](const Chain& a, const Chain& b) {
            if (a.qc != b.qc) return a.qc < b.qc;
            return a.qs < b.qs;
        });
    }
    for (const Chain& c : kept) {
        if (P.chunk_mode != 2) {
            WindowAcc& w = win_of((uint64_t)c.chunk);
            w.any = true; w.qc = c.qc;
            w.anchors += (uint64_t)c.n_anchors;
            w.lo = std::min(w.lo, c.qs); w.hi = std::max(w.hi, c.qe);
            w.seeds_in_chains += count_in(Q->pos_by_contig[c.qc], c.qs, c.qe);
        } else {
            // split the path over fixed windows of the query contig
            std::vector<std::pair<uint32_t, WindowAcc>> parts;
            for (int ai : c.path) {
                uint32_t wid = A[ai].qp / F;
                WindowAcc* pw = nullptr;
                for (auto& pr : parts) if (pr.first == wid) pw = &pr.second;
                if (!pw) { parts.emplace_back(wid, WindowAcc()); pw = &parts.back().second; }
                pw->anchors++; pw->lo = std::min(pw->lo, A[ai].qp); pw->hi = std::max(pw->hi, A[ai].qp);
            }
            for (auto& pr : parts) {
                WindowAcc& w = win_of(((uint64_t)c.qc << 32) | pr.first);
                w.any = true; w.qc = c.qc;
                w.anchors += pr.second.anchors;
                w.lo = std::min(w.lo, pr.second.lo); w.hi = std::max(w.hi, pr.second.hi);
                w.seeds_in_chains += count_in(Q->pos_by_contig[c.qc], pr.second.lo, pr.second.hi);
            }
        }
    }
    if (P.denom_mode == 3 && P.chunk_mode != 2) {
        for (int ch = 0; ch < n_chunks; ch++) {
            for (size_t t = 0; t < wkey.size(); t++) if (wkey[t] == (uint64_t)ch) {
                W[t].a_lo = A[chunk_begin[ch]].qp; W[t].a_hi = A[chunk_begin[ch + 1] - 1].qp;
            }
        }
    }

    struct Est { double ani; double anchors; double seeds; };
    std::vector<Est> ests;
    for (size_t t = 0; t < W.size(); t++) {
        const WindowAcc& w = W[t];
        if (!w.any || (int64_t)w.anchors < P.min_window_anchors) continue;
        uint64_t seeds;
        const auto& qp = Q->pos_by_contig[w.qc];
        switch (P.denom_mode) {
            case 1: seeds = w.seeds_in_chains; break;
            case 2: {
                uint32_t wid = (P.chunk_mode == 2) ? (uint32_t)(wkey[t] & 0xffffffffu) : w.lo / F;
                seeds = count_in(qp, wid * F, wid * F + F - 1);
            } break;
            case 3: seeds = count_in(qp, w.a_lo, w.a_hi); break;
            default: seeds = count_in(qp, w.lo, w.hi);
        }
        if (seeds == 0) continue;
        double ratio = (double)w.anchors / (double)seeds;
        if (ratio > 1.0) ratio = 1.0;
        ests.push_back(Est{std::pow(ratio, 1.0 / (double)k), (double)w.anchors, (double)seeds});
    }
    out->n_windows = (int64_t)ests.size();
    if (ests.empty()) return;
    std::stable_sort(ests.begin(), ests.end(), [](const Est& a, const Est& b) { return a.ani < b.ani; });

    // ---- AF ----
    double cov_q = 0, cov_r = 0;
    if (P.af_mode == 0) {
        for (const Chain& c : kept) { cov_q += (double)(c.qe - c.qs) + P.af_ext; cov_r += (double)(c.re - c.rs) + P.af_ext; }
    } else {
        for (const WindowAcc& w : W) if (w.any) cov_q += (double)(w.hi - w.lo) + P.af_ext;
        for (const Chain& c : kept) cov_r += (double)(c.re - c.rs) + P.af_ext;
    }
    double af_q = std::min(1.0, cov_q / (double)Q->total_len);
    double af_r = std::min(1.0, cov_r / (double)R->total_len);

    // ---- aggregate ANI (A.6 step 6) ----
    size_t n = ests.size(), lo = 0, hi = n;
    if (P.robust) { lo = n / 10; hi = n * 9 / 10; if (hi <= lo) { lo = 0; hi = n; } }
    double ani;
    if (P.median) ani = ests[n / 2].ani;
    else {
        double sa = 0, ss = 0, su = 0, sw2 = 0;
        for (size_t i = lo; i < hi; i++) { sa += ests[i].anchors; ss += ests[i].seeds; su += ests[i].ani; sw2 += ests[i].ani * ests[i].seeds; }
        if (P.mean_mode == 0) ani = su / (double)(hi - lo);
        else if (P.mean_mode == 1) ani = sw2 / ss;
        else ani = std::pow(std::min(1.0, sa / ss), 1.0 / (double)k);
    }
    if (af_q < P.frac_cover_cutoff && af_r < P.frac_cover_cutoff) ani = -1.0;
    // ---- feature vector of the learned-ANI regression (skani::regression; order documented in DESIGN.md) ----
    {
        double su = 0;
        for (const Est& e : ests) su += e.ani;
        const double mean_u = su / (double)n;
        double dev = 0;
        for (const Est& e : ests) dev += (e.ani - mean_u) * (e.ani - mean_u);
        auto quant = [](const std::vector<uint32_t>& lens, int q) -> float {
            if (lens.empty()) return 0.f;
            std::vector<uint32_t> v(lens);
            std::sort(v.begin(), v.end());
            return (float)v[(v.size() - 1) * (size_t)q / 100];
        };
        float* x = out->features;
        x[0] = (float)(ani * 100.0); x[1] = (float)(std::sqrt(dev / (double)n) * 100.0);
        x[2] = quant(R->contig_lengths, 90); x[3] = quant(R->contig_lengths, 50); x[4] = quant(R->contig_lengths, 10);
        x[5] = quant(Q->contig_lengths, 90); x[6] = quant(Q->contig_lengths, 50); x[7] = quant(Q->contig_lengths, 10);
        x[8] = (float)(cov_q / (double)kept.size()); x[9] = (float)cov_q;
    }
    if (sw) std::swap(af_q, af_r);
    out->ani_f64 = ani; out->af_query_f64 = af_q; out->af_ref_f64 = af_r;
    out->ani = (float)ani; out->af_query = (float)af_q; out->af_ref = (float)af_r;
}

}  // namespace

extern "C" {

uint64_t orc_mm_hash64(uint64_t x) { return mm_hash64(x); }

void orc_chain_params_default(orc_chain_params_t* p) {
    std::memset(p, 0, sizeof *p);
    p->fragment_length = 20000;
    p->anchor_score = 20.0;
    p->min_anchors = 3;
    p->min_score = 45.0;
    p->max_gap = 300.0;
    p->index_band = 100;       // D_CHAIN_BAND
    p->bp_band = 2500;
    p->frac_cover_cutoff = 0.15;
    p->robust = 0; p->median = 0;
    // Structural choices frozen by the golden fit (oracle/fit_goldens.py, DESIGN.md "Oracle"):
    p->chunk_mode = 1;         // a window opens at its first anchor and covers fragment_length bp
    p->order_by_ref = 0;       // DP in (q_contig, q_pos, r_contig, r_pos) order
    p->count_mode = 1;         // a chain contributes every anchor of its component
    p->mean_mode = 1;          // ANI = seed-weighted mean of window ANI
    p->af_mode = 0;            // AF = sum of kept chain spans (+ af_ext each)
    p->af_ext = 198.0;         // FITTED: end-of-chain allowance that reproduces both golden AFs
    p->overlap_tol = 0.0;
    p->overlap_side = 0;
    p->switch_mode = 0;
    p->denom_mode = 0;
    p->min_window_anchors = 0;
    p->strict_dr = 1;
}

orc_sketch_t* orc_sketch_new(const uint8_t* const* contigs, const uint64_t* lens, uint32_t n,
                             int32_t k, int32_t c, int32_t marker_c, int32_t seed) {
    auto* s = new orc_sketch();
    s->k = k; s->c = c; s->marker_c = marker_c; s->has_seeds = seed != 0;
    uint32_t contig_count = 0;
    for (uint32_t i = 0; i < n; i++) {
        if (lens[i] >= MIN_LENGTH_CONTIG) {
            s->contig_lengths.push_back((uint32_t)lens[i]);
            s->total_len += lens[i];
            fmh_seeds(contigs[i], lens[i], k, c, marker_c, contig_count, seed != 0, s->seeds, s->markers);
            contig_count++;
        }
    }
    std::sort(s->seeds.begin(), s->seeds.end(), [](const Seed& a, const Seed& b) {
        if (a.kmer != b.kmer) return a.kmer < b.kmer;
        if (a.contig != b.contig) return a.contig < b.contig;
        return a.pos < b.pos;
    });
    std::sort(s->markers.begin(), s->markers.end());
    s->markers.erase(std::unique(s->markers.begin(), s->markers.end()), s->markers.end());
    s->pos_by_contig.assign(contig_count, {});
    for (const Seed& sd : s->seeds) s->pos_by_contig[sd.contig].push_back(sd.pos);
    for (auto& v : s->pos_by_contig) std::sort(v.begin(), v.end());
    return s;
}

// CPU-baseline helper: sketch many genomes, one OpenMP task per genome (the axis skani's own rayon driver uses)
void orc_sketch_batch(const uint8_t* const* contigs, const uint64_t* lens, const uint32_t* genome_contig_start,
                      uint32_t n_genomes, int32_t k, int32_t c, int32_t marker_c, int32_t seed, int32_t threads,
                      orc_sketch_t** out) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t g = 0; g < (int64_t)n_genomes; g++) {
        const uint32_t a = genome_contig_start[g], b = genome_contig_start[g + 1];
        out[g] = orc_sketch_new(contigs + a, lens + a, b - a, k, c, marker_c, seed);
    }
}

void orc_sketch_free(orc_sketch_t* s) { delete s; }
uint64_t orc_sketch_n_seeds(const orc_sketch_t* s) { return s->seeds.size(); }
uint64_t orc_sketch_n_markers(const orc_sketch_t* s) { return s->markers.size(); }
uint32_t orc_sketch_n_contigs(const orc_sketch_t* s) { return (uint32_t)s->contig_lengths.size(); }
uint64_t orc_sketch_total_len(const orc_sketch_t* s) { return s->total_len; }
void orc_sketch_seeds(const orc_sketch_t* s, uint64_t* kmer, uint32_t* pos, uint32_t* contig, uint8_t* canonical) {
    for (size_t i = 0; i < s->seeds.size(); i++) {
        kmer[i] = s->seeds[i].kmer; pos[i] = s->seeds[i].pos;
        contig[i] = s->seeds[i].contig; canonical[i] = s->seeds[i].canonical;
    }
}
void orc_sketch_markers(const orc_sketch_t* s, uint64_t* out) {
    std::memcpy(out, s->markers.data(), s->markers.size() * sizeof(uint64_t));
}
void orc_sketch_contig_lengths(const orc_sketch_t* s, uint32_t* out) {
    std::memcpy(out, s->contig_lengths.data(), s->contig_lengths.size() * sizeof(uint32_t));
}

// check_markers_quickly.  The cut-off is formed as  p21 * |small|  with
// p21 = screen_val^21 computed once per query by repeated IEEE multiplication, so that the
// product path can form the identical double (one rounding per multiply, no libm pow).
// x^21 in the order of Rust's f64::powi (LLVM's powi expansion and compiler-rt's __powidf2: square and multiply from the
// lowest exponent bit, one IEEE rounding per product), so that the cut-off is the double skani forms if it calls powi
static inline double pow21(double x) {
    double r = 1.0;
    for (int b = MARKER_K;;) {
        if (b & 1) r *= x;
        b >>= 1;
        if (b == 0) break;
        x *= x;
    }
    return r;
}

int32_t orc_screen(const orc_sketch_t* q, const orc_sketch_t* r, double screen_val, int32_t rescue_small,
                   uint64_t* shared) {
    const auto& a = q->markers.size() <= r->markers.size() ? q->markers : r->markers;  // smaller
    const auto& b = q->markers.size() <= r->markers.size() ? r->markers : q->markers;
    uint64_t cnt = 0;
    size_t i = 0, j = 0;
    while (i < a.size() && j < b.size()) {
        if (a[i] < b[j]) i++; else if (a[i] > b[j]) j++; else { cnt++; i++; j++; }
    }
    if (shared) *shared = cnt;
    if (screen_val == 0.0) return 1;
    if (rescue_small && a.size() < 20) return 1;
    double cutoff = pow21(screen_val) * (double)a.size();
    return (double)cnt > cutoff ? 1 : 0;
}

void orc_chain(const orc_sketch_t* ref, const orc_sketch_t* query, const orc_chain_params_t* p, orc_result_t* out) {
    chain_impl(ref, query, *p, out);
}

// fit tooling: kept chains as rows of 10 doubles (chunk, qs, qe, rs, re, score, n_anchors, path_len, rev, chunk_first_qp)
int64_t orc_chain_dump(const orc_sketch_t* ref, const orc_sketch_t* query, const orc_chain_params_t* p,
                       double* rows, int64_t max_rows) {
    std::vector<double> sink; orc_result_t r;
    g_dump = &sink; chain_impl(ref, query, *p, &r); g_dump = nullptr;
    int64_t n = (int64_t)sink.size() / 10;
    if (rows) std::memcpy(rows, sink.data(), sizeof(double) * 10 * (size_t)std::min(n, max_rows));
    return n;
}

int64_t orc_query(const orc_sketch_t* query, const orc_sketch_t* const* refs, uint64_t n_refs,
                  double screen_val, int32_t rescue_small, const orc_chain_params_t* p,
                  int32_t threads, uint32_t* hit_idx, orc_result_t* out, uint64_t* n_screened_in) {
    std::vector<orc_result_t> res(n_refs);
    std::vector<uint8_t> pass(n_refs, 0);
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < (int64_t)n_refs; i++) {
        if (orc_screen(query, refs[i], screen_val, rescue_small, nullptr)) {
            pass[i] = 1;
            chain_impl(refs[i], query, *p, &res[i]);
        }
    }
    int64_t nh = 0; uint64_t ns = 0;
    for (uint64_t i = 0; i < n_refs; i++) {
        if (!pass[i]) continue;
        ns++;
        if (res[i].ani > 0.1f) { hit_idx[nh] = (uint32_t)i; out[nh] = res[i]; nh++; }   // reference lib.rs:654
    }
    if (n_screened_in) *n_screened_in = ns;
    return nh;
}

// pyskani's query loop for MANY queries against one database (the all-vs-all benchmark): every (query, ref) pair is
// screened (lib.rs:617-637), survivors are chained (lib.rs:640-657).  Both phases are parallel over pairs — the axis
// skani's own rayon drivers use for `skani search` / `skani triangle`.  Hits come back ordered by (query, ref).
int64_t orc_query_many(const orc_sketch_t* const* queries, uint64_t n_queries, const orc_sketch_t* const* refs, uint64_t n_refs,
                       double screen_val, int32_t rescue_small, const orc_chain_params_t* p, int32_t threads,
                       uint32_t* hit_q, uint32_t* hit_r, orc_result_t* out, uint64_t cap, uint64_t* n_screened_in) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    const int64_t n_pairs = (int64_t)(n_queries * n_refs);
    std::vector<uint8_t> pass((size_t)n_pairs, 0);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n_pairs; i++)
        pass[i] = (uint8_t)orc_screen(queries[i / (int64_t)n_refs], refs[i % (int64_t)n_refs], screen_val, rescue_small, nullptr);
    std::vector<int64_t> in;
    for (int64_t i = 0; i < n_pairs; i++) if (pass[i]) in.push_back(i);
    std::vector<orc_result_t> res(in.size());
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t j = 0; j < (int64_t)in.size(); j++)
        chain_impl(refs[in[j] % (int64_t)n_refs], queries[in[j] / (int64_t)n_refs], *p, &res[j]);
    int64_t nh = 0;
    for (size_t j = 0; j < in.size(); j++) {
        if (!(res[j].ani > 0.1f)) continue;                       // reference lib.rs:654
        if ((uint64_t)nh < cap) { hit_q[nh] = (uint32_t)(in[j] / (int64_t)n_refs); hit_r[nh] = (uint32_t)(in[j] % (int64_t)n_refs); out[nh] = res[j]; }
        nh++;
    }
    if (n_screened_in) *n_screened_in = in.size();
    return nh;
}

}  // extern "C"
