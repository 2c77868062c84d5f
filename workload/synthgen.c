/* synthgen.c — synthetic genomes for the benchmark configurations of BASELINE.json.
 *
 * Neutral workload code: neither the product (pyskani_b200) nor the oracle.  Both arms of bench.py and the tools draw
 * their inputs from here, so that they see byte-identical genomes.
 *
 *   wl_random_genome  i.i.d. uniform ACGT from a counter-based generator (splitmix64 of seed and block index), so any
 *                     slice can be produced independently and on any number of threads.
 *   wl_mutate         per base: substitution with probability 0.9 d (uniform over the three other bases), indel event
 *                     with probability 0.1 d (half insertions, half deletions, geometric length with mean 3, inserted
 *                     bases uniform).  Expected ANI of a mutant against its parent is about 1 - d.  Events are placed by
 *                     geometric skips, so the cost is proportional to the number of events, not the genome length.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static const char ACGT[4] = {'A', 'C', 'G', 'T'};

void wl_random_genome(uint8_t* out, uint64_t n, uint64_t seed) {
    const uint64_t key = splitmix64(seed ^ 0x5EED5EED5EED5EEDull);
    for (uint64_t blk = 0; blk * 32 < n; blk++) {
        uint64_t x = splitmix64(key + blk);
        const uint64_t end = (blk + 1) * 32 < n ? (blk + 1) * 32 : n;
        for (uint64_t i = blk * 32; i < end; i++, x >>= 2) out[i] = (uint8_t)ACGT[x & 3];
    }
}

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t* r) { r->s += 0x9E3779B97F4A7C15ull; uint64_t x = r->s; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }
static inline double rng_unit(rng_t* r) { return ((double)(rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }   /* (0, 1) */
static inline uint64_t rng_geometric(rng_t* r, double log1mp) { return (uint64_t)(log(rng_unit(r)) / log1mp); }         /* failures before a success */

static inline int code_of(uint8_t c) { return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; }

/* Returns the length of the mutant written to out (at most cap bytes; a mutant that would not fit is truncated). */
uint64_t wl_mutate(const uint8_t* in, uint64_t n, double d, uint64_t seed, uint8_t* out, uint64_t cap) {
    if (d <= 0.0) { const uint64_t m = n < cap ? n : cap; memcpy(out, in, m); return m; }
    rng_t r = { splitmix64(seed) };
    const double log1md = log(1.0 - d), log_len = log(2.0 / 3.0);      /* lengths: P(l) = (1/3)(2/3)^(l-1), mean 3 */
    uint64_t i = 0, o = 0;
    while (i < n && o < cap) {
        uint64_t skip = rng_geometric(&r, log1md);                     /* unchanged bases before the next event */
        if (skip > n - i) skip = n - i;
        if (skip > cap - o) skip = cap - o;
        memcpy(out + o, in + i, skip);
        i += skip; o += skip;
        if (i >= n || o >= cap) break;
        const double u = rng_unit(&r);
        if (u < 0.9) {                                                  /* substitution */
            out[o++] = (uint8_t)ACGT[(code_of(in[i]) + 1 + (int)(rng_next(&r) % 3)) & 3];
            i++;
        } else {
            const uint64_t len = 1 + rng_geometric(&r, log_len);
            if (u < 0.95) {                                             /* insertion in front of base i */
                uint64_t x = 0; int have = 0;
                for (uint64_t t = 0; t < len && o < cap; t++) {
                    if (!have) { x = rng_next(&r); have = 32; }
                    out[o++] = (uint8_t)ACGT[x & 3]; x >>= 2; have--;
                }
            } else {                                                    /* deletion of [i, i + len) */
                i += len < n - i ? len : n - i;
            }
        }
    }
    return o;
}
