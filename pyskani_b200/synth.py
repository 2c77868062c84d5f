"""Synthetic genomes for the benchmark configurations of BASELINE.json.

random_genome: i.i.d. uniform ACGT.  mutate: point substitutions with probability 0.9*d per base and indel
events with probability 0.1*d per base (half insertions, half deletions, geometric length with mean 3),
so the expected ANI of a mutant against its parent is about 1 - d.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", np.uint8)


def random_genome(n, seed):
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, n, dtype=np.uint8)]


def mutate(genome, d, seed):
    """genome: uint8 ASCII array. Returns a new uint8 ASCII array."""
    rng = np.random.default_rng(seed)
    n = len(genome)
    lut = np.zeros(256, np.uint8)
    lut[_ACGT] = np.arange(4, dtype=np.uint8)
    codes = lut[genome]
    sub = rng.random(n) < 0.9 * d
    ns = int(sub.sum())
    codes[sub] = (codes[sub] + rng.integers(1, 4, ns, dtype=np.uint8)) & 3
    ev = np.flatnonzero(rng.random(n) < 0.1 * d)
    if len(ev):
        is_ins = rng.random(len(ev)) < 0.5
        length = rng.geometric(1.0 / 3.0, len(ev))
        # deletions: drop [p, p + len)
        keep = np.ones(n, bool)
        for p, l in zip(ev[~is_ins], length[~is_ins]):
            keep[p:p + l] = False
        # insertions: random bases before position p
        ins_pos = np.repeat(ev[is_ins], length[is_ins])
        ins_val = rng.integers(0, 4, len(ins_pos), dtype=np.uint8)
        # map insertion points to the coordinates after deletion
        new_index = np.cumsum(keep) - keep
        codes2 = codes[keep]
        codes = np.insert(codes2, new_index[ins_pos], ins_val)
    return _ACGT[codes]


def fragment(genome, seed, lo=1000, hi=50000):
    """Cuts a genome into contigs with log-uniform lengths in [lo, hi], shuffles them and reverse-complements
    half of them (MAG-style query of BASELINE.json config 4)."""
    rng = np.random.default_rng(seed)
    n = len(genome)
    cuts, p = [], 0
    while p < n:
        l = int(np.exp(rng.uniform(np.log(lo), np.log(hi))))
        cuts.append((p, min(n, p + l)))
        p += l
    order = rng.permutation(len(cuts))
    comp = np.zeros(256, np.uint8)
    comp[_ACGT] = np.frombuffer(b"TGCA", np.uint8)
    out = []
    for i in order:
        a, b = cuts[i]
        s = genome[a:b]
        if rng.random() < 0.5:
            s = comp[s[::-1]]
        out.append(np.ascontiguousarray(s))
    return out
