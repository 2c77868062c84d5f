"""fmh_seeds restatement: C++ oracle vs an independent numpy restatement, plus edge cases.

Follows SURVEY.md A.3/A.4 (skani v0.3.0 seeding.rs; call site reference lib.rs:165-171) and the
driver rules of Database::_sketch (reference lib.rs:140-185).
"""
import numpy as np
import pytest

import oracle

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def np_hash(x):
    x = x.astype(np.uint64).copy()
    with np.errstate(over="ignore"):
        x = ~(x + (x << np.uint64(21)))          # Rust: !key.wrapping_add(key << 21)
        x ^= x >> np.uint64(24)
        x = x + (x << np.uint64(3)) + (x << np.uint64(8))
        x ^= x >> np.uint64(14)
        x = x + (x << np.uint64(2)) + (x << np.uint64(4))
        x ^= x >> np.uint64(28)
        x = x + (x << np.uint64(31))
    return x


def np_sketch(seq: bytes, k=15, c=125, marker_c=1000):
    a = np.frombuffer(seq, np.uint8)
    lut = np.zeros(256, np.uint64)
    for i, ch in enumerate("ACGT"):
        lut[ord(ch)] = i
        lut[ord(ch.lower())] = i
    b = lut[a]

    def kmers(kk):
        n = len(b) - kk + 1
        f = np.zeros(n, np.uint64)
        r = np.zeros(n, np.uint64)
        for j in range(kk):
            f = (f << np.uint64(2)) | b[j:j + n]
            r = r | ((np.uint64(3) - b[j:j + n]) << np.uint64(2 * j))
        return f, r

    fk, rk = kmers(k)
    end = np.arange(len(fk)) + (k - 1)
    sel = (end >= 20) & (np_hash(np.minimum(fk, rk)) < M64 // np.uint64(c))
    kmer, pos, canon = np.minimum(fk, rk)[sel], end[sel].astype(np.uint32), (fk < rk)[sel].astype(np.uint8)
    o = np.lexsort((pos, kmer))
    fm, rm = kmers(21)
    cm = np.minimum(fm, rm)
    markers = np.unique(cm[np_hash(cm) < M64 // np.uint64(marker_c)])
    return kmer[o], pos[o], canon[o], markers


def rand_seq(n, seed):
    rng = np.random.default_rng(seed)
    return np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].tobytes()


def test_hash_known_answers():
    xs = np.array([0, 1, 2, 0x3FFFFFFF, 12345678901234, (1 << 42) - 1], np.uint64)
    want = np_hash(xs)
    got = np.array([oracle.mm_hash64(int(x)) for x in xs], np.uint64)
    assert np.array_equal(got, want)
    # frozen values (guards both restatements against drifting together)
    assert [hex(int(v)) for v in got[[0, 1, 3, 5]]] == ['0x77cfa1eef01bca90', '0x1f9a5be4bfb13e81', '0x37927a2d4b6c153b', '0xc0016e5e0066af67']


@pytest.mark.parametrize("n,seed", [(600, 1), (5000, 2), (200_000, 3)])
def test_against_numpy(n, seed):
    s = rand_seq(n, seed)
    S = oracle.Sketch([s])
    kmer, pos, contig, canon = S.seeds()
    wk, wp, wc, wm = np_sketch(s)
    assert np.array_equal(kmer, wk) and np.array_equal(pos, wp) and np.array_equal(canon, wc)
    assert (contig == 0).all()
    assert np.array_equal(S.markers(), wm)


def test_ecoli_slice_against_numpy(ecoli):
    s = ecoli[0][1_000_000:1_300_000]
    S = oracle.Sketch([s])
    kmer, pos, _, canon = S.seeds()
    wk, wp, wc, wm = np_sketch(s)
    assert np.array_equal(kmer, wk) and np.array_equal(pos, wp) and np.array_equal(canon, wc)
    assert np.array_equal(S.markers(), wm)


def test_other_k_c():
    s = rand_seq(50_000, 7)
    S = oracle.Sketch([s], k=13, c=30, marker_c=200)
    kmer, pos, _, canon = S.seeds()
    wk, wp, wc, wm = np_sketch(s, 13, 30, 200)
    assert np.array_equal(kmer, wk) and np.array_equal(pos, wp) and np.array_equal(canon, wc)
    assert np.array_equal(S.markers(), wm)


def test_short_contigs_skipped_and_index_counts_kept_only():
    # reference lib.rs:156 (MIN_LENGTH_CONTIG gate) and lib.rs:165-173 (contig_count)
    a, b, c = rand_seq(3000, 11), rand_seq(499, 12), rand_seq(500, 13)
    S = oracle.Sketch([a, b, c])
    assert list(S.contig_lengths()) == [3000, 500]
    assert S.total_len == 3500
    _, pos, contig, _ = S.seeds()
    assert set(contig.tolist()) <= {0, 1}
    S1 = oracle.Sketch([c])
    assert np.array_equal(np.sort(pos[contig == 1]), np.sort(S1.seeds()[1]))


def test_reference_database_test_inputs_have_no_seeds():
    # reference tests/test_database.py sketches b"ATGC"*100 (400 bp): below the gate → empty sketch
    S = oracle.Sketch([b"ATGC" * 100])
    assert S.n_seeds == 0 and S.n_markers == 0 and S.total_len == 0


def test_non_acgt_encodes_as_A_and_lowercase_equals_uppercase():
    s = bytearray(rand_seq(20_000, 21))
    t = bytearray(s)
    for i in range(100, 20_000, 997):
        s[i] = ord("N")
        t[i] = ord("A")
    A_, B_ = oracle.Sketch([bytes(s)]), oracle.Sketch([bytes(t)])
    for x, y in zip(A_.seeds(), B_.seeds()):
        assert np.array_equal(x, y)
    assert np.array_equal(A_.markers(), B_.markers())
    C_ = oracle.Sketch([bytes(t).lower()])
    for x, y in zip(C_.seeds(), B_.seeds()):
        assert np.array_equal(x, y)


def test_seed_false_keeps_markers_only():
    s = rand_seq(30_000, 5)
    S = oracle.Sketch([s], seed=False)
    assert S.n_seeds == 0 and S.n_markers > 0
    assert np.array_equal(S.markers(), oracle.Sketch([s]).markers())


def test_screen_rules():
    base = rand_seq(400_000, 31)
    other = rand_seq(400_000, 32)
    A_, B_ = oracle.Sketch([base]), oracle.Sketch([other])
    assert oracle.screen(A_, A_)[0]
    ok, shared = oracle.screen(A_, B_)
    assert not ok and shared <= 1
    tiny = oracle.Sketch([rand_seq(5_000, 33)])
    assert tiny.n_markers < 20
    assert oracle.screen(tiny, B_, 0.8, True)[0]          # rescue_small
    assert not oracle.screen(tiny, B_, 0.8, False)[0]     # faster_small
    assert oracle.screen(A_, B_, 0.0, False)[0]           # screen_val == 0 → everything passes
