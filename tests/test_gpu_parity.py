"""GPU parity: libskb (through its C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): sketches (seed k-mers, positions, strands, contigs; marker sets) and the
set of pairs passing the screen are bit-exact; ANI within 1e-4 (0.01 pp) and AF within 1e-3.
"""
import numpy as np
import pytest

import oracle
from pyskani_b200 import synth

pytestmark = pytest.mark.gpu

ANI_TOL = 1e-4
AF_TOL = 1e-3


@pytest.fixture(scope="module")
def ctx():
    from pyskani_b200 import capi
    c = capi.Context(0)
    yield c


def assert_sketch_equal(gs, osk):
    e = gs.export()
    ok, op, oc, ocan = osk.seeds()
    assert len(e["kmer"]) == len(ok)
    assert np.array_equal(e["kmer"], ok)
    assert np.array_equal(e["pos"], op)
    assert np.array_equal(e["contig"], oc)
    assert np.array_equal(e["canonical"], ocan)
    assert np.array_equal(e["markers"], osk.markers())
    assert np.array_equal(e["contig_lengths"], osk.contig_lengths())
    i = gs.info()
    assert i.total_len == osk.total_len


def rand(n, seed):
    return synth.random_genome(n, seed).tobytes()


# ------------------------------------------------------------------ sketches
@pytest.mark.parametrize("n,seed", [(500, 1), (521, 2), (4096, 3), (4097, 4), (4112, 5), (100_000, 6), (1_000_003, 7)])
def test_sketch_single_contig(ctx, n, seed):
    s = rand(n, seed)
    (g,) = ctx.sketch_batch([[s]])
    assert_sketch_equal(g, oracle.Sketch([s]))


def test_sketch_ecoli(ctx, ecoli):
    ec, k12, _ = ecoli
    g1, g2 = ctx.sketch_batch([[ec], [k12]])
    assert_sketch_equal(g1, oracle.Sketch([ec]))
    assert_sketch_equal(g2, oracle.Sketch([k12]))
    assert (g1.info().n_seeds, g1.info().n_markers) == (37237, 4539)


def test_sketch_ragged_batch(ctx):
    # empty genome, genome of only short contigs, many small contigs, big + small mix
    rng = np.random.default_rng(5)
    genomes = [
        [],
        [rand(499, 1), rand(10, 2), b""],
        [rand(int(l), 100 + i) for i, l in enumerate(rng.integers(400, 9000, 40))],
        [rand(300_000, 7), rand(499, 8), rand(500, 9), rand(70_001, 10)],
        [rand(5000, 11)],
    ]
    gs = ctx.sketch_batch(genomes)
    for g, contigs in zip(gs, genomes):
        assert_sketch_equal(g, oracle.Sketch(contigs))
    assert gs[0].info().n_seeds == 0 and gs[1].info().n_contigs == 0


@pytest.mark.parametrize("ingest", ["default", "pack"])
def test_sketch_tile_and_region_boundaries(ctx, monkeypatch, ingest):
    """Contig lengths around the kernel's work units (16-base words, 2 048-base tiles, 16 384-base regions) with every
    position a seed (c = 1), so that a k-mer dropped or duplicated at any boundary changes the sketch; host pointers at
    odd addresses.  Once through the ASCII input of the seeding kernel and once through its packed-word input (every
    chunk compacted by the host threads, chunks of 16 KB so that launches start and end inside genomes' neighbourhoods)."""
    if ingest == "pack":
        monkeypatch.setenv("SKB_INGEST", "pack")
        monkeypatch.setenv("SKB_CHUNK_KB", "16")
        ctx.set_host_threads(3)
    lens = [500, 511, 512, 513, 2047, 2048, 2049, 2048 + 14, 2048 + 20, 4096, 16383, 16384, 16385, 16384 + 20, 32768 + 7,
            3 * 16384 - 1]
    big = np.frombuffer(rand(sum(lens) + 3 * len(lens) + 8, 41), np.uint8)
    contigs, off = [], 1
    for l in lens:
        contigs.append(big[off:off + l])          # views at odd offsets of one buffer
        off += l + 3
    for c, mc in ((1, 1), (7, 3)):
        gs = ctx.sketch_batch([contigs, contigs[::-1], [contigs[5]], [contigs[11], contigs[12]]], c=c, marker_c=mc)
        want = [contigs, contigs[::-1], [contigs[5]], [contigs[11], contigs[12]]]
        for g, w in zip(gs, want):
            assert_sketch_equal(g, oracle.Sketch([x.tobytes() for x in w], c=c, marker_c=mc))
        if ingest == "pack":
            st = ctx.stats()
            assert st.h2d_packed_bytes > 0 and st.h2d_raw_bytes == 0
    ctx.set_host_threads(-1)


def test_marker_sets_all_three_sort_paths(ctx):
    """Marker sets are sorted per genome in shared memory (tiles of up to 16 384 markers), by CUB's segmented sort above
    that; one batch with a tiny, a mid-size (P = 16 384 tile) and an over-size genome, plus duplicated contigs so that
    the de-duplication has work to do."""
    small = rand(30_000, 301)
    mid = rand(9_500_000, 302)
    genomes = [[small, small], [mid], [rand(600, 303)]]
    gs = ctx.sketch_batch(genomes)
    assert 8192 < gs[1].info().n_markers <= 16384
    for g, contigs in zip(gs, genomes):
        assert_sketch_equal(g, oracle.Sketch(contigs))
    big = rand(17_500_000, 304)                     # > 16 384 markers: the whole batch takes the segmented-sort path
    gs2 = ctx.sketch_batch([[small, small], [big]])
    assert gs2[1].info().n_markers > 16384
    for g, contigs in zip(gs2, [[small, small], [big]]):
        assert_sketch_equal(g, oracle.Sketch(contigs))


def test_host_copies_merged_or_separate_give_the_same_sketch(ctx, monkeypatch):
    """Contigs laid out back to back in one host buffer (16-byte aligned, the device layout's spacing) are transferred as
    merged copies; the same contigs as separate objects, or with merging switched off, must give identical sketches."""
    lens = [300_000, 262_144, 1_000_003, 700, 400_000, 262_145, 1_500_000, 263_000]
    seqs = [np.frombuffer(rand(l, 500 + i), np.uint8) for i, l in enumerate(lens)]
    big = np.zeros(sum((l + 15) // 16 * 16 + 16 for l in lens) + 64, np.uint8)
    views, off = [], 16
    for s_ in seqs:
        big[off:off + len(s_)] = s_
        views.append(big[off:off + len(s_)])
        off += (len(s_) + 15) // 16 * 16 + 16
    split = [(0, 3), (3, 5), (5, 8)]
    merged = ctx.sketch_batch([views[a:b] for a, b in split])
    separate = ctx.sketch_batch([[x.tobytes() for x in seqs[a:b]] for a, b in split])
    monkeypatch.setenv("SKB_NO_COPY_MERGE", "1")
    unmerged = ctx.sketch_batch([views[a:b] for a, b in split])
    for x, y, z in zip(merged, separate, unmerged):
        ex, ey, ez = x.export(), y.export(), z.export()
        for key in ex:
            assert np.array_equal(ex[key], ey[key]) and np.array_equal(ex[key], ez[key]), key
    assert_sketch_equal(merged[0], oracle.Sketch([x.tobytes() for x in seqs[0:3]]))


def test_ingest_routes_give_the_same_sketch(ctx, monkeypatch):
    """Large host batches reach the device partly as ASCII (copy engine) and partly as 2-bit words compacted by host threads
    (host_pack.h); which chunk takes which route depends on timing.  Every policy - all ASCII, all compacted, mixed, the
    pipeline switched off - must give the oracle's sketches bit for bit: ragged genomes, contigs below the 500 bp gate, small
    contigs that go through staging, junk bytes, lower case, lengths that are not multiples of 16."""
    from pyskani_b200 import capi
    monkeypatch.setenv("SKB_CHUNK_KB", "256")            # many chunks out of a few MB
    junk = bytearray(rand(600_000, 711))
    junk[1000:1400] = b"N" * 400
    junk[5000:5003] = b"\n-*"
    genomes = [[rand(1_200_003, 701)], [rand(70_001, 702), rand(499, 703), rand(9_000, 704).lower(), rand(300_000, 705)],
               [bytes(junk)], [rand(501, 706)], [rand(2_000_000, 707), rand(333_333, 708)], [rand(40_000, 709)] * 3]
    want = [oracle.Sketch(g) for g in genomes]
    seen_packed = seen_raw = False
    for policy, threads in (("raw", -1), ("pack", 4), ("mix", 3), ("mix", 8), ("pack", 2), ("mix", 1)):
        monkeypatch.setenv("SKB_INGEST", policy)
        ctx.set_host_threads(threads)
        for rep in range(2):
            got = ctx.sketch_batch(genomes)
            st = ctx.stats()
            for g, o in zip(got, want):
                assert_sketch_equal(g, o)
            if policy == "pack" and threads >= 2:
                assert st.h2d_raw_bytes == 0 and st.h2d_packed_bytes > 0
            if policy == "raw" or threads < 2:
                assert st.h2d_packed_bytes == 0 and st.h2d_raw_bytes > 0
            seen_packed |= st.h2d_packed_bytes > 0
            seen_raw |= st.h2d_raw_bytes > 0
        (one,) = ctx.sketch_batch(genomes[4:5], c=30, marker_c=200)
        assert_sketch_equal(one, oracle.Sketch(genomes[4], c=30, marker_c=200))
    # a call that fails while the team is at work (k > 16 is refused inside the first sub-batch) unwinds cleanly: the team
    # is stopped before the call returns and the context keeps working
    monkeypatch.setenv("SKB_INGEST", "mix")
    ctx.set_host_threads(4)
    for _ in range(3):
        with pytest.raises(capi.SkbError):
            ctx.sketch_batch(genomes, k=17)
        (again,) = ctx.sketch_batch(genomes[:1])
        assert_sketch_equal(again, want[0])
    # the automatic policy: pageable sources (these bytes objects) are left to the packing threads altogether, pinned ones
    # are shared between the copy engine and the packing threads
    monkeypatch.setenv("SKB_INGEST", "auto")
    ctx.set_host_threads(4)
    got = ctx.sketch_batch(genomes)
    st = ctx.stats()
    assert st.h2d_raw_bytes == 0 and st.h2d_packed_bytes > 0
    for g, o in zip(got, want):
        assert_sketch_equal(g, o)
    import ctypes as C
    flat = [c for g in genomes for c in g]
    total = sum((len(c) + 15) // 16 * 16 + 16 for c in flat) + 64
    hp = ctx.host_alloc(total)
    try:
        arr = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(total,))
        views, off = [], 16
        for g in genomes:
            vg = []
            for c in g:
                arr[off:off + len(c)] = np.frombuffer(c, np.uint8)
                vg.append(arr[off:off + len(c)])
                off += (len(c) + 15) // 16 * 16 + 16
            views.append(vg)
        for rep in range(3):
            got = ctx.sketch_batch(views)
            st = ctx.stats()
            seen_raw |= st.h2d_raw_bytes > 0
            for g, o in zip(got, want):
                assert_sketch_equal(g, o)
    finally:
        ctx.host_free(hp)
    ctx.set_host_threads(-1)
    assert seen_packed and seen_raw


def test_sketch_hash_comparison_variants(ctx, monkeypatch):
    """The seeding kernel compares the high words of hash and threshold and re-checks every hit exactly when it writes it;
    a hit that fails the re-check makes the host repeat the batch with the exact 64-bit comparison.  Forced exact
    comparison (SKB_SEED_EXACT) and a forced repeat (SKB_SEED_TEST_INEXACT halves the re-check thresholds) both give the
    oracle's sketch, for one genome and for a ragged batch."""
    genomes = [[rand(700_000, 61)], [rand(3_000, 62), rand(900, 63), b"ACGT" * 50], [rand(150_000, 64)]]
    want = [oracle.Sketch(g) for g in genomes]
    for env in ({}, {"SKB_SEED_EXACT": "1"}, {"SKB_SEED_TEST_INEXACT": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for g, o in zip(ctx.sketch_batch(genomes), want):
            assert_sketch_equal(g, o)
        (one,) = ctx.sketch_batch(genomes[:1], c=30, marker_c=200)
        assert_sketch_equal(one, oracle.Sketch(genomes[0], c=30, marker_c=200))
        for k in env:
            monkeypatch.delenv(k)


def test_sketch_junk_and_lowercase(ctx):
    rng = np.random.default_rng(3)
    alpha = np.frombuffer(b"ACGTacgtNnRYKM-*", np.uint8)
    s = alpha[rng.integers(0, len(alpha), 200_000)].tobytes()
    (g,) = ctx.sketch_batch([[s]])
    assert_sketch_equal(g, oracle.Sketch([s]))


@pytest.mark.parametrize("k,c,mc", [(15, 125, 1000), (13, 30, 200), (16, 70, 500), (12, 1, 1), (15, 200, 3000)])
def test_sketch_params(ctx, k, c, mc):
    s = rand(60_000, 17)
    (g,) = ctx.sketch_batch([[s]], k=k, c=c, marker_c=mc)
    assert_sketch_equal(g, oracle.Sketch([s], k=k, c=c, marker_c=mc))


def test_sketch_seed_false(ctx):
    s = rand(80_000, 19)
    (g,) = ctx.sketch_batch([[s]], seed=False)
    o = oracle.Sketch([s], seed=False)
    assert g.info().n_seeds == 0 and g.info().has_seeds == 0
    assert np.array_equal(g.export()["markers"], o.markers())


def test_sketch_low_complexity_overflow_path(ctx):
    # poly-A and short tandem repeats: seed density far from 1/c exercises the capacity retry
    s = (b"A" * 300_000) + (b"ACGTTGCA" * 20_000) + rand(50_000, 23)
    for c in (1, 125):
        (g,) = ctx.sketch_batch([[s]], c=c, marker_c=max(1, c * 8))
        assert_sketch_equal(g, oracle.Sketch([s], c=c, marker_c=max(1, c * 8)))


def test_import_roundtrip(ctx):
    contigs = [rand(120_000, 31), rand(30_000, 32)]
    (g,) = ctx.sketch_batch([contigs])
    e = g.export()
    perm = np.random.default_rng(1).permutation(len(e["kmer"]))
    g2 = ctx.import_sketch(e["kmer"][perm], e["pos"][perm], e["contig"][perm], e["canonical"][perm],
                           e["markers"][::-1].copy(), e["contig_lengths"])
    e2 = g2.export()
    for key in e:
        assert np.array_equal(e[key], e2[key]), key


# ------------------------------------------------------------------ screen
def make_family(n_mut, length, seed, divs):
    base = synth.random_genome(length, seed)
    return [base] + [synth.mutate(base, d, seed * 1000 + i + 1) for i, d in zip(range(n_mut), divs)]


def test_screen_matches_oracle(ctx):
    from pyskani_b200 import capi
    fams = []
    for f in range(3):
        fams += make_family(4, 400_000, 40 + f, [0.01, 0.05, 0.12, 0.22])
    fams.append(synth.random_genome(5_000, 99))      # < 20 markers: exercises the small-genome rescue
    contigs = [[g.tobytes()] for g in fams]
    gs = ctx.sketch_batch(contigs)
    os_ = [oracle.Sketch(c) for c in contigs]
    db = capi.Database(ctx)
    for g in gs:
        db.add(g)
    for cutoff, rescue in ((0.8, True), (0.8, False), (0.9, True), (0.95, False)):
        ok, shared = db.screen(gs, cutoff, rescue)
        for i, qo in enumerate(os_):
            for j, ro in enumerate(os_):
                want_ok, want_shared = oracle.screen(qo, ro, cutoff, rescue)
                assert shared[i, j] == want_shared, (i, j)
                assert ok[i, j] == want_ok, (i, j, cutoff, rescue)


@pytest.mark.parametrize("mode", ["index", "pairwise"])
def test_screen_modes_match_oracle(ctx, monkeypatch, mode):
    """Both screen implementations (pairwise list intersection; join through the database's marker index) give the
    oracle's intersection sizes and decisions, also after the database grew behind an already built index."""
    from pyskani_b200 import capi
    monkeypatch.setenv("SKB_SCREEN_MODE", mode)
    fams = []
    for f in range(3):
        fams += make_family(4, 300_000, 140 + f, [0.01, 0.05, 0.12, 0.22])
    fams.append(synth.random_genome(5_000, 199))
    fams.append(np.frombuffer(b"ACGT" * 200, np.uint8))            # passes the contig gate, almost no markers
    contigs = [[g.tobytes()] for g in fams]
    gs = ctx.sketch_batch(contigs)
    os_ = [oracle.Sketch(c) for c in contigs]
    db = capi.Database(ctx)
    db.add_many(gs[:7])

    def check(n_refs):
        for cutoff, rescue in ((0.8, True), (0.95, False)):
            ok, shared = db.screen(gs, cutoff, rescue)
            assert ok.shape == (len(gs), n_refs)
            for i, qo in enumerate(os_):
                for j in range(n_refs):
                    want_ok, want_shared = oracle.screen(qo, os_[j], cutoff, rescue)
                    assert shared[i, j] == want_shared and ok[i, j] == want_ok, (mode, i, j, cutoff, rescue)

    check(7)
    db.add_many(gs[7:])
    check(len(gs))
    hits_a, n_a = db.query(gs)
    monkeypatch.setenv("SKB_SCREEN_MODE", "pairwise" if mode == "index" else "index")
    hits_b, n_b = db.query(gs)
    assert hits_a == hits_b and n_a == n_b


def test_marker_index_build_paths(ctx, monkeypatch):
    """The database's marker index is built by bucket partition (histogram, scan, scatter, rank inside the bucket); a bucket
    beyond the rank kernel's capacity - here 110 identical genomes, so every marker has 110 postings - makes the build fall
    back to the radix sort.  Partition, forced sort and the pairwise kernels give the same counts and decisions (= the
    oracle's), with many queries and with few (where the join divides a query's markers among several CTAs)."""
    from pyskani_b200 import capi
    fams = []
    for f in range(3):
        fams += make_family(4, 250_000, 340 + f, [0.01, 0.05, 0.12, 0.22])
    contigs = [[g.tobytes()] for g in fams]
    clone = synth.random_genome(80_000, 399).tobytes()
    few = ctx.sketch_batch(contigs)
    many = ctx.sketch_batch(contigs + [[clone]] * 110)
    os_few = [oracle.Sketch(c) for c in contigs]
    o_clone = oracle.Sketch([clone])
    for name, gs in (("no overflow", few), ("overflow -> sort", many)):
        results = {}
        for mode, env in (("partition", {"SKB_SCREEN_MODE": "index"}), ("sort", {"SKB_SCREEN_MODE": "index", "SKB_MIDX_SORT": "1"}),
                          ("pairwise", {"SKB_SCREEN_MODE": "pairwise"})):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            db = capi.Database(ctx)
            db.add_many(gs)
            results[mode] = [db.screen(gs[:3], 0.8, True), db.screen(gs, 0.8, True), db.screen(gs[5:6], 0.95, False)]
            for k in env:
                monkeypatch.delenv(k)
        for mode in ("sort", "pairwise"):
            for (ok_a, sh_a), (ok_b, sh_b) in zip(results["partition"], results[mode]):
                assert np.array_equal(ok_a, ok_b) and np.array_equal(sh_a, sh_b), (name, mode)
        ok, shared = results["partition"][1]
        for i in (0, 5, len(os_few) - 1):
            for j in range(len(os_few)):
                want_ok, want_shared = oracle.screen(os_few[i], os_few[j], 0.8, True)
                assert shared[i, j] == want_shared and ok[i, j] == want_ok, (name, i, j)
        if len(gs) > len(few):
            want_ok, want_shared = oracle.screen(o_clone, o_clone, 0.8, True)
            n0 = len(contigs)
            assert (shared[n0:, n0:] == want_shared).all() and ok[n0:, n0:].all() == bool(want_ok)


# ------------------------------------------------------------------ chain / ANI
def check_hits(hits, oq, orefs, cutoff=0.8, rescue=True, **flags):
    idx, res, n_in = oracle.query(oq, orefs, cutoff, rescue, oracle.default_params(**flags))
    got = {h[1]: h for h in hits}
    assert sorted(got) == sorted(int(i) for i in idx)
    for i, r in zip(idx, res):
        h = got[int(i)]
        assert abs(h[2] - r.ani) <= ANI_TOL, (i, h[2], r.ani)
        assert abs(h[3] - r.af_query) <= AF_TOL and abs(h[4] - r.af_ref) <= AF_TOL, (i, h, r.af_query, r.af_ref)
        assert h[7] == r.n_anchors and h[6] == r.n_chains and h[5] == r.n_windows, (i, h, r.n_anchors, r.n_chains, r.n_windows)
    return n_in


def test_ecoli_goldens_through_gpu(ctx, ecoli):
    """reference tests/test_ani.py:28-61 — the same assertions, through the CUDA path."""
    from pyskani_b200 import capi
    ec, k12, gold = ecoli
    ref, qry = ctx.sketch_batch([[ec], [k12]])
    db = capi.Database(ctx)
    db.add(ref)
    for flags, key in (({}, "ani_no_learned"), ({"robust": True}, "ani_robust"), ({"median": True}, "ani_median")):
        hits, n_in = db.query([qry], **flags)
        assert len(hits) == 1 and n_in == 1
        h = hits[0]
        assert round(abs(h[4] - gold["af_ref"]), 4) == 0
        assert round(abs(h[3] - gold["af_query"]), 4) == 0
        assert round(abs(h[2] - gold[key]), 4) == 0, (key, h[2])
    oq, orf = oracle.Sketch([k12]), oracle.Sketch([ec])
    check_hits(db.query([qry])[0], oq, [orf])
    check_hits(db.query([qry], robust=True)[0], oq, [orf], robust=1)
    check_hits(db.query([qry], median=True)[0], oq, [orf], median=1)


def test_learned_ani_true_is_refused(ctx):
    from pyskani_b200 import capi
    s = rand(50_000, 3)
    (g,) = ctx.sketch_batch([[s]])
    db = capi.Database(ctx)
    db.add(g)
    with pytest.raises(capi.SkbError) as e:
        db.query([g], learned_ani=1)
    assert e.value.code == capi.SKB_ERR_UNSUPPORTED


def test_mutant_series_vs_oracle(ctx):
    """BASELINE.json config 2 at reduced size: one genome vs mutated copies, 1-15 % divergence + indels."""
    from pyskani_b200 import capi
    base = synth.random_genome(1_000_000, 7)
    divs = np.linspace(0.01, 0.15, 12)
    refs = [synth.mutate(base, d, 700 + i) for i, d in enumerate(divs)]
    gs = ctx.sketch_batch([[r.tobytes()] for r in refs] + [[base.tobytes()]])
    db = capi.Database(ctx)
    for g in gs[:-1]:
        db.add(g)
    oq = oracle.Sketch([base.tobytes()])
    orefs = [oracle.Sketch([r.tobytes()]) for r in refs]
    for flags in ({}, {"robust": 1}, {"median": 1}):
        hits, n_in = db.query([gs[-1]], robust=bool(flags.get("robust")), median=bool(flags.get("median")))
        n_in_o = check_hits(hits, oq, orefs, **flags)
        assert n_in == n_in_o
    # sanity of the estimate itself: ANI tracks 1 - d for the close mutants
    hits, _ = db.query([gs[-1]])
    for h in hits:
        if divs[h[1]] <= 0.08:
            assert abs(h[2] - (1 - divs[h[1]])) < 0.01


def test_fragmented_query_vs_oracle(ctx):
    """BASELINE.json config 4 at reduced size: contigs of 1-50 kbp, shuffled, half reverse-complemented."""
    from pyskani_b200 import capi
    base = synth.random_genome(800_000, 11)
    refs = [synth.mutate(base, d, 1100 + i) for i, d in enumerate((0.02, 0.06, 0.1))] + [synth.random_genome(500_000, 12)]
    qcontigs = [c.tobytes() for c in synth.fragment(synth.mutate(base, 0.03, 1199), 5, lo=400, hi=50_000)]
    rcontigs = [[c.tobytes() for c in synth.fragment(r, 50 + i, lo=2000, hi=200_000)] for i, r in enumerate(refs)]
    gs = ctx.sketch_batch(rcontigs + [qcontigs])
    db = capi.Database(ctx)
    for g in gs[:-1]:
        db.add(g)
    oq = oracle.Sketch(qcontigs)
    orefs = [oracle.Sketch(c) for c in rcontigs]
    for g, o in zip(gs, orefs + [oq]):
        assert_sketch_equal(g, o)
    hits, n_in = db.query([gs[-1]])
    assert check_hits(hits, oq, orefs) == n_in
    assert len(hits) == 3


def test_all_vs_all_small(ctx, monkeypatch):
    """BASELINE.json config 3 at reduced size: families of related genomes, every ordered pair; the same in many small
    chaining batches (the batch size follows the free device memory and is normally far above this workload)."""
    from pyskani_b200 import capi
    genomes = []
    for f in range(4):
        genomes += make_family(3, 300_000, 60 + f, [0.02, 0.07, 0.13])
    contigs = [[g.tobytes()] for g in genomes]
    gs = ctx.sketch_batch(contigs)
    db = capi.Database(ctx)
    for g in gs:
        db.add(g)
    os_ = [oracle.Sketch(c) for c in contigs]
    hits, n_in = db.query(gs)
    total_in = 0
    for qi, oq in enumerate(os_):
        total_in += check_hits([h for h in hits if h[0] == qi], oq, os_)
    assert total_in == n_in
    # only intra-family pairs can pass the screen
    assert all(h[0] // 4 == h[1] // 4 for h in hits)
    monkeypatch.setenv("SKB_CHAIN_BATCH_MSEEDS", "1")           # <= 2^20 query seeds per batch: 64 pairs -> several batches
    big = ctx.sketch_batch([[synth.random_genome(2_000_000, 75).tobytes()], [synth.mutate(synth.random_genome(2_000_000, 75), 0.03, 76).tobytes()]])
    db2 = capi.Database(ctx)
    db2.add_many(list(gs) + list(big))
    hits_small, n_small = db2.query(list(gs) + list(big))
    monkeypatch.delenv("SKB_CHAIN_BATCH_MSEEDS")
    hits_one, n_one = db2.query(list(gs) + list(big))
    assert hits_small == hits_one and n_small == n_one and n_one >= n_in + 4
    # a chaining arena that cannot be allocated halves the batch size of the context and the remaining pairs are planned again
    from pyskani_b200 import capi as _capi
    tight = _capi.Context(0)
    gt = tight.sketch_batch(contigs)
    dbt = _capi.Database(tight)
    dbt.add_many(gt)
    monkeypatch.setenv("SKB_TEST_CHAIN_ARENA_MB", "12")
    hits_tight, n_tight = dbt.query(gt)
    monkeypatch.delenv("SKB_TEST_CHAIN_ARENA_MB")
    assert (hits_tight, n_tight) == (hits, n_in)


def test_walk_groups(ctx, monkeypatch):
    """The window walk handles several pairs of one query per CTA once a batch holds more pairs than SMs.  Force that
    grouping on a small all-vs-all (single-contig and fragmented queries) and compare with the ungrouped run and the oracle."""
    from pyskani_b200 import capi
    genomes = []
    for f in range(3):
        fam = make_family(4, 250_000, 170 + f, [0.01, 0.04, 0.08, 0.12])
        genomes += [[g.tobytes()] for g in fam[:3]]
        genomes.append([c.tobytes() for c in synth.fragment(fam[3], 180 + f, 2_000, 30_000)])
    gs = ctx.sketch_batch(genomes)
    db = capi.Database(ctx)
    db.add_many(gs)
    monkeypatch.setenv("SKB_WALK_GROUP", "1")
    want = db.query(gs)
    for g in ("3", "7"):
        monkeypatch.setenv("SKB_WALK_GROUP", g)
        assert db.query(gs) == want
    os_ = [oracle.Sketch(c) for c in genomes]
    total = 0
    for qi, oq in enumerate(os_):
        total += check_hits([h for h in want[0] if h[0] == qi], oq, os_)
    assert total == want[1] and total >= 36


def test_pack_unpack_device_roundtrip(ctx):
    """skb_sketch_pack / skb_sketch_unpack (the multi-GPU exchange path): sketches rebuilt from a packed device payload on
    ANOTHER context are identical to the originals and answer queries identically."""
    from pyskani_b200 import capi
    fam = make_family(3, 200_000, 210, [0.02, 0.06, 0.1])
    genomes = [[g.tobytes()] for g in fam]
    genomes.append([c.tobytes() for c in synth.fragment(fam[1], 211, 2_000, 20_000)])
    genomes.append([b"ACGT" * 50])                                   # below the contig gate: an empty sketch travels too
    gs = ctx.sketch_batch(genomes)
    pb, mb = ctx.pack_size(gs)
    assert pb % 16 == 0 and mb > 0
    other = capi.Context(0)
    buf = ctx.dev_alloc(pb)
    try:
        meta = ctx.pack(gs, buf, pb)
        got = other.unpack(meta, buf, pb)
        with pytest.raises(capi.SkbError):
            other.unpack(meta[:10], buf, pb)
        with pytest.raises(capi.SkbError):
            other.unpack(meta, buf, pb - 16)
    finally:
        ctx.dev_free(buf)
    assert len(got) == len(gs)
    for a, b in zip(gs, got):
        ia, ib = a.info(), b.info()
        assert (ia.n_seeds, ia.n_markers, ia.n_contigs, ia.total_len, ia.k, ia.c, ia.marker_c, ia.has_seeds) == \
               (ib.n_seeds, ib.n_markers, ib.n_contigs, ib.total_len, ib.k, ib.c, ib.marker_c, ib.has_seeds)
        ea, eb = a.export(), b.export()
        for key in ea:
            assert np.array_equal(ea[key], eb[key]), key
    db_a, db_b = capi.Database(ctx), capi.Database(other)
    db_a.add_many(gs); db_b.add_many(got)
    assert db_a.query(gs) == db_b.query(got)
    assert len(other.unpack(ctx.pack([], None, 0), None, 0)) == 0
    # finished sketches of another context of the SAME device may be used (the pipelined all-vs-all queries on a second
    # context): same answers either way
    assert db_a.query(got[:2]) == db_b.query(got[:2]) == db_a.query(gs[:2])
    ok_a, sh_a = db_a.screen(got[:2])
    ok_b, sh_b = db_b.screen(gs[:2])
    assert np.array_equal(ok_a, ok_b) and np.array_equal(sh_a, sh_b)


def test_repeat_rich_window_takes_the_wide_dp_path(ctx):
    """A tandem repeat with every position a seed (c = 1) puts millions of anchors into one 20 kb window: the DP then
    runs its wide-score variant (packed score would overflow), keeps its per-anchor state in global memory (> 256
    anchors) and needs predecessors beyond the 32 held in registers for most anchors."""
    from pyskani_b200 import capi
    rng = np.random.default_rng(97)
    unit = synth.random_genome(20, 98)
    core = np.tile(unit, 450)
    q = np.concatenate([synth.random_genome(4000, 99), core, synth.random_genome(4000, 100)])
    r = np.concatenate([synth.random_genome(3000, 101), core, synth.random_genome(5000, 102)])
    r[rng.integers(0, len(r), 40)] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 40)]
    genomes = [[q.tobytes()], [r.tobytes()]]
    gs = ctx.sketch_batch(genomes, c=1, marker_c=50)
    os_ = [oracle.Sketch(g, c=1, marker_c=50) for g in genomes]
    for g, o in zip(gs, os_):
        assert_sketch_equal(g, o)
    db = capi.Database(ctx)
    db.add_many(gs)
    hits, n_in = db.query(gs)
    total = 0
    for qi, oq in enumerate(os_):
        total += check_hits([h for h in hits if h[0] == qi], oq, os_)
    assert total == n_in and n_in >= 2
    assert max(h[7] for h in hits) * 20 >= (1 << 22)           # anchors of a pair: the wide path was needed


def test_empty_and_degenerate_queries(ctx):
    from pyskani_b200 import capi
    s = rand(100_000, 77)
    full, empty, unrelated = ctx.sketch_batch([[s], [b"ATGC" * 100], [rand(100_000, 78)]])
    db = capi.Database(ctx)
    assert db.query([full])[0] == []          # empty database
    db.add(full)
    hits, n_in = db.query([empty])             # reference tests/test_database.py input: below the contig gate
    assert hits == [] and n_in == 1            # rescue_small lets it through the screen, chaining finds nothing
    hits, n_in = db.query([empty], faster_small=True)
    assert hits == [] and n_in == 0
    hits, n_in = db.query([unrelated])
    assert hits == [] and n_in == 0
    hits, n_in = db.query([full])
    assert len(hits) == 1 and abs(hits[0][2] - 1.0) < 1e-6 and hits[0][3] > 0.99


def test_all_vs_all_driver_single_rank(ctx):
    """pyskani_b200.parallel.all_vs_all with the CUDA backend (world size 1) against the oracle's query loop,
    including the export -> import path every other rank's sketches take in a multi-GPU run."""
    from pyskani_b200 import parallel
    genomes = []
    for f in range(3):
        base = synth.random_genome(200_000, 900 + f)
        genomes += [[base.tobytes()], [synth.mutate(base, 0.04, 950 + f).tobytes()]]
    be = parallel.CudaBackend(0)
    table = parallel.all_vs_all(genomes, be)
    os_ = [oracle.Sketch(g) for g in genomes]
    want = []
    for qi, oq in enumerate(os_):
        idx, res, _ = oracle.query(oq, os_)
        want += [(qi, int(i), r.ani, r.af_query, r.af_ref) for i, r in zip(idx, res)]
    want = np.asarray(want, np.float64)
    assert table.shape == want.shape and np.array_equal(table[:, :2], want[:, :2])
    assert np.abs(table[:, 2] - want[:, 2]).max() <= ANI_TOL and np.abs(table[:, 3:] - want[:, 3:]).max() <= AF_TOL
    # imported sketches behave exactly like freshly sketched ones
    local = be.sketch(genomes)
    imported = [be.import_(be.export(s)) for s in local]
    a = be.query(imported, local)
    b = be.query(local, local)
    assert len(a) and np.array_equal(a, b)


def test_anchor_capacity_rerun(ctx, monkeypatch):
    """The anchor arrays are sized from an estimate; a batch that overflows it is rerun with the exact size."""
    from pyskani_b200 import capi
    base = synth.random_genome(300_000, 41)
    refs = [synth.mutate(base, d, 410 + i) for i, d in enumerate((0.02, 0.08))]
    gs = ctx.sketch_batch([[r.tobytes()] for r in refs] + [[base.tobytes()]])
    db = capi.Database(ctx)
    for g in gs[:-1]:
        db.add(g)
    want = db.query([gs[-1]])
    monkeypatch.setenv("SKB_FORCE_ANCHOR_EST", "100")
    got = db.query([gs[-1]])
    assert got == want and len(want[0]) == 2


def test_anchor_join_with_unequal_genome_sizes(ctx):
    """Anchor join of a query far smaller than its references (a tile of its k-mers spans many reference buckets) beside a
    query of the references' size: hits, anchor, chain and window counts equal the oracle's."""
    from pyskani_b200 import capi
    base = synth.random_genome(1_200_000, 51)
    small = synth.mutate(base[:150_000], 0.03, 52)
    refs = [synth.mutate(base, d, 520 + i) for i, d in enumerate((0.01, 0.06))] + [synth.random_genome(400_000, 53)]
    contigs = [[r.tobytes()] for r in refs] + [[small.tobytes()], [base.tobytes()]]
    gs = ctx.sketch_batch(contigs)
    db = capi.Database(ctx)
    for g in gs[:3]:
        db.add(g)
    hits, _ = db.query(gs[3:])
    assert len(hits) == 4                                         # both queries hit both relatives, not the stranger
    osk = oracle.sketch_batch(contigs)
    for qi in (0, 1):
        check_hits([h for h in hits if h[0] == qi], osk[3 + qi], osk[:3])


def test_large_genomes_take_the_global_memory_paths(ctx):
    """Genomes beyond the shared-memory fast paths: > 49 152 seeds (window walk falls back to global memory) and
    > ~20 000 markers (the screen falls back to the warp-per-pair kernel)."""
    from pyskani_b200 import capi
    base = synth.random_genome(26_000_000, 71)
    refs = [synth.mutate(base, 0.03, 72), synth.random_genome(7_000_000, 73)]
    contigs = [[r.tobytes()] for r in refs] + [[base.tobytes()]]
    gs = ctx.sketch_batch(contigs)
    os_ = [oracle.Sketch(c) for c in contigs]
    assert gs[-1].info().n_seeds > 49152 and gs[-1].info().n_markers > 21000
    for g, o in zip(gs, os_):
        assert_sketch_equal(g, o)
    db = capi.Database(ctx)
    for g in gs[:-1]:
        db.add(g)
    ok, shared = db.screen([gs[-1]], 0.8, True)
    for j in range(2):
        want_ok, want_shared = oracle.screen(os_[-1], os_[j], 0.8, True)
        assert shared[0, j] == want_shared and ok[0, j] == want_ok
    hits, n_in = db.query([gs[-1]])
    assert check_hits(hits, os_[-1], os_[:-1]) == n_in
    assert len(hits) == 1
    # and the other way round: small query against the large reference
    db2 = capi.Database(ctx)
    db2.add(gs[-1])
    hits2, _ = db2.query([gs[0], gs[1]])
    for qi in range(2):
        check_hits([h for h in hits2 if h[0] == qi], os_[qi], [os_[-1]])
