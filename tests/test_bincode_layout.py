"""the on-disk layout at byte level.

The product's writer / reader (pyskani_b200/ext/skani_module.cpp) is compared with tests/bincode_ref.py, an independent
bincode-1.3 encoder / decoder written from SURVEY.md Appendix C.  CPU part: the extension's `_encode_sketch` /
`_decode_sketch` / `_encode_index` hooks (the very functions save / flush / load / open call) on host arrays.  GPU part:
the files a real Database writes - `markers.bin`, `<name>.sketch`, `sketches.db`, `index.db` - decoded by the reference
decoder and compared with the device sketches; the product's own order (k-mers and markers ascending) makes the files
byte-reproducible, so they must also equal the reference encoder's output byte for byte.
Field order inside SketchParams / Sketch / SeedPosition is recalled from skani v0.3.0 (not in /root/reference): this
pins the writer to that description, not to a real skani file (none is available here)."""
import os

import numpy as np
import pytest

from tests import bincode_ref as B


def toy_sketch(rng, name, n_kmers, n_contigs, n_markers):
    kmers = np.unique(rng.integers(0, 1 << 30, n_kmers).astype(np.uint64))
    seeds = {}
    for km in kmers:
        m = 1 if rng.random() < 0.8 else int(rng.integers(2, 4))
        lst = sorted((int(rng.integers(0, n_contigs)), int(rng.integers(20, 1 << 22))) for _ in range(m))
        seeds[int(km)] = [(pos, bool(rng.integers(0, 2)), ci) for ci, pos in lst]
    lens = [int(x) for x in rng.integers(500, 1 << 22, n_contigs)]
    markers = sorted(set(int(x) for x in rng.integers(0, 1 << 42, n_markers)))
    return dict(name=name, seeds=seeds, contigs=["%s_%d" % (name, i) for i in range(n_contigs)], total_len=sum(lens),
                contig_lengths=lens, markers=markers)


def flat(seeds):
    kmer, pos, contig, canon = [], [], [], []
    for km in sorted(seeds):
        for p, c, ci in seeds[km]:
            kmer.append(km); pos.append(p); contig.append(ci); canon.append(int(c))
    return kmer, pos, contig, canon


@pytest.fixture(scope="module")
def ext():
    import pyskani_b200._skani as e
    return e


@pytest.mark.parametrize("c,k,mc", [(125, 15, 1000), (30, 14, 200)])
def test_writer_bytes_equal_the_reference_encoder(ext, c, k, mc):
    rng = np.random.default_rng(c)
    for t in (toy_sketch(rng, "g1", 400, 1, 50), toy_sketch(rng, "dir/g two", 1500, 7, 200), toy_sketch(rng, "empty", 0, 0, 0)):
        kmer, pos, contig, canon = flat(t["seeds"])
        want = B.encode_params(c, k, mc) + B.encode_sketch(t["name"], t["seeds"], t["contigs"], t["total_len"], t["contig_lengths"],
                                                           t["markers"], c, k, mc)
        got = ext._encode_sketch(t["name"], c, k, mc, True, kmer, pos, contig, canon, t["contigs"], t["total_len"], t["contig_lengths"],
                                 t["markers"], False, True)
        assert got == want
        # markers-only form (markers.bin elements, lib.rs:187-197): Option tag 0, no params in front
        want_m = B.encode_sketch(t["name"], None, t["contigs"], t["total_len"], t["contig_lengths"], t["markers"], c, k, mc)
        got_m = ext._encode_sketch(t["name"], c, k, mc, True, kmer, pos, contig, canon, t["contigs"], t["total_len"], t["contig_lengths"],
                                   t["markers"], True, False)
        assert got_m == want_m
        # sizes follow Appendix C: 10 bytes per SeedPosition, 16 per map entry header
        n_pos, n_km = len(kmer), len(t["seeds"])
        assert len(got) - len(got_m) - len(B.encode_params(c, k, mc)) == 8 + 16 * n_km + 10 * n_pos


def test_reader_accepts_any_hash_order(ext):
    """the reference serialises HashMap / HashSet in arbitrary order: the reader must not depend on ours"""
    rng = np.random.default_rng(7)
    t = toy_sketch(rng, "shuffled", 800, 3, 120)
    seed_order = list(t["seeds"]); rng.shuffle(seed_order)
    marker_order = list(t["markers"]); rng.shuffle(marker_order)
    raw = B.encode_params(125, 15, 1000) + B.encode_sketch(t["name"], t["seeds"], t["contigs"], t["total_len"], t["contig_lengths"],
                                                           t["markers"], 125, 15, 1000, seed_order, marker_order)
    d = ext._decode_sketch(raw, True)
    assert d["consumed"] == len(raw) and d["params"] == (125, 15, 1000) and d["sketch_params"] == (125, 15, 1000)
    assert d["file_name"] == "shuffled" and d["has_seeds"] and d["contigs"] == t["contigs"] and d["total_len"] == t["total_len"]
    assert d["contig_lengths"] == t["contig_lengths"] and sorted(d["markers"]) == t["markers"]
    got = {}
    for km, p, ci, cn in zip(d["kmer"], d["pos"], d["contig"], d["canonical"]):
        got.setdefault(km, []).append((p, bool(cn), ci))
    assert got == t["seeds"]
    with pytest.raises(ValueError):
        ext._decode_sketch(raw[:-3], True)
    with pytest.raises(ValueError):
        ext._decode_sketch(b"\x05" * 40, True)


def test_index_bytes(ext):
    entries = [("b", 1000, 77), ("a", 0, 1000), ("name with space", 1077, 5)]
    assert ext._encode_index(sorted(entries, key=lambda e: e[1])) == B.encode_index(entries)
    assert B.decode_index_file(B.encode_index(entries)) == sorted(entries, key=lambda e: e[1])


# ------------------------------------------------------------------------------------------------ real files (GPU)
def export_truth(genomes):
    """sketches of the genomes through the C ABI, as the dict form bincode_ref uses"""
    from pyskani_b200 import capi
    ctx = capi.Context(0)
    out = {}
    for name, contigs in genomes.items():
        (g,) = ctx.sketch_batch([contigs])
        e = g.export()
        seeds = {}
        for km, p, ci, cn in zip(e["kmer"].tolist(), e["pos"].tolist(), e["contig"].tolist(), e["canonical"].tolist()):
            seeds.setdefault(km, []).append((p, bool(cn), ci))
        kept = [i for i, c in enumerate(contigs) if len(c) >= 500]
        out[name] = dict(name=name, seeds=seeds, contigs=["%s_%d" % (name, i) for i in kept], total_len=int(g.info().total_len),
                         contig_lengths=e["contig_lengths"].tolist(), markers=e["markers"].tolist())
    return out


def check_sketch(p, s, truth, markers_only=False, c=125, k=15, mc=1000):
    assert (p["c"], p["k"], p["marker_c"], p["use_syncs"], p["use_aa"], p["orf_size"]) == (c, k, mc, 0, 0, 30)
    assert len(p["encoding"]) == 64 and p["letters"] == B.CODON_TABLE.encode()
    assert s["file_name"] == truth["name"] and s["contigs"] == truth["contigs"] and s["total_len"] == truth["total_len"]
    assert s["contig_lengths"] == truth["contig_lengths"] and sorted(s["markers"]) == truth["markers"]
    assert (s["marker_c"], s["c"], s["k"], s["contig_order"], s["amino_acid"], s["repetitive_kmers"]) == (mc, c, k, 0, 0, 0)
    if markers_only:
        assert s["seeds"] is None
    else:
        assert s["seeds"] == truth["seeds"]


@pytest.mark.gpu
def test_files_written_by_a_database(tmp_path):
    import pyskani_b200 as pyskani
    from pyskani_b200 import synth
    base = synth.random_genome(300_000, 5150)
    # a repeat gives k-mers with several positions (SmallVec longer than 1); short contigs are dropped from the contig list
    rep = base[1000:31000]
    genomes = {"alpha": [base.tobytes()],
               "beta gamma": [synth.mutate(base, 0.03, 5151)[:120_000].tobytes(), b"ACGT" * 20, rep.tobytes(), rep.tobytes()]}
    truth = export_truth(genomes)
    assert any(len(v) > 1 for v in truth["beta gamma"]["seeds"].values())

    def sketch_all(db):
        for name, contigs in genomes.items():
            db.sketch(name, *contigs)

    # --- separated: <name>.sketch written at sketch() time, markers.bin at flush()
    sep = tmp_path / "sep"
    db = pyskani.Database(str(sep), format="separated")
    sketch_all(db)
    db.flush()
    assert sorted(os.listdir(sep)) == ["alpha.sketch", "beta gamma.sketch", "markers.bin"]
    for name, t in truth.items():
        raw = (sep / (name + ".sketch")).read_bytes()
        p, s = B.decode_sketch_file(raw)
        check_sketch(p, s, t)
        assert raw == B.encode_params(125, 15, 1000) + B.encode_sketch(t["name"], t["seeds"], t["contigs"], t["total_len"],
                                                                       t["contig_lengths"], t["markers"], 125, 15, 1000)
    mraw = (sep / "markers.bin").read_bytes()
    p, sk = B.decode_markers_file(mraw)
    assert [s["file_name"] for s in sk] == list(genomes)
    for s in sk:
        check_sketch(p, s, truth[s["file_name"]], markers_only=True)
    assert mraw == B.encode_params(125, 15, 1000) + B.u64(2) + b"".join(
        B.encode_sketch(t["name"], None, t["contigs"], t["total_len"], t["contig_lengths"], t["markers"], 125, 15, 1000) for t in truth.values())

    # --- consolidated: sketches.db grows at sketch() time, index.db + markers.bin at flush()
    con = tmp_path / "con"
    with pyskani.Database(str(con), format="consolidated") as db:
        sketch_all(db)
    assert sorted(os.listdir(con)) == ["index.db", "markers.bin", "sketches.db"]
    blob = (con / "sketches.db").read_bytes()
    index = B.decode_index_file((con / "index.db").read_bytes())
    assert [e[0] for e in index] == list(genomes) and index[0][1] == 0 and index[1][1] == index[0][2] and index[1][1] + index[1][2] == len(blob)
    for name, off, length in index:
        p, s = B.decode_sketch_file(blob[off:off + length])
        check_sketch(p, s, truth[name])
    assert (con / "markers.bin").read_bytes() == mraw
    assert blob == b"".join((sep / (n + ".sketch")).read_bytes() for n in genomes)

    # --- save(): the reference's swapped names (lib.rs:696-699) and the strict variant produce the same bytes as above
    mem = pyskani.Database()
    sketch_all(mem)
    for fmt, strict, separate in ((None, False, True), ("separated", False, False), ("separated", True, True), ("consolidated", True, False)):
        out = tmp_path / ("save_%s_%s" % (fmt, strict))
        mem.save(str(out), format=fmt, strict_format=strict)
        assert (out / "markers.bin").read_bytes() == mraw
        if separate:
            assert sorted(os.listdir(out)) == ["alpha.sketch", "beta gamma.sketch", "markers.bin"]
            for n in genomes:
                assert (out / (n + ".sketch")).read_bytes() == (sep / (n + ".sketch")).read_bytes()
        else:
            assert (out / "sketches.db").read_bytes() == blob and (out / "index.db").read_bytes() == (con / "index.db").read_bytes()

    # --- a file in the reference's arbitrary hash order loads and answers like the original
    rng = np.random.default_rng(3)
    shuf = tmp_path / "shuffled"
    shuf.mkdir()
    recs = []
    for t in truth.values():
        so = list(t["seeds"]); rng.shuffle(so)
        mo = list(t["markers"]); rng.shuffle(mo)
        (shuf / (t["name"] + ".sketch")).write_bytes(B.encode_params(125, 15, 1000) + B.encode_sketch(
            t["name"], t["seeds"], t["contigs"], t["total_len"], t["contig_lengths"], t["markers"], 125, 15, 1000, so, mo))
        recs.append(B.encode_sketch(t["name"], None, t["contigs"], t["total_len"], t["contig_lengths"], t["markers"], 125, 15, 1000, None, mo))
    (shuf / "markers.bin").write_bytes(B.encode_params(125, 15, 1000) + B.u64(len(recs)) + b"".join(recs))
    q = synth.mutate(base, 0.02, 5152).tobytes()
    want = [(h.reference_name, h.identity, h.query_fraction, h.reference_fraction) for h in mem.query("q", q, learned_ani=False)]
    got = [(h.reference_name, h.identity, h.query_fraction, h.reference_fraction) for h in pyskani.Database.load(str(shuf)).query("q", q, learned_ani=False)]
    assert len(want) == 2 and got == want
