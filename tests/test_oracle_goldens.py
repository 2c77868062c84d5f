"""The CPU oracle against the reference's own known-answer test.

Reference: src/pyskani/tests/test_ani.py:28-61 — K-12 queried against a database holding EC590,
asserting reference_fraction 0.9246, query_fraction 0.9189 and identity 0.9946 (learned_ani=False),
0.9977 (robust), 0.9995 (median), each with assertAlmostEqual(places=4).  The learned-model golden
(0.9939) needs skani's embedded GBDT weights, which are not available here.
"""
import numpy as np
import pytest

import oracle


@pytest.fixture(scope="module")
def sketches(ecoli):
    ec, k12, gold = ecoli
    return oracle.Sketch([ec]), oracle.Sketch([k12]), gold


def almost(a, b, places=4):  # unittest.assertAlmostEqual semantics
    return round(abs(a - b), places) == 0


def test_fixture_identity(ecoli):
    import hashlib
    ec, k12, _ = ecoli
    assert len(ec) == 4617703 and hashlib.sha256(ec).hexdigest().startswith("6ec76febfe69cd16")
    assert len(k12) == 4646332 and hashlib.sha256(k12).hexdigest().startswith("1dd7c87af0a25051")


def test_sketch_sizes_match_survey_probe(sketches):
    # SURVEY.md Appendix B (probed on the same fixtures)
    R, Q, _ = sketches
    assert (R.n_seeds, R.n_markers) == (37237, 4539)
    assert (Q.n_seeds, Q.n_markers) == (37384, 4551)


def test_screen_passes(sketches):
    R, Q, _ = sketches
    ok, shared = oracle.screen(Q, R, 0.8, True)
    assert ok and shared == 4085
    assert shared > 0.8 ** 21 * R.n_markers


def test_no_learned_ani(sketches):
    R, Q, g = sketches
    r = oracle.chain(R, Q)
    assert r.n_anchors == 43220
    assert almost(r.af_ref, g["af_ref"]), r.af_ref
    assert almost(r.af_query, g["af_query"]), r.af_query
    assert almost(r.ani, g["ani_no_learned"]), r.ani


def test_robust(sketches):
    R, Q, g = sketches
    r = oracle.chain(R, Q, robust=1)
    assert almost(r.af_ref, g["af_ref"]) and almost(r.af_query, g["af_query"])
    assert almost(r.ani, g["ani_robust"]), r.ani


def test_median(sketches):
    R, Q, g = sketches
    r = oracle.chain(R, Q, median=1)
    assert almost(r.af_ref, g["af_ref"]) and almost(r.af_query, g["af_query"])
    assert almost(r.ani, g["ani_median"]), r.ani


def test_query_loop_one_hit(sketches):
    R, Q, g = sketches
    idx, res, n_in = oracle.query(Q, [R], 0.8, True)
    assert list(idx) == [0] and n_in == 1
    assert almost(res[0].ani, g["ani_no_learned"])
