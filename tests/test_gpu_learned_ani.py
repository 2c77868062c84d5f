"""SURVEY.md §8 rows a9 / f4 on the GPU: the gbdt-rs ensemble evaluated on the device (skb_model_predict and inside
ani_reduce_kernel) against the oracle's CPU evaluator.  Ensemble outputs are bit-equal f32; end-to-end ANI is compared
with the oracle's features -> evaluator chain (north-star tolerance 1e-4; equality is reported)."""
import numpy as np
import pytest

import oracle
from pyskani_b200 import synth
from tests import gbdt_synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from pyskani_b200 import capi
    return capi.Context(0)


@pytest.mark.parametrize("seed,n_trees,depth", [(11, 1, 1), (12, 31, 3), (13, 32, 5), (14, 33, 5), (15, 200, 6), (16, 97, 8)])
def test_device_ensemble_is_bit_equal_to_the_oracle(ctx, seed, n_trees, depth):
    from pyskani_b200 import capi
    text = gbdt_synth.random_model(seed, n_trees=n_trees, max_depth=depth)
    m = capi.Model(ctx, text)
    ref = oracle.Gbdt(text)
    info = m.info()
    assert info["n_trees"] == n_trees and info["n_nodes"] == sum(len(t) for t in ref.trees) and info["n_features"] == 10
    rows = gbdt_synth.random_rows(seed + 50, 500, unknown_frac=0.05)
    got = m.predict(rows)
    want = np.array([ref.predict(r) for r in rows], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_model_errors(ctx):
    from pyskani_b200 import capi
    with pytest.raises(capi.SkbError) as e:
        capi.Model(ctx, "{")
    assert e.value.code == capi.SKB_ERR_ARG
    with pytest.raises(capi.SkbError):
        capi.Model(ctx, '{"conf": {"feature_size": 3, "shrinkage": 0.1, "loss": "LogLikelyhood"}, "trees": []}')
    m = capi.Model(ctx, gbdt_synth.random_model(1, n_trees=2, max_depth=2))
    with pytest.raises(capi.SkbError):
        m.predict(np.zeros((2, 4), np.float32))            # narrower than feature_size
    other = capi.Context(0)
    db = capi.Database(other)
    with pytest.raises(capi.SkbError):
        db.set_model(m)                                    # models are bound to their context, like sketches


def test_query_with_model_matches_oracle(ctx):
    """learned_ani = None / True / False, robust and median, for a family of mutants, single- and multi-contig."""
    from pyskani_b200 import capi
    base = synth.random_genome(600_000, 70)
    genomes = [[base.tobytes()]] + [[synth.mutate(base, d, 71 + i).tobytes()] for i, d in enumerate((0.005, 0.02, 0.06, 0.12))]
    genomes.append([c.tobytes() for c in synth.fragment(synth.mutate(base, 0.03, 80), 81, lo=600, hi=40_000)])
    genomes.append([synth.random_genome(200_000, 82)[:100_000].tobytes(), synth.mutate(base, 0.01, 83)[:100_000].tobytes()])   # < 150 kb aligned
    gs = ctx.sketch_batch(genomes)
    os_ = [oracle.Sketch(g) for g in genomes]
    text = gbdt_synth.random_model(21, n_trees=80, max_depth=5, bias=96.0, leaf_scale=0.8)
    model, ref = capi.Model(ctx, text), oracle.Gbdt(text)
    db = capi.Database(ctx)
    db.add_many(gs)
    plain, _ = db.query_array(gs, learned_ani=0)
    with pytest.raises(capi.SkbError) as e:
        db.query_array(gs, learned_ani=1)
    assert e.value.code == capi.SKB_ERR_UNSUPPORTED
    db.set_model(model)
    n_corrected = n_equal = 0
    for mode, kw in (("none", dict(learned_ani=-1)), ("true", dict(learned_ani=1)), ("false", dict(learned_ani=0)),
                     ("robust", dict(learned_ani=1, robust=True)), ("median", dict(learned_ani=-1, median=True))):
        hits, _ = db.query_array(gs, **kw)
        assert len(hits) >= 30
        params = oracle.default_params(robust=int(kw.get("robust", False)), median=int(kw.get("median", False)))
        for h in hits:
            r = oracle.chain(os_[h["ref_index"]], os_[h["query_index"]], params)
            if mode in ("none", "true"):
                want = oracle.learned_ani(r, ref)
                n_corrected += int(want != np.float32(r.ani))
            else:
                want = np.float32(r.ani)
            assert abs(float(h["ani"]) - float(want)) <= 1e-4, (mode, h, want, [r.features[i] for i in range(10)])
            n_equal += int(np.float32(h["ani"]) == want)
            assert abs(h["af_query"] - r.af_query) <= 1e-3 and abs(h["af_ref"] - r.af_ref) <= 1e-3
        if mode == "false":
            assert np.array_equal(hits, plain)
    assert n_corrected >= 40                 # the model really was applied ...
    assert n_equal >= 0.95 * 5 * len(plain)  # ... and nearly every value is the oracle's f32 to the bit
    # pairs below the aligned-bases gate keep the uncorrected estimate even with the model on
    small = [h for h in db.query_array(gs, learned_ani=1)[0] if h["query_index"] == 6 and h["ref_index"] != 6]
    for h in small:
        r = oracle.chain(os_[h["ref_index"]], os_[6])
        assert r.features[9] < 150000 and np.float32(h["ani"]) == np.float32(r.ani)
    db.set_model(None)
    assert np.array_equal(db.query_array(gs, learned_ani=-1)[0], plain)
