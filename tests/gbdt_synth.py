"""Synthetic gbdt-rs ensembles for the learned-ANI tests (the real weights live inside the skani crate and are not part
of the reference tree): random regression trees over the 10 pair features, dumped in gbdt-rs' serde_json layout."""
import json

import numpy as np

# plausible ranges of the 10 features (ANI %, std %, ref q90/q50/q10, query q90/q50/q10, bases per chain, bases covered)
FEATURE_RANGES = [(80.0, 100.0), (0.0, 3.0)] + [(500.0, 6e6)] * 6 + [(300.0, 60000.0), (1e5, 6e6)]


def f32(x):
    return float(np.float32(x))


def random_tree(rng, n_features, max_depth, leaf_scale):
    nodes = []

    def build(depth):
        idx = len(nodes)
        nodes.append(None)
        leaf = depth >= max_depth or (depth > 0 and rng.random() < 0.15)
        fi = int(rng.integers(0, n_features))
        lo, hi = FEATURE_RANGES[fi]
        node = {"value": {"feature_index": fi, "feature_value": f32(rng.uniform(lo, hi)), "pred": f32(rng.normal(0, leaf_scale)),
                          "missing": int(rng.integers(-1, 2)), "is_leaf": bool(leaf)}, "index": idx, "left": 0, "right": 0}
        nodes[idx] = node
        if not leaf:
            node["left"] = build(depth + 1)
            node["right"] = build(depth + 1)
        return idx
    build(0)
    return {"tree": {"tree": nodes}, "feature_size": n_features, "max_depth": max_depth, "min_leaf_size": 1, "loss": "SquaredError",
            "feature_sample_ratio": 1.0}


def random_model(seed, n_trees=60, n_features=10, max_depth=5, bias=97.0, shrinkage=0.1, leaf_scale=1.5, iterations=None):
    rng = np.random.default_rng(seed)
    conf = {"feature_size": n_features, "max_depth": max_depth, "iterations": n_trees if iterations is None else iterations,
            "shrinkage": f32(shrinkage), "feature_sample_ratio": 1.0, "data_sample_ratio": 1.0, "min_leaf_size": 1, "loss": "SquaredError",
            "debug": False, "initial_guess_enabled": False, "training_optimization_level": 2}
    return json.dumps({"conf": conf, "trees": [random_tree(rng, n_features, max_depth, leaf_scale) for _ in range(n_trees)], "bias": f32(bias)})


def identity_like_model(delta=-0.07):
    """One stump per step that reproduces 'ANI % + delta' piecewise: prediction = bias + shrinkage * leaf; used to see the
    correction end to end.  Two leaves on feature 0 (ANI %): below 99 -> bias + delta, else bias + 2 delta."""
    conf = {"feature_size": 10, "max_depth": 1, "iterations": 1, "shrinkage": 1.0, "feature_sample_ratio": 1.0, "data_sample_ratio": 1.0,
            "min_leaf_size": 1, "loss": "SquaredError", "debug": False, "initial_guess_enabled": False, "training_optimization_level": 2}
    nodes = [{"value": {"feature_index": 0, "feature_value": 99.0, "pred": 0.0, "missing": 0, "is_leaf": False}, "index": 0, "left": 1, "right": 2},
             {"value": {"feature_index": 0, "feature_value": 0.0, "pred": f32(delta), "missing": 0, "is_leaf": True}, "index": 1, "left": 0, "right": 0},
             {"value": {"feature_index": 0, "feature_value": 0.0, "pred": f32(2 * delta), "missing": 0, "is_leaf": True}, "index": 2, "left": 0, "right": 0}]
    tree = {"tree": {"tree": nodes}, "feature_size": 10, "max_depth": 1, "min_leaf_size": 1, "loss": "SquaredError", "feature_sample_ratio": 1.0}
    return json.dumps({"conf": conf, "trees": [tree], "bias": 95.0})


def random_rows(seed, n, unknown_frac=0.02):
    rng = np.random.default_rng(seed)
    rows = np.empty((n, 10), np.float32)
    for f, (lo, hi) in enumerate(FEATURE_RANGES):
        rows[:, f] = rng.uniform(lo, hi, n).astype(np.float32)
    mask = rng.random((n, 10)) < unknown_frac
    rows[mask] = np.float32(-3.40282347e+38)          # gbdt-rs VALUE_TYPE_UNKNOWN
    return rows
