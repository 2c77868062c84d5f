"""Sweep the oracle's structural knobs against the reference's five goldens (test_ani.py:28-61).

Build-container tool (uses tests/golden/ecoli_pair.npz).  Prints max |delta| over
AF_ref, AF_query, ANI(no learned), ANI(robust), ANI(median).
"""
import itertools, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from tests.fixtures import ecoli_pair

ec, k12, G = ecoli_pair()
R = oracle.Sketch([ec]); Q = oracle.Sketch([k12])


def evaluate(**kw):
    out = {}
    for name, flags in (("mean", {}), ("robust", {"robust": 1}), ("median", {"median": 1})):
        r = oracle.chain(R, Q, **kw, **flags)
        out[name] = r.ani_f64
        out["af_q"], out["af_r"] = r.af_query_f64, r.af_ref_f64
        out["nw"], out["nc"] = r.n_windows, r.n_chains
    d = {
        "af_r": out["af_r"] - G["af_ref"], "af_q": out["af_q"] - G["af_query"],
        "mean": out["mean"] - G["ani_no_learned"], "robust": out["robust"] - G["ani_robust"],
        "median": out["median"] - G["ani_median"],
    }
    return out, d


def show(kw):
    out, d = evaluate(**kw)
    worst = max(abs(v) for v in d.values())
    print(f"{worst:.5f} | " + " ".join(f"{k}={out[k]:.5f}({d[k]:+.5f})" for k in ("af_r", "af_q", "mean", "robust", "median"))
          + f" nw={out['nw']} nc={out['nc']} | {kw}")
    return worst


if __name__ == "__main__":
    grid = dict(
        chunk_mode=[0, 1, 2], order_by_ref=[0, 1], index_band=[20, 50, 100], bp_band=[2500, 5000],
        mean_mode=[0, 1, 2], count_mode=[0, 1],
    )
    keys = list(grid)
    res = []
    for vals in itertools.product(*[grid[k] for k in keys]):
        kw = dict(zip(keys, vals))
        out, d = evaluate(**kw)
        res.append((max(abs(d[k]) for k in ("mean", "robust", "median")), kw, out, d))
    res.sort(key=lambda t: t[0])
    for w, kw, out, d in res[:25]:
        print(f"{w:.5f} | " + " ".join(f"{k}={out[k]:.5f}({d[k]:+.5f})" for k in ("af_r", "af_q", "mean", "robust", "median")) + f" | {kw}")
