"""Latency of the reference's own call pattern: Database.query(name, sequence) = sketch ONE genome + screen + chain against
a resident database (BASELINE.json configs[0] shape, synthetic data).  Prints per-call wall times through the ctypes ABI
and through the pyskani-compatible extension."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyskani_b200 import capi, synth
import pyskani_b200 as pyskani

n_refs = int(sys.argv[1]) if len(sys.argv) > 1 else 100
glen = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
base = synth.random_genome(glen, 1)
divs = np.linspace(0.01, 0.15, n_refs)
refs = [synth.mutate(base, float(d), 100 + i) if i < 8 else synth.random_genome(glen, 1000 + i) for i, d in enumerate(divs)]
q = base.tobytes()

ctx = capi.Context(0)
sk = ctx.sketch_batch([[r.tobytes()] for r in refs])
db = capi.Database(ctx)
db.add_many(sk)
for it in range(8):
    t0 = time.perf_counter()
    (qs,) = ctx.sketch_batch([[q]])
    t1 = time.perf_counter()
    st = ctx.stats()
    hits, n_in = db.query([qs])
    t2 = time.perf_counter()
    st2 = ctx.stats()
    print("capi: sketch %.3f ms (device: seed %.3f, total %.3f) | query %.3f ms (device %.3f) | %d hits" % (
        1e3 * (t1 - t0), st.seed_ms, st.total_ms, 1e3 * (t2 - t1), st2.total_ms, len(hits)))

pdb = pyskani.Database()
for i, r in enumerate(refs):
    pdb.sketch("ref%d" % i, r.tobytes())
for it in range(5):
    t0 = time.perf_counter()
    hits = pdb.query("q", q, learned_ani=False)
    t1 = time.perf_counter()
    print("extension: Database.query %.3f ms, %d hits, best identity %.4f" % (1e3 * (t1 - t0), len(hits), max(h.identity for h in hits)))
