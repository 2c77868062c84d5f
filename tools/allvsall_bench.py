"""All-vs-all micro-benchmark (BASELINE.json config 3 shape, scaled): F families x 10 genomes, every ordered pair."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyskani_b200 import capi, synth
from concurrent.futures import ThreadPoolExecutor

F = int(sys.argv[1]) if len(sys.argv) > 1 else 20
glen = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
divs = [0.01, 0.02, 0.03, 0.05, 0.07, 0.09, 0.11, 0.13, 0.15]
t0 = time.time()
def fam(f):
    base = synth.random_genome(glen, 10_000 + f)
    return [base] + [synth.mutate(base, d, 20_000 + 100 * f + i) for i, d in enumerate(divs)]
with ThreadPoolExecutor(16) as ex:
    genomes = [g for fam_ in ex.map(fam, range(F)) for g in fam_]
print("generated", len(genomes), "genomes in %.1f s" % (time.time() - t0))
ctx = capi.Context(0)
t0 = time.perf_counter()
sk = ctx.sketch_batch([[g] for g in genomes])
t1 = time.perf_counter()
st = ctx.stats()
print("sketch: wall %.1f ms  (seed %.2f ms, total dev %.2f ms)  %.1f Gbp/s e2e" % (1e3 * (t1 - t0), st.seed_ms, st.total_ms, len(genomes) * glen / (t1 - t0) / 1e9))
db = capi.Database(ctx)
for s in sk:
    db.add(s)
for it in range(3):
    t0 = time.perf_counter()
    hits, n_in = db.query(sk)
    t1 = time.perf_counter()
    st = ctx.stats()
    n = len(genomes)
    print("query: wall %.1f ms  screen %.2f ms  chain %.2f ms | %d pairs screened (%.2e pairs/s), %d chained (%.0f pairs/s), %d hits" % (
        1e3 * (t1 - t0), st.screen_ms, st.chain_ms, n * n, n * n / (t1 - t0), n_in, n_in / max(1e-9, st.chain_ms / 1e3), len(hits)))
fam_ok = all(h[0] // 10 == h[1] // 10 for h in hits)
print("all hits intra-family:", fam_ok)
