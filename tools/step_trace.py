"""Wall-clock vs device time of the two calls of a device-resident bench step (where do the host gaps go?)."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyskani_b200 import capi, synth
n_g, glen = 101, 5_000_000
ctx = capi.Context(0)
base = synth.random_genome(glen, 1)
genomes = [synth.mutate(base, 0.01 + 0.14 * j / 99, 2 + j) for j in range(100)] + [base]
lens = np.array([len(g) for g in genomes], np.uint64)
offs = np.zeros(n_g, np.uint64); cur = 64
for i, l in enumerate(lens):
    offs[i] = cur; cur += (int(l) + 15) // 16 * 16 + 16
buf = np.zeros(cur + 64, np.uint8)
for g, o in zip(genomes, offs): buf[int(o):int(o) + len(g)] = g
d = ctx.dev_alloc(buf.size); ctx.memcpy_h2d(d, buf.ctypes.data, buf.size)
gs = np.arange(n_g + 1, dtype=np.uint32)
for it in range(6):
    t0 = time.perf_counter()
    sk = ctx.sketch_batch_device(d, gs, offs, lens)
    t1 = time.perf_counter(); st = ctx.stats()
    db = capi.Database(ctx); db.add_many(sk[:-1])
    t2 = time.perf_counter()
    hits, n_in = db.query([sk[-1]])
    t3 = time.perf_counter(); st2 = ctx.stats()
    del db; del sk
    t4 = time.perf_counter()
    print("sketch wall %.3f dev %.3f (seed %.3f) | add %.3f | query wall %.3f dev %.3f (screen %.3f chain %.3f) | free %.3f | total %.3f" % (
        1e3*(t1-t0), st.total_ms, st.seed_ms, 1e3*(t2-t1), 1e3*(t3-t2), st2.total_ms, st2.screen_ms, st2.chain_ms, 1e3*(t4-t3), 1e3*(t4-t0)))
os.environ["SKB_TRACE"] = "1"
sk = ctx.sketch_batch_device(d, gs, offs, lens)
