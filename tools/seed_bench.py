"""Kernel micro-benchmark: seeding + index build on device-resident random genomes (not a bench.py number)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyskani_b200 import capi

n_g = int(sys.argv[1]) if len(sys.argv) > 1 else 100
glen = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
ctx = capi.Context(0)
rng = np.random.default_rng(0)
stride = (glen + 15) // 16 * 16 + 16
buf = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 64 + stride * n_g + 64, dtype=np.uint8)]
d = ctx.dev_alloc(buf.size)
ctx.memcpy_h2d(d, buf.ctypes.data, buf.size)
offs = 64 + stride * np.arange(n_g, dtype=np.uint64)
lens = np.full(n_g, glen, np.uint64)
gs = np.arange(n_g + 1, dtype=np.uint32)
res = []
for it in range(6):
    t0 = time.perf_counter()
    sk = ctx.sketch_batch_device(d, gs, offs, lens)
    wall = 1e3 * (time.perf_counter() - t0)
    st = ctx.stats()
    res.append((st.seed_ms, st.index_ms, st.total_ms, wall))
    tot = sum(s.info().n_seeds for s in sk)
    del sk
for r in res:
    print("seed %.3f ms  index %.3f ms  total %.3f ms  wall %.3f ms" % r)
best = min(r[0] for r in res)
print("bases %d seeds %d  seed kernel %.1f Gbp/s" % (n_g * glen, tot, n_g * glen / best / 1e6))
