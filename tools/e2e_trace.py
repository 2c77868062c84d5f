"""Host-side timeline of one host-buffer sketch call (SKB_TRACE=1)."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyskani_b200 import capi
n_g, glen = 101, 5_000_000
ctx = capi.Context(0)
stride = (glen + 15) // 16 * 16 + 16
nbytes = 64 + stride * n_g + 64
h = ctx.host_alloc(nbytes)
arr = np.ctypeslib.as_array(C.cast(h, C.POINTER(C.c_uint8)), shape=(nbytes,))
arr[:] = np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(0).integers(0, 4, nbytes, dtype=np.uint8)]
ptrs = (C.c_void_p * n_g)(*[h + 64 + stride * i for i in range(n_g)])
lens = (C.c_uint64 * n_g)(*[glen] * n_g)
gs = (C.c_uint32 * (n_g + 1))(*range(n_g + 1))
p = capi.SketchParams(15, 125, 1000)
L = capi.lib()
for it in range(4):
    out = (C.c_void_p * n_g)()
    if it == 3: os.environ["SKB_TRACE"] = "1"
    t0 = time.perf_counter()
    ctx.check(L.skb_sketch_batch(ctx._h, C.byref(p), 1, n_g, gs, ptrs, lens, out))
    print("call wall ms", 1e3 * (time.perf_counter() - t0), "stats", ctx.stats().seed_ms, ctx.stats().index_ms, ctx.stats().total_ms)
    sk = [capi.Sketch(ctx, out[i]) for i in range(n_g)]
    t0 = time.perf_counter(); del sk; print("free ms", 1e3 * (time.perf_counter() - t0))
