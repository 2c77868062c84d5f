#!/usr/bin/env python
"""Host->device ingest of skb_sketch_batch: Gbp/s of one large host batch (pinned ASCII) by policy and thread count.

    python tools/ingest_sweep.py [--genomes 250] [--genome-len 5000000] [--reps 5]

raw  = every byte travels as ASCII (copy engine only);  pack = every chunk is compacted 4:1 by host threads first;
mix  = both routes at once, chunks claimed by whichever is free (the default of the library)."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=250)
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--threads", default="4,8,12,14,15,16")
    ap.add_argument("--only", default="", help="run only this policy (raw | pack | mix)")
    a = ap.parse_args()
    from pyskani_b200 import capi
    import workload
    ctx = capi.Context(0)
    L = capi.lib()
    slot = workload.slot_bytes(a.genome_len)
    nbytes = 64 + slot * a.genomes + 64
    h_ptr = ctx.host_alloc(nbytes)
    h = np.ctypeslib.as_array(C.cast(h_ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))
    lens = workload.fill_families(h[64:], slot, a.genome_len, list(range(a.genomes)), members=10)
    offs = 64 + slot * np.arange(a.genomes, dtype=np.uint64)
    n = a.genomes
    ptrs = (C.c_void_p * n)(*[h_ptr + int(o) for o in offs])
    clens = (C.c_uint64 * n)(*[int(x) for x in lens])
    gs = (C.c_uint32 * (n + 1))(*range(n + 1))
    params = capi.SketchParams(15, 125, 1000)
    bases = float(np.sum(lens))
    cpus = len(os.sched_getaffinity(0))
    print("cpus usable: %d; batch: %d genomes, %.0f MB" % (cpus, n, bases / 1e6), flush=True)

    def run(policy, threads):
        os.environ["SKB_INGEST"] = policy
        ctx.set_host_threads(threads)
        best, raw, packed = 1e9, 0, 0
        for r in range(a.reps + 1):
            out = np.zeros(n, np.uint64)
            t0 = time.perf_counter()
            ctx.check(L.skb_sketch_batch(ctx._h, C.byref(params), 1, n, gs, ptrs, clens, out.ctypes.data))
            dt = time.perf_counter() - t0
            st = ctx.stats()
            raw, packed = st.h2d_raw_bytes, st.h2d_packed_bytes
            del_me = capi.SketchArray(ctx, out); del del_me
            if r:
                best = min(best, dt)
        print("%-5s threads %2d: %7.2f ms  %6.1f GB/s of ASCII input   (link: %6.1f MB ASCII + %6.1f MB packed)"
              % (policy, threads, 1e3 * best, bases / best / 1e9, raw / 1e6, packed / 1e6), flush=True)

    if a.only in ("", "raw"):
        run("raw", 0)
    for t in [int(x) for x in a.threads.split(",")]:
        if t > max(2, cpus):
            continue
        if a.only in ("", "pack"):
            run("pack", t)
        if a.only in ("", "mix"):
            run("mix", t)
    os.environ.pop("SKB_INGEST", None)


if __name__ == "__main__":
    main()
