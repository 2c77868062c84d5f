#!/usr/bin/env python
"""Host-side timeline of one skb_db_query of Q queries against a fresh database of N genomes (the per-rank shape of a
multi-GPU all-vs-all: N = 1000, Q = 1000 / ranks).  Run with SKB_TRACE=1 to see the marks of run_screen / db_query.

    SKB_TRACE=1 python tools/screen_trace.py [--genomes 1000] [--queries 125] [--reps 3]"""
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
    ap.add_argument("--genomes", type=int, default=1000)
    ap.add_argument("--queries", type=int, default=125)
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from pyskani_b200 import capi
    import workload
    ctx = capi.Context(0)
    slot = workload.slot_bytes(a.genome_len)
    parts = []
    trace = os.environ.pop("SKB_TRACE", None)
    for b0 in range(0, a.genomes, 250):
        n = min(250, a.genomes - b0)
        buf = np.zeros(64 + slot * n + 64, np.uint8)
        lens = workload.fill_families(buf[64:], slot, a.genome_len, list(range(b0, b0 + n)), members=10)
        parts.append(ctx.sketch_batch([[buf[64 + slot * j: 64 + slot * j + int(lens[j])]] for j in range(n)]))
    gs = capi.SketchArray.concat(ctx, [capi.SketchArray.of(ctx, p) for p in parts])
    if trace:
        os.environ["SKB_TRACE"] = trace
    for r in range(a.reps):
        db = capi.Database(ctx)                 # a fresh database per repetition, as in a step of the bench
        db.add_many(gs)
        t0 = time.perf_counter()
        h, k = db.query_array(gs[:a.queries])
        dt = time.perf_counter() - t0
        st = ctx.stats()
        print("rep %d: %d queries x %d genomes: %.3f ms (screen %.3f ms, chain %.3f ms on the device), %d pairs chained, %d hits"
              % (r, a.queries, a.genomes, 1e3 * dt, st.screen_ms, st.chain_ms, k, len(h)), file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
