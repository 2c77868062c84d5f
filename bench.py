#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on the configuration it is quoted on (configs[2]):

    all-vs-all ANI of 1 000 synthetic 5 Mbp genomes (100 families x 10: a random root + 9 mutants at 1-15 % divergence,
    10 % of the events indels) at 1 / 2 / 4 / 8 B200.

One step = the whole job once, through the product's multi-GPU path (pyskani_b200.parallel):
    every rank sketches its share of the genomes (FracMinHash seeding + index build)            [skb_sketch_batch*]
    -> the sketch database is exchanged ONCE over NVLink, device to device                      [skb_exchange_* + NCCL]
    -> every rank screens its genomes against all 1 000, chains the survivors, computes ANI/AF  [skb_db_query]
    -> the hit table is gathered on rank 0 (host memory).
`--gpus N` is STRONG scaling: the same 10^6 ordered pairs are split over N ranks by query.

Reported (one JSON line, rank 0):
    value        ANI pairs/s (all 10^6 ordered pairs / step time) with the ASCII genomes resident in HBM
    e2e          the same with the genomes in pinned HOST memory: H2D copies inside the timed region, hits on the host
    sketch_gbps, screen_pairs_per_s, chained_pairs_per_s, phase_ms, exchange {ms, bytes, busbw}
    roofline     the seeding kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
    parity       after the timed region, a fixed sample is compared with the CPU oracle (sketches bit-exact, every
                 screen decision of 10 queries, >= 200 chained pairs: ANI <= 1e-4, AF <= 1e-3; at N > 1 the multi-GPU
                 hit table must equal a single-GPU run).  A mismatch exits non-zero.
    cpu_baseline the CPU port of skani's algorithm (oracle/; the Rust reference cannot be built here) on a bounded
                 sample of the same workload, all host threads (N = 1 only)
    configs1     secondary block: configs[1], 1 genome vs 100 mutated copies on one GPU (the round-1 headline)
`--impl reference` times that CPU port alone, on the same workload definition (the reference arm of this tier).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_BASE = 1.0 + 16.0 / 125.0 + 8.0 / 1000.0   # 1 B read + seeds + markers written
METRIC = "ANI pairs/s, all-vs-all (sketch + exchange + screen + chain + ANI)"


_JSON_FD = None


def claim_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries loaded later (NCCL's version banner, a warning of the CUDA
    runtime) write to file descriptor 1 on their own, so the descriptor is pointed at stderr for the whole run and the result
    line goes through a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    fd = _JSON_FD if _JSON_FD is not None else 1
    while data:
        data = data[os.write(fd, data):]


def workload_string(a):
    return ("configs[2]: all-vs-all of %d synthetic %d bp genomes (%d families x %d: root + mutants at 1-15%% divergence, indels)"
            % (a.families * a.members, a.genome_len, a.families, a.members))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--families", type=int, default=100)
    ap.add_argument("--members", type=int, default=10)
    ap.add_argument("--sketch-batch", type=int, default=250, help="genomes per skb_sketch_batch call")
    ap.add_argument("--e2e-batch", type=int, default=0, help="genomes per skb_sketch_batch call of the host-buffer run (default: --sketch-batch)")
    ap.add_argument("--pipeline", action="store_true",
                    help="host-buffer run at 1 GPU through parallel.all_vs_all_pipelined (queries on a second context under the "
                         "ingest of the next batch); off by default: it gained 3 ms on one box and lost 3-9 ms on two others")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries of the CPU sample (default: about 96, a multiple of the threads)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--skip-configs1", action="store_true")
    ap.add_argument("--skip-python-api", action="store_true")
    ap.add_argument("--n-refs", type=int, default=100, help="configs[1] block: mutated copies")
    return ap.parse_args()


def bind_to_gpu_cpus(index, uuid=None):
    """One process per GPU: run on the CPUs NVML reports as local to this GPU, so that the pinned input buffers (first
    touched below) sit on the GPU's own NUMA node and the host->device copies of eight ranks do not cross the socket link.
    Returns a short description for the JSON line; a no-op when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = None
        if uuid is not None:
            u = str(uuid)
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID((u if u.startswith("GPU-") else "GPU-" + u).encode())
            except Exception:
                h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1 and 64 * w + b < n_cpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return "unchanged (no NVML-local CPU in this process's cpuset)"
        os.sched_setaffinity(0, allowed)
        return "cpus %d-%d (%d) local to the GPU" % (allowed[0], allowed[-1], len(allowed))
    except Exception as e:       # noqa: BLE001 - purely an optimisation
        return "unchanged (%s)" % type(e).__name__


class ClockSampler:
    """SM clock + throttle reasons during the timed regions.  NVML is polled from a thread (two cheap queries per
    sample); `nvidia-smi -lms` is only the fallback because each of its polls stalls PCIe traffic for milliseconds,
    which distorts the host->device leg of the e2e measurement."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, uuid=None, period_s=0.01):
        self.index, self.uuid, self.period = index, uuid, period_s
        self.sm, self.mask, self.max_mhz, self.mode = [], 0, None, None
        self._stop = threading.Event()
        self.proc = None

    def start(self):
        if os.environ.get("BENCH_NO_SAMPLER"):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(self.uuid)).encode() if not str(self.uuid).startswith("GPU-") else str(self.uuid).encode())
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.mask |= int(reasons(h))
                    except Exception:
                        pass
                    self._stop.wait(self.period)
            self.thr = threading.Thread(target=loop, daemon=True)
            self.thr.start()
            self.mode = "nvml"
            return
        except Exception:
            pass
        try:
            fields = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.rows = []
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={fields}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=lambda: [self.rows.append([x.strip() for x in l.split(",")]) for l in self.proc.stdout], daemon=True)
            self.thr.start()
            self.mode = "nvidia-smi"
        except OSError:
            self.proc = None

    def stop(self):
        if self.mode == "nvml":
            self._stop.set()
            self.thr.join(timeout=2)
            reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.sm), "source": "nvml, %.0f ms period" % (1e3 * self.period)}
        if self.mode == "nvidia-smi" and self.proc:
            self.proc.terminate()
            self.thr.join(timeout=2)
            sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
            mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                    "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 200"}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}


# ------------------------------------------------------------------------------------------------ workload
class HostGenomes:
    """The genomes `ids` as ASCII in consecutive 16-aligned slots of one uint8 buffer (ordinary or pinned memory)."""

    def __init__(self, ids, genome_len, members, buf=None, threads=None):
        import workload
        self.ids = list(ids)
        self.slot = workload.slot_bytes(genome_len)
        self.lead = 64                                             # guard in front of the first contig (device layout)
        nbytes = self.lead + self.slot * len(self.ids) + 64
        self.buf = buf if buf is not None else np.zeros(nbytes, np.uint8)
        assert self.buf.size >= nbytes
        self.nbytes = nbytes
        self.lens = workload.fill_families(self.buf[self.lead:], self.slot, genome_len, self.ids, members=members,
                                           divergences=workload.FAMILY_DIVERGENCES[:members] if members <= len(workload.FAMILY_DIVERGENCES)
                                           else tuple(0.15 * m / (members - 1) for m in range(members)), threads=threads)
        self.offs = np.array([self.lead + j * self.slot for j in range(len(self.ids))], np.uint64)

    def view(self, j):
        o = int(self.offs[j])
        return self.buf[o:o + int(self.lens[j])]


def cpu_sample_size(threads, want=0):
    if want:
        return want
    return threads * max(1, round(96 / threads))


def cpu_sample_step(oracle, host, sample, db_sketches, threads):
    """A bounded sample of the all-vs-all job on the CPU port: sketch the sample's genomes, query them against the full
    database (whose sketches exist already).  Returns (sketch s, query s, hits, screened-in)."""
    t0 = time.perf_counter()
    sk = oracle.sketch_batch([[host.view(j)] for j in sample], threads=threads)
    t1 = time.perf_counter()
    hq, hr, res, n_in = oracle.query_many(sk, db_sketches, 0.8, True, threads=threads)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, len(hq), n_in


def run_reference(args, rank, world):
    """Reference arm: the CPU port of the path (oracle/), all host threads, rank 0 only.  Imports nothing of the product."""
    if rank != 0:
        return
    import oracle
    oracle.lib()
    threads = args.cpu_threads or (os.cpu_count() or 1)
    n_total = args.families * args.members
    host = HostGenomes(range(n_total), args.genome_len, args.members)
    db = oracle.sketch_batch([[host.view(j)] for j in range(n_total)], threads=threads)      # the database side, not timed
    n_s = min(n_total, cpu_sample_size(threads, args.cpu_sample))
    sample = list(range(n_s))
    for _ in range(args.warmup):
        cpu_sample_step(oracle, host, sample, db, threads)
    times, sk_t, q_t = [], [], []
    for _ in range(args.steps):
        a, b, nh, n_in = cpu_sample_step(oracle, host, sample, db, threads)
        times.append(a + b); sk_t.append(a); q_t.append(b)
    ms = 1e3 * float(np.mean(times))
    pairs = n_s * n_total
    val = pairs / (ms / 1e3)
    sample_txt = ("per step: sketch %d of the %d genomes and query them against all %d (= %d of the %d ordered pairs, same %.1f %% "
                  "of them screened in), CPU port of skani's algorithm (oracle/) with OpenMP over genomes and pairs; the Rust "
                  "reference itself cannot be built in this image" % (n_s, n_total, n_total, pairs, n_total * n_total,
                                                                      100.0 * n_in / max(pairs, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_string(args), "k": 15, "c": 125, "marker_c": 1000, "cpu_only": True},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample_txt,
                         "sketch_gbps": float(sum(int(host.lens[j]) for j in sample) / np.mean(sk_t) / 1e9),
                         "chained_pairs_per_s": float(n_in / np.mean(q_t)), "hits_per_step": nh},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ parity gate
def parity_gate(args, capi, ctx, host_all, table, world, threads, oracle_db):
    """Rank 0, after the timed region: a fixed sample of the job against the CPU oracle.  Raises on any mismatch."""
    import oracle
    from pyskani_b200 import parallel
    n_total = len(host_all.ids)
    members = args.members
    out = {}
    # (a) sketches: bit-exact seeds (k-mer, position, contig, strand) and marker sets for 8 genomes
    sample_sk = sorted(set([0, 1, members - 1, n_total // 3, n_total // 2, n_total // 2 + 3, n_total - 2, n_total - 1]))
    gs = ctx.sketch_batch([[host_all.view(j)] for j in sample_sk])
    for j, g in zip(sample_sk, gs):
        e = g.export()
        ok, op, oc, ocan = oracle_db[j].seeds()
        if not (np.array_equal(e["kmer"], ok) and np.array_equal(e["pos"], op) and np.array_equal(e["contig"], oc)
                and np.array_equal(e["canonical"], ocan) and np.array_equal(e["markers"], oracle_db[j].markers())):
            raise SystemExit("PARITY FAILURE: sketch of genome %d differs from the oracle" % j)
    out["sketches_bit_exact"] = len(sample_sk)
    del gs
    # a single-GPU database of everything on rank 0 (not timed): screen sample + reference table for N > 1
    be = parallel.CudaBackend(0, ctx=ctx)
    full = capi.SketchArray.concat(ctx, [ctx.sketch_batch([[host_all.view(j)] for j in range(b0, min(n_total, b0 + args.sketch_batch))])
                                         for b0 in range(0, n_total, args.sketch_batch)])
    db = capi.Database(ctx)
    db.add_many(full)
    # (b) every screen decision of 10 queries (and the marker intersection sizes)
    q_screen = [int(x) for x in np.linspace(0, n_total - 1, 10)]
    ok_gpu, shared_gpu = db.screen(capi.SketchArray.of(ctx, [full[q] for q in q_screen]))
    for i, q in enumerate(q_screen):
        for r in range(n_total):
            ok, sh = oracle.screen(oracle_db[q], oracle_db[r])
            if ok != bool(ok_gpu[i, r]) or sh != int(shared_gpu[i, r]):
                raise SystemExit("PARITY FAILURE: screen of (%d, %d): GPU %s/%d, oracle %s/%d" % (q, r, ok_gpu[i, r], shared_gpu[i, r], ok, sh))
    out["screen_decisions"] = len(q_screen) * n_total
    # (c) chained pairs of the TIMED run's hit table: sample queries = whole families
    n_q = min(n_total, max(4 * members, 40))
    q_chain = list(range(n_q // 2)) + list(range(n_total - (n_q - n_q // 2), n_total))
    hq, hr, res, n_in = oracle.query_many([oracle_db[q] for q in q_chain], oracle_db, 0.8, True, threads=threads)
    want = {(q_chain[int(a)], int(b)): r for a, b, r in zip(hq, hr, res)}
    sel = np.isin(table[:, 0].astype(np.int64), q_chain)
    got = {(int(r[0]), int(r[1])): r for r in table[sel]}
    if set(got) != set(want):
        raise SystemExit("PARITY FAILURE: hit set of the sample queries differs from the oracle: only GPU %s, only oracle %s"
                         % (sorted(set(got) - set(want))[:5], sorted(set(want) - set(got))[:5]))
    worst_ani = worst_af = 0.0
    for key, r in want.items():
        g = got[key]
        worst_ani = max(worst_ani, abs(g[2] - r.ani)); worst_af = max(worst_af, abs(g[3] - r.af_query), abs(g[4] - r.af_ref))
    if worst_ani > 1e-4 or worst_af > 1e-3:
        raise SystemExit("PARITY FAILURE: ANI/AF of the sample differ from the oracle by %.2e / %.2e" % (worst_ani, worst_af))
    out.update({"chained_pairs": int(n_in), "hits_compared": len(want), "max_abs_ani_err": worst_ani, "max_abs_af_err": worst_af,
                "tolerance": {"ani": 1e-4, "af": 1e-3}})
    # (d) N > 1: the gathered table of the multi-GPU path against a single-GPU run of the same job
    if world > 1:
        single = be.query(full, full)
        single = single[np.lexsort((single[:, 1], single[:, 0]))]
        if single.shape != table.shape or not np.array_equal(single, table):
            raise SystemExit("PARITY FAILURE: the %d-GPU hit table differs from the single-GPU table (%s vs %s rows)"
                             % (world, table.shape, single.shape))
        out["multi_gpu_table_equals_single_gpu"] = True
    return out


# ------------------------------------------------------------------------------------------------ configs[1] block
def configs1_block(args, capi, ctx, torch, stream):
    """BASELINE.json configs[1]: one 5 Mbp genome against 100 mutated copies (1-15 %), device-resident, one GPU."""
    import workload
    n_refs = args.n_refs
    slot = workload.slot_bytes(args.genome_len)
    buf = np.zeros(64 + slot * (n_refs + 1) + 64, np.uint8)
    base = workload.random_genome(args.genome_len, workload.SEED0 + 999_983)
    lens = []
    from concurrent.futures import ThreadPoolExecutor
    def make(j):
        d = 0.01 + 0.14 * j / max(1, n_refs - 1)
        return workload.mutate_into(base.ctypes.data, len(base), d, workload.SEED0 + 999_984 + j, buf.ctypes.data + 64 + j * slot, slot - 16)
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        lens = list(ex.map(make, range(n_refs)))
    buf[64 + n_refs * slot:64 + n_refs * slot + len(base)] = base
    lens.append(len(base))
    offs = np.array([64 + j * slot for j in range(n_refs + 1)], np.uint64)
    lens = np.array(lens, np.uint64)
    d_ptr = ctx.dev_alloc(buf.size)
    ctx.memcpy_h2d(d_ptr, buf.ctypes.data, buf.size)
    gstart = np.arange(n_refs + 2, dtype=np.uint32)

    def step():
        sk = ctx.sketch_batch_device(d_ptr, gstart, offs, lens)
        db = capi.Database(ctx)
        db.add_many(sk[:-1])
        hits, n_in = db.query_array([sk[-1]])
        return len(hits)
    for _ in range(max(3, args.warmup)):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync()
    steps = max(5, args.steps)
    e0.record(stream)
    for _ in range(steps):
        nh = step()
    e1.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ctx.dev_free(d_ptr)
    return {"workload": "configs[1]: 1 synthetic %d bp genome vs %d mutated copies (1-15%% divergence, indels), device-resident, 1 GPU"
                        % (args.genome_len, n_refs),
            "value": n_refs / (ms / 1e3), "unit": "pairs/s", "ms_per_step": ms, "steps": steps, "hits_per_query": nh}


# ------------------------------------------------------------------------------------------------ the B200 arm
def python_api_block(args, host, n_total, want_hits):
    """The same all-vs-all through the pyskani-compatible extension, as a Python caller would run it: genomes are `bytes`
    objects (pageable memory), Database.sketch_many(...) then Database.query_many(...) - which, like the reference's
    query(), sketches every query genome again - plus a sample of the reference's own one-genome call, Database.query()."""
    import pyskani_b200 as pyskani
    items = [("g%d" % j, host.view(j).tobytes()) for j in range(n_total)]
    best, n_hits = None, 0
    for _ in range(2):
        db = pyskani.Database()
        t0 = time.perf_counter()
        db.sketch_many(items)
        t1 = time.perf_counter()
        hits = db.query_many(items, learned_ani=False)
        t2 = time.perf_counter()
        n_hits = sum(len(h) for h in hits)
        if best is None or t2 - t0 < best[0]:
            best = (t2 - t0, t1 - t0, t2 - t1)
        if _ == 0:
            lat = []
            for j in range(0, min(n_total, 24)):
                ta = time.perf_counter()
                db.query(items[j][0], items[j][1], learned_ani=False)
                lat.append(time.perf_counter() - ta)
        del db
    if n_hits != want_hits:
        raise SystemExit("the extension returned %d hits, the C ABI path %d" % (n_hits, want_hits))
    return {"value": n_total * n_total / best[0], "unit": "pairs/s", "ms_per_step": 1e3 * best[0],
            "sketch_many_ms": 1e3 * best[1], "query_many_ms": 1e3 * best[2], "hits": n_hits,
            "host_bytes_per_step": 2 * int(sum(len(b) for _, b in items)),
            "single_query_call_ms": 1e3 * float(np.median(lat)),
            "note": "pyskani_b200.Database.sketch_many + query_many on Python bytes (pageable host memory); query_many sketches the "
                    "1 000 query genomes again, as the reference's query() does, so 10 GB of ASCII are ingested per step; "
                    "single_query_call_ms = median of Database.query(name, bytes) against the 1 000-genome database (the "
                    "reference's own call pattern: sketch one genome, screen 1 000, chain ~9)"}


def main():
    args = parse()
    claim_stdout()
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keeps NCCL's version banner off stdout: rank 0 prints ONE JSON line
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from pyskani_b200 import capi, parallel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pyskani_b200 has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_cpus(local_rank, getattr(torch.cuda.get_device_properties(local_rank), "uuid", None))
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    dgroup = dist if world > 1 else None

    ctx = capi.Context(local_rank)
    L = capi.lib()
    backend = parallel.CudaBackend(local_rank, ctx=ctx)
    stream = torch.cuda.ExternalStream(ctx.stream, device=device)

    # ---------------- synthetic inputs (not timed): this rank's share, ASCII, in pinned host memory and in HBM
    n_total = args.families * args.members
    plan = parallel.partition_by_size([args.genome_len] * n_total, world)
    mine = plan[rank]
    import workload
    slot = workload.slot_bytes(args.genome_len)
    my_bytes = 64 + slot * len(mine) + 64
    h_ptr = ctx.host_alloc(my_bytes)
    h_arr = np.ctypeslib.as_array(C.cast(h_ptr, C.POINTER(C.c_uint8)), shape=(my_bytes,))
    h_arr[:64] = 0; h_arr[-64:] = 0
    host = HostGenomes(mine, args.genome_len, args.members, buf=h_arr)
    my_bases = int(host.lens.sum())
    d_ptr = ctx.dev_alloc(my_bytes)
    ctx.memcpy_h2d(d_ptr, h_ptr, my_bytes)
    n_mine = len(mine)
    batches = [(b0, min(n_mine, b0 + args.sketch_batch)) for b0 in range(0, n_mine, args.sketch_batch)]
    params = capi.SketchParams(15, 125, 1000)
    host_ptr_arrays = []
    eb = args.e2e_batch or args.sketch_batch
    batches_h = [(b0, min(n_mine, b0 + eb)) for b0 in range(0, n_mine, eb)]
    for b0, b1 in batches_h:
        n = b1 - b0
        host_ptr_arrays.append(((C.c_void_p * n)(*[h_ptr + int(host.offs[j]) for j in range(b0, b1)]),
                                (C.c_uint64 * n)(*[int(host.lens[j]) for j in range(b0, b1)]),
                                (C.c_uint32 * (n + 1))(*range(n + 1))))

    def sketch_device(tm):
        parts, seed_ms, tot_ms = [], 0.0, 0.0
        for b0, b1 in batches:
            parts.append(ctx.sketch_batch_device(d_ptr, np.arange(b1 - b0 + 1, dtype=np.uint32), host.offs[b0:b1], host.lens[b0:b1]))
            st = ctx.stats()
            seed_ms += st.seed_ms; tot_ms += st.total_ms
        tm["seed_kernel_ms"] = seed_ms; tm["sketch_device_ms"] = tot_ms
        return capi.SketchArray.concat(ctx, parts)

    def sketch_host(tm):
        parts, raw, packed = [], 0, 0
        for part in sketch_host_batches(tm):
            parts.append(part)
        return capi.SketchArray.concat(ctx, parts)

    def sketch_host_batches(tm):
        """the host-buffer sketch calls, one batch per iteration"""
        raw, packed = 0, 0
        for (b0, b1), (ptrs, lens, gs) in zip(batches_h, host_ptr_arrays):
            out = np.zeros(b1 - b0, np.uint64)
            ctx.check(L.skb_sketch_batch(ctx._h, C.byref(params), 1, b1 - b0, gs, ptrs, lens, out.ctypes.data))
            st = ctx.stats()
            raw += st.h2d_raw_bytes; packed += st.h2d_packed_bytes
            tm["h2d_ascii_bytes"] = float(raw); tm["h2d_packed_bytes"] = float(packed)
            yield capi.SketchArray(ctx, out)

    # one GPU, host buffers: the queries of a batch run on a second context of the same device while the next batch is
    # still crossing PCIe (parallel.all_vs_all_pipelined); with several ranks the exchange comes first
    pipelined = world == 1 and args.pipeline and len(batches_h) > 1
    backend_q = parallel.CudaBackend(local_rank, ctx=capi.Context(local_rank)) if pipelined else None
    if os.environ.get("BENCH_TEAM"):
        ctx.set_host_threads(int(os.environ["BENCH_TEAM"]))      # tuning hook: threads of the ingest team
    elif pipelined:
        # the query thread spins on its own waits: one CPU less for the ingest team (measured: 16 -> 15 threads, -4 ms)
        ctx.set_host_threads(max(2, min(32, len(os.sched_getaffinity(0))) - 1))
    if pipelined and not os.environ.get("BENCH_NO_PRIORITY"):
        ctx.set_priority(True)          # seeding / index kernels go ahead of the query context's
        stream = torch.cuda.ExternalStream(ctx.stream, device=device)

    def step_pipelined(_unused):
        tm = {}
        t0 = time.perf_counter()
        table = parallel.all_vs_all_pipelined(sketch_host_batches(tm), backend_q, timings=tm, sort=False)
        if len(table):
            ids = np.asarray(mine, np.float64)
            table[:, 0] = ids[table[:, 0].astype(np.int64)]
            table[:, 1] = ids[table[:, 1].astype(np.int64)]
        tm["step_wall_ms"] = 1e3 * (time.perf_counter() - t0)
        return table, tm

    def step(sketch_fn):
        if sketch_fn is sketch_host and pipelined:
            return step_pipelined(None)
        tm = {}
        t0 = time.perf_counter()
        sk = sketch_fn(tm)
        tm["sketch_ms"] = 1e3 * (time.perf_counter() - t0)
        # the hit table arrives on rank 0 rank after rank; ordering it by (query, ref) for the comparisons below is
        # presentation, not part of the job (the reference returns its hits in hash-set order, lib.rs:640)
        table = parallel.query_and_gather(backend, sk, mine, plan, dgroup, device, None, tm, sort=False)
        tm["step_wall_ms"] = 1e3 * (time.perf_counter() - t0)
        return table, tm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def timed(sketch_fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        tms, table = [], None
        for _ in range(steps):
            table, tm = step(sketch_fn)
            tms.append(tm)
            if os.environ.get("BENCH_STEP_LOG") and rank == 0:
                print("step: " + " ".join("%s=%.2f" % (k, v) for k, v in tm.items() if k.endswith("_ms")), file=sys.stderr, flush=True)
        e1.record(stream)
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        ms = torch.tensor([e0.elapsed_time(e1), wall_ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return max(float(ms[0]), float(ms[1])) / steps, tms, table

    def phase_stats(tms, keys):
        """mean over steps, then max (times) or sum (counts) over ranks"""
        out = {}
        for k, how in keys:
            v = float(np.mean([t.get(k, 0.0) for t in tms]))
            if world > 1:
                x = torch.tensor([v], dtype=torch.float64, device=device)
                dist.all_reduce(x, op=dist.ReduceOp.MAX if how == "max" else dist.ReduceOp.SUM)
                v = float(x[0])
            out[k] = v
        return out

    sampler = ClockSampler(local_rank, uuid=getattr(torch.cuda.get_device_properties(local_rank), 'uuid', None))
    sampler.start()
    for _ in range(args.warmup):
        step(sketch_device)
    k0 = ctx.stats().kernels_launched
    ms_per_step, tms, table = timed(sketch_device, args.steps)
    k1 = ctx.stats().kernels_launched
    keys = [("seed_kernel_ms", "max"), ("sketch_device_ms", "max"), ("sketch_ms", "max"), ("exchange_ms", "max"), ("exchange_sizes_ms", "max"),
            ("exchange_pack_enqueue_ms", "max"), ("exchange_allgather_ms", "max"), ("exchange_adopt_ms", "max"), ("query_ms", "max"),
            ("screen_ms", "max"), ("chain_ms", "max"), ("gather_ms", "max"), ("screened_in", "sum"), ("local_hits", "sum"),
            ("exchange_bytes_in", "max"), ("exchange_bytes_total", "max")]
    ph = phase_stats(tms, keys)

    for _ in range(max(2, args.warmup)):          # also lets every context settle on its ingest policy (six large calls)
        step(sketch_host)
    e2e_ms, tms_h, table_h = timed(sketch_host, args.steps)
    ph_h = phase_stats(tms_h, [("sketch_ms", "max"), ("exchange_ms", "max"), ("query_ms", "max"), ("gather_ms", "max"),
                               ("h2d_ascii_bytes", "sum"), ("h2d_packed_bytes", "sum")])
    clocks = sampler.stop()      # sampled from the first warm-up step to the end of the e2e region
    # the floor of the host->device leg: all ranks copy their pinned bytes at the same time
    barrier()
    t0 = time.perf_counter()
    ctx.memcpy_h2d(d_ptr, h_ptr, my_bytes)
    h2d_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    launches = torch.tensor([float(k1 - k0)], dtype=torch.float64, device=device)
    bases_t = torch.tensor([float(my_bases)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(h2d_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
        dist.all_reduce(bases_t, op=dist.ReduceOp.SUM)
    h2d_floor_ms = 1e3 * float(h2d_s[0])
    total_bases = int(bases_t[0])

    pairs_total = n_total * n_total
    value = pairs_total / (ms_per_step / 1e3)
    e2e_value = pairs_total / (e2e_ms / 1e3)

    if rank == 0:
        table, table_h = parallel.sort_hits(table), parallel.sort_hits(table_h)
        if table_h.shape != table.shape or not np.array_equal(table_h, table):
            raise SystemExit("host-buffer path and device-resident path returned different hit tables")
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # roofline of the dominant kernel (seed_scan_kernel), this rank's launches: algorithmic bytes / time between CUDA
        # events recorded around the seeding launches on the library's own stream
        seed_ms = float(np.mean([t["seed_kernel_ms"] for t in tms]))
        achieved = ALG_BYTES_PER_BASE * my_bases / (seed_ms / 1e3) / 1e9
        traffic, traffic_note = None, "no ncu capture committed for this launch shape"
        tpath = os.path.join(ROOT, "profiles", "r2_seed_scan_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            # the capture describes one seeding launch of this workload: valid when one of this rank's launches covers exactly its bases
            if int(tj.get("bases_per_launch", -1)) in [int(host.lens[b0:b1].sum()) for b0, b1 in batches]:
                traffic, traffic_note = tj["dram_bytes_per_launch"], tj.get("source", tpath)
        ex_ms = ph.get("exchange_allgather_ms", 0.0)
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_string(args), "k": 15, "c": 125, "marker_c": 1000,
                       "l2": "inputs (%.0f MB of ASCII per rank and step) exceed the 126 MB L2" % (my_bases / 1e6),
                       "parallelism": "queries split over %d rank(s); sketch database exchanged once per step%s" % (
                           world, " (" + tms[-1].get("exchange_collective", "") + ")" if world > 1 else " (single GPU: no exchange)"),
                       "sketch_batch": args.sketch_batch, "cpu_affinity": numa},
            "pairs_per_step": pairs_total, "hits_per_step": int(len(table)),
            "sketch_gbps": total_bases / (ph["sketch_ms"] / 1e3) / 1e9,
            "seed_kernel_gbps": total_bases / (ph["seed_kernel_ms"] / 1e3) / 1e9,
            "screen_pairs_per_s": pairs_total / (max(ph["screen_ms"], 1e-6) / 1e3),
            "chained_pairs_per_step": int(ph["screened_in"]),
            "chained_pairs_per_s": ph["screened_in"] / (max(ph["chain_ms"], 1e-6) / 1e3),
            "phase_ms": {k: round(v, 4) for k, v in ph.items() if k.endswith("_ms")},
            "exchange": None if world == 1 else {
                "ms": ph["exchange_ms"], "allgather_ms": ex_ms, "sizes_ms": ph["exchange_sizes_ms"], "pack_enqueue_ms": ph["exchange_pack_enqueue_ms"],
                "adopt_ms": ph["exchange_adopt_ms"], "bytes_total": int(ph["exchange_bytes_total"]),
                "bytes_in_per_gpu": int(ph["exchange_bytes_in"]), "collective": tms[-1].get("exchange_collective"),
                "busbw_gbs": ph["exchange_bytes_in"] / (max(ex_ms, 1e-6) / 1e3) / 1e9,
                "note": "ms = host time until the database is usable (sizes + pack + heads + adopt); the body all-gather (allgather_ms, "
                        "device-timed from the first collective to the end of the second) continues under the marker screen of the "
                        "query. busbw = bytes received per GPU / allgather_ms (NCCL's definition for all-gather). Peers receive the "
                        "reference side of every sketch only (no position-order seeds): half of the sketch bytes"},
            "e2e": {"value": e2e_value, "unit": "pairs/s",
                    "h2d_bytes_per_step": int(ph_h["h2d_ascii_bytes"] + ph_h["h2d_packed_bytes"]),
                    "d2h_bytes_per_step": int(len(table)) * C.sizeof(capi.Hit), "ms_per_step": e2e_ms,
                    "phase_ms": {k: round(v, 4) for k, v in ph_h.items() if k.endswith("_ms")},
                    "host_input_bytes_per_step": total_bases,
                    "sketch_batch": eb,
                    "pipeline": ("queries of a sketched batch (new batch x database so far, both directions) run on a second context "
                                 "of the same device under the ingest of the next batch; phase_ms.query_ms is what was left after "
                                 "the last batch") if pipelined else "sketch all, then query",
                    "ingest": {"h2d_ascii_bytes": int(ph_h["h2d_ascii_bytes"]), "h2d_packed_bytes": int(ph_h["h2d_packed_bytes"]),
                               "bases_sent_packed": int(4 * ph_h["h2d_packed_bytes"]),
                               "note": "skb_sketch_batch moves large host batches two ways at once: the copy engine pulls chunks of "
                                       "ASCII from the caller's pinned buffer while host threads compact other chunks to 2-bit words "
                                       "(4 bases per byte) in pinned staging memory; chunks are claimed by whichever route is free. "
                                       "h2d_bytes_per_step counts the bytes that actually crossed the link"},
                    "h2d_floor_ms": h2d_floor_ms, "h2d_floor_gbs_per_gpu": my_bytes / (h2d_floor_ms / 1e3) / 1e9,
                    "e2e_over_floor": e2e_ms / h2d_floor_ms,
                    "note": "floor = all ranks copying their pinned ASCII input bytes to the device at the same time, nothing else "
                            "running (what a copy-everything ingest cannot beat)"},
            "gpu_launches": int(launches[0]),
            "roofline": {"bound": "hbm", "kernel": "seed_scan_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read + write)",
                         "traffic_source": traffic_note, "launches_per_step_per_rank": len(batches),
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_BASE * my_bases / max(1, len(batches)), "peak_source": peak_src,
                         "note": "algorithmic bytes = 1.136 B/base x bases per launch; the kernel is integer-issue bound (two exact "
                                 "64-bit hashes per base), see DESIGN.md section 4"},
            "issue_roofline": {
                "bound": "integer ALU pipe", "kernel": "seed_scan_kernel", "unit": "Gbp/s",
                "achieved": my_bases / (seed_ms / 1e3) / 1e9,
                "peak": 148 * 4 * 32 / (40.25 * 2.0) * 1.965,
                "frac": (my_bases / (seed_ms / 1e3) / 1e9) / (148 * 4 * 32 / (40.25 * 2.0) * 1.965),
                "note": "two exact 64-bit mm_hash64 evaluations per base cost 37.25 ALU-pipe instructions per position in the hash "
                        "loop plus 3 IMAD.WIDE that take an ALU slot each besides their two FMA slots (profiles/r2_seed_scan_sass_mix.txt, "
                        "derived from the shipped cubin by tools/sass_mix.py; slot costs measured by tools/micro/int_pipes.cu, "
                        "profiles/r2_int_pipes.txt); the pipe issues one warp instruction per 2 cycles per SM sub-partition: "
                        "148 SMs x 4 x 32 lanes / (40.25 x 2) x 1.965 GHz. ncu: ALU pipe 75-81 % busy over the whole kernel "
                        "(76 instructions per position incl. pack, scan and write-out)"},
            "clocks": clocks,
        }
        threads = args.cpu_threads or (os.cpu_count() or 1)
        if not (args.skip_parity and (args.skip_cpu_baseline or world > 1)):
            import oracle
            oracle.lib()
            os.sched_setaffinity(0, all_cpus)      # the CPU legs may use every core the process was given
            host_all = host if world == 1 else HostGenomes(range(n_total), args.genome_len, args.members)
            odb = oracle.sketch_batch([[host_all.view(j)] for j in range(n_total)], threads=threads)
            if not args.skip_parity:
                line["parity"] = parity_gate(args, capi, ctx, host_all, table, world, threads, odb)
                line["parity_checked"] = True
            if world == 1 and not args.skip_cpu_baseline:
                n_s = min(n_total, cpu_sample_size(threads, args.cpu_sample))
                sample = list(range(n_s))
                cpu_sample_step(oracle, host_all, sample, odb, threads)
                a, b, nh, n_in = cpu_sample_step(oracle, host_all, sample, odb, threads)
                line["cpu_baseline"] = {
                    "value": n_s * n_total / (a + b), "unit": "pairs/s", "cores": threads, "kind": "port",
                    "sketch_gbps": float(sum(int(host_all.lens[j]) for j in sample)) / a / 1e9, "chained_pairs_per_s": n_in / b,
                    "sample": "sketch %d of the %d genomes and query them against all %d (%d ordered pairs, %d chained) on the CPU "
                              "port of skani's algorithm (oracle/), OpenMP over genomes and pairs; database sketches prepared "
                              "outside the timed sample" % (n_s, n_total, n_total, n_s * n_total, n_in)}
            del odb
        if not args.skip_configs1:
            line["configs1"] = configs1_block(args, capi, ctx, torch, stream)
        if world == 1 and not args.skip_python_api:
            line["python_api"] = python_api_block(args, host, n_total, int(len(table)))
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
