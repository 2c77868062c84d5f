#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its single-GPU configuration (configs[1]):

    one synthetic 5 Mbp genome queried against 100 mutated copies (1-15 % divergence, 10 % of the events
    indels), per GPU.  One step = the whole hot path once: sketch the 100 references and the query
    (FracMinHash seeding + index build), marker screen, anchor lookup + chaining, ANI/AF.

Reported (one JSON line, rank 0):
    value        ANI pairs/s with the ASCII sequences already resident in HBM (device-timed, max over ranks)
    e2e          the same through the host-buffer C-ABI calls: pinned host ASCII -> H2D -> ... -> hits on the host
    sketch_gbps  sketching throughput alone (Gbp/s, device-resident input)
    roofline     the seeding kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
    cpu_baseline the CPU oracle (a port of skani's algorithm; the Rust reference cannot be built here) on
                 the same workload, all host threads
`--impl reference` times that CPU port alone (the reference arm of this tier).
Multi-GPU (`torchrun ... bench.py --gpus N`): each rank owns an independent 1-vs-100 family (weak scaling,
no data-path collective — SURVEY.md §8e; the sketch-DB all-gather only exists for all-vs-all workloads).
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
NCU_SEED_TRAFFIC_BYTES = 508_747_008 + 16_123_392         # profiles/r1_seed_scan_kernel_ncu_full.txt (read + write, one launch)


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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--n-refs", type=int, default=100)
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


def make_family(genome_len, n_refs, rank):
    """base seed 0x5EED0000 (+ rank family), mutant j at d_j = 1 % + 14 % * j / (n-1)."""
    from concurrent.futures import ThreadPoolExecutor
    from pyskani_b200 import synth
    seed0 = 0x5EED0000 + 100_000 * rank
    base = synth.random_genome(genome_len, seed0)
    divs = [0.01 + 0.14 * j / max(1, n_refs - 1) for j in range(n_refs)]
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        refs = list(ex.map(lambda j: synth.mutate(base, divs[j], seed0 + 1 + j), range(n_refs)))
    return base, refs, divs


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


def cpu_step(base, refs, threads):
    """The same step on the CPU oracle: sketch all genomes, then pyskani's query loop."""
    import oracle
    t0 = time.perf_counter()
    sk = oracle.sketch_batch([[r] for r in refs] + [[base]], threads=threads)
    t1 = time.perf_counter()
    idx, res, n_in = oracle.query(sk[-1], sk[:-1], 0.8, True, threads=threads)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, len(idx)


def run_reference(args, rank, world):
    """Reference arm: the CPU port of the path (oracle/), all host threads, rank 0 only."""
    if rank != 0:
        return
    import oracle
    oracle.lib()
    threads = args.cpu_threads or (os.cpu_count() or 1)
    base, refs, _ = make_family(args.genome_len, args.n_refs, 0)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_step(base, refs, threads)
    times = []
    for _ in range(args.steps):
        a, b, nh = cpu_step(base, refs, threads)
        times.append(a + b)
    ms = 1e3 * float(np.mean(times))
    val = args.n_refs / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "ANI pairs/s (sketch + screen + chain + ANI, 1 x 5 Mbp query vs 100 mutated refs)",
        "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": "configs[1]: 1 synthetic %d bp genome vs %d mutated copies (1-15%% divergence, indels)"
                               % (args.genome_len, args.n_refs), "cpu_only": True},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": "the full step (sketch %d genomes + 1 x %d query), CPU port of skani's algorithm "
                                   "(oracle/); the Rust reference itself cannot be built in this image"
                                   % (args.n_refs + 1, args.n_refs)},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from pyskani_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pyskani_b200 has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_cpus(local_rank, getattr(torch.cuda.get_device_properties(local_rank), "uuid", None))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = capi.Context(local_rank)
    L = capi.lib()
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    # ---------------- synthetic inputs (not timed)
    base, refs, divs = make_family(args.genome_len, args.n_refs, rank)
    genomes = refs + [base]
    n_g = len(genomes)
    lens = np.array([len(g) for g in genomes], np.uint64)
    total_bases = int(lens.sum())
    # host side: one pinned buffer, contigs at 16-aligned offsets (what a caller holding FASTA records would pass)
    offs = np.zeros(n_g, np.uint64)
    cur = 64
    for i, l in enumerate(lens):
        offs[i] = cur
        cur += (int(l) + 15) // 16 * 16 + 16
    buf_bytes = cur + 64
    h_ptr = ctx.host_alloc(buf_bytes)
    h_arr = np.ctypeslib.as_array(C.cast(h_ptr, C.POINTER(C.c_uint8)), shape=(buf_bytes,))
    for g, o in zip(genomes, offs):
        h_arr[int(o):int(o) + len(g)] = g
    d_ptr = ctx.dev_alloc(buf_bytes)
    ctx.memcpy_h2d(d_ptr, h_ptr, buf_bytes)
    gstart = np.arange(n_g + 1, dtype=np.uint32)

    host_ptrs = (C.c_void_p * n_g)(*[h_ptr + int(o) for o in offs])
    host_lens = (C.c_uint64 * n_g)(*[int(l) for l in lens])
    gs_c = (C.c_uint32 * (n_g + 1))(*range(n_g + 1))
    params = capi.SketchParams(15, 125, 1000)

    def step_device():
        """inputs resident in HBM"""
        sk = ctx.sketch_batch_device(d_ptr, gstart, offs, lens)
        st = ctx.stats()
        db = capi.Database(ctx)
        db.add_many(sk[:-1])
        hits, n_in = db.query([sk[-1]])
        st2 = ctx.stats()
        return len(hits), st.seed_ms, st.total_ms, st2.total_ms

    def step_host():
        """inputs in (pinned) host memory: H2D inside the call, hits come back to host memory"""
        out = (C.c_void_p * n_g)()
        ctx.check(L.skb_sketch_batch(ctx._h, C.byref(params), 1, n_g, gs_c, host_ptrs, host_lens, out))
        sk = [capi.Sketch(ctx, out[i]) for i in range(n_g)]
        db = capi.Database(ctx)
        db.add_many(sk[:-1])
        hits, n_in = db.query([sk[-1]])
        return len(hits)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        outs, per_step = [], []
        for _ in range(steps):
            ts = time.perf_counter()
            outs.append(fn())
            per_step.append(1e3 * (time.perf_counter() - ts))
        timed.last_per_step = per_step
        e1.record(stream)
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        dev_ms = e0.elapsed_time(e1)
        ms = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1]), outs

    sampler = ClockSampler(local_rank, uuid=getattr(torch.cuda.get_device_properties(local_rank), 'uuid', None))
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    k0 = ctx.stats().kernels_launched
    dev_ms, wall_ms, outs = timed(step_device, args.steps)
    k1 = ctx.stats().kernels_launched
    n_hits = outs[-1][0]
    seed_ms = float(np.mean([o[1] for o in outs]))
    sketch_ms = float(np.mean([o[2] for o in outs]))
    query_ms = float(np.mean([o[3] for o in outs]))

    for _ in range(max(1, args.warmup // 2)):
        step_host()
    e2e_dev_ms, e2e_wall_ms, outs_h = timed(step_host, args.steps)
    e2e_per_step = list(timed.last_per_step)
    clocks = sampler.stop()      # sampled from the first warm-up step to the end of the e2e region
    # plain pinned-host -> device copy of the same bytes, for context (the e2e floor on this box)
    t0 = time.perf_counter()
    ctx.memcpy_h2d(d_ptr, h_ptr, buf_bytes)
    h2d_gbs = buf_bytes / (time.perf_counter() - t0) / 1e9

    pairs_total = args.n_refs * world
    ms_per_step = dev_ms / args.steps
    value = pairs_total / (ms_per_step / 1e3)
    e2e_ms = max(e2e_dev_ms, e2e_wall_ms) / args.steps
    e2e_value = pairs_total / (e2e_ms / 1e3)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = ALG_BYTES_PER_BASE * total_bases / (seed_ms / 1e3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of one seed_scan_kernel launch, from the committed
        # `ncu --set full` capture of this command (profiles/r1_seed_scan_kernel_ncu_full.txt); only valid for the
        # default workload, which is the one that capture ran
        traffic = NCU_SEED_TRAFFIC_BYTES if (args.genome_len, args.n_refs) == (5_000_000, 100) else None
        line = {
            "metric": "ANI pairs/s (sketch + screen + chain + ANI, 1 x 5 Mbp query vs 100 mutated refs)",
            "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": "configs[1]: 1 synthetic %d bp genome vs %d mutated copies (1-15%% divergence, indels) per GPU"
                                   % (args.genome_len, args.n_refs),
                       "k": 15, "c": 125, "marker_c": 1000, "l2": "inputs (%.0f MB ASCII per step) exceed the 126 MB L2" % (total_bases / 1e6),
                       "parallelism": "independent family per GPU, no collective", "cpu_affinity": numa},
            "hits_per_query": n_hits,
            "sketch_gbps": world * total_bases / (sketch_ms / 1e3) / 1e9,
            "seed_kernel_gbps": world * total_bases / (seed_ms / 1e3) / 1e9,
            "query_pairs_per_s": pairs_total / (query_ms / 1e3),
            "phase_ms": {"seed_kernel": seed_ms, "sketch_total": sketch_ms, "query_total": query_ms, "wall_per_step": wall_ms / args.steps},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": total_bases,
                    "d2h_bytes_per_step": int(n_hits) * C.sizeof(capi.Hit), "ms_per_step": e2e_ms,
                    "pinned_h2d_copy_gbs": h2d_gbs, "per_step_wall_ms": [round(x, 3) for x in e2e_per_step]},
            "gpu_launches": int(k1 - k0),
            "roofline": {"bound": "hbm", "kernel": "seed_scan_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read + write)",
                         "algorithmic_bytes": ALG_BYTES_PER_BASE * total_bases, "peak_source": peak_src,
                         "note": "algorithmic bytes = 1.136 B/base x %d bases per launch; the kernel is integer-issue bound, see DESIGN.md" % total_bases},
            # the bound that actually holds for this kernel: two exact 64-bit hashes per base cost
            # ~54 ALU-pipe instructions per position (SASS count); the ALU pipe issues one warp instruction per 2 cycles
            # per SM sub-partition.  ncu: sm__pipe_alu_cycles_active 81 % (profiles/r1_seed_scan_kernel_ncu_full.txt)
            "issue_roofline": {"bound": "integer ALU pipe", "unit": "Gbp/s", "alu_inst_per_base": 54,
                               "peak": 148 * 4 * 32 / (54 * 2) * (clocks.get("sm_mhz") or 1965.0) / 1e3,
                               "achieved": total_bases / (seed_ms / 1e3) / 1e9},
            "clocks": clocks,
        }
        line["issue_roofline"]["frac"] = line["issue_roofline"]["achieved"] / line["issue_roofline"]["peak"]
        if not args.skip_cpu_baseline:
            import oracle
            oracle.lib()
            os.sched_setaffinity(0, all_cpus)      # the CPU baseline may use every core the process was given
            threads = args.cpu_threads or (os.cpu_count() or 1)
            cpu_step(base, refs, threads)
            a, b, nh = cpu_step(base, refs, threads)
            line["cpu_baseline"] = {"value": args.n_refs / (a + b), "unit": "pairs/s", "cores": threads, "kind": "port",
                                    "sketch_gbps": total_bases / a / 1e9, "query_pairs_per_s": args.n_refs / b,
                                    "sample": "one full step (sketch %d genomes + 1 x %d query) on the CPU port of skani's "
                                              "algorithm (oracle/), OpenMP over genomes and pairs" % (n_g, args.n_refs)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
