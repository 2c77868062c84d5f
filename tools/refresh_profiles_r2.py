"""Rebuilds profiles/r2_* from what tools/profile_round2.sh left in gpurun_out/ (run here, after gpurun merged the files)."""
import collections, contextlib, csv, io, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, ROOT)
import summarize_ncu as S


def capture(fn, *a):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()


WL = "python bench.py --steps 1 --warmup 1 (all-vs-all of 1 000 x 5 Mbp genomes on one B200: 4 sketch calls of 250 genomes, one chaining batch per step)"
rep = os.path.join(G, "r2_seed_scan_packed.ncu-rep")
if os.path.exists(rep):
    open(os.path.join(P, "r2_seed_scan_kernel_packed_ncu_full.txt"), "w").write(
        capture(S.full, rep, "ncu --set full --clock-control none --import-source on -k regex:seed_scan_kernel -s 20 -c 1: SKB_INGEST=pack python "
                             "tools/ingest_sweep.py --only pack (seed_scan_kernel<false, true>: the launch of one ~16 MB chunk that the host "
                             "threads compacted to 2-bit words; launches of this path are small and wait for their chunk); round 2"))
for k in ["seed_scan_kernel", "chain_dp_thread_kernel", "match_count_kernel", "anchor_fill_kernel", "window_walk_smem_kernel", "marker_join_kernel", "marker_rank_kernel"]:
    rep = os.path.join(G, "r2_%s.ncu-rep" % k)
    if os.path.exists(rep):
        open(os.path.join(P, "r2_%s_ncu_full.txt" % k), "w").write(
            capture(S.full, rep, "ncu --set full --clock-control none --import-source on -k regex:%s -s 1 -c 1: %s; round 2" % (k, WL)))

# one device-resident step cut from the launch list (every step ends with its ani_reduce launch: one chaining batch per step)
rows = list(csv.reader(open(os.path.join(G, "r2_launches_allvsall.csv"))))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
H = rows[hi]; kn, mv, mn = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Name')
L = [(r[kn], float(r[mv].replace(',', '')) / 1e3) for r in rows[hi + 1:] if len(r) > mv and r[mn] == 'gpu__time_duration.sum']
ends = [i for i, (n, t) in enumerate(L) if 'ani_reduce' in n]
step = L[ends[0] + 1:ends[1] + 1]


def short(n):
    if 'DeviceRadixSort' in n or 'RadixSort' in n and 'skb::' not in n.split('(')[0]: return 'CUB radix sort'
    if 'DeviceScan' in n: return 'CUB scan (anchor offsets, region counts)'
    if 'DeviceSelect' in n or 'DeviceCompact' in n: return 'CUB select'
    m = re.search(r'skb::(?:<unnamed>::)?(\w+)', n)
    if m: return m.group(1)
    return 'CUB other'


agg = collections.OrderedDict()
for n, t in step:
    a = agg.setdefault(short(n), [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
bench = json.loads(open(os.path.join(G, "r2_bench_n1.json")).read().strip().splitlines()[-1])
out = ["# One device-resident step of bench.py (all-vs-all of 1 000 x 5 Mbp genomes, one B200), cut from the launch list of",
       "# `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 1 --warmup 1` (second step).",
       "# Cold-cache, serialised launch durations: compare shares, not absolutes.",
       "# total %.1f us over %d launches; the same step measured with CUDA events in bench.py: %.2f ms" % (tot, len(step), bench["ms_per_step"]),
       "# (memsets and host round trips are not kernels and do not appear here)"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%10.1f us %4d launches %5.1f %%  %s" % (v[1], v[0], 100 * v[1] / tot, k))
open(os.path.join(P, "r2_launches_allvsall.txt"), "w").write("\n".join(out) + "\n")

# DRAM traffic of the first seed_scan launch of a step (the launch bench.py's roofline object describes)
rep = os.path.join(G, "r2_seed_scan_first.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    Hh, U, V = r[0], r[1], r[2]
    def val(name):
        i = Hh.index(name)
        x = float(V[i].replace(',', ''))
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[U[i]]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    import workload, numpy as np
    slot = workload.slot_bytes(5_000_000)
    buf = np.zeros(slot * 250, np.uint8)
    bases = int(workload.fill_families(buf, slot, 5_000_000, range(250)).sum())
    json.dump({"kernel": "seed_scan_kernel", "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
               "bases_per_launch": bases, "algorithmic_bytes_per_launch": 1.136 * bases,
               "source": "profiles/r2_seed_scan_traffic.json: ncu --set full --clock-control none -k regex:seed_scan_kernel -c 1 python bench.py "
                         "--steps 1 --warmup 1 (first seeding launch of a step: genomes 0..249 of the workload)"},
              open(os.path.join(P, "r2_seed_scan_traffic.json"), "w"), indent=1)

for src, dst in (("r2_bench_n1.json", "r2_bench_line_n1.json"), ("r2_bench_reference.json", "r2_bench_line_reference.json"),
                 ("bench_n2.json", "r2_bench_line_n2.json"), ("bench_n4.json", "r2_bench_line_n4.json"), ("bench_n8.json", "r2_bench_line_n8.json")):
    if os.path.exists(os.path.join(G, src)):
        line = open(os.path.join(G, src)).read().strip().splitlines()[-1]
        json.loads(line)
        open(os.path.join(P, dst), "w").write(line + "\n")

mem = open(os.path.join(G, "r2_memcheck.log")).read().strip().splitlines()
race = open(os.path.join(G, "r2_racecheck.log")).read().strip().splitlines()
open(os.path.join(P, "r2_sanitizer.txt"), "w").write(
    "# compute-sanitizer on the GPU parity tests (B200, end of round 2: thread-per-window DP, k-order anchor join, exchange block,\n"
    "# learned-ANI evaluator, seeding changes, packed-input seeding variant + host ingest, bucket-partition marker index, split join,\n"
    "# pipelined all-vs-all on two contexts); commands in tools/profile_round2.sh\n"
    "memcheck : tests/test_gpu_parity.py tests/test_gpu_exchange.py tests/test_gpu_learned_ani.py -k 'not ecoli and not mutant_series and not large_genomes and not two_gpu'\n"
    "           -> %s ; %s\n"
    "racecheck: tests/test_gpu_parity.py tests/test_gpu_learned_ani.py -k 'ragged or fragmented or walk_groups or all_vs_all_small or repeat_rich or hash_comparison or query_with_model or marker_index_build or screen_modes'\n"
    "           -> %s ; %s\n" % (mem[-2].strip(), mem[-1].strip("= "), race[-2].strip(), race[-1].strip("= ")))
print(open(os.path.join(P, "r2_launches_allvsall.txt")).read())
