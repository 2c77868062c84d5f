"""Rebuilds profiles/r1_* from what tools/profile_round.sh left in gpurun_out/ (run here, after gpurun merged the files)."""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, os.path.join(ROOT, "tools"))
import summarize_ncu as S
import contextlib, io

def capture(fn, *a):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()

KERNELS = ["seed_scan_kernel", "chain_dp_kernel", "window_walk_smem_kernel", "marker_screen_smem_kernel", "marker_join_kernel",
           "bucket_scatter_kernel", "bucket_rank_kernel", "region_gather_kernel", "match_count_kernel", "anchor_fill_kernel",
           "marker_sort_smem_kernel", "ani_reduce_kernel"]
for k in KERNELS:
    rep = os.path.join(G, "r1e_%s.ncu-rep" % k)
    if not os.path.exists(rep):
        continue
    wl = ("python tools/scale_bench.py --families 40 --members 10 (400 x 400 all-vs-all of 5 Mbp genomes)" if k == "marker_join_kernel"
          else "python bench.py --steps 2 --warmup 1 (configs[1], 505 Mbp sketched / 100 pairs chained per step)")
    open(os.path.join(P, "r1_%s_ncu_full.txt" % k), "w").write(
        capture(S.full, rep, "ncu --set full --clock-control none -k %s -c 1: %s; round 1, end of round" % (k, wl)))

open(os.path.join(P, "r1_launches_bench.txt"), "w").write(capture(
    S.launches, os.path.join(G, "launches_r1e.csv"),
    "ncu --metrics gpu__time_duration.sum --clock-control none -c 600: python bench.py --steps 2 --warmup 1 --skip-cpu-baseline "
    "(3 device-resident + 3 host-buffer steps of configs[1]; a host-buffer step launches seed_scan_kernel once per 16 MB copy chunk); "
    "round 1, end of round"))
open(os.path.join(P, "r1_launches_allvsall_batch.txt"), "w").write(capture(
    S.launches, os.path.join(G, "launches_ava.csv"),
    "ncu --metrics gpu__time_duration.sum --clock-control none -k <this library's query kernels>: python tools/scale_bench.py "
    "--families 40 --members 10 (400 x 400 all-vs-all, 3 454 pairs chained in 3 batches + warm-up); round 1, end of round"))

# one device-resident step cut from the launch list
rows = list(csv.reader(open(os.path.join(G, "launches_r1e.csv"))))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
H = rows[hi]; kn, mv, mn = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Name')
L = [(r[kn], float(r[mv].replace(',', '')) / 1e3) for r in rows[hi + 1:] if len(r) > mv and r[mn] == 'gpu__time_duration.sum']
idx = [i for i, (n, t) in enumerate(L) if 'seed_scan_kernel' in n]
step = L[idx[1]:idx[2]]
def short(n):
    m = re.search(r'skb::(?:<unnamed>::)?(\w+)\(', n)
    if m: return m.group(1)
    if 'RadixSort' in n: return 'CUB segmented/device radix sort (markers, window keys)'
    if 'Scan' in n: return 'CUB scan'
    if 'Select' in n or 'Compact' in n: return 'CUB unique/select'
    return 'CUB other'
agg = collections.OrderedDict()
for n, t in step:
    a = agg.setdefault(short(n), [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(v[1] for v in agg.values())
bench = json.loads(open(os.path.join(G, "bench_r1_final.json")).read().strip().splitlines()[-1])
out = ["# One device-resident step of bench.py configs[1] (505 Mbp sketched, 100 pairs chained), cut from the ncu launch list",
       "# (second step of gpurun_out/launches_r1e.csv). Cold-cache, serialised launch durations: compare shares, not absolutes.",
       "# total %.1f us over %d launches; the same step measured with CUDA events in bench.py: %.2f ms (the marker kernels overlap the"
       % (tot, len(step), bench["ms_per_step"]), "# bucket kernels there, and host round trips add gaps)"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%10.1f us %5.1f %%  %3d x  %s" % (v[1], 100 * v[1] / tot, v[0], k))
open(os.path.join(P, "r1_device_step_kernel_shares.txt"), "w").write("\n".join(out) + "\n")

open(os.path.join(P, "r1_bench_line_final.json"), "w").write(json.dumps(bench) + "\n")
ref = open(os.path.join(G, "bench_r1_reference.json")).read().strip().splitlines()[-1]
open(os.path.join(P, "r1_bench_line_reference.json"), "w").write(ref + "\n")
with open(os.path.join(P, "r1_scale_configs.txt"), "w") as f:
    f.write("# Full-size runs of BASELINE.json configs[2], [3], [4] on ONE B200 (tools/scale_bench.py, genomes generated on the device)\n")
    for name, cmd in (("scale_cfg2", "--families 100 --members 10"), ("scale_cfg3", "--families 500 --members 10 --mag-queries 500 --repeat 2"),
                      ("scale_cfg4", "--families 100 --members 100")):
        f.write("\n$ python tools/scale_bench.py %s\n" % cmd)
        f.write(open(os.path.join(G, name + ".log")).read())
    if os.path.exists(os.path.join(G, "scale_cfg4_2gpu.log")):
        f.write("\n# Two GPUs (gpurun --gpus 2, tools/profile_round_2gpu.sh): families dealt to the ranks, sketches exchanged once on the device\n")
        for name, cmd in (("scale_cfg2_2gpu", "--families 100 --members 10"), ("scale_cfg4_2gpu", "--families 100 --members 100")):
            f.write("\n$ torchrun --nproc-per-node 2 tools/scale_bench.py %s\n" % cmd)
            f.write(open(os.path.join(G, name + ".log")).read())
        f.write("\n$ torchrun --nproc-per-node 2 tools/allvsall_multi.py 40 1000000   (host genomes, pyskani_b200.parallel.all_vs_all)\n")
        f.write(open(os.path.join(G, "allvsall_multi_2gpu.log")).read())
    if os.path.exists(os.path.join(G, "scale_cfg4_8gpu.log")):
        f.write("\n# Eight GPUs (gpurun --gpus 8 -- bash tools/profile_round_ngpu.sh 8): BASELINE.json's target configuration on one 8 x B200 box\n")
        f.write("\n$ torchrun --nproc-per-node 8 tools/scale_bench.py --families 100 --members 100\n")
        f.write(open(os.path.join(G, "scale_cfg4_8gpu.log")).read())
if os.path.exists(os.path.join(G, "bench_r1_8gpu.json")):
    open(os.path.join(P, "r1_bench_line_8gpu.json"), "w").write(open(os.path.join(G, "bench_r1_8gpu.json")).read().strip().splitlines()[-1] + "\n")
if os.path.exists(os.path.join(G, "bench_r1_2gpu.json")):
    open(os.path.join(P, "r1_bench_line_2gpu.json"), "w").write(open(os.path.join(G, "bench_r1_2gpu.json")).read().strip().splitlines()[-1] + "\n")
print(open(os.path.join(P, "r1_device_step_kernel_shares.txt")).read())
