"""Turns gpurun_out/*.ncu-rep / launch CSVs into the text summaries kept under profiles/."""
import csv, io, subprocess, sys, collections

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum']


def full(rep, title):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U, V = rows[0], rows[1], rows[2]
    print("# " + title)
    for i, h in enumerate(H):
        if h in KEYS or ('warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')):
            print(f"{h:95s} {V[i]:>18s} {U[i]}")


def launches(path, title):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    H = rows[hi]; kn, mv, mn = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Name')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
            continue
        name = r[kn].split('(')[0]
        name = name[-70:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += float(r[mv].replace(',', '')) / 1e3
    tot = sum(v[1] for v in agg.values())
    print("# " + title)
    print(f"# total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.1f} us {v[0]:5d} launches {100 * v[1] / tot:5.1f} %  {k}")


if __name__ == "__main__":
    kind, path, title = sys.argv[1], sys.argv[2], sys.argv[3]
    (full if kind == "full" else launches)(path, title)
