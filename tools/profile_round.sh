#!/bin/bash
# End-of-round evidence: launch list + ncu --set full captures of the hot kernels on the bench workload, the bench line,
# and the full-size runs of BASELINE.json configs[2..4].  Run through gpurun; outputs land in gpurun_out/.
set -x
B="python bench.py --steps 2 --warmup 1 --skip-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1e.csv $B > gpurun_out/bench_under_ncu.log 2>&1
for k in seed_scan_kernel chain_dp_kernel window_walk_smem_kernel marker_screen_smem_kernel bucket_scatter_kernel bucket_rank_kernel region_gather_kernel match_count_kernel anchor_fill_kernel marker_sort_smem_kernel; do
  ncu --set full --clock-control none --import-source on -k $k -s 1 -c 1 -f -o gpurun_out/r1e_$k $B > /dev/null 2>&1
done
# all-vs-all batch shape: only this library's query kernels are profiled (torch's generator kernels run unprofiled)
K='regex:chain_dp|window_walk|match_count|anchor_fill|ani_reduce|window_keys|marker_join|screen_decide|marker_postings|marker_index'
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/launches_ava.csv python tools/scale_bench.py --families 40 --members 10 > gpurun_out/ava.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k marker_join_kernel -c 1 -f -o gpurun_out/r1e_marker_join_kernel python tools/scale_bench.py --families 40 --members 10 > /dev/null 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1_final.json 2>gpurun_out/bench_r1_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2>gpurun_out/bench_r1_reference.err
timeout 600 python tools/scale_bench.py --families 100 --members 10 --json gpurun_out/scale_cfg2.json > gpurun_out/scale_cfg2.log 2>&1
timeout 600 python tools/scale_bench.py --families 500 --members 10 --mag-queries 500 --repeat 2 --json gpurun_out/scale_cfg3.json > gpurun_out/scale_cfg3.log 2>&1
timeout 900 python tools/scale_bench.py --families 100 --members 100 --json gpurun_out/scale_cfg4.json > gpurun_out/scale_cfg4.log 2>&1
tail -3 gpurun_out/scale_cfg2.log gpurun_out/scale_cfg3.log gpurun_out/scale_cfg4.log
tail -1 gpurun_out/bench_r1_final.json | cut -c1-200
