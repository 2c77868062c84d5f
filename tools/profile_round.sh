set -x
B="python bench.py --steps 2 --warmup 1 --skip-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1e.csv $B > gpurun_out/bench_under_ncu.log 2>&1
for k in seed_scan_kernel chain_dp_kernel window_walk_smem_kernel marker_screen_smem_kernel bucket_scatter_kernel bucket_rank_kernel region_gather_kernel match_count_kernel anchor_fill_kernel; do
  ncu --set full --clock-control none --import-source on -k $k -s 1 -c 1 -f -o gpurun_out/r1e_$k $B > /dev/null 2>&1
done
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1_final.json 2>gpurun_out/bench_r1_final.err
tail -1 gpurun_out/bench_r1_final.json | cut -c1-300
