#!/bin/bash
# Round-2 evidence on ONE GPU: launch list and ncu --set full captures of the hot kernels on the bench workload (all-vs-all of
# 1 000 x 5 Mbp genomes), the packed-input variant of the seeding kernel (host ingest path), compute-sanitizer over the parity
# tests, the bench lines of both arms.  Run through gpurun; outputs land in gpurun_out/ and are turned into profiles/r2_* by
# tools/refresh_profiles_r2.py.
set -x
B="python bench.py --steps 1 --warmup 1 --skip-parity --skip-cpu-baseline --skip-configs1 --skip-python-api"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_allvsall.csv $B > gpurun_out/ncu_bench.log 2>&1
for k in seed_scan_kernel chain_dp_thread_kernel match_count_kernel anchor_fill_kernel window_walk_smem_kernel marker_join_kernel marker_rank_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r2_$k $B > /dev/null 2>&1
done
# the first seed_scan launch of a step (250 genomes): the launch bench.py's roofline object describes
ncu --set full --clock-control none -k regex:seed_scan_kernel -c 1 -f -o gpurun_out/r2_seed_scan_first $B > /dev/null 2>&1
# the packed-input variant: every chunk of a 1.25 GB host batch compacted by the host threads (SKB_INGEST=pack); launches are
# per ~16 MB chunk, the 20th is captured
SKB_INGEST=pack ncu --set full --clock-control none --import-source on -k regex:seed_scan_kernel -s 20 -c 1 -f -o gpurun_out/r2_seed_scan_packed \
    python tools/ingest_sweep.py --threads 16 --reps 1 --only pack > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_exchange.py tests/test_gpu_learned_ani.py -q -x \
    -k "not ecoli and not mutant_series and not large_genomes and not two_gpu" > gpurun_out/r2_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_learned_ani.py -q -x \
    -k "ragged or fragmented or walk_groups or all_vs_all_small or repeat_rich or hash_comparison or query_with_model or marker_index_build or screen_modes" > gpurun_out/r2_racecheck.log 2>&1
tail -n 4 gpurun_out/r2_memcheck.log gpurun_out/r2_racecheck.log
python tools/ingest_sweep.py > gpurun_out/r2_ingest_sweep.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
tail -c 400 gpurun_out/r2_bench_n1.json; tail -c 400 gpurun_out/r2_bench_reference.json
