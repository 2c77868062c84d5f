#!/bin/bash
# Last pass of round 2 on ONE GPU: the whole GPU test suite, both bench arms, the launch list of one step and one ncu capture of
# the seeding kernel of the final build.  Run through gpurun; tools/refresh_profiles_r2.py turns gpurun_out/ into profiles/.
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2f_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
tail -c 300 gpurun_out/r2f_bench_reference.json; head -c 600 gpurun_out/r2f_bench_n1.json
B="python bench.py --steps 1 --warmup 1 --skip-parity --skip-cpu-baseline --skip-configs1 --skip-python-api"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_allvsall.csv $B > gpurun_out/r2f_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:seed_scan_kernel -s 1 -c 1 -f -o gpurun_out/r2f_seed_scan_kernel $B > /dev/null 2>&1
ls -la gpurun_out/r2f_*
