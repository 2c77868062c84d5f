#!/bin/bash
# Two-GPU evidence (gpurun --gpus 2): weak-scaling bench line, all-vs-all with the device-resident sketch exchange.
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r1_2gpu.json
$T --master-port 29512 tools/allvsall_multi.py 40 1000000 2>&1 | grep -E "^world|identical" > gpurun_out/allvsall_multi_2gpu.log
$T --master-port 29513 tools/scale_bench.py --families 100 --members 10 --json gpurun_out/scale_cfg2_2gpu.json 2>&1 | grep -E "^(context|sketch calls|database|query|properties)" > gpurun_out/scale_cfg2_2gpu.log
$T --master-port 29514 tools/scale_bench.py --families 100 --members 100 --json gpurun_out/scale_cfg4_2gpu.json 2>&1 | grep -E "^(context|sketch calls|database|query|properties)" > gpurun_out/scale_cfg4_2gpu.log
tail -n 3 gpurun_out/allvsall_multi_2gpu.log gpurun_out/scale_cfg2_2gpu.log gpurun_out/scale_cfg4_2gpu.log; cut -c1-200 gpurun_out/bench_r1_2gpu.json
