#!/bin/bash
# N-GPU evidence (gpurun --gpus N -- bash tools/profile_round_ngpu.sh N): weak-scaling bench line + configs[4] all-vs-all
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 --skip-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_r1_${N}gpu.json
$T --master-port 29524 tools/scale_bench.py --families 100 --members 100 --json gpurun_out/scale_cfg4_${N}gpu.json 2>&1 | grep -E "^(context|sketch calls|database|query|properties)" > gpurun_out/scale_cfg4_${N}gpu.log
cat gpurun_out/scale_cfg4_${N}gpu.log; cut -c1-200 gpurun_out/bench_r1_${N}gpu.json
