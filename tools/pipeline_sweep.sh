#!/bin/bash
# e2e of bench.py on one GPU with host buffers: sketch-all-then-query against the pipelined all-vs-all, on the same box
run() { # name, bench args...
  name=$1; shift
  timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-configs1 --skip-parity "$@" > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err
  tail -c 200 gpurun_out/r2f_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r2f_$name.json')); print('$name', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['e2e']['phase_ms'])"
}
run plain_a --no-pipeline
run piped_a
run plain_b --no-pipeline
run piped_b
