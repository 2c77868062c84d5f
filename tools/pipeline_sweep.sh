#!/bin/bash
# e2e of bench.py on one GPU with host buffers: sketch-all-then-query against the pipelined all-vs-all, alternating on the
# same box (the boxes differ by several ms among themselves)
run() { # name, env/bench args...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 10 --warmup 4 --skip-cpu-baseline --skip-configs1 --skip-parity --skip-python-api $EXTRA > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err
  tail -c 200 gpurun_out/r2f_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/r2f_$name.json')); print('$name', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['e2e']['phase_ms'])"
}
for rep in 1 2 3; do
  EXTRA= run plain_$rep X=1
  EXTRA=--pipeline run piped_$rep X=1
  EXTRA=--pipeline run piped16_$rep BENCH_TEAM=16
done
