#!/bin/bash
# pipelined bench matrix: soft-NMS variant x lanes
mkdir -p gpurun_out
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline $BENCH_ARGS > gpurun_out/rd_$name.json 2> gpurun_out/rd_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rd_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms_per_step', d['ms_per_step'], 'img/s', d['value'], 'k1 launch_ms', d['roofline'].get('launch_ms'), 'lanes', d['config'].get('pipeline_depth'))
except Exception as e:
    print('$name failed', e)
PY
}
BENCH_ARGS="--pipeline 4" run def_p4 BOD_K3_SLIM=0
BENCH_ARGS="--pipeline 8" run def_p8 BOD_K3_SLIM=0
BENCH_ARGS="--pipeline 4" run slim_p4 BOD_K3_SLIM=1
BENCH_ARGS="--pipeline 8" run slim_p8 BOD_K3_SLIM=1
BENCH_ARGS="--pipeline 4" run t1024_p4 BOD_K3_THREADS=1024
BENCH_ARGS="--pipeline 4" run t256_p4 BOD_K3_THREADS=256
