#!/bin/bash
# diagnostics build: slim soft-NMS variant in the pipelined step
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline $BENCH_ARGS > gpurun_out/re_$name.json 2> gpurun_out/re_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/re_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms_per_step', d['ms_per_step'], 'k1 launch_ms', d['roofline'].get('launch_ms'))
except Exception as e:
    print('$name failed', e)
PY
}
BOD_K3_SLIM=1 timeout 300 python scripts/diag_stages.py 2>&1 | tail -1
BENCH_ARGS="--pipeline 4" run slim_p4 BOD_K3_SLIM=1
BENCH_ARGS="--pipeline 6" run slim_p6 BOD_K3_SLIM=1
BENCH_ARGS="--pipeline 8" run slim_p8 BOD_K3_SLIM=1
BENCH_ARGS="--pipeline 8" run slim_p8_nok4 BOD_K3_SLIM=1 BOD_DEBUG_SKIP=4
BENCH_ARGS="--pipeline 8" run slim_p8_nok2k4 BOD_K3_SLIM=1 BOD_DEBUG_SKIP=5
