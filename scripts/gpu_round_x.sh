#!/bin/bash
# consecutive runs no longer chained through the caller's stream (BOD_HOLD_INPUTS=1), with one and two head streams
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rx_$name.json 2> gpurun_out/rx_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rx_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(18), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'), 'path', d['config']['path_roofline_frac'])
except Exception as e:
    print('$name failed', e, open('gpurun_out/rx_$name.err').read()[-300:])
PY
}
run b4 --batch 4
BOD_HOLD_INPUTS=1 run b4_hold --batch 4
BOD_HOLD_INPUTS=1 BOD_HEADS=2 run b4_hold_h2 --batch 4
run b8 --batch 8
BOD_HOLD_INPUTS=1 BOD_HEADS=2 run b8_hold_h2 --batch 8
BOD_HOLD_INPUTS=1 BOD_HEADS=2 run b1k8_hold_h2 --workload bdd_covar_b1_k8
BOD_HOLD_INPUTS=1 BOD_HEADS=2 run kitti8_hold_h2 --workload kitti_covar_b64_n20_k4 --batch 8
run kitti8 --workload kitti_covar_b64_n20_k4 --batch 8
BOD_HOLD_INPUTS=1 run b32_hold
BOD_HOLD_INPUTS=1 BOD_HEADS=2 BOD_GRAPHS=0 run b32_hold_h2_streams
