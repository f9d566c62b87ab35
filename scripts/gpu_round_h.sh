#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/rh_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/rh_tests.log; tail -3 gpurun_out/rh_tests.log
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rh_$name.json 2> gpurun_out/rh_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rh_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms_per_step', d['ms_per_step'], 'img/s', d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1 ms', d['roofline'].get('launch_ms'))
except Exception as e:
    print('$name failed', e, open('gpurun_out/rh_$name.err').read()[-400:])
PY
}
run b32
BOD_GRAPHS=0 run b32_streams
run b4 --batch 4
run b8 --batch 8
run b1k8 --workload bdd_covar_b1_k8
run kendall --workload bdd_kendall_b8_k8
run kitti8 --workload kitti_covar_b64_n20_k4 --batch 8
run k8 --workload bdd_covar_b32_k8
