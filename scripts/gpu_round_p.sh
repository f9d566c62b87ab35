#!/bin/bash
# launch clock instead of timing events around the moments kernel; ring cap of pipelined contexts: full suite + every workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/rp_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/rp_tests.log; tail -3 gpurun_out/rp_tests.log
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rp_$name.json 2> gpurun_out/rp_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rp_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(14), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'), 'path', d['config']['path_roofline_frac'])
except Exception as e:
    print('$name failed', e, open('gpurun_out/rp_$name.err').read()[-300:])
PY
}
run b32
run b4 --batch 4
run b8 --batch 8
run b16 --batch 16
run b1k8 --workload bdd_covar_b1_k8
run k8 --workload bdd_covar_b32_k8
run kendall --workload bdd_kendall_b8_k8
run kitti --workload kitti_covar_b64_n20_k4
run kitti8 --workload kitti_covar_b64_n20_k4 --batch 8
run kraw --workload kitti_raw_b64_n20_k4
run stress --workload stress_b16_n40_k11
