#!/bin/bash
# long runs (graph replay): two moments kernels in flight (BOD_K1_OVERLAP=1, held inputs)
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rz2_$name.json 2> gpurun_out/rz2_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rz2_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(18), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'), 'path', d['config']['path_roofline_frac'])
except Exception as e:
    print('$name failed', e, open('gpurun_out/rz2_$name.err').read()[-300:])
PY
}
run b32
BOD_K1_OVERLAP=1 run b32_ov
BOD_K1_OVERLAP=1 run b32_ov_p6 --pipeline 6
run kitti --workload kitti_covar_b64_n20_k4
BOD_K1_OVERLAP=1 run kitti_ov --workload kitti_covar_b64_n20_k4
BOD_K1_OVERLAP=1 run k8_ov --workload bdd_covar_b32_k8
run k8 --workload bdd_covar_b32_k8
BOD_K1_OVERLAP=1 run kraw_ov --workload kitti_raw_b64_n20_k4
run kraw --workload kitti_raw_b64_n20_k4
