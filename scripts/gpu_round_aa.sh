#!/bin/bash
# scan fused into the moments kernel (the CTA that finishes an image's last tile scans it): full suite, then with and without it
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/raa_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/raa_tests.log; tail -3 gpurun_out/raa_tests.log
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline "$@" > gpurun_out/raa_$name.json 2> gpurun_out/raa_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/raa_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(18), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'), 'path', d['config']['path_roofline_frac'], 'verified', d.get('verified'), 'launches', d.get('gpu_launches'))
except Exception as e:
    print('$name failed', e, open('gpurun_out/raa_$name.err').read()[-300:])
PY
}
run b32
BOD_FUSED_SCAN=0 run b32_noscanfuse
run b32_2
BOD_FUSED_SCAN=0 run b32_noscanfuse_2
run b4 --batch 4
BOD_FUSED_SCAN=0 run b4_noscanfuse --batch 4
run b1k8 --workload bdd_covar_b1_k8
BOD_FUSED_SCAN=0 run b1k8_noscanfuse --workload bdd_covar_b1_k8
run kitti --workload kitti_covar_b64_n20_k4
BOD_FUSED_SCAN=0 run kitti_noscanfuse --workload kitti_covar_b64_n20_k4
