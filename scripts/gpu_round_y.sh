#!/bin/bash
# bod_set_input_hold in the product: streaming tests, small-batch and default lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streaming or pipelined or launch_clock" 2>&1 | tail -2
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/ry_$name.json 2> gpurun_out/ry_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ry_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(18), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'), 'path', d['config']['path_roofline_frac'])
except Exception as e:
    print('$name failed', e, open('gpurun_out/ry_$name.err').read()[-300:])
PY
}
run b32
run b32_nohold --no-input-hold
run b4 --batch 4
run b4_p8 --batch 4 --pipeline 8
run b4_p16 --batch 4 --pipeline 16
run b2 --batch 2
run b8 --batch 8
run b8_p12 --batch 8 --pipeline 12
run b16 --batch 16
run kendall --workload bdd_kendall_b8_k8
run kitti8 --workload kitti_covar_b64_n20_k4 --batch 8
run b1k8 --workload bdd_covar_b1_k8
timeout 600 python bench.py --workload bdd_covar_b32_k11 --batch 4 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/ry_b4_full.json 2> gpurun_out/ry_b4_full.err; python -c "
import json; d=json.loads(open('gpurun_out/ry_b4_full.json').read().strip().splitlines()[-1]); print('b4 full verified', d.get('verified'), d['value'])"
