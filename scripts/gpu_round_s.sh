#!/bin/bash
# PDQ bench with the CPU oracle leg in the same run; one-image runs with more lanes; small-batch lines with the launch clock
mkdir -p gpurun_out
timeout 600 python scripts/pdq_bench.py 64 60 20 > gpurun_out/pdq_bench_r2.json 2> gpurun_out/pdq_bench_r2.err; tail -c 1500 gpurun_out/pdq_bench_r2.json; tail -3 gpurun_out/pdq_bench_r2.err
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rs_$name.json 2> gpurun_out/rs_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rs_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(14), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'), 'path', d['config']['path_roofline_frac'])
except Exception as e:
    print('$name failed', e, open('gpurun_out/rs_$name.err').read()[-300:])
PY
}
run b1k8_p8 --workload bdd_covar_b1_k8 --pipeline 8
run b1k8_p12 --workload bdd_covar_b1_k8 --pipeline 12
run b1k8_p16 --workload bdd_covar_b1_k8 --pipeline 16
run b1k8_p16_nf --workload bdd_covar_b1_k8 --pipeline 16 --no-stream-fetch
run b4 --batch 4
run b4_p12 --batch 4 --pipeline 12
run b8 --batch 8
run b16 --batch 16
run b16_p8 --batch 16 --pipeline 8
run kitti8 --workload kitti_covar_b64_n20_k4 --batch 8
