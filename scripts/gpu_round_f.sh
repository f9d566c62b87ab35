#!/bin/bash
# graph replay: pipelined parity tests, then small-batch and full-batch pipelined throughput with and without graphs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined or per_level or batch_composition or full_size" > gpurun_out/rf_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/rf_tests.log
tail -4 gpurun_out/rf_tests.log
run() {
  name=$1; shift
  env $ENVV timeout 600 python bench.py --steps 300 --warmup 40 --no-e2e --no-cpu-baseline "$@" > gpurun_out/rf_$name.json 2> gpurun_out/rf_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rf_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms_per_step', d['ms_per_step'], 'img/s', d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1 ms', d['roofline'].get('launch_ms'))
except Exception as e:
    print('$name failed', e, open('gpurun_out/rf_$name.err').read()[-400:])
PY
}
ENVV="BOD_GRAPHS=1" run g_b32
ENVV="BOD_GRAPHS=0" run s_b32
ENVV="BOD_GRAPHS=1" run g_b4_p8 --batch 4 --pipeline 8
ENVV="BOD_GRAPHS=1" run g_b4_p16 --batch 4 --pipeline 16
ENVV="BOD_GRAPHS=0" run s_b4_p8 --batch 4 --pipeline 8
ENVV="BOD_GRAPHS=1" run g_b8_p8 --batch 8 --pipeline 8
ENVV="BOD_GRAPHS=1" run g_b1k8_p16 --workload bdd_covar_b1_k8 --pipeline 16
ENVV="BOD_GRAPHS=1" run g_b1k8_p8 --workload bdd_covar_b1_k8 --pipeline 8
ENVV="BOD_GRAPHS=0" run s_b1k8_p8 --workload bdd_covar_b1_k8 --pipeline 8
