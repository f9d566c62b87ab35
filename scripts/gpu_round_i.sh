#!/bin/bash
# round 2, session 3: K1 consumer loop v2 (packed pairs, unrolled rounds) against the previous build; K2 in 96-thread CTAs;
# two head streams for small runs; soft-NMS phase timers (timers-only diagnostics build)
mkdir -p gpurun_out
LIB=bayes_od_rc_b200/lib/libbayesod.so
V=bayes_od_rc_b200/lib/variants
use() { cp $V/lib_$1.so $LIB; }
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/ri_$name.json 2> gpurun_out/ri_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ri_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms_per_step', d['ms_per_step'], 'img/s', d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1 ms', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'), 'serial', d.get('serial',{}).get('stage_ms'))
except Exception as e:
    print('$name failed', e, open('gpurun_out/ri_$name.err').read()[-400:])
PY
}
use new
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/ri_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ri_tests.log; tail -3 gpurun_out/ri_tests.log
use old; run old_b32
use new; run new_b32
BOD_K1_UNROLL=0 run new_b32_generic
use k2_96; run k2_96_b32
use new
run b4 --batch 4
BOD_HEADS=2 run b4_h2 --batch 4
run b8 --batch 8
BOD_HEADS=2 run b8_h2 --batch 8
run b1k8 --workload bdd_covar_b1_k8
BOD_HEADS=2 run b1k8_h2 --workload bdd_covar_b1_k8
run k8 --workload bdd_covar_b32_k8
run kitti --workload kitti_covar_b64_n20_k4
run kendall --workload bdd_kendall_b8_k8
BOD_HEADS=2 run kendall_h2 --workload bdd_kendall_b8_k8
use diag
BOD_K3_DEBUG=1 timeout 300 python scripts/diag_stages.py 2>&1 | tail -16 > gpurun_out/ri_diag.log 2>&1
cat gpurun_out/ri_diag.log
use new
