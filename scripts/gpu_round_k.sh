#!/bin/bash
# K = 4 ring depth vs unrolling in the pipelined step; K1 chained behind the previous K1 (BOD_K1_CHAIN); K1 alone on fewer
# SMs; cost of each tail kernel in the pipelined step (BOD_DEBUG_SKIP, diagnostics build)
mkdir -p gpurun_out
LIB=bayes_od_rc_b200/lib/libbayesod.so
V=bayes_od_rc_b200/lib/variants
use() { cp $V/lib_$1.so $LIB; }
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rk_$name.json 2> gpurun_out/rk_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rk_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms_per_step', d['ms_per_step'], 'img/s', d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1 ms', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'), 'serial', d.get('serial',{}).get('stage_ms'))
except Exception as e:
    print('$name failed', e, open('gpurun_out/rk_$name.err').read()[-400:])
PY
}
KIT="--workload kitti_covar_b64_n20_k4"
use new3
run b32
BOD_K1_CHAIN=1 run b32_chain
BOD_K1_CHAIN=1 run b32_chain_p6 --pipeline 6
run kitti $KIT
BOD_K1_CHAIN=1 run kitti_chain $KIT
BOD_K1_UNROLL=0 run kitti_generic13 $KIT
BOD_K1_UNROLL=0 BOD_K1_NS=10 run kitti_generic10 $KIT
BOD_K1_UNROLL=0 BOD_K1_CHAIN=1 run kitti_generic13_chain $KIT
BOD_K1_SMS=116 run b32_sms116 --pipeline 1
BOD_K1_SMS=132 run b32_sms132 --pipeline 1
use diag
for m in 0 1 2 4 3 6 7; do BOD_DEBUG_SKIP=$m run b32_skip$m; done
for m in 0 1 2 4 7; do BOD_DEBUG_SKIP=$m run kitti_skip$m $KIT; done
use new3
