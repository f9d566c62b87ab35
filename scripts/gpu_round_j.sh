#!/bin/bash
# reproducibility of round i's two outliers (old build at B = 32, new build on the KITTI shape), A/B/A/B on one box
mkdir -p gpurun_out
LIB=bayes_od_rc_b200/lib/libbayesod.so
V=bayes_od_rc_b200/lib/variants
use() { cp $V/lib_$1.so $LIB; }
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rj_$name.json 2> gpurun_out/rj_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rj_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms_per_step', d['ms_per_step'], 'img/s', d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1 ms', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'), 'serial', d.get('serial',{}).get('stage_ms'))
except Exception as e:
    print('$name failed', e, open('gpurun_out/rj_$name.err').read()[-400:])
PY
}
KIT="--workload kitti_covar_b64_n20_k4"
use new;  run new_b32_1
use old;  run old_b32_1
use new;  run new_kitti_1 $KIT
use old;  run old_kitti_1 $KIT
use new;  BOD_K1_UNROLL=0 run new_kitti_generic $KIT
use new;  BOD_GRAPHS=0 run new_kitti_streams $KIT
use old;  BOD_GRAPHS=0 run old_kitti_streams $KIT
use new;  run new_b32_2
use old;  run old_b32_2
use new;  run new_kitti_2 $KIT
use old;  run old_kitti_2 $KIT
use new;  run new_kitti_raw --workload kitti_raw_b64_n20_k4
use old;  run old_kitti_raw --workload kitti_raw_b64_n20_k4
use new;  run new_stress --workload stress_b16_n40_k11
use old;  run old_stress --workload stress_b16_n40_k11
use new2; run new2_b1k8 --workload bdd_covar_b1_k8
use new2; run new2_b4 --batch 4
use new2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "softmax_rows or pipelined or streaming or philox" 2>&1 | tail -3
