#!/bin/bash
# device timelines of the pipelined step (diagnostics build); unrolled K = 4 moments kernel with a shallower ring
mkdir -p gpurun_out
LIB=bayes_od_rc_b200/lib/libbayesod.so
V=bayes_od_rc_b200/lib/variants
use() { cp $V/lib_$1.so $LIB; }
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rm_$name.json 2> gpurun_out/rm_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rm_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(22), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'))
except Exception as e:
    print('$name failed', e, open('gpurun_out/rm_$name.err').read()[-300:])
PY
}
use diag
timeout 300 python scripts/timeline.py > gpurun_out/rm_tl_b32.log 2>&1; cat gpurun_out/rm_tl_b32.log
BOD_GRAPHS=0 timeout 300 python scripts/timeline.py > gpurun_out/rm_tl_b32_streams.log 2>&1; cat gpurun_out/rm_tl_b32_streams.log
TL_WORKLOAD=kitti_covar_b64_n20_k4 timeout 300 python scripts/timeline.py > gpurun_out/rm_tl_kitti.log 2>&1; cat gpurun_out/rm_tl_kitti.log
TL_WORKLOAD=kitti_covar_b64_n20_k4 BOD_K1_UNROLL=0 timeout 300 python scripts/timeline.py > gpurun_out/rm_tl_kitti_generic.log 2>&1; cat gpurun_out/rm_tl_kitti_generic.log
TL_B=4 TL_LANES=8 timeout 300 python scripts/timeline.py > gpurun_out/rm_tl_b4.log 2>&1; cat gpurun_out/rm_tl_b4.log
use new5
KIT="--workload kitti_covar_b64_n20_k4"
BOD_K1_NS=5 run kitti_u5 $KIT
BOD_K1_NS=4 run kitti_u4 $KIT
BOD_K1_NS=5 run kraw_u5 --workload kitti_raw_b64_n20_k4
BOD_K1_UNROLL=0 BOD_K1_NS=10 run kraw_g10 --workload kitti_raw_b64_n20_k4
BOD_K1_UNROLL=0 BOD_K1_NS=8 run kitti_g8 $KIT
