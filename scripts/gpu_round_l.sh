#!/bin/bash
# more lanes for the tail-heavy KITTI shape; the posterior kernel with a small grid (BOD_K2_CTAS) so that it runs beside the
# next moments kernel instead of in front of it, in 128- and 96-thread CTAs
mkdir -p gpurun_out
LIB=bayes_od_rc_b200/lib/libbayesod.so
V=bayes_od_rc_b200/lib/variants
use() { cp $V/lib_$1.so $LIB; }
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rl_$name.json 2> gpurun_out/rl_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rl_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(22), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'), 'k2 serial', d.get('serial',{}).get('stage_ms',{}).get('posterior'))
except Exception as e:
    print('$name failed', e, open('gpurun_out/rl_$name.err').read()[-300:])
PY
}
KIT="--workload kitti_covar_b64_n20_k4"
use new4
run kitti_p4 $KIT
run kitti_p6 $KIT --pipeline 6
run kitti_p8 $KIT --pipeline 8
BOD_K1_UNROLL=0 run kitti_gen_p8 $KIT --pipeline 8
run kraw_p8 --workload kitti_raw_b64_n20_k4 --pipeline 8
for n in 1 2 3; do BOD_K2_CTAS=$n run b32_k2c$n; done
BOD_K2_CTAS=2 run b32_k2c2_p6 --pipeline 6
BOD_K2_CTAS=2 run kitti_k2c2_p8 $KIT --pipeline 8
use k2_96
for n in 1 2 6; do BOD_K2_CTAS=$n run b32_k296_c$n; done
BOD_K2_CTAS=1 run b32_k296_c1_p6 --pipeline 6
BOD_K2_CTAS=1 run kitti_k296_c1_p8 $KIT --pipeline 8
BOD_K2_CTAS=2 run kitti_k296_c2_p8 $KIT --pipeline 8
use new4
