#!/bin/bash
# boundary work out of the way of the next moments kernel: scan with the tail (BOD_SCAN_TAIL=1 in graph mode), no timing
# events around the moments kernel, posterior kernel as a small co-resident grid (96-thread CTAs, BOD_K2_CTAS), fusion
# kernel with a capped grid (BOD_K4_GRIDX); timelines of the candidates, then bench numbers
mkdir -p gpurun_out
LIB=bayes_od_rc_b200/lib/libbayesod.so
V=bayes_od_rc_b200/lib/variants
use() { cp $V/lib_$1.so $LIB; }
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rn_$name.json 2> gpurun_out/rn_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rn_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(26), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'))
except Exception as e:
    print('$name failed', e, open('gpurun_out/rn_$name.err').read()[-300:])
PY
}
tl() { name=$1; shift; env "$@" timeout 300 python scripts/timeline.py > gpurun_out/rn_tl_$name.log 2>&1; echo "== $name: $@"; tail -5 gpurun_out/rn_tl_$name.log | cut -c1-330; }
export BOD_NO_STAGE_EVENTS=1
use diag96
tl a BOD_NO_K1_EVENTS=1 BOD_SCAN_TAIL=1 BOD_K2_CTAS=1
tl b BOD_NO_K1_EVENTS=1 BOD_SCAN_TAIL=1 BOD_K2_CTAS=1 BOD_K4_GRIDX=5
tl c BOD_NO_K1_EVENTS=1 BOD_SCAN_TAIL=1 BOD_K2_CTAS=2 BOD_K4_GRIDX=5
tl d BOD_NO_K1_EVENTS=1 BOD_SCAN_TAIL=1 BOD_K2_CTAS=1 BOD_K4_GRIDX=5 BOD_GRAPHS=0
use diag
tl e BOD_NO_K1_EVENTS=1 BOD_SCAN_TAIL=1
tl f BOD_NO_K1_EVENTS=1
unset BOD_NO_STAGE_EVENTS
use k2_96
BOD_NO_K1_EVENTS=1 BOD_SCAN_TAIL=1 BOD_K2_CTAS=1 run a
BOD_NO_K1_EVENTS=1 BOD_SCAN_TAIL=1 BOD_K2_CTAS=1 BOD_K4_GRIDX=5 run b
BOD_NO_K1_EVENTS=1 BOD_SCAN_TAIL=1 BOD_K2_CTAS=2 BOD_K4_GRIDX=5 run c
BOD_NO_K1_EVENTS=1 BOD_SCAN_TAIL=1 BOD_K2_CTAS=1 BOD_K4_GRIDX=5 run b_p6 --pipeline 6
BOD_NO_K1_EVENTS=1 BOD_SCAN_TAIL=1 BOD_K2_CTAS=1 BOD_K4_GRIDX=5 BOD_GRAPHS=0 run d
use new6
BOD_NO_K1_EVENTS=1 BOD_SCAN_TAIL=1 run e
BOD_NO_K1_EVENTS=1 run f
run base
