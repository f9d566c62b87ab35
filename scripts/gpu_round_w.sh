#!/bin/bash
# posterior kernel: guarded table-driven exp against the library's binary64 exp; full suite on the new default
mkdir -p gpurun_out
LIB=bayes_od_rc_b200/lib/libbayesod.so
V=bayes_od_rc_b200/lib/variants
use() { cp $V/lib_$1.so $LIB; }
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rw_$name.json 2> gpurun_out/rw_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rw_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(18), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'k2 serial', d.get('serial',{}).get('stage_ms',{}).get('posterior'), 'path', d['config']['path_roofline_frac'])
except Exception as e:
    print('$name failed', e, open('gpurun_out/rw_$name.err').read()[-300:])
PY
}
use exptab
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/rw_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/rw_tests.log; tail -3 gpurun_out/rw_tests.log
KIT="--workload kitti_covar_b64_n20_k4"
for v in libexp exptab libexp exptab; do
  use $v
  run ${v}_b32
  run ${v}_kitti $KIT
done
use exptab
