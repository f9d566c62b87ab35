#!/bin/bash
# posterior kernel prefetch through L1 (cp.async.ca) against L2-only (cp.async.cg, round u)
mkdir -p gpurun_out
LIB=bayes_od_rc_b200/lib/libbayesod.so
V=bayes_od_rc_b200/lib/variants
use() { cp $V/lib_$1.so $LIB; }
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/rv_$name.json 2> gpurun_out/rv_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/rv_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(18), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'k2 serial', d.get('serial',{}).get('stage_ms',{}).get('posterior'), 'path', d['config']['path_roofline_frac'])
except Exception as e:
    print('$name failed', e, open('gpurun_out/rv_$name.err').read()[-300:])
PY
}
KIT="--workload kitti_covar_b64_n20_k4"
for v in k2p1ca k2p2ca k2p3ca k2p2; do
  use $v
  run ${v}_b32
  run ${v}_kitti $KIT
done
use k2p2ca
run k2p2ca_kraw --workload kitti_raw_b64_n20_k4
run k2p2ca_k8 --workload bdd_covar_b32_k8
run k2p2ca_kendall --workload bdd_kendall_b8_k8
