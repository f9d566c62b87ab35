#!/bin/bash
# one graph per run (posterior kernel right behind the scan, "head done" as an event node), with and without the timing
# event nodes around the moments kernel
mkdir -p gpurun_out
LIB=bayes_od_rc_b200/lib/libbayesod.so
V=bayes_od_rc_b200/lib/variants
use() { cp $V/lib_$1.so $LIB; }
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/ro_$name.json 2> gpurun_out/ro_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ro_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(26), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'alone', d['roofline'].get('alone',{}).get('launch_ms'))
except Exception as e:
    print('$name failed', e, open('gpurun_out/ro_$name.err').read()[-300:])
PY
}
tl() { name=$1; shift; env "$@" timeout 300 python scripts/timeline.py > gpurun_out/ro_tl_$name.log 2>&1; echo "== $name: $@"; tail -5 gpurun_out/ro_tl_$name.log | cut -c1-330; }
use diag
tl m1 BOD_MERGED_GRAPH=1 BOD_NO_K1_EVENTS=1 BOD_NO_STAGE_EVENTS=1
tl m2 BOD_MERGED_GRAPH=1
use new7
run base
BOD_MERGED_GRAPH=1 run merged
BOD_MERGED_GRAPH=1 BOD_NO_K1_EVENTS=1 run merged_noev
BOD_MERGED_GRAPH=1 BOD_NO_K1_EVENTS=1 run merged_noev_k8 --workload bdd_covar_b32_k8
run base_k8 --workload bdd_covar_b32_k8
BOD_MERGED_GRAPH=1 BOD_NO_K1_EVENTS=1 BOD_K1_NS=5 run merged_noev_kitti --workload kitti_covar_b64_n20_k4
BOD_K1_NS=5 run base_kitti --workload kitti_covar_b64_n20_k4
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined or streaming" 2>&1 | tail -2
BOD_MERGED_GRAPH=1 BOD_GRAPHS=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined or streaming" 2>&1 | tail -2
