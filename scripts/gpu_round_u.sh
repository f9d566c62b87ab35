#!/bin/bash
# posterior kernel: gathers kK2Prefetch samples ahead as cp.async copies (0 = one sample ahead in registers, the round-1 form)
mkdir -p gpurun_out
LIB=bayes_od_rc_b200/lib/libbayesod.so
V=bayes_od_rc_b200/lib/variants
use() { cp $V/lib_$1.so $LIB; }
run() {
  name=$1; shift
  timeout 600 python bench.py --steps 300 --warmup 20 --no-e2e --no-cpu-baseline --no-verify "$@" > gpurun_out/ru_$name.json 2> gpurun_out/ru_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ru_$name.json').read().strip().splitlines()[-1])
    print('$name'.ljust(18), d['ms_per_step'], d['value'], 'lanes', d['config'].get('pipeline_depth'), 'k1', d['roofline'].get('launch_ms'), 'k2 serial', d.get('serial',{}).get('stage_ms',{}).get('posterior'), 'path', d['config']['path_roofline_frac'])
except Exception as e:
    print('$name failed', e, open('gpurun_out/ru_$name.err').read()[-300:])
PY
}
use k2p3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or synthetic_batch or full_size or random_config or per_level or host_path" 2>&1 | tail -2
KIT="--workload kitti_covar_b64_n20_k4"
for v in k2p0 k2p2 k2p3 k2p4 k2p3mb6; do
  use $v
  run ${v}_b32
  run ${v}_kitti $KIT
done
use k2p3
run k2p3_kraw --workload kitti_raw_b64_n20_k4
run k2p3_stress --workload stress_b16_n40_k11
run k2p3_b4 --batch 4
