#!/bin/bash
# One bench.py JSON line per BASELINE.json workload (roofline, cpu_baseline, e2e, verification included), for profiles/.
mkdir -p gpurun_out
for wl in bdd_covar_b32_k11 bdd_covar_b32_k8 bdd_kendall_b8_k8 kitti_covar_b64_n20_k4 kitti_raw_b64_n20_k4 stress_b16_n40_k11 bdd_covar_b1_k8; do
  timeout 900 python bench.py --workload $wl --steps ${EV_STEPS:-100} --warmup 10 > gpurun_out/ev_$wl.json 2> gpurun_out/ev_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ev_$wl.json').read().strip().splitlines()[-1])
    print('$wl', 'img/s', d['value'], 'ms', d['ms_per_step'], 'path frac', d['config']['path_roofline_frac'], 'k1 frac', d['roofline']['frac'], 'alone', d['roofline'].get('alone',{}).get('frac'), 'verified', d.get('verified'), 'e2e', d.get('e2e',{}).get('value'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'S', d['config']['mean_survivors'])
except Exception as e:
    print('$wl failed', e, open('gpurun_out/ev_$wl.err').read()[-500:])
PY
done
