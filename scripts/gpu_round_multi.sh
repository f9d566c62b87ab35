#!/bin/bash
# multi-GPU bench (weak + strong scaling blocks); N = number of GPUs of this box
N=${1:-8}
mkdir -p gpurun_out
for wl in bdd_covar_b32_k11 kitti_covar_b64_n20_k4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $wl --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/multi_${wl}_n$N.json 2> gpurun_out/multi_${wl}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/multi_${wl}_n$N.json').read().strip().splitlines()[-1])
    print('$wl N=$N weak', d['value'], 'ms', d['ms_per_step'], 'strong', d.get('strong'), 'e2e', d.get('e2e',{}).get('value'))
except Exception as e:
    print('$wl failed', e, open('gpurun_out/multi_${wl}_n$N.err').read()[-800:])
PY
done
