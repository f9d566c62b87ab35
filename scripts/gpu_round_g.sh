#!/bin/bash
# bench.py contract run (both arms) + tiny workload sanity
mkdir -p gpurun_out
timeout 600 python bench.py --workload tiny --steps 20 --warmup 3 > gpurun_out/rg_tiny.json 2> gpurun_out/rg_tiny.err; tail -c 600 gpurun_out/rg_tiny.err
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/rg_bench.json 2> gpurun_out/rg_bench.err; tail -c 600 gpurun_out/rg_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/rg_ref.json 2> gpurun_out/rg_ref.err; tail -c 600 gpurun_out/rg_ref.err
python - <<'PY'
import json
for f in ('rg_tiny','rg_bench','rg_ref'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','verified','steps')}, d['config'].get('workload'))
        if 'verification' in d: print('   ', d['verification'])
        if 'e2e' in d: print('   e2e', d['e2e'].get('value'), d['e2e'].get('h2d_gbs'), d['e2e'].get('h2d_probe_gbs'), d['e2e'].get('pcie_frac'))
        if 'roofline' in d: print('   roofline', d['roofline'].get('frac'), d['roofline'].get('launch_ms'), d['config'].get('path_roofline_frac'))
    except Exception as e:
        print(f, 'failed', e)
PY
