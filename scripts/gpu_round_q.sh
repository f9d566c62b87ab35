#!/bin/bash
# round-2 evidence: full GPU suite, ncu launch list + --set full capture of every stage kernel, one full bench line per workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/rq_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/rq_tests.log; tail -3 gpurun_out/rq_tests.log
bash scripts/profile_gpu.sh r2 main
EV_STEPS=200 bash scripts/evidence_r2.sh
