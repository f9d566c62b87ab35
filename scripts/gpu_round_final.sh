#!/bin/bash
# last GPU session of round 2: full suite, sanitizer passes, one full bench line per workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/rfinal_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/rfinal_tests.log; tail -3 gpurun_out/rfinal_tests.log
SAN_SKIP_PDQ=1 bash scripts/sanitize_gpu.sh
EV_STEPS=200 bash scripts/evidence_r2.sh
