#!/bin/bash
# GPU check of a K3 change: parity subset, then stage timings per CTA size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "synthetic_batch or softnms or golden or prefilter or full" > gpurun_out/ra_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ra_tests.log
tail -5 gpurun_out/ra_tests.log
for nt in ${K3_NTS:-512 1024 256}; do
  BOD_K3_THREADS=$nt timeout 300 python scripts/diag_stages.py 2>&1 | tail -1
  BOD_K3_THREADS=$nt DIAG_K=8 timeout 300 python scripts/diag_stages.py 2>&1 | tail -1
done > gpurun_out/ra_diag.log 2>&1
cat gpurun_out/ra_diag.log
