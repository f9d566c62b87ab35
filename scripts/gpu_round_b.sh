#!/bin/bash
# diagnostic build (BOD_EXTRA_NVCC_K3_SOFTNMS=-DBOD_DIAGNOSTICS BOD_EXTRA_NVCC_BOD_API=-DBOD_DIAGNOSTICS, --force): soft-NMS phase counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "synthetic_batch or softnms or golden or prefilter or full" > gpurun_out/rb_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/rb_tests.log
tail -3 gpurun_out/rb_tests.log
for nt in ${K3_NTS:-512 1024 256}; do
  BOD_K3_DEBUG=1 BOD_K3_THREADS=$nt timeout 300 python scripts/diag_stages.py 2>&1 | tail -16
done > gpurun_out/rb_diag.log 2>&1
cat gpurun_out/rb_diag.log
