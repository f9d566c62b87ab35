#!/bin/bash
# parity subset + one ncu --set full capture of the soft-NMS kernel (source-level sampling)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "synthetic_batch or softnms or golden or prefilter or full" > gpurun_out/rc_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/rc_tests.log
tail -3 gpurun_out/rc_tests.log
ncu --set full --clock-control none --import-source on -k regex:"k3_softnms" -s 3 -c 1 \
    -o gpurun_out/prof_k3_r2b -f python scripts/diag_stages.py > gpurun_out/ncu_k3_r2b.log 2>&1
tail -2 gpurun_out/ncu_k3_r2b.log
