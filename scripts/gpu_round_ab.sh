#!/bin/bash
# soft-NMS: why a round's batch ends (counters of the diagnostics build), BDD and raw KITTI shapes
mkdir -p gpurun_out
cp bayes_od_rc_b200/lib/libbayesod.so /tmp/lib_release.so
cp bayes_od_rc_b200/lib/variants/lib_diag.so bayes_od_rc_b200/lib/libbayesod.so
BOD_K3_DEBUG=1 DIAG_WORKLOAD=kitti_raw_b64_n20_k4 timeout 300 python scripts/diag_stages.py 2>&1 | tail -16 > gpurun_out/rab_diag_kraw.log; head -5 gpurun_out/rab_diag_kraw.log | cut -c1-260
BOD_K3_DEBUG=1 timeout 300 python scripts/diag_stages.py 2>&1 | tail -16 > gpurun_out/rab_diag_bdd.log; head -5 gpurun_out/rab_diag_bdd.log | cut -c1-260
cp /tmp/lib_release.so bayes_od_rc_b200/lib/libbayesod.so
