#!/bin/bash
mkdir -p gpurun_out
cp bayes_od_rc_b200/lib/libbayesod.so /tmp/lib_release.so
for v in tau097 tau098 tau099; do
  cp bayes_od_rc_b200/lib/variants/lib_$v.so bayes_od_rc_b200/lib/libbayesod.so
  echo "$v bdd   $(timeout 300 python scripts/diag_stages.py 2>&1 | tail -1 | cut -c1-200)"
  echo "$v kraw  $(DIAG_WORKLOAD=kitti_raw_b64_n20_k4 timeout 300 python scripts/diag_stages.py 2>&1 | tail -1 | cut -c1-200)"
done
cp /tmp/lib_release.so bayes_od_rc_b200/lib/libbayesod.so
