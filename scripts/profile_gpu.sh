#!/bin/bash
# Runs ON THE GPU BOX (gpurun): the two ncu passes of /opt/skills/guides/B200_PROFILING.md for bench.py.
#   (1) launch list: every kernel of a short pipelined bench run with its duration (cold-cache, serialised)
#   (2) one `--set full` capture of each stage kernel of one step (one step at a time, so that the kernels are
#       the ones of a single run), with source correlation
# Outputs go to gpurun_out/ (scratch); scripts/summarize_ncu.py turns them into profiles/*.
set -u
R=${1:-r1}
ONLY=${2:-all}
mkdir -p gpurun_out
if [ "$ONLY" = "pdq" ] || [ "$ONLY" = "all" ]; then
  # PDQ kernels (csrc/kp_pdq.cu): one --set full capture of the first launches of scripts/pdq_bench.py
  PDQ_REPS=1 ncu --set full --import-source on --clock-control none -k regex:pdq_ -c 8 -o gpurun_out/prof_pdq_${R} -f \
      python scripts/pdq_bench.py 16 60 20 --no-cpu > gpurun_out/ncu_pdq_${R}.log 2>&1
  tail -1 gpurun_out/ncu_pdq_${R}.log
  [ "$ONLY" = "pdq" ] && exit 0
fi
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k1_moments|k3_softnms|k4_fusion|k2_posterior|scan_tiles" \
    -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_${R}.log 2>&1
[ "$ONLY" = "launches" ] && exit 0
ncu --set full --clock-control none --import-source on \
    -k regex:"k1_moments|k3_softnms|k4_fusion|k2_posterior|scan_tiles" -s 25 -c 5 \
    -o gpurun_out/prof_all_${R} -f python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --pipeline 1 \
    > gpurun_out/ncu_full_${R}.log 2>&1
tail -2 gpurun_out/ncu_full_${R}.log
