#!/bin/bash
# compute-sanitizer passes over one small invocation of every kernel (run on the GPU box):
#   memcheck + racecheck + initcheck + synccheck on __graft_entry__.smoke() (K1, scan, K2, K3, K4) and on a
#   small PDQ call (P1, P2, P3, P5).  Writes gpurun_out/sanitize_*.log; a clean run ends with "ERROR SUMMARY: 0 errors".
mkdir -p gpurun_out
cat > /tmp/san_pdq.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from bayes_od_rc_b200 import pdq
g = np.load('tests/golden/pdq_borders.npz')
e = pdq.PdqEngine(g['img_size'])
hm = e.heatmaps(g['boxes'], g['covs'])
assert np.abs(hm - g['heatmaps']).max() < 2e-7
e.losses([0, len(g['boxes'])], g['boxes'], g['covs'], [0, len(g['gt_boxes'])], g['gt_boxes'])
print('pdq ok')
PY
for tool in memcheck racecheck initcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_${tool}_smoke.log 2>&1
  echo "$tool smoke: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/sanitize_${tool}_smoke.log | tr '\n' ' ')"
  if [ "${SAN_ONLY:-}" != "pdq" ] && [ "$tool" != "initcheck" ]; then   # pipelined / graph-replay / held-input path (round 2)
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/san_stream.py > gpurun_out/sanitize_${tool}_stream.log 2>&1
    echo "$tool stream: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|stream ok' gpurun_out/sanitize_${tool}_stream.log | tr '\n' ' ')"
  fi
  [ "${SAN_SKIP_PDQ:-}" = "1" ] && continue
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_pdq.py > gpurun_out/sanitize_${tool}_pdq.log 2>&1
  echo "$tool pdq: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|pdq ok' gpurun_out/sanitize_${tool}_pdq.log | tr '\n' ' ')"
done
# optional: memcheck over the whole GPU test suite (slow: ~10 min).  bash scripts/sanitize_gpu.sh suite
if [ "${1:-}" = "suite" ]; then
  timeout 1500 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/sanitize_memcheck_suite.log 2>&1
  echo "memcheck suite: $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_memcheck_suite.log | tr '\n' ' ')"
fi
