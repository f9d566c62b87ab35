#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/rr_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/rr_tests.log; tail -3 gpurun_out/rr_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
