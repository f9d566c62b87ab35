#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streaming or pipelined or launch_clock" > gpurun_out/rz_tests.log 2>&1; tail -60 gpurun_out/rz_tests.log | cut -c1-220
