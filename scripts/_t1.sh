mkdir -p gpurun_out
timeout 300 python scripts/diag_stages.py 2>&1 | tail -1
DIAG_K=8 timeout 300 python scripts/diag_stages.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "synthetic_batch or softnms or golden or prefilter or full" 2>&1 | tail -3
