mkdir -p gpurun_out
BOD_K3_DEBUG=1 timeout 300 python scripts/diag_stages.py 2>&1 | tail -16 > gpurun_out/rc_diag.log 2>&1
cat gpurun_out/rc_diag.log
