mkdir -p gpurun_out
timeout 300 python scripts/diag_stages.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:k3_softnms -s 6 -c 1 -o gpurun_out/prof_k3_r2c -f python scripts/diag_stages.py > gpurun_out/ncu_k3_r2c.log 2>&1
tail -2 gpurun_out/ncu_k3_r2c.log
