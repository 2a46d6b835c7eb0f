set -e
python profiles/tools/quick_step.py 0 > gpurun_out/plain_quick.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"r_bwd_kernel|r_fwd_kernel|r_cores_kernel" -s 6 -c 3 -o gpurun_out/r2b_rows python profiles/tools/quick_step.py 0 > gpurun_out/ncu_quick.log 2>&1
tail -3 gpurun_out/plain_quick.log; tail -5 gpurun_out/ncu_quick.log
