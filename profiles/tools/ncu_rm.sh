# usage: bash profiles/tools/ncu_rm.sh [kernel regex]   (default: the right-grouped mma.sync row kernels)
set -e
K=${1:-"rm_fwd_kernel|rm_bwd_kernel"}
python profiles/tools/quick_step.py 1024 > gpurun_out/plain_quick_rm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 2 -c 2 -o gpurun_out/r2c_rm -f python profiles/tools/quick_step.py 1024 > gpurun_out/ncu_quick_rm.log 2>&1
tail -2 gpurun_out/plain_quick_rm.log; tail -3 gpurun_out/ncu_quick_rm.log
