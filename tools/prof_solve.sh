#!/bin/bash
# One `ncu --set full` capture of the solve kernel (config 3, batch 4096) -> gpurun_out/prof_solve_$1.ncu-rep
tag=${1:-x}
mkdir -p gpurun_out
timeout 420 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f \
  -o gpurun_out/prof_solve_$tag python tools/quick_gpu.py 3 4096 0 > gpurun_out/ncu_solve_$tag.log 2>&1
echo "ncu exit $?"; tail -4 gpurun_out/ncu_solve_$tag.log
