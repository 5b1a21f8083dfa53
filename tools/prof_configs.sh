#!/bin/bash
# `ncu --set full` captures of the solve kernel at configs 2, 4 and 5 (full batch) -> gpurun_out/prof_cfg<k>_$1.ncu-rep
tag=${1:-x}
mkdir -p gpurun_out
for spec in "2 1024" "4 8192" "5 2048"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_cfg$1_$tag python tools/quick_gpu.py $1 $2 0 > gpurun_out/ncu_cfg$1_$tag.log 2>&1
  echo "cfg$1 ncu exit $?"; tail -2 gpurun_out/ncu_cfg$1_$tag.log
done
