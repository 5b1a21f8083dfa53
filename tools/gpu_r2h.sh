#!/bin/bash
# round-2 GPU session H: config-5 profile of the current build, tight-delta experiment
mkdir -p gpurun_out
timeout 420 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/prof_cfg5_r2h python tools/quick_gpu.py 5 2048 0 > gpurun_out/ncu_cfg5_r2h.log 2>&1; echo "ncu cfg5 rc $?"
timeout 900 python tools/tight_delta_experiment.py 2>&1 | tee gpurun_out/tight_delta_r2h.txt | cut -c1-260
