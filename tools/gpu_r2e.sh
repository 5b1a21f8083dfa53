#!/bin/bash
# round-2 GPU session E: corridor + multi tests, all configs, config-5 profile, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_corridor.py tests/test_gpu_multi.py tests/test_gpu_sample.py -q -s > gpurun_out/tests_r2e.log 2>&1; echo "tests rc $?"; grep -E "resolution|passed|failed|Error" gpurun_out/tests_r2e.log | head
timeout 300 python tools/all_configs.py > gpurun_out/all_configs_r2e.log 2>&1; tail -5 gpurun_out/all_configs_r2e.log | cut -c1-220
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc $?"; cut -c1-400 gpurun_out/bench_r2e.json
timeout 420 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/prof_cfg5_r2e python tools/quick_gpu.py 5 2048 0 > gpurun_out/ncu_cfg5_r2e.log 2>&1; echo "ncu cfg5 rc $?"
timeout 420 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/prof_cfg4_r2e python tools/quick_gpu.py 4 8192 0 > gpurun_out/ncu_cfg4_r2e.log 2>&1; echo "ncu cfg4 rc $?"
