#!/bin/bash
# round-2 GPU session G: config-5 pair compaction (bit-identity + timing), device-side queue order, full GPU test-suite
mkdir -p gpurun_out
timeout 600 python tools/cfg5_compare.py gpurun_out_in/lib_before_cfg5.so 2>&1 | tee gpurun_out/cfg5_compare_r2g.txt
timeout 300 python tools/all_configs.py > gpurun_out/all_configs_r2g.log 2>&1; tail -5 gpurun_out/all_configs_r2g.log | cut -c1-200
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc $?"; cut -c1-300 gpurun_out/bench_r2g.json
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2g.log 2>&1; echo "pytest rc $?"; tail -6 gpurun_out/pytest_gpu_r2g.log
