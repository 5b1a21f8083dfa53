#!/bin/bash
# round-2 GPU session I: exact closed-form piece location (config 5), N3 fits, full test-suite, all configs
mkdir -p gpurun_out
timeout 600 python tools/cfg5_compare.py gpurun_out_in/lib_prev.so 2>&1 | tee gpurun_out/cfg5_compare_r2i.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2i.log 2>&1; echo "pytest rc $?"; tail -6 gpurun_out/pytest_gpu_r2i.log
timeout 300 python tools/all_configs.py > gpurun_out/all_configs_r2i.log 2>&1; tail -5 gpurun_out/all_configs_r2i.log | cut -c1-200
