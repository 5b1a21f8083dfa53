#!/bin/bash
# round-2 GPU session D: cold code out of line (A/B), warps per CTA, multi-GPU C-ABI tests on one device
mkdir -p gpurun_out
timeout 900 bash tools/ab.sh > gpurun_out/ab_r2d.log 2>&1; cat gpurun_out/ab_r2d.log
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -q -x > gpurun_out/multi_r2d.log 2>&1; echo "multi+parity rc $?"; tail -5 gpurun_out/multi_r2d.log
timeout 300 python tools/all_configs.py > gpurun_out/all_configs_r2d.log 2>&1; tail -8 gpurun_out/all_configs_r2d.log
