#!/bin/bash
# round-2 GPU session F: config-5 code-size fix (bit-identity + timing), queue-order experiment
mkdir -p gpurun_out
timeout 600 python tools/cfg5_compare.py gpurun_out_in/lib_before_cfg5.so 2>&1 | tee gpurun_out/cfg5_compare_r2f.txt
timeout 600 python tools/order_experiment.py 2>&1 | tee gpurun_out/order_r2f.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "cfg5 or 5 or helped" > gpurun_out/parity_r2f.log 2>&1; echo "parity rc $?"; tail -3 gpurun_out/parity_r2f.log
