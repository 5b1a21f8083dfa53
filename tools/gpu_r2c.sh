#!/bin/bash
# round-2 GPU session C: Gram-form two-loop - microbench, parity tests, A/B, latency
mkdir -p gpurun_out
timeout 300 ./gpurun_out_in/two_loop_bench > gpurun_out/two_loop_bench_r2c.txt 2>&1; cat gpurun_out/two_loop_bench_r2c.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -s -x > gpurun_out/parity_r2c.log 2>&1; echo "parity rc $?"; grep -E "^cfg|passed|failed|Error" gpurun_out/parity_r2c.log | head -30
timeout 600 python -m pytest tests/test_gpu_strict.py tests/test_gpu_sample.py -q -x > gpurun_out/strict_r2c.log 2>&1; echo "strict rc $?"; tail -3 gpurun_out/strict_r2c.log
timeout 900 bash tools/ab.sh > gpurun_out/ab_r2c.log 2>&1; cat gpurun_out/ab_r2c.log
timeout 200 python tools/small_batches.py > gpurun_out/small_r2c.log 2>&1; cat gpurun_out/small_r2c.log
cp gpurun_out_in/lib_vr8.so /tmp/x.so; cp dftpav_b200/csrc/libdftpav_b200.so /tmp/main.so; cp /tmp/x.so dftpav_b200/csrc/libdftpav_b200.so
echo "--- VR8 latency"; timeout 200 python tools/small_batches.py 2>&1 | tail -8
cp /tmp/main.so dftpav_b200/csrc/libdftpav_b200.so
