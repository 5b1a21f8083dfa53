#!/bin/bash
# round-2 GPU session A: strict + parity tests, bench, A/B of the build variants in gpurun_out_in/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_strict.py -q -s -x > gpurun_out/strict_r2a.log 2>&1; echo "strict rc $?"
tail -5 gpurun_out/strict_r2a.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -s > gpurun_out/parity_r2a.log 2>&1; echo "parity rc $?"
tail -15 gpurun_out/parity_r2a.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc $?"
cut -c1-1200 gpurun_out/bench_r2a.json
timeout 900 bash tools/ab.sh > gpurun_out/ab_r2a.log 2>&1; cat gpurun_out/ab_r2a.log
timeout 300 python tools/small_batches.py > gpurun_out/small_r2a.log 2>&1; cat gpurun_out/small_r2a.log
