#!/bin/bash
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2u.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r2u.log
for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | tail -1 > gpurun_out/bench_r2u_$i.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r2u_$i.json')); print('value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['last_step_breakdown_rank0'])"; done
bash tools/sanitize.sh r2
