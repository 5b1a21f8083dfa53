#!/bin/bash
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_corridor.py tests/test_gpu_frontend.py -m gpu -q -x > gpurun_out/pytest_gpu_r2w.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_r2w.log
for i in 1 2 3; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | tail -1 > gpurun_out/bench_r2w_$i.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r2w_$i.json')); print('value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], 'e2e_ms', d['e2e']['ms_per_step'], d['e2e']['last_step_breakdown_rank0'], d['e2e']['h2d_bytes_per_step'])"; done
CUDA_LAUNCH_BLOCKING=1 timeout 120 python -c "
import sys; sys.path.insert(0,'.')
from dftpav_b200 import synth
from dftpav_b200.optimizer import BatchOptimizer
b,prm=synth.make_config(3,B=700,seed=1); o=BatchOptimizer(prm,device=0); s=o.solve(b); print('launch-blocking solve ok', float(s.success.mean()))"
