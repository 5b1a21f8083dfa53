#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_corridor.py tests/test_gpu_minco_fit.py tests/test_gpu_sample.py -q -s 2>&1 | grep -E "resolution|passed|failed" 
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r2l.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['last_step_breakdown_rank0'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2l.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_r2l.log 2>&1; echo "ncu list exit $?"
CUDA_LAUNCH_BLOCKING=1 timeout 120 python -c "
import sys; sys.path.insert(0,'.')
from dftpav_b200 import synth
from dftpav_b200.optimizer import BatchOptimizer
b,p = synth.make_config(3,B=64,seed=1); o=BatchOptimizer(p,device=0); s=o.solve(b); print('CUDA_LAUNCH_BLOCKING=1 solve ok', s.success.mean())"
