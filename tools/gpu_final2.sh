#!/bin/bash
tag=r2y
timeout 300 python tools/lib_ab.py 4 8192 gpurun_out_in/lib_old.so 2>&1 | grep -v Warn | tail -4 | tee gpurun_out/lib_ab_cfg4_$tag.txt
timeout -k 10 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | tail -1 > gpurun_out/bench_${tag}_$i.json; python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}_$i.json')); print('value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"; done
timeout -k 10 400 python tools/all_configs.py > gpurun_out/all_configs_$tag.log 2>&1; cp gpurun_out/all_configs_latest.json gpurun_out/all_configs_$tag.json; tail -4 gpurun_out/all_configs_$tag.log | cut -c1-120
