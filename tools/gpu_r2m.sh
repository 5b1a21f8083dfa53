#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/lone_phases.py 2>&1 | tee gpurun_out/lone_phases_r2m.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2m.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_r2m.json')); print('value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
