#!/bin/bash
# arrival-aware longest-first queue with the plane-free key: identity vs the previous build, throughput, e2e
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/order_identity_r2q.txt
import numpy as np, sys
sys.path.insert(0, '.')
from dftpav_b200 import synth
from dftpav_b200.optimizer import BatchOptimizer
for cfg, B in ((3, 1024), (2, 1024), (5, 768)):
    b, prm = synth.make_config(cfg, B=B, seed=5)
    ref = BatchOptimizer(prm, device=0, variant='gpurun_out_in/lib_old.so')
    new = BatchOptimizer(prm, device=0)
    r = ref.solve(b); s = new.solve(b)
    new.upload(b); s2 = new.optimize() if hasattr(new, 'optimize') else None
    same = lambda a, c: np.array_equal(a.cost, c.cost) and np.array_equal(a.x, c.x) and np.array_equal(a.iters, c.iters)
    print('cfg', cfg, 'solve identical' if same(r, s) else 'solve DIFFERENT', int((s.cost != r.cost).sum()))
PY
for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | tail -1 > gpurun_out/bench_r2q_$i.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r2q_$i.json')); print('value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['last_step_breakdown_rank0'])"; done
bash tools/ab.sh 2>&1 | tee gpurun_out/ab_r2q.txt
