#!/bin/bash
# staged (cp.async ring) two-loop: microbench, bit-identity vs the previous build, lone-problem phases, throughput A/B
timeout 60 gpurun_out_in/two_loop_bench 2>&1 | tee gpurun_out/two_loop_bench_r2p.txt
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/twoloop_ring_identity.txt
import glob, numpy as np, sys
sys.path.insert(0, '.')
from dftpav_b200 import synth
from dftpav_b200.optimizer import BatchOptimizer
for cfg, B in ((3, 512), (2, 256), (4, 256), (1, 64)):
    b, prm = synth.make_config(cfg, B=B, seed=5)
    ref = BatchOptimizer(prm, device=0, variant='gpurun_out_in/lib_old.so').solve(b)
    for f in ['']:
        s = BatchOptimizer(prm, device=0, variant=f).solve(b)
        print('cfg', cfg, f or 'product', 'identical' if (np.array_equal(s.cost, ref.cost) and np.array_equal(s.x, ref.x) and np.array_equal(s.iters, ref.iters)) else 'DIFFERENT', int((s.cost != ref.cost).sum()), 'iters max', int(s.iters.max()))
PY
timeout 120 python tools/lone_phases.py gpurun_out_in/fine_ring4.so 2>&1 | grep -v warning | tail -16 | tee gpurun_out/lone_phases_r2p.txt
bash tools/ab.sh 2>&1 | tee gpurun_out/ab_r2p.txt
