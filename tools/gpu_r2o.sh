#!/bin/bash
# two-loop prefetch distance: bit-identity vs the main build, lone-problem phase split, batch throughput
python - <<'PY' 2>&1 | tee gpurun_out/twoloop_ahead_identity.txt
import glob, numpy as np, sys
sys.path.insert(0, '.')
from dftpav_b200 import synth
from dftpav_b200.optimizer import BatchOptimizer
b, prm = synth.make_config(3, B=512, seed=5)
ref = BatchOptimizer(prm, device=0).solve(b)
for f in sorted(glob.glob('gpurun_out_in/lib_*.so')):
    s = BatchOptimizer(prm, device=0, variant=f).solve(b)
    print(f, 'identical' if (np.array_equal(s.cost, ref.cost) and np.array_equal(s.x, ref.x) and np.array_equal(s.iters, ref.iters)) else 'DIFFERENT', int((s.cost != ref.cost).sum()))
PY
for f in gpurun_out_in/fine_*.so; do echo "== $f"; python tools/lone_phases.py $f 2>&1 | grep -v warning | tail -8; done | tee gpurun_out/lone_phases_r2o.txt
bash tools/ab.sh 2>&1 | tee gpurun_out/ab_r2o.txt
