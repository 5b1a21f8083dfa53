"""One long problem alone on the GPU (latency regime): the problem with the most evaluations of a 148-problem batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftpav_b200 import synth
from dftpav_b200.optimizer import BatchOptimizer
big, prm = synth.make_config(3, B=148, seed=0)
opt = BatchOptimizer(prm, device=0)
sol = opt.solve(big)
p = int(np.argmax(sol.evals))
one = big.slice(p, p + 1, prm.traj_resolution, prm.des_traj_resolution)
opt.upload(one)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    opt.optimize()
    print(f"problem {p}: evals {sol.evals[p]} iters {sol.iters[p]}: {opt.stats().solve_ms:.3f} ms")
