"""Kernel time of small batches (latency regime): which problem sets the time, and what the helpers buy."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftpav_b200 import synth
from dftpav_b200.optimizer import BatchOptimizer
big, prm = synth.make_config(3, B=148, seed=0)
opt = BatchOptimizer(prm, device=0)
sol = opt.solve(big)
order = np.argsort(-sol.evals)
print("B=148 evals max", sol.evals.max(), "iters of it", sol.iters[order[0]])
def t(batch):
    opt.upload(batch); ms = []
    for _ in range(3): opt.optimize(); ms.append(opt.stats().solve_ms)
    return min(ms)
print("full 148: %.2f ms" % t(big))
for p in (int(order[0]), int(order[1]), int(order[74]), 0):
    one = big.slice(p, p + 1, prm.traj_resolution, prm.des_traj_resolution)
    print(f"problem {p} alone: evals {sol.evals[p]} iters {sol.iters[p]}: %.2f ms" % t(one))
for B in (2, 8, 37):
    print(f"first {B}: %.2f ms" % t(big.slice(0, B, prm.traj_resolution, prm.des_traj_resolution)))
