"""Does the order in which the persistent warps take the problems matter?  input order vs descending f(x0) vs the oracle
order (descending evaluation count of a previous solve = perfect longest-first) for configurations 3 and 4."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftpav_b200 import synth
from dftpav_b200.optimizer import BatchOptimizer
for cfg, B in ((3, 4096), (4, 8192), (2, 1024)):
    batch, prm = synth.make_config(cfg, B=B, seed=0)
    opt = BatchOptimizer(prm, device=0)
    opt.upload(batch)
    f0, _ = opt.evaluate(opt.initial_x(batch))
    def t():
        ms = []
        for _ in range(4):
            opt.optimize(); ms.append(opt.stats().solve_ms)
        return float(np.median(ms[1:]))
    base = t(); sol = opt.download()
    work = sol.evals + 0.01 * sol.iters.astype(float) ** 2
    opt.set_order(np.argsort(-f0)); lpt_f0 = t()
    opt.set_order(np.argsort(-work)); lpt_true = t()
    opt.set_order(np.argsort(work)); spt = t()
    opt.set_order(None)
    print(f"cfg{cfg} B={B}: input order {base:.2f} ms, by f(x0) {lpt_f0:.2f} ms, perfect longest-first {lpt_true:.2f} ms, shortest-first {spt:.2f} ms")
    opt.close()
