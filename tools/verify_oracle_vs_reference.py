"""Wide randomized check (run in the build container, where oracle/_ref exists): the oracle against the reference's own
optimizer source compiled in place, bit for bit, on many more problems than the unit tests use."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftpav_b200 import synth
from oracle import oracle_py as O

assert O.ref_optimizer_lib() is not None
t0 = time.time()
tot_eval = tot_solve = bad = 0
for cfg in (1, 2, 3, 4, 5):
    for seed in (101, 202):
        B = 40
        batch, prm = synth.make_config(cfg, B=B, seed=seed)
        rng = np.random.default_rng(seed)
        for p in range(B):
            x0 = O.initial_x(prm, batch, p)
            for sc in (0.0, 0.01, 0.1, 0.5):
                x = x0 + sc * rng.normal(size=x0.shape)
                fr, gr, cr, dr = O.ref_evaluate(prm, batch, p, x)
                fo, go, co, do = O.evaluate(prm, batch, p, x)
                ok = (fr == fo) and np.array_equal(gr, go) and np.array_equal(cr, co) and np.array_equal(dr, do)
                tot_eval += 1
                bad += 0 if ok else 1
        r = O.ref_solve_batch(prm, batch, 0, B)
        s = O.optimize(prm, batch, 0, B)
        same = np.array_equal(r["cost"], s.cost[:B]) and np.array_equal(r["iters"], s.iters[:B]) and \
            np.array_equal(r["evals"], s.evals[:B]) and np.array_equal(r["status"], s.status[:B]) and \
            np.array_equal(r["success"], s.success[:B])
        tot_solve += B
        bad += 0 if same else 1
        print(f"cfg {cfg} seed {seed}: {4*B} evaluations and {B} complete solves bit-identical: {same}", flush=True)
print(f"{tot_eval} evaluations, {tot_solve} complete OptimizeTrajectory runs, mismatches: {bad}, {time.time()-t0:.0f} s")
