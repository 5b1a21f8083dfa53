"""Small workload that touches every kernel of the library once (run under compute-sanitizer by tools/sanitize.sh)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftpav_b200 import synth, capi
from dftpav_b200.optimizer import BatchOptimizer

lib = capi.load_library()
for cfg, B in ((3, 40), (5, 14), (4, 6), (2, 20)):
    batch, prm = synth.make_config(cfg, B=B, seed=1)
    opt = BatchOptimizer(prm, device=0)
    sol = opt.solve(batch)                 # overlapped upload, helpers from the first evaluation (B < warps)
    again = opt.solve_stepwise(batch)
    assert np.array_equal(sol.x, again.x)
    f, g = opt.evaluate(opt.initial_x(batch))
    if cfg == 3:
        t = np.linspace(0, 3.0, 50)
        opt.sample(np.zeros(50, dtype=np.int32), t)
        opt.sample_uniform(0.05, 700)
        opt.pack_polytraj()
        grid, origin, res, states = synth.make_obstacle_map(batch, prm, resolution=0.2, clutter=40, seed=2)
        opt.set_map(grid, origin, res)
        opt.corridor_rectangles(states[:300])
        opt.collision_check()
        fe = synth.make_frontend_paths(12, seed=3, trapezoid_duration=lib.dftpav_trapezoid_duration)
        opt.frontend_resample(fe, fe.shot_time.copy(), 0.5, 8, 16)
        opt.frontend_evaluate_pos(fe, np.zeros(10, dtype=np.int32), np.linspace(0, 5, 10))
    print("config", cfg, "ok: success", float(sol.success.mean()), flush=True)
    opt.close()
# a batch big enough for the longest-first queue (B > 4 x SMs)
batch, prm = synth.make_config(3, B=640, seed=2)
opt = BatchOptimizer(prm, device=0)
sol = opt.solve(batch)
print("queue path ok", float(sol.success.mean()))
