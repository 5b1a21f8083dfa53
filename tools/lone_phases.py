"""Where does the time of ONE long problem go (the critical path that sets the batch time)?  Cycle split of the owner warp."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftpav_b200 import synth, capi
from dftpav_b200.optimizer import BatchOptimizer
FINE = len(sys.argv) > 1   # argv[1] = path of a -DDFTPAV_PHASE_FINE=1 build
names = ["rhs", "minco product", "coeffs+jerk", "sweep (incl. team sync)", "adj rhs", "adj product", "assembly", "L-BFGS logic + two-loop"]
big, prm = synth.make_config(3, B=148, seed=0)
opt = BatchOptimizer(prm, device=0, variant=sys.argv[1] if FINE else "")
sol = opt.solve(big)
p = int(np.argmax(sol.evals))
one = big.slice(p, p + 1, prm.traj_resolution, prm.des_traj_resolution)
lib = capi.load_library(sys.argv[1] if FINE else "")
lib.dftpav_debug_phase_clock.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong)]
opt.upload(one)
opt.optimize()
lib.dftpav_debug_phase_clock(opt._h, 1, None)
opt.optimize()
ms = opt.stats().solve_ms
out = (C.c_ulonglong * 16)()
lib.dftpav_debug_phase_clock(opt._h, 2, out)
fine = list(out)[8:]
out = (C.c_ulonglong * 8)()
lib.dftpav_debug_phase_clock(opt._h, 0, out)
tot = sum(out)
print(f"problem {p}: evals {sol.evals[p]} iters {sol.iters[p]}: kernel {ms:.3f} ms = {ms*1.965e6:.0f} cycles at 1965 MHz; instrumented owner cycles {tot}")
for n, v in zip(names, out):
    print(f"   {n:26s} {v:10d}  {100*v/max(1,tot):5.1f}%   {v/1.965e6:.3f} ms")
if FINE:
    for n, v in zip(["line-search trial logic", "next trial point", "end-of-iteration tests", "history update", "two-loop recursion",
                     "line-search head"], fine):
        print(f"      {n:26s} {v:10d}  {100*v/max(1,tot):5.1f}%   {v/1.965e6:.3f} ms")
