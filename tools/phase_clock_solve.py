"""Cycle split of the solve kernel: evaluation phases vs L-BFGS logic (diagnostic)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dftpav_b200 import synth, capi
from dftpav_b200.optimizer import BatchOptimizer
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
names = ["rhs", "banded solve", "coeffs+jerk", "sweep", "adj rhs", "adj solve", "assembly", "L-BFGS logic"]
batch, prm = synth.make_config(cfg, B=B, seed=0)
opt = BatchOptimizer(prm, device=0)
opt.upload(batch)
lib = capi.load_library()
lib.dftpav_debug_phase_clock.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong)]
opt.optimize()
lib.dftpav_debug_phase_clock(opt._h, 1, None)
opt.optimize()
st = opt.stats()
out = (C.c_ulonglong * 8)()
lib.dftpav_debug_phase_clock(opt._h, 0, out)
sol = opt.download()
ev = int(sol.evals.sum()) + B  # + ordering evaluation
tot = sum(out)
print(f"cfg{cfg} B={B}: solve {st.solve_ms:.2f} ms, {ev} evaluations; warp-cycles per evaluation: total {tot/ev:.0f}")
for n, v in zip(names, out):
    print(f"   {n:14s} {v/ev:9.0f}  {100*v/max(1,tot):5.1f}%")
