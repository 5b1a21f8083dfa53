"""Per-phase cycle breakdown of one evaluation (diagnostic)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftpav_b200 import synth, capi
from dftpav_b200.optimizer import BatchOptimizer
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
names = ["rhs", "banded solve", "coeffs+jerk", "sweep", "adj rhs", "adj solve", "assembly", "-"]
for B in [int(a) for a in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["148", "2368"])]:
    batch, prm = synth.make_config(cfg, B=B, seed=0)
    opt = BatchOptimizer(prm, device=0)
    opt.upload(batch)
    x0 = opt.initial_x(batch)
    lib = capi.load_library()
    lib.dftpav_debug_phase_clock.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong)]
    opt.evaluate(x0)
    lib.dftpav_debug_phase_clock(opt._h, 1, None)
    opt.evaluate(x0)
    out = (C.c_ulonglong * 8)()
    lib.dftpav_debug_phase_clock(opt._h, 0, out)
    tot = sum(out)
    print(f"cfg{cfg} B={B}: cycles per evaluation per warp (avg over {B} warps): total {tot/B:.0f}")
    for n, v in zip(names, out):
        print(f"   {n:14s} {v/B:9.0f}  {100*v/max(1,tot):5.1f}%")
    opt.close()
