"""Queue key A/B on one config (resident batch): plane-free f(x0) vs the whole f(x0).  usage: cfg_ab.py <cfg> <B>"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftpav_b200 import synth, capi
from dftpav_b200.optimizer import BatchOptimizer
cfg, B = int(sys.argv[1]), int(sys.argv[2])
batch, prm = synth.make_config(cfg, B=B, seed=0)
opt = BatchOptimizer(prm, device=0)
lib = capi.load_library()
lib.dftpav_debug_set_order_key.argtypes = [C.c_void_p, C.c_int]
opt.upload(batch)
for rep in range(2):
    for mode, name in ((1, "plane-free key"), (2, "full f(x0) key")):
        lib.dftpav_debug_set_order_key(opt._h, mode)
        ms = []
        for _ in range(4):
            opt.optimize(); ms.append(opt.stats().solve_ms)
        print(cfg, name, "solve_ms", ["%.2f" % m for m in ms], "median %.2f" % float(np.median(ms[1:])), flush=True)
opt.set_order(np.arange(B))
ms = []
for _ in range(3):
    opt.optimize(); ms.append(opt.stats().solve_ms)
print(cfg, "input order", "solve_ms", ["%.2f" % m for m in ms], flush=True)
