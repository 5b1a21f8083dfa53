"""A/B of the product against an OLDER library build on one config (resident batch).  usage: lib_ab.py <cfg> <B> <old.so>"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


class _Tolerant(ctypes.CDLL):  # an older build lacks the newest entry points: give the loader something to set argtypes on
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name.startswith("dftpav_"):
                return type("Missing", (), {})()
            raise


ctypes.CDLL = _Tolerant
from dftpav_b200 import synth, capi
from dftpav_b200.optimizer import BatchOptimizer
cfg, B, old = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
batch, prm = synth.make_config(cfg, B=B, seed=0)
for rep in range(2):
    for name, variant in (("product", ""), ("other", old)):
        opt = BatchOptimizer(prm, device=0, variant=variant)
        opt.upload(batch)
        ms = []
        for _ in range(4):
            opt.optimize(); ms.append(opt.stats().solve_ms)
        print(cfg, name, "solve_ms", ["%.2f" % m for m in ms], "median %.2f" % float(np.median(ms[1:])), flush=True)
        opt.close()
