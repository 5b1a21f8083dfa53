"""config 5 (moving obstacles): results and timing of the in-tree library against another build (bit-identity expected when
only code layout changed).  usage: cfg5_compare.py <other_lib.so>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftpav_b200 import synth
from dftpav_b200.optimizer import BatchOptimizer
batch, prm = synth.make_config(5, B=2048, seed=0)
res = {}
for name, variant in (("in-tree", ""), ("other", sys.argv[1])):
    opt = BatchOptimizer(prm, device=0, variant=variant)
    opt.upload(batch)
    x0 = opt.initial_x(batch)
    f, g = opt.evaluate(x0); ev_ms = opt.stats().solve_ms
    f, g = opt.evaluate(x0); ev_ms = min(ev_ms, opt.stats().solve_ms)
    ms = []
    for _ in range(3):
        opt.optimize(); ms.append(opt.stats().solve_ms)
    sol = opt.download()
    res[name] = (f, g, sol)
    print(f"{name:8s}: eval kernel {ev_ms:.2f} ms ({2048/ev_ms*1e3:.3e} evals/s), solve {min(ms):.1f} ms ({2048/min(ms)*1e3:.0f} solves/s, "
          f"{opt.stats().total_evals/min(ms)*1e3:.3e} evals/s)")
    opt.close()
a, b = res["in-tree"], res["other"]
print("eval f identical:", np.array_equal(a[0], b[0]), " g identical:", np.array_equal(a[1], b[1]),
      " solve x identical:", np.array_equal(a[2].x, b[2].x), " cost identical:", np.array_equal(a[2].cost, b[2].cost))
