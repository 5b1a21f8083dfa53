"""Throughput of every BASELINE config at its full batch size on one GPU (device-resident batch)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftpav_b200 import synth
from dftpav_b200.optimizer import BatchOptimizer
out = {}
for cfg, B in ((2, 1024), (3, 4096), (4, 8192), (5, 16384 // 8)):
    batch, prm = synth.make_config(cfg, B=B, seed=0)
    opt = BatchOptimizer(prm, device=0)
    opt.upload(batch)
    ms = []
    for _ in range(4):
        opt.optimize(); ms.append(opt.stats().solve_ms)
    sol = opt.download(); st = opt.stats()
    best = float(np.median(ms[1:]))
    out[cfg] = dict(batch=B, solve_ms=best, solves_per_s=B / best * 1e3, evals_per_s=st.total_evals / best * 1e3,
                    iters_mean=float(sol.iters.mean()), evals_mean=float(sol.evals.mean()), success=float(sol.success.mean()),
                    alg_GBps=st.algorithmic_bytes / best / 1e6, status=dict(zip(*[a.tolist() for a in np.unique(sol.status, return_counts=True)])))
    print(cfg, json.dumps(out[cfg]))
    opt.close()
json.dump(out, open("gpurun_out/all_configs_latest.json", "w"), indent=1)
