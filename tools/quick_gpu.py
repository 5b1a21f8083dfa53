"""Quick GPU sanity + timing (development helper, run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dftpav_b200 import synth, capi
from dftpav_b200.optimizer import BatchOptimizer
from oracle import oracle_py as O

cfgs = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [3]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
check = int(sys.argv[3]) if len(sys.argv) > 3 else 16
for cfg in cfgs:
    batch, prm = synth.make_config(cfg, B=B, seed=0)
    opt = BatchOptimizer(prm, device=0)
    t = time.time(); opt.upload(batch); up = time.time() - t
    x0 = opt.initial_x(batch)
    f, g = opt.evaluate(x0)
    st = opt.stats()
    print(f"cfg{cfg} B={B} upload {up*1e3:.1f} ms (dev {st.upload_ms:.2f}) eval kernel {st.solve_ms:.3f} ms -> {B/st.solve_ms*1e3:.3e} evals/s")
    wf = wg = 0
    for p in range(min(check, B)):
        n = batch.n_of(p)
        fo, go, _, _ = O.evaluate(prm, batch, p, x0[p, :n])
        wf = max(wf, abs(f[p]-fo)/max(1, abs(fo))); wg = max(wg, np.abs(g[p,:n]-go).max()/max(1, np.abs(go).max()))
    print(f"   eval parity: f rel {wf:.2e} g rel {wg:.2e}")
    for rep in range(3):
        opt.optimize(); st = opt.stats()
        print(f"   solve kernel {st.solve_ms:.2f} ms -> {B/st.solve_ms*1e3:.3e} solves/s")
    sol = opt.download(); st = opt.stats()
    print(f"   iters mean {sol.iters.mean():.1f} evals mean {sol.evals.mean():.1f} evals/s {st.total_evals/st.solve_ms*1e3:.3e} alg GB/s {st.algorithmic_bytes/st.solve_ms/1e6:.1f} status {np.unique(sol.status, return_counts=True)}")
    if check:
        ref = O.optimize(prm, batch, 0, min(check, B), threads=0)
        c = min(check, B)
        rel = np.abs(sol.cost[:c]-ref.cost[:c])/np.maximum(1, np.abs(ref.cost[:c]))
        print(f"   solve parity over {c}: worst cost rel {rel.max():.2e}; iters equal {np.mean(sol.iters[:c]==ref.iters[:c]):.2f}; status equal {np.array_equal(sol.status[:c], ref.status[:c])}")
    opt.close()
