"""N>1 path on CPU: two gloo ranks shard a batch, solve their slices (the oracle stands in for the GPU solver, injected
through the same ``solve_sharded`` plumbing the product uses), all-reduce the 1-int convergence flag and gather to
rank 0.  The gathered result must be bit-identical to the single-process solve (problems are independent)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from dftpav_b200 import synth
    from dftpav_b200.sharding import global_flag, solve_sharded
    from oracle import oracle_py

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch, prm = synth.make_config(3, B=7, seed=4)
    K, Kd = prm.traj_resolution, prm.des_traj_resolution
    sol, (lo, hi), ok, gathered = solve_sharded(batch, K, Kd, lambda sub: oracle_py.optimize(prm, sub, threads=1), gather=True)
    assert ok is True
    # a failing rank flips the global flag everywhere
    assert global_flag(rank != 1) is False
    # ... and the flag is derived from the solves, not a constant: one problem of rank 1's slice that OptimizeTrajectory
    # rejects (initTs < mini_T) makes the global flag false on BOTH ranks (the old test `status > -1024` was
    # true for every L-BFGS error code)
    bad, _ = synth.make_config(3, B=7, seed=4)
    bad.init_T[5] = 0.01
    _, _, ok_bad = solve_sharded(bad, K, Kd, lambda sub: oracle_py.optimize(prm, sub, threads=1))
    assert ok_bad is False
    if rank == 0:
        np.savez(out_path, lo=lo, hi=hi, **gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path, oracle):
    from dftpav_b200 import synth

    out = str(tmp_path / "gathered.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    batch, prm = synth.make_config(3, B=7, seed=4)
    ref = oracle.optimize(prm, batch, threads=1)
    np.testing.assert_array_equal(got["x"], ref.x)
    np.testing.assert_array_equal(got["cost"], ref.cost)
    np.testing.assert_array_equal(got["status"], ref.status)
    np.testing.assert_array_equal(got["iters"], ref.iters)
    assert (int(got["lo"]), int(got["hi"])) == (0, 4)
