"""Multi-GPU entry points of the C-ABI (dftpav_create_multi / dftpav_solve_multi) and the torchrun
sharding plumbing with the real solver.  Shards are contiguous slices solved concurrently by one host thread per device;
results must be BIT-IDENTICAL to the single-device solve of the whole batch; the only collective is the 1-int flag."""
import os
import sys

import numpy as np
import pytest
import torch  # noqa: F401  (before the library: it brings its own NCCL, which dftpav_create_multi then reuses)

from dftpav_b200 import capi, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_identical(a, b):
    for name in ("x", "cost", "status", "success", "iters", "evals", "coeffs", "piece_dT"):
        np.testing.assert_array_equal(getattr(a, name), getattr(b, name), err_msg=name)


@pytest.mark.parametrize("cfg", [3, 4, 5])
def test_shards_on_one_device_are_bit_identical(cfg, gpu_opt_factory):
    """Three handles on cuda:0 (duplicate device ids: the NCCL communicator cannot be formed, the flag falls back to the host
    reduction) exercise the sharding - pointer offsets of every per-segment / per-sample / obstacle array, strides, result
    placement - on a single-GPU box."""
    from dftpav_b200.optimizer import MultiGpuOptimizer

    batch, prm = synth.make_config(cfg, B=50, seed=13)
    ref = gpu_opt_factory(prm).solve(batch)
    multi = MultiGpuOptimizer(prm, devices=(0, 0, 0))
    assert multi.n_dev == 3 and not multi.uses_nccl()
    sol, ok = multi.solve(batch)
    _check_identical(sol, ref)
    assert ok == bool(np.all(ref.success != 0))
    # a rejected problem in the middle shard flips the flag and nothing else
    batch.init_T[int(batch.seg_ptr[20])] = 0.01
    sol2, ok2 = multi.solve(batch)
    assert ok2 is False and sol2.status[20] == capi.PROBLEM_REJECTED
    keep = np.arange(50) != 20
    np.testing.assert_array_equal(sol2.x[keep], ref.x[keep])
    multi.close()


def test_mixed_shapes_across_shards(gpu_opt_factory):
    """Shards with different shapes (different strides per shard) land correctly in the batch-wide arrays."""
    from dftpav_b200.batch import TrajBatch
    from dftpav_b200.optimizer import MultiGpuOptimizer

    prm = capi.default_params()
    b2, _ = synth.make_config(2, B=9, seed=5)
    b4, _ = synth.make_config(4, B=7, seed=5)
    b4.planes = np.concatenate([b4.planes.reshape(-1, 4, 4), np.zeros((b4.planes.shape[0] // 4, 2, 4))], axis=1).reshape(-1, 4)
    mixed = TrajBatch(seg_ptr=np.concatenate([b2.seg_ptr, b2.seg_ptr[-1] + b4.seg_ptr[1:]]),
                      piece_num=np.concatenate([b2.piece_num, b4.piece_num]), singul=np.concatenate([b2.singul, b4.singul]),
                      ini_state=np.concatenate([b2.ini_state, b4.ini_state]), fin_state=np.concatenate([b2.fin_state, b4.fin_state]),
                      init_T=np.concatenate([b2.init_T, b4.init_T]), inner_pts=np.concatenate([b2.inner_pts, b4.inner_pts]),
                      planes=np.concatenate([b2.planes, b4.planes]), t_now=np.zeros(16), uniform_H=6)
    ref = gpu_opt_factory(prm).solve(mixed)
    multi = MultiGpuOptimizer(prm, devices=(0, 0))
    sol, _ = multi.solve(mixed)
    _check_identical(sol, ref)
    multi.close()


def test_two_gpus_with_nccl(gpu_opt_factory):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from dftpav_b200.optimizer import MultiGpuOptimizer

    batch, prm = synth.make_config(3, B=600, seed=3)
    ref = gpu_opt_factory(prm).solve(batch)
    multi = MultiGpuOptimizer(prm, devices=(0, 1))
    assert multi.uses_nccl(), "libnccl.so.2 could not be loaded / ncclCommInitAll failed"
    sol, ok = multi.solve(batch)
    _check_identical(sol, ref)
    assert ok == bool(np.all(ref.success != 0))
    multi.close()


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from dftpav_b200 import synth as S
    from dftpav_b200.optimizer import BatchOptimizer
    from dftpav_b200.sharding import solve_sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    batch, prm = S.make_config(3, B=200, seed=4)
    opt = BatchOptimizer(prm, device=rank)
    sol, (lo, hi), ok, gathered = solve_sharded(batch, prm.traj_resolution, prm.des_traj_resolution, opt.solve, gather=True)
    if rank == 0:
        np.savez(out_path, ok=ok, **gathered)
    dist.barrier()
    opt.close()
    dist.destroy_process_group()


def test_torchrun_style_sharding_with_the_real_solver(tmp_path, gpu_opt_factory):
    """sharding.solve_sharded (the plumbing bench.py uses: one process per GPU, NCCL all-reduce of the flag) with the real
    BatchOptimizer on 2 GPUs: the gathered result equals the single-GPU solve bit for bit."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    out = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, 29600 + os.getpid() % 1000, out), nprocs=2, join=True)
    got = np.load(out)
    batch, prm = synth.make_config(3, B=200, seed=4)
    ref = gpu_opt_factory(prm).solve(batch)
    for name in ("x", "cost", "status", "iters", "evals"):
        np.testing.assert_array_equal(got[name], getattr(ref, name), err_msg=name)
