"""Multi-GPU plumbing: the batch shards embarrassingly (problems are independent).

One process per GPU (torchrun); rank r owns the contiguous slice [lo, hi) of the batch.  The only collective on the
path is the all-reduce (MIN) of a single int: "every problem of my slice finished with a usable status" — the global
convergence flag of BASELINE.json's north_star.  Results are gathered to rank 0 only when asked (not on the hot path).
The solve itself is injected (``solve_fn``): the product passes ``BatchOptimizer.solve``; CPU tests (gloo) pass the
oracle, so the same plumbing is exercised without a GPU.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_bounds(B: int, world: int) -> np.ndarray:
    """Contiguous, balanced partition: returns world+1 offsets."""
    base, rem = divmod(B, world)
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def my_shard(B: int, rank: int, world: int) -> Tuple[int, int]:
    b = shard_bounds(B, world)
    return int(b[rank]), int(b[rank + 1])


def global_flag(local_ok: bool, device: Optional[str] = None) -> bool:
    """All-reduce(MIN) of the 1-int convergence flag over the default process group (NCCL on GPUs, gloo in CPU tests)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return bool(local_ok)
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([1 if local_ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t.item()))


def solve_sharded(batch, K: int, Kd: int, solve_fn: Callable, gather: bool = False):
    """Solves this rank's slice with ``solve_fn(sub_batch) -> solution`` (attributes x, cost, status, success, iters,
    evals); returns (local_solution, (lo, hi), all_ok[, gathered dict on rank 0])."""
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = my_shard(batch.B, rank, world)
    sub = batch.slice(lo, hi, K, Kd)
    sol = solve_fn(sub)
    # the local half of the convergence flag: every problem of my slice finished AND was accepted (flag_success of
    # traj_optimizer.cpp:176-201: an accepted L-BFGS return code and a final cost below the limit)
    ok = global_flag(bool(sol.status.shape[0] == hi - lo and np.all(np.asarray(sol.success) != 0)))
    if not gather:
        return sol, (lo, hi), ok
    payload = {k: np.asarray(getattr(sol, k)) for k in ("x", "cost", "status", "success", "iters", "evals")}
    gathered = [None] * world if rank == 0 else None
    if world > 1:
        dist.gather_object(payload, gathered, dst=0)
    else:
        gathered = [payload]
    out = None
    if rank == 0:
        out = {k: np.concatenate([g[k] for g in gathered], axis=0) for k in payload}
    return sol, (lo, hi), ok, out
