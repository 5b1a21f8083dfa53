"""Collision re-check of solved trajectories on the GPU (dftpav_collision_check, SURVEY.md section 8f row N2: the consumer
loop of traj_server_ros.cpp:385-397) against the CPU restatement (oracle/corridor_oracle.cpp: first colliding sample of
every gear segment).  Index-for-index equal; the only libm calls on the path are atan2 / cos / sin of the heading, whose CUDA
and glibc values can differ by an ulp and move a footprint vertex by ~1e-15 m - never across a cell border in practice."""
import numpy as np
import pytest

from dftpav_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _segments(batch, sol):
    """(problem, coefficient rows [6N, 2], piece duration, singul) of every gear segment, in segment order."""
    out = []
    for p in range(batch.B):
        row = 0
        for g in range(int(batch.seg_ptr[p]), int(batch.seg_ptr[p + 1])):
            N = int(batch.piece_num[g])
            out.append((p, sol.coeffs[p, row: row + 6 * N], float(sol.piece_dT[g]), int(batch.singul[g])))
            row += 6 * N
    return out


@pytest.mark.parametrize("cfg,B", [(3, 48), (4, 12)])
def test_first_colliding_sample_matches_the_oracle(cfg, B, oracle, gpu_opt_factory):
    batch, prm = synth.make_config(cfg, B=B, seed=4)
    opt = gpu_opt_factory(prm)
    sol = opt.solve(batch)
    segs = _segments(batch, sol)
    # a map around the trajectories: free, some clutter away from the paths, and single occupied cells planted ON every eighth trajectory
    pts = np.concatenate([batch.ini_state[:, 0:2], batch.fin_state[:, 0:2]])
    res = 0.2
    lo = pts.min(axis=0) - 30.0
    size = np.ceil((pts.max(axis=0) + 30.0 - lo) / res).astype(int) + 1
    grid = np.full((size[1], size[0]), 127, dtype=np.uint8)
    planted = set()
    for k, (p, c, dT, sg) in enumerate(segs):
        if p % 8 == 0:
            N = c.shape[0] // 6
            t = 0.97 * N * dT  # near the segment's end: the synthetic problems share their start, not their goals
            st, _ = oracle.traj_state(c, dT, sg, t)
            cx, cy = np.round((st[0:2] - lo) / res).astype(int)
            grid[cy, cx] = capi.OCCUPIED
            planted.add(p)
    opt.set_map(grid, lo, res)
    col, first = opt.collision_check(dt=0.05)
    ms = opt.stats().solve_ms
    want = np.array([oracle.collision_check_segment(grid, lo, res, c, dT, sg) for (_, c, dT, sg) in segs])
    np.testing.assert_array_equal(first, want)
    per_problem = np.zeros(batch.B, dtype=np.int32)
    for (p, _, _, _), f in zip(segs, want):
        per_problem[p] |= int(f >= 0)
    np.testing.assert_array_equal(col, per_problem)
    assert all(col[p] == 1 for p in planted) and (want >= 0).any() and (want == -1).any()
    print(f"config {cfg}: {len(segs)} segments re-checked in {ms:.3f} ms on the device, {int(col.sum())} of {batch.B} problems collide")
    # an empty map: nothing collides; a cap that is too small is refused
    opt.set_map(np.full_like(grid, 127), lo, res)
    col0, first0 = opt.collision_check()
    assert not col0.any() and np.all(first0 == -1)
    with pytest.raises(Exception):
        opt.collision_check(max_samples=10)
