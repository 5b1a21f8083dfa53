"""Host logic: batch container, generators (prefix-stable seeding), slicing, algorithmic-bytes bookkeeping."""
import numpy as np

from dftpav_b200 import capi, synth
from dftpav_b200.batch import batch_from_problems, samples_per_segment
from dftpav_b200.sharding import my_shard, shard_bounds


def test_sizes_match_baseline_table():
    # n, samples, algorithmic bytes per evaluation
    expect = {1: (15, 168, 21840), 2: (31, 304, 58960), 3: (47, 440, 57168), 4: (67, 640, 83280), 5: (47, 440, 82128)}
    for cfg, (n, M, bytes_eval) in expect.items():
        batch, prm = synth.make_config(cfg, B=2, seed=0)
        assert batch.n_of(0) == n
        assert batch.total_samples(prm.traj_resolution, prm.des_traj_resolution) == 2 * M
        assert synth.algorithmic_bytes_per_eval(batch, prm, 0) == bytes_eval


def test_generators_are_prefix_stable_and_seeded():
    a, _ = synth.make_config(3, B=4, seed=5)
    b, _ = synth.make_config(3, B=9, seed=5)
    c, _ = synth.make_config(3, B=4, seed=6)
    Q = a.planes.shape[0]
    np.testing.assert_array_equal(a.planes, b.planes[:Q])
    np.testing.assert_array_equal(a.inner_pts, b.inner_pts[: a.inner_pts.shape[0]])
    assert not np.array_equal(a.inner_pts, c.inner_pts)


def test_slice_is_self_contained(oracle):
    for cfg in (3, 4, 5):
        batch, prm = synth.make_config(cfg, B=5, seed=2)
        sub = batch.slice(2, 5, prm.traj_resolution, prm.des_traj_resolution)
        assert sub.B == 3
        for p in range(3):
            x = oracle.initial_x(prm, batch, 2 + p)
            np.testing.assert_array_equal(x, oracle.initial_x(prm, sub, p))
            f0 = oracle.evaluate(prm, batch, 2 + p, x)[0]
            assert f0 == oracle.evaluate(prm, sub, p, x)[0]


def test_batch_from_problems_roundtrip(oracle):
    prm = capi.default_params()
    batch, _ = synth.make_config(2, B=2, seed=3)
    K, Kd = prm.traj_resolution, prm.des_traj_resolution
    M = samples_per_segment(16, K, Kd)
    probs = []
    for p in range(2):
        hp = [batch.planes.reshape(2, M, 6, 4)[p, q].T for q in range(M)]
        probs.append(dict(iniStates=[batch.ini_state[p].reshape(3, 2).T], finStates=[batch.fin_state[p].reshape(3, 2).T],
                          initInnerPts=[batch.inner_pts.reshape(2, 15, 2)[p].T], initTs=[batch.init_T[p]], hPolys=[hp],
                          singuls=[1], now=0.0))
    b2 = batch_from_problems(probs)
    assert b2.uniform_H == 6
    for p in range(2):
        x = oracle.initial_x(prm, batch, p)
        assert oracle.evaluate(prm, batch, p, x)[0] == oracle.evaluate(prm, b2, p, x)[0]


def test_shard_bounds():
    b = shard_bounds(10, 4)
    assert list(b) == [0, 3, 6, 8, 10]
    assert my_shard(4096, 7, 8) == (3584, 4096)
