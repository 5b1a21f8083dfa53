"""The oracle against the reference's OWN optimizer: `oracle/_ref/libref_optimizer.so` is src/traj_optimizer.cpp (with
plan_manage/traj_optimizer.h, poly_traj_utils.hpp, traj_container.hpp, lbfgs.hpp) compiled where it lies, unmodified,
against a test-only Eigen-compatible layer and inert stubs of its ROS / protobuf / RViz dependencies (oracle/Makefile,
oracle/ref_optimizer_driver.cpp).  Problems are handed to the reference's public API (setParam, setSurroundTrajs,
OptimizeTrajectory) and read back through getMinJerkOptPtr().

  * costFunctionCallback: cost, gradient and coefficients of the oracle are BIT-IDENTICAL to the reference's at the
    initial point and at perturbed points of every BASELINE config (corridor, feasibility, gear coupling, moving
    obstacles);
  * complete solves: final x, cost, L-BFGS return code, success flag, iteration and evaluation counts BIT-IDENTICAL.

Live when the library is present (this container and, prebuilt, the GPU box) and always against
tests/golden/optimizer_reference.npz recorded from it by tests/make_golden.py."""
import os

import numpy as np
import pytest

import make_golden as MG  # tests/make_golden.py (same directory)

GOLD = os.path.join(os.path.dirname(__file__), "golden", "optimizer_reference.npz")


def test_oracle_matches_recorded_reference_optimizer(oracle):
    g = np.load(GOLD)
    for cfg, p in MG.OPT_CASES:
        batch, prm, xs = MG.optimizer_inputs(cfg, p)
        for k, x in enumerate(xs):
            f, gr, c, dT = oracle.evaluate(prm, batch, p, x)
            key = f"ev_{cfg}_{p}_{k}"
            assert f == float(g[key + "_f"]), (cfg, p, k, f, float(g[key + "_f"]))
            np.testing.assert_array_equal(gr, g[key + "_g"], err_msg=key)
            np.testing.assert_array_equal(c, g[key + "_c"], err_msg=key)
            np.testing.assert_array_equal(dT, g[key + "_dT"], err_msg=key)
        sol = oracle.optimize(prm, batch, p, p + 1, threads=1)
        n = batch.n_of(p)
        key = f"sv_{cfg}_{p}_"
        np.testing.assert_array_equal(sol.x[p, :n], g[key + "x"], err_msg=key)
        assert sol.cost[p] == float(g[key + "cost"])
        assert int(sol.status[p]) == int(g[key + "status"]) and int(sol.success[p]) == int(g[key + "success"])
        assert int(sol.iters[p]) == int(g[key + "iters"]) and int(sol.evals[p]) == int(g[key + "evals"])
        rows = g[key + "coeffs"].shape[0]
        np.testing.assert_array_equal(np.asarray(sol.coeffs[p]).reshape(-1, 2)[:rows], g[key + "coeffs"], err_msg=key)
        np.testing.assert_array_equal(np.asarray(sol.piece_dT)[batch.seg_ptr[p]:batch.seg_ptr[p + 1]], g[key + "piece_dT"])


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_oracle_matches_reference_optimizer_live(oracle, cfg):
    if oracle.ref_optimizer_lib() is None:
        pytest.skip("oracle/_ref/libref_optimizer.so not available (needs /root/reference or the prebuilt copy)")
    from dftpav_b200 import synth
    batch, prm = synth.make_config(cfg, B=3, seed=1000 + cfg)
    rng = np.random.default_rng(cfg)
    for p in range(batch.B):
        x0 = oracle.initial_x(prm, batch, p)
        for x in (x0, x0 + 0.02 * rng.normal(size=x0.shape), x0 + 0.2 * rng.normal(size=x0.shape)):
            fr, gr, cr, dr = oracle.ref_evaluate(prm, batch, p, x)
            fo, go, co, do = oracle.evaluate(prm, batch, p, x)
            assert fr == fo, (cfg, p, fr, fo)
            np.testing.assert_array_equal(go, gr)
            np.testing.assert_array_equal(co, cr)
            np.testing.assert_array_equal(do, dr)
    # one complete OptimizeTrajectory per config (the eager test layer makes the reference build slow)
    p = 1
    r = oracle.ref_solve(prm, batch, p)
    sol = oracle.optimize(prm, batch, p, p + 1, threads=1)
    n = batch.n_of(p)
    np.testing.assert_array_equal(sol.x[p, :n], r["x"])
    assert sol.cost[p] == r["cost"] and int(sol.status[p]) == r["status"] and int(sol.success[p]) == r["success"]
    assert int(sol.iters[p]) == r["iters"] and int(sol.evals[p]) == r["evals"]
