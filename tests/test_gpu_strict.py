"""Strict-order verification build (libdftpav_b200_strict.so = the product sources with -DDFTPAV_STRICT=1 -fmad=false) against
the CPU oracle: COMPLETE SOLVES MUST BE BIT-IDENTICAL.

The production kernels differ from the reference in rounding-level ways only (FMA, parallel reductions, A^-1 products, 1/ys,
j*step, rsqrt, vertex weight 2: DESIGN.md section 4), but the reference algorithm is chaotic at the ulp level, so a production
solve cannot be compared bit by bit.  The strict build removes every one of those differences (dftpav_strict.cuh) while
sharing the whole solver state machine of solve_kernel - fetch, first evaluation, Lewis-Overton trials, convergence tests,
history ring indices, revert on failure, write-back - the upload path, the plane layout and the batch plumbing with the
production build.  sqrt and division are correctly rounded on both sides and no libm call is on the path of configurations
1-3, so x, cost, status, iteration and evaluation counts and the coefficients of every problem must agree to the last bit.
This proves the L-BFGS state machine, and turns "the fast build is statistically indistinguishable" (test_gpu_parity.py)
into "the fast build differs from a bit-exact implementation only by the listed rounding-level substitutions".
"""
import numpy as np
import pytest

from dftpav_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _assert_identical(sol, ref, B, what):
    for name in ("status", "iters", "evals", "success"):
        a, b = getattr(sol, name)[:B], getattr(ref, name)[:B]
        bad = np.nonzero(a != b)[0]
        assert bad.size == 0, (what, name, bad[:8], a[bad[:8]], b[bad[:8]])
    for name in ("cost", "x", "coeffs", "piece_dT"):
        a, b = getattr(sol, name), getattr(ref, name)
        assert np.array_equal(a, b), (what, name, float(np.abs(a - b).max()))


@pytest.mark.parametrize("cfg,B", [(1, 64), (2, 512), (3, 512)])
def test_complete_solves_bit_identical(cfg, B, oracle, gpu_strict_factory):
    """>= 512 complete OptimizeTrajectory runs per configuration (config 1 is the single-problem plumbing case: 64 seeds)."""
    batch, prm = synth.make_config(cfg, B=B, seed=21)
    opt = gpu_strict_factory(prm)
    sol = opt.solve(batch)
    ref = oracle.optimize(prm, batch, threads=0)
    print(f"cfg{cfg}: {B} solves, iterations {ref.iters.min()}..{ref.iters.max()}, evaluations {ref.evals.sum()}, "
          f"statuses {sorted(set(ref.status.tolist()))}")
    _assert_identical(sol, ref, B, f"cfg{cfg}")


def test_one_evaluation_bit_identical(oracle, gpu_strict_factory):
    """cost and gradient of costFunctionCallback at the initial and at perturbed points: equal bits."""
    for cfg in (1, 2, 3):
        batch, prm = synth.make_config(cfg, B=1 if cfg == 1 else 16, seed=5)
        opt = gpu_strict_factory(prm)
        opt.upload(batch)
        x0 = opt.initial_x(batch)
        rng = np.random.default_rng(cfg)
        for scale in (0.0, 0.05):
            x = x0 + scale * rng.normal(size=x0.shape) * (x0 != 0)
            f, g = opt.evaluate(x)
            for p in range(batch.B):
                n = batch.n_of(p)
                fo, go, _, _ = oracle.evaluate(prm, batch, p, x[p, :n])
                assert f[p] == fo, (cfg, p, f[p], fo)
                assert np.array_equal(g[p, :n], go), (cfg, p, np.abs(g[p, :n] - go).max())


# every branch of lbfgs_optimize / line_search_lewisoverton that the default parameters never reach, to the last bit
BRANCH_CASES = {
    "ring_wrap_mem4": dict(lbfgs_mem_size=4),                      # end = (end + 1) % m wraps, bound == m (lbfgs.hpp:716-739)
    "ring_wrap_mem8": dict(lbfgs_mem_size=8),
    "no_past_test": dict(lbfgs_past=0),                            # delta test off (lbfgs.hpp:326, 642-659)
    "help_eps": dict(help_eps=1.0e-4),                             # the header default of traj_optimizer.h:120
    "max_iterations": dict(lbfgs_max_iterations=7),                # LBFGSERR_MAXIMUMITERATION counts as success
    "max_linesearch": dict(lbfgs_max_linesearch=2),                # LBFGSERR_MAXIMUMLINESEARCH counts as success
    "width_too_small": dict(lbfgs_machine_prec=0.2),               # LBFGSERR_WIDTHTOOSMALL (lbfgs.hpp:355-359)
    "minimum_step": dict(lbfgs_min_step=0.3),                      # LBFGSERR_MINIMUMSTEP (lbfgs.hpp:370-376)
    "maximum_step": dict(lbfgs_max_step=1.0e-3),                   # step clamped once, then LBFGSERR_MAXIMUMSTEP
    "tight_delta": dict(lbfgs_delta=1.0e-9),                       # long solves: > 256 iterations wrap the default ring too
}


@pytest.mark.parametrize("case", sorted(BRANCH_CASES))
def test_lbfgs_branches_bit_identical(case, oracle, gpu_strict_factory):
    batch, prm = synth.make_config(3, B=128, seed=33)
    for k, v in BRANCH_CASES[case].items():
        setattr(prm, k, v)
    opt = gpu_strict_factory(prm)
    sol = opt.solve(batch)
    ref = oracle.optimize(prm, batch, threads=0)
    st = sorted(set(ref.status.tolist()))
    print(f"{case}: statuses {st}, iterations {ref.iters.min()}..{ref.iters.max()}")
    want = {"max_iterations": capi.LBFGSERR_MAXIMUMITERATION, "max_linesearch": capi.LBFGSERR_MAXIMUMLINESEARCH,
            "width_too_small": capi.LBFGSERR_WIDTHTOOSMALL, "minimum_step": capi.LBFGSERR_MINIMUMSTEP,
            "maximum_step": capi.LBFGSERR_MAXIMUMSTEP}.get(case)
    if want is not None:
        assert want in st, (case, st)  # the case really reaches the branch it is named after
    if case.startswith("ring_wrap"):
        assert ref.iters.max() > prm.lbfgs_mem_size + 2
    _assert_identical(sol, ref, batch.B, case)


def test_large_and_multi_segment_shapes(oracle, gpu_strict_factory):
    """n > 96 decision variables (the register two-loop of the production build stops there) and gear_opt
    off; gear segments use sin / cos (CUDA libm vs glibc: not guaranteed bit-equal), so configuration 4 is reported as a
    fraction and gated loosely."""
    from dftpav_b200.synth import make_custom

    batch, prm = make_custom(B=12, pieces=[56], seed=3)
    assert batch.n_of(0) > 96
    opt = gpu_strict_factory(prm)
    sol = opt.solve(batch)
    ref = oracle.optimize(prm, batch, threads=0)
    _assert_identical(sol, ref, batch.B, "N=56")
    b4, p4 = synth.make_config(4, B=64, seed=8)
    o4 = gpu_strict_factory(p4)
    s4 = o4.solve(b4)
    r4 = oracle.optimize(p4, b4, threads=0)
    same = (s4.cost == r4.cost) & (s4.iters == r4.iters) & (s4.status == r4.status)
    rel = np.abs(s4.cost - r4.cost) / np.maximum(1.0, np.abs(r4.cost))
    # measured: 0.000 - every evaluation calls sin / cos of the gear angles, the CUDA and glibc implementations differ by an
    # ulp now and then, and the algorithm amplifies that like any other rounding difference.  What is
    # gated here is therefore only what the production build is gated on for this configuration.
    print(f"cfg4 (sin/cos on the path): bit-identical solves {same.mean():.3f}, rel cost median {np.median(rel):.1e}")
    assert np.mean(s4.success == r4.success) >= 0.95 and abs(s4.cost.mean() / r4.cost.mean() - 1.0) <= 5e-3
    # the first evaluation (x0 is packed on the host with glibc's atan2) differs by libm rounding only
    o4.upload(b4)
    x0 = o4.initial_x(b4)
    f, g = o4.evaluate(x0)
    for p in range(8):
        fo, go, _, _ = oracle.evaluate(p4, b4, p, x0[p, : b4.n_of(p)])
        assert abs(f[p] - fo) <= 1e-13 * abs(fo)
