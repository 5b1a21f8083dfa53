"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

What is gated, and why (details in DESIGN.md, "Parity"):
  * dftpav_eval (one costFunctionCallback): |df|/max(1,|f|) <= 1e-12, |dg|_inf/max(1,|g|_inf) <= 1e-10 — strict;
    measured ~1e-14.  All arithmetic is fp64 on both sides; differences come from FMA contraction, reduction order,
    1/x products and libm (sin/cos/exp/log) only.
  * L-BFGS logic: per-iteration traces (cost, step, #evaluations) agree for the first iterations of every problem and
    stay together for most of each solve.
  * final results: the north_star tolerance is 1e-5 relative on cost / waypoints / times.  The reference algorithm is
    chaotic at the ulp level (tests/test_oracle_golden.py::test_reference_algorithm_is_chaotic_at_ulp_level: the
    oracle misses 1e-5 against ITSELF under a 1e-15 input perturbation in a sizeable fraction of problems), so the
    gate is: the fraction of problems inside 1e-5 is at least the oracle's own self-consistency fraction minus a
    margin, no systematic cost bias, success flags equal.
"""
import os

import numpy as np
import pytest

from dftpav_b200 import capi, synth

pytestmark = pytest.mark.gpu

EVAL_F_TOL = 1e-12
EVAL_G_TOL = 1e-10
SOLVE_TOL = 1e-5  # north_star: final waypoints/times/cost within 1e-5 relative
# allowed shortfall of the inside-1e-5 fraction against the oracle's own self-consistency fraction (round 1: 0.2 everywhere)
MARGIN = {1: 0.0, 2: 0.08, 3: 0.05, 4: 0.08, 5: 0.08}
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _eval_parity(oracle, opt, batch, prm, x):
    f, g = opt.evaluate(x)
    worst_f, worst_g = 0.0, 0.0
    for p in range(batch.B):
        n = batch.n_of(p)
        fo, go, _, _ = oracle.evaluate(prm, batch, p, x[p, :n])
        worst_f = max(worst_f, abs(f[p] - fo) / max(1.0, abs(fo)))
        worst_g = max(worst_g, np.abs(g[p, :n] - go).max() / max(1.0, np.abs(go).max()))
    return worst_f, worst_g


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_eval_parity_at_x0_and_perturbed(cfg, oracle, gpu_opt_factory):
    B = 1 if cfg == 1 else 24
    batch, prm = synth.make_config(cfg, B=B, seed=3)
    opt = gpu_opt_factory(prm)
    opt.upload(batch)
    x0 = opt.initial_x(batch)
    for p in range(batch.B):  # x0 packing itself must agree with the oracle
        np.testing.assert_array_equal(x0[p, : batch.n_of(p)], oracle.initial_x(prm, batch, p))
    rng = np.random.default_rng(cfg)
    for scale in (0.0, 0.05, 0.5):
        x = x0 + scale * rng.normal(size=x0.shape) * (x0 != 0)
        wf, wg = _eval_parity(oracle, opt, batch, prm, x)
        assert wf <= EVAL_F_TOL, (cfg, scale, wf)
        assert wg <= EVAL_G_TOL, (cfg, scale, wg)


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_eval_against_committed_golden(cfg, gpu_opt_factory):
    """Same check against tests/golden/oracle_eval.npz (fixtures recorded from the oracle in the build container)."""
    gold = np.load(os.path.join(GOLD, "oracle_eval.npz"))
    batch, prm = synth.make_config(cfg, B=2, seed=9)
    opt = gpu_opt_factory(prm)
    d = opt.upload(batch)
    x = np.zeros((2, d.n_stride))
    for p in range(2):
        x[p, : batch.n_of(p)] = gold[f"ev_{cfg}_{p}_x"]
    f, g = opt.evaluate(x)
    for p in range(2):
        n = batch.n_of(p)
        fo, go = float(gold[f"ev_{cfg}_{p}_f"]), gold[f"ev_{cfg}_{p}_g"]
        assert abs(f[p] - fo) <= EVAL_F_TOL * max(1.0, abs(fo))
        assert np.abs(g[p, :n] - go).max() <= EVAL_G_TOL * max(1.0, np.abs(go).max())


def _self_consistency(oracle, prm, batch, ref):
    """Fraction of problems for which the oracle agrees with ITSELF to 1e-5 under a 1e-15 relative input perturbation."""
    import copy

    b2 = copy.copy(batch)
    rng = np.random.default_rng(123)
    b2.inner_pts = batch.inner_pts * (1.0 + 1e-15 * np.sign(rng.normal(size=batch.inner_pts.shape)))
    pert = oracle.optimize(prm, b2, threads=0)
    rel = np.abs(pert.cost - ref.cost) / np.maximum(1.0, np.abs(ref.cost))
    return float(np.mean(rel <= SOLVE_TOL)), rel


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_eval_parity_against_the_reference_build(cfg, oracle, gpu_opt_factory):
    """The CUDA path against the reference's OWN code: oracle/_ref/libref_optimizer.so is src/traj_optimizer.cpp compiled in
    place (test-only Eigen layer, oracle/Makefile) and travels to the GPU box prebuilt.  One costFunctionCallback of the
    reference vs dftpav_eval at the initial and at a perturbed point: |df| <= 1e-12 max(1,|f|), |dg|_inf <= 1e-10
    max(1,|g|_inf), coefficients to 1e-12 (the oracle itself is bit-identical to this build, so this is the same gate as
    test_eval_parity_at_x0_and_perturbed with the restatement taken out of the loop)."""
    if oracle.ref_optimizer_lib() is None:
        pytest.skip("oracle/_ref/libref_optimizer.so did not travel")
    B = 4
    batch, prm = synth.make_config(cfg, B=B, seed=77)
    opt = gpu_opt_factory(prm)
    opt.upload(batch)
    x0 = opt.initial_x(batch)
    rng = np.random.default_rng(cfg)
    for scale in (0.0, 0.03):
        x = x0 + scale * rng.normal(size=x0.shape)
        f, g = opt.evaluate(x)
        for p in range(B):
            n = batch.n_of(p)
            fr, gr, cr, _ = oracle.ref_evaluate(prm, batch, p, x[p, :n])
            assert abs(f[p] - fr) <= EVAL_F_TOL * max(1.0, abs(fr)), (cfg, p, f[p], fr)
            assert np.abs(g[p, :n] - gr).max() <= EVAL_G_TOL * max(1.0, np.abs(gr).max()), (cfg, p)


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_solve_parity(cfg, oracle, gpu_opt_factory):
    B = 4 if cfg == 1 else 256
    CAP = 48
    batch, prm = synth.make_config(cfg, B=B, seed=11)
    opt = gpu_opt_factory(prm)
    opt.enable_trace(CAP)
    sol = opt.solve(batch)
    gtrace = opt.get_trace()
    opt.enable_trace(0)
    ref = oracle.optimize(prm, batch, threads=0, trace_cap=CAP)
    rel_cost = np.abs(sol.cost - ref.cost) / np.maximum(1.0, np.abs(ref.cost))
    rel_x = np.abs(sol.x - ref.x).max(axis=1) / np.maximum(1.0, np.abs(ref.x).max(axis=1))
    inside = (rel_cost <= SOLVE_TOL) & (rel_x <= SOLVE_TOL) & (sol.status == ref.status)
    self_frac, self_rel = _self_consistency(oracle, prm, batch, ref)
    # ---- L-BFGS logic: traces agree at the start of every solve
    agree = np.zeros(B, dtype=int)
    for p in range(B):
        kmax = min(CAP, int(min(sol.iters[p], ref.iters[p])))
        a = 0
        for k in range(kmax):
            fo, fg = ref.trace[p, k, 0], gtrace[p, k, 0]
            if abs(fo - fg) <= 1e-9 * max(1.0, abs(fo)) and ref.trace[p, k, 2] == gtrace[p, k, 2]:
                a += 1
            else:
                break
        agree[p] = a
    print(f"cfg{cfg}: inside-1e-5 fraction gpu-vs-oracle {inside.mean():.3f} (oracle-vs-perturbed-oracle {self_frac:.3f}); "
          f"cost rel median {np.median(rel_cost):.1e} p90 {np.quantile(rel_cost, .9):.1e} max {rel_cost.max():.1e} "
          f"(self: median {np.median(self_rel):.1e} max {self_rel.max():.1e}); iterations equal {np.mean(sol.iters == ref.iters):.2f}; "
          f"trace agreement: min {agree.min()} median {int(np.median(agree))} iterations; evals gpu/ref {sol.evals.sum()}/{ref.evals.sum()}")
    assert agree.min() >= min(3, int(ref.iters.min())), agree
    # most problems agree on a long prefix of the iteration history (a third of the oracle's iterations or more)
    assert np.mean(agree >= np.minimum(ref.iters, CAP) // 3) >= 0.7
    # ---- final results: statistically indistinguishable from the oracle's own ulp sensitivity (tests/test_gpu_strict.py
    #      holds the bit-exact proof of the algorithm; these gates bound what the rounding-level substitutions of the
    #      production build may do)
    from scipy.stats import ks_2samp

    assert np.mean(sol.success == ref.success) >= 0.97
    assert np.mean(sol.status == ref.status) >= 0.9
    assert inside.mean() >= max(0.0, self_frac - MARGIN[cfg]), (inside.mean(), self_frac)
    assert rel_cost.max() <= max(0.25, 3 * self_rel.max())
    if B >= 64:
        ks = ks_2samp(np.log10(rel_cost + 1e-17), np.log10(self_rel + 1e-17))
        print(f"cfg{cfg}: two-sample KS of log10(rel cost diff), gpu-vs-oracle against oracle-vs-perturbed-oracle: "
              f"D = {ks.statistic:.3f}, p = {ks.pvalue:.3f}")
        assert ks.statistic <= 0.25, ks
    bias = sol.cost.mean() / ref.cost.mean() - 1.0
    print(f"cfg{cfg}: cost bias {bias:+.2e}, evaluations gpu/oracle {sol.evals.sum() / ref.evals.sum():.4f}")
    assert abs(bias) <= (1e-3 if B >= 64 else 1e-2), bias
    assert abs(sol.evals.sum() / ref.evals.sum() - 1.0) <= (0.05 if B >= 64 else 0.2)
    # the coefficients returned are those of the last evaluated point: re-evaluating at the final x reproduces them
    # whenever the line search did not fail (quirk Q8)
    for p in range(min(B, 8)):
        if sol.status[p] >= 0:
            n = batch.n_of(p)
            _, _, c, dT = oracle.evaluate(prm, batch, p, sol.x[p, :n])
            rows = batch.rows_of(p)
            np.testing.assert_allclose(sol.coeffs[p, :rows], c, rtol=1e-9, atol=1e-9)


def test_ragged_corridor_and_mixed_shapes(oracle, gpu_opt_factory):
    """Variable plane count per sample (hPoly column count is variable in the reference API) and several shapes
    in one batch (grouped into one launch per shape)."""
    from dftpav_b200.batch import TrajBatch

    prm = capi.default_params()
    b2, _ = synth.make_config(2, B=6, seed=5)
    b3, _ = synth.make_config(3, B=5, seed=5)
    K, Kd = prm.traj_resolution, prm.des_traj_resolution
    M2 = b2.total_samples(K, Kd)
    cnt2 = np.where(np.arange(M2) % 2 == 0, 6, 4).astype(np.int32)
    pl2 = b2.planes.reshape(M2, 6, 4)
    planes2 = np.concatenate([pl2[q, : cnt2[q]] for q in range(M2)], axis=0)
    M3 = b3.total_samples(K, Kd)
    cnt = np.concatenate([cnt2, np.full(M3, 4, dtype=np.int32)])
    mixed = TrajBatch(seg_ptr=np.concatenate([b2.seg_ptr, b2.seg_ptr[-1] + b3.seg_ptr[1:]]),
                      piece_num=np.concatenate([b2.piece_num, b3.piece_num]),
                      singul=np.concatenate([b2.singul, b3.singul]),
                      ini_state=np.concatenate([b2.ini_state, b3.ini_state]),
                      fin_state=np.concatenate([b2.fin_state, b3.fin_state]),
                      init_T=np.concatenate([b2.init_T, b3.init_T]),
                      inner_pts=np.concatenate([b2.inner_pts, b3.inner_pts]),
                      planes=np.concatenate([planes2, b3.planes]), t_now=np.zeros(11), plane_cnt=cnt)
    opt = gpu_opt_factory(prm)
    opt.upload(mixed)
    assert opt.stats().n_shapes == 2
    x0 = opt.initial_x(mixed)
    wf, wg = _eval_parity(oracle, opt, mixed, prm, x0)
    assert wf <= EVAL_F_TOL and wg <= EVAL_G_TOL, (wf, wg)
    sol = opt.solve(mixed)
    ref = oracle.optimize(prm, mixed, threads=0)
    assert np.array_equal(sol.success, ref.success)
    assert abs(sol.cost.mean() / ref.cost.mean() - 1.0) < 2e-2


def test_bad_problem_is_rejected_alone(oracle, gpu_opt_factory):
    """Reference preconditions (traj_optimizer.cpp:26-48): the reference rejects that ONE call (`return false`), so a bad
    problem is reported with status DFTPAV_PROBLEM_REJECTED / success 0 and the rest of the batch is solved as if it were
    not there."""
    batch, prm = synth.make_config(3, B=6, seed=0)
    good = gpu_opt_factory(prm).solve(batch)
    batch.init_T[2] = 0.01  # < mini_T
    opt = gpu_opt_factory(prm)
    sol = opt.solve(batch)
    assert sol.status[2] == capi.PROBLEM_REJECTED and sol.success[2] == 0 and sol.evals[2] == 0
    assert np.all(sol.x[2] == 0.0) and sol.cost[2] == 0.0
    keep = [0, 1, 3, 4, 5]
    np.testing.assert_array_equal(sol.x[keep], good.x[keep])
    np.testing.assert_array_equal(sol.cost[keep], good.cost[keep])
    np.testing.assert_array_equal(sol.status[keep], good.status[keep])
    ref = oracle.optimize(prm, batch, threads=0)
    assert ref.status[2] == capi.PROBLEM_REJECTED and ref.success[2] == 0
    f, g = opt.evaluate(opt.initial_x(batch))
    assert f[2] == 0.0 and np.all(g[2] == 0.0) and np.all(np.isfinite(f[keep]))


def test_malformed_batches_are_refused(gpu_opt_factory):
    """Inconsistent CSR arrays -> DFTPAV_ERR_INVALID_ARG from the validation pass, nothing is dereferenced blindly."""
    from dftpav_b200.optimizer import DftpavError

    def expect_invalid(mutate, cfg=5):
        batch, prm = synth.make_config(cfg, B=4, seed=0)
        mutate(batch)
        with pytest.raises(DftpavError) as e:
            gpu_opt_factory(prm).upload(batch)
        assert e.value.code == capi.ERR_INVALID_ARG, e.value

    expect_invalid(lambda b: b.obs_idx.__setitem__(0, 10 ** 6))          # obstacle index outside the pool
    expect_invalid(lambda b: b.obs_idx.__setitem__(1, -1))
    expect_invalid(lambda b: b.obs_ptr.__setitem__(2, b.obs_ptr[3] + 1))  # not monotone
    expect_invalid(lambda b: b.obs_piece_ptr.__setitem__(1, b.obs_piece_ptr[0]))  # obstacle without pieces
    expect_invalid(lambda b: b.seg_ptr.__setitem__(0, 1), cfg=3)
    expect_invalid(lambda b: b.piece_num.__setitem__(1, 0), cfg=3)


@pytest.mark.parametrize("field,value,code", [
    ("lbfgs_mem_size", 0, capi.LBFGSERR_INVALID_MEMSIZE), ("lbfgs_g_epsilon", -1.0, capi.LBFGSERR_INVALID_GEPSILON),
    ("lbfgs_past", -1, capi.LBFGSERR_INVALID_TESTPERIOD), ("lbfgs_delta", -1.0, capi.LBFGSERR_INVALID_DELTA),
    ("lbfgs_min_step", -1.0, capi.LBFGSERR_INVALID_MINSTEP), ("lbfgs_max_step", 1e-40, capi.LBFGSERR_INVALID_MAXSTEP),
    ("lbfgs_f_dec_coeff", 1.5, capi.LBFGSERR_INVALID_FDECCOEFF), ("lbfgs_s_curv_coeff", 1e-5, capi.LBFGSERR_INVALID_SCURVCOEFF),
    ("lbfgs_machine_prec", 0.0, capi.LBFGSERR_INVALID_MACHINEPREC), ("lbfgs_max_linesearch", 0, capi.LBFGSERR_INVALID_MAXLINESEARCH)])
def test_invalid_lbfgs_parameters_return_the_reference_codes(field, value, code, oracle, gpu_opt_factory):
    """lbfgs_optimize checks its parameter block before anything else (lbfgs.hpp:456-501): every solve returns that code, the
    decision vector stays the initial packing, OptimizeTrajectory reports failure."""
    batch, prm = synth.make_config(3, B=3, seed=1)
    setattr(prm, field, value)
    opt = gpu_opt_factory(prm)
    sol = opt.solve(batch)
    ref = oracle.optimize(prm, batch, threads=0)
    assert np.all(sol.status == code) and np.all(ref.status == code), (sol.status, ref.status)
    assert np.all(sol.success == 0) and np.all(sol.evals == 0)
    np.testing.assert_array_equal(sol.x, opt.initial_x(batch))


def test_determinism_and_sharding(gpu_opt_factory):
    """Same problems, different batch composition / launch shape -> bit-identical results (each problem's
    arithmetic is confined to one warp)."""
    batch, prm = synth.make_config(3, B=96, seed=2)
    opt = gpu_opt_factory(prm)
    full = opt.solve(batch)
    again = opt.solve_stepwise(batch)
    np.testing.assert_array_equal(full.x, again.x)
    half = batch.slice(40, 96, prm.traj_resolution, prm.des_traj_resolution)
    part = opt.solve(half)
    np.testing.assert_array_equal(full.x[40:96], part.x)
    np.testing.assert_array_equal(full.cost[40:96], part.cost)
    np.testing.assert_array_equal(full.iters[40:96], part.iters)


@pytest.mark.parametrize("cfg,B", [(3, 1500), (2, 1031), (5, 700)])
def test_queue_order_is_invisible(cfg, B, gpu_opt_factory):
    """Batches above 4 problems per SM are served longest-predicted-first through per-chunk lists (chunk_order_kernel,
    fetch_by_chunk): every problem must be solved exactly once and no result may depend on the order.  Compared: the
    overlapped call (32 upload chunks), upload + optimize (one chunk), a caller-supplied reversed order (plain counter
    queue), and the two halves solved as separate batches."""
    batch, prm = synth.make_config(cfg, B=B, seed=9)
    opt = gpu_opt_factory(prm)
    a = opt.solve(batch)
    b = opt.solve_stepwise(batch)
    opt.set_order(np.arange(B - 1, -1, -1))
    c = opt.solve_stepwise(batch)
    opt.set_order(None)
    assert np.all(a.evals >= 1)  # nobody was skipped
    for other in (b, c):
        np.testing.assert_array_equal(a.x, other.x)
        np.testing.assert_array_equal(a.cost, other.cost)
        np.testing.assert_array_equal(a.status, other.status)
        np.testing.assert_array_equal(a.evals, other.evals)
    h = B // 2
    for lo, hi in ((0, h), (h, B)):
        part = opt.solve(batch.slice(lo, hi, prm.traj_resolution, prm.des_traj_resolution))
        np.testing.assert_array_equal(a.x[lo:hi], part.x)
        np.testing.assert_array_equal(a.iters[lo:hi], part.iters)


@pytest.mark.parametrize("cfg", [3, 4, 5])
def test_helped_and_unhelped_solves_are_bit_identical(cfg, gpu_opt_factory):
    """Drain-phase helpers: which warps of the CTA swept which virtual ranks must not change a single
    bit.  Tiny batches run with full teams from the first evaluation (B = 1: one owner + helpers; B = 2, 3, 5: owners
    outliving each other, helpers re-attaching), against the same problems solved inside a batch large enough that the
    queue keeps every warp busy for most of the run; and through both entry points (dftpav_solve copies the planes while
    the kernel runs, dftpav_upload_batch + dftpav_optimize before)."""
    big, prm = synth.make_config(cfg, B=600, seed=5)
    opt = gpu_opt_factory(prm)
    full = opt.solve(big)
    for lo, hi in ((0, 1), (7, 9), (10, 13), (20, 25), (100, 250)):
        sub = big.slice(lo, hi, prm.traj_resolution, prm.des_traj_resolution)
        for part in (opt.solve(sub), opt.solve_stepwise(sub)):
            np.testing.assert_array_equal(full.x[lo:hi], part.x)
            np.testing.assert_array_equal(full.cost[lo:hi], part.cost)
            np.testing.assert_array_equal(full.evals[lo:hi], part.evals)
            np.testing.assert_array_equal(full.coeffs[lo:hi], part.coeffs)
    # the evaluation hook sweeps alone (every virtual rank in order): same bits as a helped evaluation of the solver
    one = big.slice(3, 4, prm.traj_resolution, prm.des_traj_resolution)
    s1 = opt.solve_stepwise(one)
    if s1.status[0] >= 0:
        f, _ = opt.evaluate(s1.x)
        assert f[0] == s1.cost[0]


def test_odd_piece_counts_and_other_resolutions(oracle, gpu_opt_factory):
    """Odd N puts the two tiles of the last piece into different sweep passes unless an empty tile is inserted
    (ShapeDev::tpad); resolutions other than the reference's (K > 16 or Kd > 32) switch the helpers off.  Evaluation parity
    and bit-identical small/large batches in all cases."""
    from dftpav_b200.synth import make_custom

    for pieces, K, Kd in (([7], 16, 32), ([9], 16, 32), ([8], 16, 16), ([8], 8, 16), ([6], 24, 40), ([5], 16, 48)):
        batch, prm = make_custom(B=40, pieces=pieces, seed=2)
        if (K, Kd) != (16, 32):
            # the generator sizes the corridor list from the resolutions: regenerate with them
            prm.traj_resolution, prm.des_traj_resolution = K, Kd
            from dftpav_b200 import synth as S

            batch = S._single_segment(90, 40, 2, pieces[0], prm, length=0.8 * 3.0 * pieces[0], f_lo=0.5, f_hi=1.0, corridor="rect")
        opt = gpu_opt_factory(prm)
        opt.upload(batch)
        x0 = opt.initial_x(batch)
        wf, wg = _eval_parity(oracle, opt, batch, prm, x0 + 0.05 * np.random.default_rng(1).normal(size=x0.shape))
        assert wf <= EVAL_F_TOL and wg <= EVAL_G_TOL, (pieces, K, Kd, wf, wg)
        full = opt.solve(batch)
        part = opt.solve(batch.slice(5, 7, K, Kd))
        np.testing.assert_array_equal(full.x[5:7], part.x)
        np.testing.assert_array_equal(full.cost[5:7], part.cost)


def test_large_shapes(oracle, gpu_opt_factory):
    """Up to the compiled-in limits: n > 96 decision variables (the register two-loop covered 96 and the rest
    of the direction was silently never updated), 4 gear segments, tables that do not fit densely (banded fallback)."""
    from dftpav_b200.synth import make_custom

    for pieces, B in (([56], 24), ([64], 8), ([40, 30, 40, 20], 8), ([64, 64, 64, 64], 4)):
        batch, prm = make_custom(B=B, pieces=pieces, seed=4)
        opt = gpu_opt_factory(prm)
        opt.upload(batch)
        x0 = opt.initial_x(batch)
        wf, wg = _eval_parity(oracle, opt, batch, prm, x0)
        assert wf <= EVAL_F_TOL and wg <= 10 * EVAL_G_TOL, (pieces, wf, wg)
        if sum(pieces) > 130:
            continue  # the CPU oracle needs minutes for complete solves of this size
        CAP = 12
        opt.enable_trace(CAP)
        sol = opt.solve(batch)
        tr = opt.get_trace()
        opt.enable_trace(0)
        ref = oracle.optimize(prm, batch, threads=0, trace_cap=CAP)
        # the first iterations (history of 1 .. CAP pairs, all n components of the direction) agree with the oracle
        for p in range(B):
            for k in range(min(6, int(sol.iters[p]), int(ref.iters[p]))):
                assert abs(tr[p, k, 0] - ref.trace[p, k, 0]) <= 1e-8 * max(1.0, abs(ref.trace[p, k, 0])), (pieces, p, k)
                assert tr[p, k, 2] == ref.trace[p, k, 2], (pieces, p, k)
        assert np.mean(sol.success == ref.success) >= 0.9
        assert abs(sol.cost.mean() / ref.cost.mean() - 1.0) < 2e-2


def test_banded_fallback(oracle, gpu_opt_factory):
    """The banded LU substitutions (what large shapes fall back to) forced on a BASELINE shape: evaluation parity with the
    oracle - whose arithmetic they follow more closely than the dense products do - and complete solves."""
    for cfg in (3, 4):
        batch, prm = synth.make_config(cfg, B=64, seed=6)
        opt = gpu_opt_factory(prm)
        opt.force_banded(True)
        opt.upload(batch)
        x0 = opt.initial_x(batch)
        wf, wg = _eval_parity(oracle, opt, batch, prm, x0 + 0.03 * np.random.default_rng(2).normal(size=x0.shape) * (x0 != 0))
        assert wf <= EVAL_F_TOL and wg <= EVAL_G_TOL, (cfg, wf, wg)
        sol = opt.solve(batch)
        ref = oracle.optimize(prm, batch, threads=0)
        assert np.mean(sol.success == ref.success) >= 0.97
        assert abs(sol.cost.mean() / ref.cost.mean() - 1.0) < 5e-3
        opt.force_banded(False)


# L-BFGS / cost parameters away from the reference defaults (none of them was varied by a GPU test in round 1); the
# bit-exact versions of these cases live in tests/test_gpu_strict.py, here the production build is gated statistically
PARAM_CASES = {
    "ring_wrap_mem4": dict(lbfgs_mem_size=4), "ring_wrap_mem8": dict(lbfgs_mem_size=8), "no_past_test": dict(lbfgs_past=0),
    "help_eps": dict(help_eps=1.0e-4), "max_iterations": dict(lbfgs_max_iterations=7),
    "max_linesearch": dict(lbfgs_max_linesearch=2), "width_too_small": dict(lbfgs_machine_prec=0.2),
    "minimum_step": dict(lbfgs_min_step=0.3),
}


@pytest.mark.parametrize("case", sorted(PARAM_CASES))
def test_parameter_variations(case, oracle, gpu_opt_factory):
    cfg = 4 if case == "no_gear_opt" else 3
    batch, prm = synth.make_config(cfg, B=128, seed=33)
    for k, v in PARAM_CASES[case].items():
        setattr(prm, k, v)
    opt = gpu_opt_factory(prm)
    sol = opt.solve(batch)
    ref = oracle.optimize(prm, batch, threads=0)
    rel = np.abs(sol.cost - ref.cost) / np.maximum(1.0, np.abs(ref.cost))
    print(f"{case}: statuses gpu {sorted(set(sol.status.tolist()))} oracle {sorted(set(ref.status.tolist()))}; "
          f"inside 1e-5 {np.mean(rel <= SOLVE_TOL):.3f}; status equal {np.mean(sol.status == ref.status):.3f}; "
          f"iterations equal {np.mean(sol.iters == ref.iters):.3f}")
    assert np.mean(sol.status == ref.status) >= 0.85
    assert np.mean(sol.success == ref.success) >= 0.95
    assert abs(sol.cost.mean() / ref.cost.mean() - 1.0) <= 5e-3
    assert abs(sol.evals.sum() / ref.evals.sum() - 1.0) <= 0.1
    if case in ("max_iterations", "max_linesearch"):  # short, deterministic-length solves: nothing chaotic yet
        assert np.mean(rel <= SOLVE_TOL) >= 0.95


def test_gear_opt_off(oracle, gpu_opt_factory):
    batch, prm = synth.make_config(4, B=48, seed=3)
    prm.gear_opt = 0
    opt = gpu_opt_factory(prm)
    opt.upload(batch)
    x0 = opt.initial_x(batch)
    wf, wg = _eval_parity(oracle, opt, batch, prm, x0)
    assert wf <= EVAL_F_TOL and wg <= EVAL_G_TOL
    sol = opt.solve(batch)
    ref = oracle.optimize(prm, batch, threads=0)
    tv0 = 2 * (12 - 1) + 2 * (8 - 1) + 2 * (12 - 1) + 3
    np.testing.assert_array_equal(sol.x[:, tv0:tv0 + 6], opt.initial_x(batch)[:, tv0:tv0 + 6])  # gears / angles never move
    assert np.mean(sol.success == ref.success) >= 0.95


def test_full_size_properties(gpu_opt_factory):
    """BASELINE.json full size (config 3, batch 4096): size-independent properties of every solve."""
    batch, prm = synth.make_config(3, B=4096, seed=0)
    opt = gpu_opt_factory(prm)
    opt.upload(batch)
    x0 = opt.initial_x(batch)
    f0, _ = opt.evaluate(x0)
    opt.optimize()
    sol = opt.download()
    N = 24
    assert np.all(np.isfinite(sol.cost)) and np.all(np.isfinite(sol.x))
    valid = {capi.LBFGS_CONVERGENCE, capi.LBFGS_STOP, capi.LBFGSERR_MAXIMUMLINESEARCH, capi.LBFGSERR_MAXIMUMITERATION,
             capi.LBFGSERR_WIDTHTOOSMALL, capi.LBFGSERR_MINIMUMSTEP}
    assert set(np.unique(sol.status).tolist()) <= valid
    ok = sol.status >= 0
    assert ok.mean() > 0.98
    assert np.all(sol.cost[ok] <= f0[ok] * (1 + 1e-12))                       # descent
    assert np.all(sol.success == ((sol.cost < prm.cost_limit) & (ok | (sol.status == capi.LBFGSERR_MAXIMUMLINESEARCH))))
    assert np.all(sol.evals >= sol.iters + 1) and sol.iters.max() < 12000
    # MINCO identities on the returned coefficients: waypoints of x are interpolated, start state reproduced, C2 at knots
    c = sol.coeffs.reshape(4096, N, 6, 2)
    dT = sol.piece_dT[:, None]
    pw = dT[:, :, None] ** np.arange(6)[None, None, :]                        # [B,1,6]
    end = (c * pw[..., None]).sum(axis=2)                                     # position at the end of every piece
    wp = sol.x[:, : 2 * (N - 1)].reshape(4096, N - 1, 2)
    assert np.abs(end[ok, : N - 1] - wp[ok]).max() < 1e-6
    assert np.abs(c[ok, 1:, 0] - end[ok, : N - 1]).max() < 1e-6               # C0 at the knots
    vel_end = (c[:, :, 1:] * (np.arange(1, 6)[None, None, :, None] * pw[:, :, :5, None])).sum(axis=2)
    assert np.abs(c[ok, 1:, 1] - vel_end[ok, : N - 1]).max() < 1e-6           # C1 at the knots
    assert np.abs(c[ok, 0, 0] - batch.ini_state[ok, 0:2]).max() < 1e-9        # start position
    # the optimised durations stay positive and above mini_T
    assert np.all(sol.piece_dT * N > prm.mini_T)
    # a second run is bit-identical (fixed-order reductions, no atomics on the data path)
    opt.optimize()
    again = opt.download()
    np.testing.assert_array_equal(sol.x, again.x)
