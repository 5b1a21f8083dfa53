"""Generates the fixtures under tests/golden/ (run in the build container, where /root/reference exists).

  lbfgs_reference_header.npz : iterates of the UNMODIFIED reference L-BFGS (include/geo_utils2d/lbfgs.hpp compiled in
                               place into oracle/_ref against the test-only Eigen shim) on analytic functions and on
                               real trajectory problems (cost function = the oracle's).  The GPU box has no
                               /root/reference: there the oracle's L-BFGS restatement is checked against these.
  oracle_eval.npz            : cost / gradient / coefficients of the oracle at fixed points of every BASELINE config
                               (regression pin of the restatement itself).
  solve_small.npz            : final results of full oracle solves on small batches of every config.
  minco_reference_header.npz : outputs of the UNMODIFIED reference MINCO header (plan_utils/poly_traj_utils.hpp compiled
                               in place into oracle/_ref/libref_minco.so against oracle/eigen_shim2): coefficients, jerk
                               energy, gdC / gdT initialisation, calGrads_PT outputs, trajectory samples.
  optimizer_reference.npz    : outputs of the UNMODIFIED reference optimizer (src/traj_optimizer.cpp compiled in place into
                               oracle/_ref/libref_optimizer.so): costFunctionCallback values / gradients / coefficients at
                               fixed points and complete OptimizeTrajectory results, all five configs.
  traj_state_reference.npz   : the reference header's Trajectory / Piece evaluators (getPos ... getSteer, GetState) on the
                               coefficient sets of tests/test_traj_state_oracle.py::_cases (row N2 of SURVEY.md section 8f).
usage: python tests/make_golden.py [--reference-only]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dftpav_b200 import capi, synth  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
ANALYTIC = [(0, 10), (1, 30), (2, 12), (0, 50)]
LB_SETTINGS = [(256, 3, 1e-4, 1e-16), (8, 3, 1e-8, 1e-16), (5, 0, 0.0, 1e-10)]
CAP = 600


def minco_cases():
    """Deterministic inputs of the MINCO fixture: (N, P, dT, head, tail, extra_gdC, extra_gdT)."""
    rng = np.random.default_rng(20240917)
    out = []
    for N, dT in ((2, 0.8), (3, 0.4), (8, 1.3), (24, 0.9), (24, 2.2), (40, 0.55)):
        P = np.cumsum(rng.normal(size=(N - 1, 2)) * 2.0, axis=0)
        head = rng.normal(size=(3, 2)) * np.array([[5.0], [3.0], [1.0]])
        tail = rng.normal(size=(3, 2)) * np.array([[5.0], [3.0], [1.0]])
        tail[0] += P[-1]
        out.append((N, P, dT, head, tail, rng.normal(size=(6 * N, 2)) * 10.0, float(rng.normal())))
    return out


def traj_times(case):
    N, dT = case[0], case[2]
    return [0.0, 0.37 * dT, dT, 1.5 * dT, (N - 0.25) * dT, N * dT, N * dT + 0.9]


def analytic_x0(n):
    return np.linspace(-1.2, 1.0, n)


def run_ref(n, x0, fgptr, user, mem, past, delta, geps, max_iter=12000):
    R = O.ref_lib()
    assert R is not None, "oracle/_ref is not built"
    dp = capi.c_double_p
    x = x0.copy()
    f = C.c_double()
    ev = C.c_int()
    tr = np.zeros((CAP, 3))
    ret = R.ref_lbfgs(n, x.ctypes.data_as(dp), C.cast(C.byref(f), dp), fgptr, user, mem, past, geps, 1e-32, delta, max_iter,
                      1, C.byref(ev), tr.ctypes.data_as(dp), CAP)
    return ret, f.value, x, ev.value, tr


OPT_CASES = [(cfg, p) for cfg in (1, 2, 3, 4, 5) for p in (0, 1)]   # (config, problem) pairs of the optimizer fixture


def optimizer_inputs(cfg, p):
    """Deterministic inputs of the reference-optimizer fixture: the batch and two evaluation points."""
    batch, prm = synth.make_config(cfg, B=2, seed=21)
    x0 = O.initial_x(prm, batch, p)
    xs = [x0, x0 + 0.04 * np.random.default_rng(100 * cfg + p).normal(size=x0.shape)]
    return batch, prm, xs


def make_optimizer():
    assert O.ref_optimizer_lib() is not None, "oracle/_ref/libref_optimizer.so missing: run `make -C oracle ref`"
    out = {}
    for cfg, p in OPT_CASES:
        batch, prm, xs = optimizer_inputs(cfg, p)
        for k, x in enumerate(xs):
            f, g, c, dT = O.ref_evaluate(prm, batch, p, x)
            key = f"ev_{cfg}_{p}_{k}"
            out[key + "_f"], out[key + "_g"], out[key + "_c"], out[key + "_dT"] = f, g, c, dT
        r = O.ref_solve(prm, batch, p)
        for name, v in r.items():
            out[f"sv_{cfg}_{p}_{name}"] = v
    np.savez_compressed(os.path.join(GOLD, "optimizer_reference.npz"), **out)


def make_minco():
    R = O.ref_minco_lib()
    assert R is not None, "oracle/_ref/libref_minco.so missing: run `make -C oracle ref` where /root/reference exists"
    out = {}
    for i, case in enumerate(minco_cases()):
        got = O.minco_run(R.ref_minco_run, *case)
        for k, v in got.items():
            out[f"run{i}_{k}"] = v
        for j, t in enumerate(traj_times(case)):
            out[f"traj{i}_{j}"] = O.traj_eval(R.ref_traj_eval, case[0], case[1], case[2], case[3], case[4], t)
    np.savez_compressed(os.path.join(GOLD, "minco_reference_header.npz"), **out)


def make_traj_state():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_traj_state_oracle as T

    outs, gss = [], []
    for c, dT, sg, t, clamp in T._cases():
        a, g = O.traj_state(c, dT, sg, t, clamp, use_ref=True)
        outs.append(a)
        gss.append(g)
    np.savez_compressed(os.path.join(GOLD, "traj_state_reference.npz"), out=np.array(outs), gs=np.array(gss))


def main():
    make_minco()
    make_optimizer()
    make_traj_state()
    if "--reference-only" in sys.argv or "--minco-only" in sys.argv:
        return
    os.makedirs(GOLD, exist_ok=True)
    L = O.lib()
    out = {}
    tf = C.cast(L.oracle_testfunc_fg, C.c_void_p)
    for fid, n in ANALYTIC:
        for si, (mem, past, delta, geps) in enumerate(LB_SETTINGS):
            idv = C.c_int(fid)
            ret, f, x, ev, tr = run_ref(n, analytic_x0(n), tf, C.cast(C.byref(idv), C.c_void_p), mem, past, delta, geps)
            key = f"an_{fid}_{n}_{si}"
            out[key + "_ret"], out[key + "_f"], out[key + "_x"], out[key + "_ev"], out[key + "_tr"] = ret, f, x, ev, tr
    for cfg in (1, 2, 3, 4, 5):
        batch, prm = synth.make_config(cfg, B=3, seed=7)
        cb = batch.as_c()
        for p in range(batch.B):
            h = L.oracle_problem_create(C.byref(prm), C.byref(cb), p)
            n = L.oracle_problem_n(h)
            x0 = np.zeros(n)
            L.oracle_problem_initial_x(h, x0.ctypes.data_as(capi.c_double_p))
            ret, f, x, ev, tr = run_ref(n, x0, C.cast(L.oracle_problem_fg, C.c_void_p), h, prm.lbfgs_mem_size, prm.lbfgs_past,
                                        prm.lbfgs_delta, prm.lbfgs_g_epsilon, prm.lbfgs_max_iterations)
            L.oracle_problem_free(h)
            key = f"pr_{cfg}_{p}"
            out[key + "_ret"], out[key + "_f"], out[key + "_x"], out[key + "_ev"], out[key + "_tr"] = ret, f, x, ev, tr
    np.savez_compressed(os.path.join(GOLD, "lbfgs_reference_header.npz"), **out)

    ev = {}
    rng = np.random.default_rng(2024)
    for cfg in (1, 2, 3, 4, 5):
        batch, prm = synth.make_config(cfg, B=2, seed=9)
        for p in range(batch.B):
            x0 = O.initial_x(prm, batch, p)
            x = x0 + 0.05 * rng.normal(size=x0.shape)
            f, g, c, dT, terms = O.evaluate(prm, batch, p, x, want_terms=True)
            key = f"ev_{cfg}_{p}"
            ev[key + "_x"], ev[key + "_f"], ev[key + "_g"], ev[key + "_c"], ev[key + "_dT"], ev[key + "_terms"] = x, f, g, c, dT, terms
    np.savez_compressed(os.path.join(GOLD, "oracle_eval.npz"), **ev)

    sv = {}
    for cfg in (1, 2, 3, 4, 5):
        batch, prm = synth.make_config(cfg, B=4, seed=13)
        sol = O.optimize(prm, batch, threads=1)
        for k in ("x", "cost", "status", "success", "iters", "evals", "coeffs", "piece_dT"):
            sv[f"sv_{cfg}_{k}"] = getattr(sol, k)
    np.savez_compressed(os.path.join(GOLD, "solve_small.npz"), **sv)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
