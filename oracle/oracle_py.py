"""ctypes loader for the CPU oracle (oracle/libdftpav_oracle.so) and the reference-header build
(oracle/_ref/libref_lbfgs.so).

TEST INFRASTRUCTURE ONLY — import from tests/, __graft_entry__.smoke() and bench.py's CPU legs, never from the
product package ``dftpav_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from dftpav_b200 import capi
from dftpav_b200.batch import TrajBatch

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libdftpav_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref_lbfgs.so")

FG_T = C.CFUNCTYPE(C.c_double, C.c_void_p, capi.c_double_p, capi.c_double_p, C.c_int)

_lib = None
_ref = None


def build(force: bool = False) -> None:
    """Compiles the oracle (and oracle/_ref when /root/reference is present) with oracle/Makefile."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(
            os.path.getmtime(os.path.join(HERE, f)) for f in ("dftpav_oracle.cpp", "corridor_oracle.cpp")):
        subprocess.check_call(["make", "-C", HERE, "libdftpav_oracle.so"])
    if os.path.exists("/root/reference/src/Plan/traj_planner/include/geo_utils2d/lbfgs.hpp") and (
            force or not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-C", HERE, "ref"])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        P, Bt = C.POINTER(capi.Params), C.POINTER(capi.Batch)
        dp, ip = capi.c_double_p, capi.c_int32_p
        L.oracle_eval.argtypes = [P, Bt, C.c_int, dp, dp, dp, dp, dp, dp]
        L.oracle_problem_dims.argtypes = [P, Bt, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_optimize.argtypes = [P, Bt, C.c_int, C.c_int, C.c_int, C.POINTER(capi.Result), dp, C.c_int]
        L.oracle_lbfgs.argtypes = [C.c_int, dp, dp, C.c_void_p, C.c_void_p, P, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   dp, C.c_int]
        L.oracle_minco_matrix.argtypes = [C.c_int, dp, dp]
        _minco_sig(L.oracle_minco_run, L.oracle_traj_eval)
        L.oracle_minco_generate.argtypes = [C.c_int, dp, C.c_double, dp, dp, dp]
        L.oracle_testfunc.argtypes = [C.c_int, dp, dp, C.c_int]
        L.oracle_testfunc.restype = C.c_double
        L.oracle_problem_create.argtypes = [P, Bt, C.c_int]
        L.oracle_problem_create.restype = C.c_void_p
        L.oracle_problem_free.argtypes = [C.c_void_p]
        L.oracle_problem_free.restype = None
        L.oracle_problem_n.argtypes = [C.c_void_p]
        L.oracle_problem_initial_x.argtypes = [C.c_void_p, dp]
        L.oracle_problem_initial_x.restype = None
        L.oracle_problem_fg.restype = C.c_double
        L.oracle_num_threads.argtypes = []
        L.oracle_corridor_rectangles.argtypes = [C.POINTER(capi.GridMap), C.POINTER(capi.CorridorParams), C.c_int64, dp, dp, dp, dp]
        L.oracle_traj_state.argtypes = [C.c_int, dp, C.c_double, C.c_int, C.c_double, C.c_int, C.c_double, dp, dp]
        _lib = L
    return _lib


REF_MINCO_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_minco.so")
_ref_minco = None


def _minco_sig(fn_run, fn_eval):
    dp = capi.c_double_p
    fn_run.argtypes = [C.c_int, dp, C.c_double, dp, dp, dp, C.c_double, dp, dp, dp, dp, dp, dp, dp, dp]
    fn_run.restype = C.c_int
    fn_eval.argtypes = [C.c_int, dp, C.c_double, dp, dp, C.c_int, C.c_double, dp]
    fn_eval.restype = C.c_int


def ref_minco_lib():
    """The reference MINCO header (plan_utils/poly_traj_utils.hpp) compiled in place against oracle/eigen_shim2
    (None if it is not built here and no prebuilt copy travelled)."""
    global _ref_minco
    if _ref_minco is None:
        if not os.path.exists(REF_MINCO_SO):
            try:
                build()
            except Exception:
                return None
        if not os.path.exists(REF_MINCO_SO):
            return None
        try:
            R = C.CDLL(REF_MINCO_SO)
        except OSError:
            return None
        _minco_sig(R.ref_minco_run, R.ref_traj_eval)
        if hasattr(R, "ref_traj_state"):
            R.ref_traj_state.argtypes = [C.c_int, capi.c_double_p, C.c_double, C.c_int, C.c_double, C.c_int,
                                         capi.c_double_p, capi.c_double_p]
        _ref_minco = R
    return _ref_minco


def minco_run(fn, N, P, dT, head, tail, extra_gdC, extra_gdT):
    """Calls ref_minco_run / oracle_minco_run (same signature).  P [N-1][2], head/tail [3][2] (p, v, a rows),
    extra_gdC [6N][2].  Returns a dict of arrays."""
    dp = capi.c_double_p
    P = np.ascontiguousarray(P, dtype=np.float64)
    head = np.ascontiguousarray(head, dtype=np.float64)
    tail = np.ascontiguousarray(tail, dtype=np.float64)
    ex = np.ascontiguousarray(extra_gdC, dtype=np.float64)
    c = np.zeros((6 * N, 2)); gdC0 = np.zeros((6 * N, 2)); gdP = np.zeros((N - 1, 2)); gdH = np.zeros((3, 2)); gdTl = np.zeros((3, 2))
    en = C.c_double(); g0 = C.c_double(); gT = C.c_double()
    rc = fn(N, P.ctypes.data_as(dp), float(dT), head.ctypes.data_as(dp), tail.ctypes.data_as(dp), ex.ctypes.data_as(dp),
            float(extra_gdT), c.ctypes.data_as(dp), C.cast(C.byref(en), dp), gdC0.ctypes.data_as(dp),
            C.cast(C.byref(g0), dp), gdP.ctypes.data_as(dp), gdH.ctypes.data_as(dp), gdTl.ctypes.data_as(dp),
            C.cast(C.byref(gT), dp))
    assert rc == 0
    return dict(c=c, energy=np.array(en.value), gdC0=gdC0, gdT0=np.array(g0.value), gdP=gdP, gdHead=gdH, gdTail=gdTl,
                gdT=np.array(gT.value))


def traj_eval(fn, N, P, dT, head, tail, t):
    dp = capi.c_double_p
    P = np.ascontiguousarray(P, dtype=np.float64)
    head = np.ascontiguousarray(head, dtype=np.float64)
    tail = np.ascontiguousarray(tail, dtype=np.float64)
    out = np.zeros(17)
    rc = fn(N, P.ctypes.data_as(dp), float(dT), head.ctypes.data_as(dp), tail.ctypes.data_as(dp), 1, float(t), out.ctypes.data_as(dp))
    assert rc == 0
    return out


def traj_state(c: np.ndarray, dT: float, singul: int, t: float, clamp: int = 0, wheel_base: float = 2.85,
               use_ref: bool = False):
    """Trajectory / Piece evaluators on coefficients c [6N, 2] (row 6i+k = coefficient of t^k of piece i): returns
    (out[12], getstate[6]) from the oracle's restatement or (use_ref) from the reference's own header (oracle/_ref)."""
    dp = capi.c_double_p
    c = np.ascontiguousarray(c, dtype=np.float64)
    N = c.shape[0] // 6
    out = np.zeros(12)
    gs = np.zeros(6)
    if use_ref:
        R = ref_minco_lib()
        assert R is not None and hasattr(R, "ref_traj_state")
        rc = R.ref_traj_state(N, c.ctypes.data_as(dp), float(dT), int(singul), float(t), int(clamp), out.ctypes.data_as(dp),
                              gs.ctypes.data_as(dp))
    else:
        rc = lib().oracle_traj_state(N, c.ctypes.data_as(dp), float(dT), int(singul), float(t), int(clamp), float(wheel_base),
                                     out.ctypes.data_as(dp), gs.ctypes.data_as(dp))
    assert rc == 0
    return out, gs


def corridor_rectangles(grid: np.ndarray, origin_xy, resolution: float, states: np.ndarray, cos_sin=None, params=None,
                        occupied: int = capi.OCCUPIED):
    """CPU restatement of TrajPlanner::getRectangleConst (oracle/corridor_oracle.cpp): returns (hpolys [n,4,4], expand [n,4])."""
    dp = capi.c_double_p
    grid = np.ascontiguousarray(grid, dtype=np.uint8)
    states = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, 3)
    n = states.shape[0]
    m = capi.GridMap(grid.shape[1], grid.shape[0], float(origin_xy[0]), float(origin_xy[1]), float(resolution),
                     grid.ctypes.data_as(C.POINTER(C.c_uint8)), occupied)
    cp = params if params is not None else capi.default_corridor_params()
    out = np.zeros((n, 4, 4))
    ex = np.zeros((n, 4))
    cs = None if cos_sin is None else np.ascontiguousarray(cos_sin, dtype=np.float64).reshape(n, 2)
    rc = lib().oracle_corridor_rectangles(C.byref(m), C.byref(cp), n, states.ctypes.data_as(dp),
                                          cs.ctypes.data_as(dp) if cs is not None else None, out.ctypes.data_as(dp),
                                          ex.ctypes.data_as(dp))
    assert rc == 0
    return out, ex


def collision_check_segment(grid: np.ndarray, origin_xy, resolution: float, c: np.ndarray, dT: float, singul: int,
                            dt: float = 0.05, vertex_spacing: float = 0.1, max_samples: int = 2048, params=None,
                            occupied: int = capi.OCCUPIED) -> int:
    """CPU restatement of the collision re-check of one gear segment (oracle/corridor_oracle.cpp): index of the first
    colliding sample, -1 if none.  c = [6N, 2] coefficients, lowest power first (the solver's output layout)."""
    grid = np.ascontiguousarray(grid, dtype=np.uint8)
    c = np.ascontiguousarray(c, dtype=np.float64)
    N = c.shape[0] // 6
    m = capi.GridMap(grid.shape[1], grid.shape[0], float(origin_xy[0]), float(origin_xy[1]), float(resolution),
                     grid.ctypes.data_as(C.POINTER(C.c_uint8)), occupied)
    cp = params if params is not None else capi.default_corridor_params()
    fn = lib().oracle_collision_check_segment
    fn.argtypes = [C.POINTER(capi.GridMap), C.POINTER(capi.CorridorParams), C.c_int, capi.c_double_p, C.c_double, C.c_int,
                   C.c_double, C.c_double, C.c_int]
    fn.restype = C.c_int
    return int(fn(C.byref(m), C.byref(cp), N, c.ctypes.data_as(capi.c_double_p), float(dT), int(singul), float(dt),
                  float(vertex_spacing), int(max_samples)))


def trapezoid_duration(length, max_vel, max_acc, v0, v1) -> float:
    fn = lib().oracle_trapezoid_duration
    fn.argtypes = [C.c_double] * 5
    fn.restype = C.c_double
    return float(fn(length, max_vel, max_acc, v0, v1))


def trapezoid_length(curt, locallength, max_vel, max_acc, v0, v1) -> float:
    fn = lib().oracle_trapezoid_length
    fn.argtypes = [C.c_double] * 6
    fn.restype = C.c_double
    return float(fn(curt, locallength, max_vel, max_acc, v0, v1))


def frontend_evaluate_pos(fe, path: np.ndarray, t: np.ndarray) -> np.ndarray:
    """CPU restatement of KinoAstar::evaluatePos (oracle/frontend_oracle.cpp)."""
    path = np.ascontiguousarray(path, dtype=np.int32)
    t = np.ascontiguousarray(t, dtype=np.float64)
    out = np.zeros((len(t), 3))
    c = fe.as_c()
    fn = lib().oracle_frontend_evaluate_pos
    fn.argtypes = [C.POINTER(capi.Frontend), C.c_int64, capi.c_int32_p, capi.c_double_p, capi.c_double_p]
    assert fn(C.byref(c), len(t), path.ctypes.data_as(capi.c_int32_p), t.ctypes.data_as(capi.c_double_p),
              out.ctypes.data_as(capi.c_double_p)) == 0
    return out


def frontend_resample(fe, seg_duration: np.ndarray, traj_piece_duration: float, K: int, Kd: int):
    """CPU restatement of the resampling loop of TrajPlanner::RunMINCOParking: (piece_num, piece_dur, states, inner)."""
    seg_duration = np.ascontiguousarray(seg_duration, dtype=np.float64)
    SH = len(seg_duration)
    c = fe.as_c()
    fn = lib().oracle_frontend_resample
    dp = capi.c_double_p
    fn.argtypes = [C.POINTER(capi.Frontend), dp, C.c_double, C.c_int, C.c_int, capi.c_int32_p, dp, dp, dp]
    fn.restype = C.c_int64
    pn = np.zeros(SH, dtype=np.int32)
    pd = np.zeros(SH)
    n = fn(C.byref(c), seg_duration.ctypes.data_as(dp), traj_piece_duration, K, Kd, pn.ctypes.data_as(capi.c_int32_p),
           pd.ctypes.data_as(dp), None, None)
    states = np.zeros((n, 3))
    inner = np.zeros((int((pn - 1).sum()), 2))
    fn(C.byref(c), seg_duration.ctypes.data_as(dp), traj_piece_duration, K, Kd, pn.ctypes.data_as(capi.c_int32_p),
       pd.ctypes.data_as(dp), states.ctypes.data_as(dp), inner.ctypes.data_as(dp))
    return pn, pd, states, inner


REF_OPT_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_optimizer.so")
_ref_opt = None


def ref_optimizer_lib():
    """The reference's own optimizer (src/traj_optimizer.cpp + headers) compiled in place against oracle/eigen_shim2 and
    oracle/ref_stubs (None if it is not built here and no prebuilt copy travelled)."""
    global _ref_opt
    if _ref_opt is None:
        if not os.path.exists(REF_OPT_SO):
            try:
                build()
            except Exception:
                return None
        if not os.path.exists(REF_OPT_SO):
            return None
        try:
            R = C.CDLL(REF_OPT_SO)
        except OSError:
            return None
        dp, ip = capi.c_double_p, C.POINTER(C.c_int)
        P, Bt = C.POINTER(capi.Params), C.POINTER(capi.Batch)
        R.ref_opt_solve.argtypes = [P, Bt, C.c_int, dp, dp, ip, ip, ip, ip, dp, dp]
        R.ref_opt_eval.argtypes = [P, Bt, C.c_int, dp, dp, dp, dp, dp]
        R.ref_opt_solve.restype = R.ref_opt_eval.restype = C.c_int
        R.ref_opt_solve_batch.argtypes = [P, Bt, C.c_int, C.c_int, C.c_int, dp, ip, ip, ip, ip]
        R.ref_opt_solve_batch.restype = C.c_int
        _ref_opt = R
    return _ref_opt


def ref_evaluate(prm: capi.Params, batch: TrajBatch, p: int, x: np.ndarray):
    """The reference's costFunctionCallback at x -> (f, g, coeffs [rows][2], piece_dT [S])."""
    R = ref_optimizer_lib()
    n, rows, _ = problem_dims(prm, batch, p)
    S = int(batch.seg_ptr[p + 1] - batch.seg_ptr[p])
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert x.size == n
    g = np.zeros(n); c = np.zeros((rows, 2)); dT = np.zeros(S); f = C.c_double()
    cb = batch.as_c()
    dp = capi.c_double_p
    rc = R.ref_opt_eval(C.byref(prm), C.byref(cb), p, x.ctypes.data_as(dp), C.cast(C.byref(f), dp), g.ctypes.data_as(dp),
                        c.ctypes.data_as(dp), dT.ctypes.data_as(dp))
    if rc != 0:
        raise RuntimeError(f"ref_opt_eval failed: {rc}")
    return f.value, g, c, dT


def ref_solve(prm: capi.Params, batch: TrajBatch, p: int):
    """One full PolyTrajOptimizer::OptimizeTrajectory of the reference -> dict(x, cost, status, success, iters, evals,
    coeffs, piece_dT)."""
    R = ref_optimizer_lib()
    n, rows, _ = problem_dims(prm, batch, p)
    S = int(batch.seg_ptr[p + 1] - batch.seg_ptr[p])
    x = np.zeros(n); c = np.zeros((rows, 2)); dT = np.zeros(S)
    cost = C.c_double(); st = C.c_int(); ok = C.c_int(); it = C.c_int(); ev = C.c_int()
    cb = batch.as_c()
    dp = capi.c_double_p
    rc = R.ref_opt_solve(C.byref(prm), C.byref(cb), p, x.ctypes.data_as(dp), C.cast(C.byref(cost), dp), C.byref(st),
                         C.byref(ok), C.byref(it), C.byref(ev), c.ctypes.data_as(dp), dT.ctypes.data_as(dp))
    if rc != 0:
        raise RuntimeError(f"ref_opt_solve failed: {rc}")
    return dict(x=x, cost=cost.value, status=st.value, success=ok.value, iters=it.value, evals=ev.value, coeffs=c, piece_dT=dT)


def ref_solve_batch(prm: capi.Params, batch: TrajBatch, p0: int, p1: int, threads: int = 0):
    """Complete reference solves of problems p0..p1-1, OpenMP over problems -> dict of [p1-p0] arrays."""
    R = ref_optimizer_lib()
    m = p1 - p0
    cost = np.zeros(m); st = np.zeros(m, dtype=np.int32); ok = np.zeros(m, dtype=np.int32)
    it = np.zeros(m, dtype=np.int32); ev = np.zeros(m, dtype=np.int32)
    cb = batch.as_c()
    ip = C.POINTER(C.c_int)
    rc = R.ref_opt_solve_batch(C.byref(prm), C.byref(cb), p0, p1, threads, cost.ctypes.data_as(capi.c_double_p),
                               st.ctypes.data_as(ip), ok.ctypes.data_as(ip), it.ctypes.data_as(ip), ev.ctypes.data_as(ip))
    if rc != 0:
        raise RuntimeError(f"ref_opt_solve_batch failed: {rc}")
    return dict(cost=cost, status=st, success=ok, iters=it, evals=ev)


def ref_lib():
    """The reference lbfgs.hpp compiled in place (None if it could not be built and no prebuilt copy travelled)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            try:
                build()
            except Exception:
                return None
        if not os.path.exists(REF_SO):
            return None
        R = C.CDLL(REF_SO)
        dp = capi.c_double_p
        R.ref_lbfgs.argtypes = [C.c_int, dp, dp, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int), dp, C.c_int]
        _ref = R
    return _ref


def _dp(a: np.ndarray):
    return a.ctypes.data_as(capi.c_double_p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(capi.c_int32_p)


def problem_dims(prm: capi.Params, batch: TrajBatch, p: int):
    n, r, s = C.c_int(), C.c_int(), C.c_int()
    cb = batch.as_c()
    rc = lib().oracle_problem_dims(C.byref(prm), C.byref(cb), p, C.byref(n), C.byref(r), C.byref(s))
    assert rc == 0, rc
    return n.value, r.value, s.value


def evaluate(prm: capi.Params, batch: TrajBatch, p: int, x: np.ndarray, want_terms: bool = False):
    """One costFunctionCallback: returns (f, g, coeffs[rows,2], piece_dT[S]) (+ terms[5])."""
    n, rows, _ = problem_dims(prm, batch, p)
    x = np.ascontiguousarray(x[:n], dtype=np.float64)
    g = np.zeros(n)
    f = C.c_double()
    S = int(batch.seg_ptr[p + 1] - batch.seg_ptr[p])
    coeffs = np.zeros((rows, 2))
    dT = np.zeros(S)
    terms = np.zeros(5)
    cb = batch.as_c()
    rc = lib().oracle_eval(C.byref(prm), C.byref(cb), p, _dp(x), C.cast(C.byref(f), capi.c_double_p), _dp(g),
                           _dp(coeffs), _dp(dT), _dp(terms))
    if rc != 0:
        raise RuntimeError(f"oracle_eval rc={rc}")
    if want_terms:
        return f.value, g, coeffs, dT, terms
    return f.value, g, coeffs, dT


def initial_x(prm: capi.Params, batch: TrajBatch, p: int) -> np.ndarray:
    cb = batch.as_c()
    h = lib().oracle_problem_create(C.byref(prm), C.byref(cb), p)
    if not h:
        raise RuntimeError("oracle_problem_create failed (bad input)")
    n = lib().oracle_problem_n(h)
    x = np.zeros(n)
    lib().oracle_problem_initial_x(h, _dp(x))
    lib().oracle_problem_free(h)
    return x


class Solution:
    """Outputs of a batch solve, same layout as dftpav_result."""

    def __init__(self, B, n_stride, row_stride, G):
        self.x = np.zeros((B, n_stride))
        self.cost = np.zeros(B)
        self.status = np.zeros(B, dtype=np.int32)
        self.success = np.zeros(B, dtype=np.int32)
        self.iters = np.zeros(B, dtype=np.int32)
        self.evals = np.zeros(B, dtype=np.int32)
        self.coeffs = np.zeros((B, row_stride, 2))
        self.piece_dT = np.zeros(G)

    def as_c(self) -> capi.Result:
        r = capi.Result()
        r.x, r.cost, r.coeffs, r.piece_dT = _dp(self.x), _dp(self.cost), _dp(self.coeffs), _dp(self.piece_dT)
        r.status, r.success, r.iters, r.evals = _ip(self.status), _ip(self.success), _ip(self.iters), _ip(self.evals)
        return r


def optimize(prm: capi.Params, batch: TrajBatch, p0: int = 0, p1: int | None = None, threads: int = 0,
             trace_cap: int = 0):
    """Full OptimizeTrajectory for problems [p0,p1) on `threads` host threads (0 = all)."""
    if p1 is None:
        p1 = batch.B
    sol = Solution(batch.B, batch.n_stride(), batch.row_stride(), batch.G)
    trace = np.zeros((batch.B, trace_cap, 3)) if trace_cap else None
    cb = batch.as_c()
    res = sol.as_c()
    rc = lib().oracle_optimize(C.byref(prm), C.byref(cb), p0, p1, threads, C.byref(res),
                               _dp(trace) if trace is not None else None, trace_cap)
    if rc != 0:
        raise RuntimeError(f"oracle_optimize rc={rc}")
    sol.trace = trace
    return sol
