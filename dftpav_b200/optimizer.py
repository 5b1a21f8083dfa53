"""Python host binding over the C-ABI (include/dftpav_b200.h) — thin, no arithmetic of its own.

``BatchOptimizer`` mirrors the call sequence a planner makes against the reference
``plan_manage::PolyTrajOptimizer`` (traj_optimizer.h:100-120):

    reference                               this binding                     C-ABI
    ---------------------------------------------------------------------------------------------
    new PolyTrajOptimizer; setParam(cfg)    BatchOptimizer(params)           dftpav_create
    setSurroundTrajs(&trajs)                (obstacles are part of the batch) dftpav_upload_batch
    OptimizeTrajectory(ini, fin, P, T, ...) solve(batch) / optimize()        dftpav_optimize
    getMinJerkOptPtr()->at(i).getTraj(s)    download() -> Solution.coeffs    dftpav_download

The C++ adapter with the reference's exact signatures lives in dftpav_b200/host/poly_traj_optimizer.hpp.
All failures raise; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import capi
from .batch import TrajBatch


class DftpavError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dftpav error {code}: {msg}")
        self.code = code


class Solution:
    """Host copies of dftpav_result (same strides as dftpav_dims)."""

    def __init__(self, B: int, n_stride: int, row_stride: int, G: int, pinned: bool = False):
        def alloc(shape, dtype):
            if pinned:
                import torch

                t = torch.empty(shape, dtype={np.float64: torch.float64, np.int32: torch.int32}[dtype]).pin_memory()
                self._keep.append(t)
                return t.numpy()
            return np.zeros(shape, dtype=dtype)

        self._keep = []
        self.x = alloc((B, n_stride), np.float64)
        self.cost = alloc((B,), np.float64)
        self.status = alloc((B,), np.int32)
        self.success = alloc((B,), np.int32)
        self.iters = alloc((B,), np.int32)
        self.evals = alloc((B,), np.int32)
        self.coeffs = alloc((B, row_stride, 2), np.float64)
        self.piece_dT = alloc((G,), np.float64)

    def as_c(self) -> capi.Result:
        r = capi.Result()
        dp, ip = capi.c_double_p, capi.c_int32_p
        r.x = self.x.ctypes.data_as(dp)
        r.cost = self.cost.ctypes.data_as(dp)
        r.status = self.status.ctypes.data_as(ip)
        r.success = self.success.ctypes.data_as(ip)
        r.iters = self.iters.ctypes.data_as(ip)
        r.evals = self.evals.ctypes.data_as(ip)
        r.coeffs = self.coeffs.ctypes.data_as(dp)
        r.piece_dT = self.piece_dT.ctypes.data_as(dp)
        return r

    def nbytes(self) -> int:
        return sum(a.nbytes for a in (self.x, self.cost, self.status, self.success, self.iters, self.evals, self.coeffs,
                                      self.piece_dT))


class BatchOptimizer:
    def __init__(self, params: Optional[capi.Params] = None, device: int = -1, variant: str = ""):
        self._lib = capi.load_library(variant)
        self.params = params if params is not None else capi.default_params()
        self._h = C.c_void_p()
        rc = self._lib.dftpav_create(C.byref(self.params), device, C.byref(self._h))
        if rc != capi.OK:
            raise DftpavError(rc, "dftpav_create failed (no sm_100 device? there is no CPU fallback)")
        self._batch = None
        self._dims = None

    # -- plumbing ------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != capi.OK:
            raise DftpavError(rc, self._lib.dftpav_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._lib.dftpav_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self._lib.dftpav_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def stats(self) -> capi.Stats:
        s = capi.Stats()
        self._check(self._lib.dftpav_last_stats(self._h, C.byref(s)))
        return s

    # -- the call sequence ---------------------------------------------------------------------
    def upload(self, batch: TrajBatch) -> capi.Dims:
        self._batch = batch
        cb = batch.as_c()
        self._check(self._lib.dftpav_upload_batch(self._h, C.byref(cb)))
        d = capi.Dims()
        self._check(self._lib.dftpav_get_dims(self._h, C.byref(d)))
        self._dims = d
        return d

    def optimize(self):
        self._check(self._lib.dftpav_optimize(self._h))

    def new_solution(self, pinned: bool = False) -> Solution:
        d = self._dims
        return Solution(d.B, d.n_stride, d.row_stride, d.seg_total, pinned=pinned)

    def download(self, sol: Optional[Solution] = None) -> Solution:
        if sol is None:
            sol = self.new_solution()
        res = sol.as_c()
        self._check(self._lib.dftpav_download(self._h, C.byref(res)))
        return sol

    def solve(self, batch: TrajBatch, sol: Optional[Solution] = None) -> Solution:
        """The end-to-end planner call (dftpav_solve): upload, solve and read-back in ONE C-ABI call; the corridor planes are
        copied in chunks while the solver kernels already run."""
        self._batch = batch
        cb = batch.as_c()
        if sol is None:
            d = capi.Dims()
            self._check(self._lib.dftpav_batch_dims(C.byref(self.params), C.byref(cb), C.byref(d)))
            sol = Solution(d.B, d.n_stride, d.row_stride, d.seg_total)
        res = sol.as_c()
        self._check(self._lib.dftpav_solve(self._h, C.byref(cb), C.byref(res)))
        d = capi.Dims()
        self._check(self._lib.dftpav_get_dims(self._h, C.byref(d)))
        self._dims = d
        return sol

    def solve_stepwise(self, batch: TrajBatch, sol: Optional[Solution] = None) -> Solution:
        """upload, optimize, download as three C-ABI calls (the batch stays resident for further optimize / evaluate calls)."""
        self.upload(batch)
        self.optimize()
        return self.download(sol)

    def success_count(self) -> int:
        """Problems of the last solve that ended with flag_success (dftpav_success_count)."""
        n = C.c_int32(0)
        self._check(self._lib.dftpav_success_count(self._h, C.byref(n)))
        return int(n.value)

    def force_banded(self, on: bool = True):
        """Test hook: banded LU substitutions (the large-N fallback) even where the dense MINCO tables fit."""
        self._check(self._lib.dftpav_debug_force_banded(self._h, 1 if on else 0))

    def set_order(self, order: Optional[np.ndarray]):
        """Experiment hook: serve the work queue in this order of problem ids (None: input order)."""
        self._lib.dftpav_debug_set_order.argtypes = [C.c_void_p, capi.c_int32_p, C.c_int32]
        if order is None:
            self._check(self._lib.dftpav_debug_set_order(self._h, None, 0))
        else:
            o = np.ascontiguousarray(order, dtype=np.int32)
            self._check(self._lib.dftpav_debug_set_order(self._h, o.ctypes.data_as(capi.c_int32_p), o.shape[0]))

    def build_flags(self) -> int:
        return int(self._lib.dftpav_debug_build_flags())

    def evaluate(self, x: np.ndarray):
        """One cost+gradient evaluation per problem at x [B, n_stride] (dftpav_eval)."""
        d = self._dims
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(d.B, d.n_stride)
        f = np.zeros(d.B)
        g = np.zeros((d.B, d.n_stride))
        dp = capi.c_double_p
        self._check(self._lib.dftpav_eval(self._h, x.ctypes.data_as(dp), f.ctypes.data_as(dp), g.ctypes.data_as(dp)))
        return f, g

    # -- read-back: batched trajectory evaluation (Trajectory::getPos / getAngle / GetState ... on the device) ------------
    def sample(self, seg: np.ndarray, t: np.ndarray, flags: int = 0, wheel_base: float = capi.WHEEL_BASE) -> np.ndarray:
        """States [nq, 12] of segment seg[q] at local time t[q] (dftpav_sample)."""
        seg = np.ascontiguousarray(seg, dtype=np.int32)
        t = np.ascontiguousarray(t, dtype=np.float64)
        out = np.zeros((seg.shape[0], capi.STATE_DOUBLES))
        self._check(self._lib.dftpav_sample(self._h, seg.shape[0], seg.ctypes.data_as(capi.c_int32_p),
                                            t.ctypes.data_as(capi.c_double_p), flags, wheel_base,
                                            out.ctypes.data_as(capi.c_double_p)))
        return out

    def sample_uniform(self, dt: float, max_per_seg: int):
        """(x, y, yaw) of every segment at t = 0, dt, 2 dt, ... < duration (dftpav_sample_uniform): ([G, max, 3], counts [G])."""
        G = self._dims.seg_total
        out = np.zeros((G, max_per_seg, 3))
        cnt = np.zeros(G, dtype=np.int32)
        self._check(self._lib.dftpav_sample_uniform(self._h, dt, max_per_seg, out.ctypes.data_as(capi.c_double_p),
                                                    cnt.ctypes.data_as(capi.c_int32_p)))
        return out, cnt

    def pack_polytraj(self):
        """msg/PolyTraj.msg fields of all segments: (coef_x [P, 6], coef_y [P, 6], duration [P]) float32, highest power first."""
        P = int(np.sum(self._batch.piece_num))
        cx = np.zeros((P, 6), dtype=np.float32)
        cy = np.zeros((P, 6), dtype=np.float32)
        du = np.zeros(P, dtype=np.float32)
        fp = C.POINTER(C.c_float)
        self._check(self._lib.dftpav_pack_polytraj(self._h, cx.ctypes.data_as(fp), cy.ctypes.data_as(fp), du.ctypes.data_as(fp)))
        return cx, cy, du

    # -- batched MINCO fits (ConverSurroundTrajFromPoints / initial guess) ----------------------------------------------------
    def minco_fit(self, piece_ptr: np.ndarray, dT: np.ndarray, head: np.ndarray, tail: np.ndarray, inner_pts: np.ndarray):
        """coeffs [sum N, 12] (Piece::coeffMat, highest power first) of n fits (dftpav_minco_fit)."""
        pp = np.ascontiguousarray(piece_ptr, dtype=np.int32)
        n = pp.shape[0] - 1
        dT = np.ascontiguousarray(dT, dtype=np.float64)
        head = np.ascontiguousarray(head, dtype=np.float64).reshape(n, 6)
        tail = np.ascontiguousarray(tail, dtype=np.float64).reshape(n, 6)
        inner = np.ascontiguousarray(inner_pts, dtype=np.float64).reshape(-1)
        out = np.zeros((int(pp[-1]), 12))
        dp = capi.c_double_p
        self._check(self._lib.dftpav_minco_fit(self._h, n, pp.ctypes.data_as(capi.c_int32_p), dT.ctypes.data_as(dp),
                                               head.ctypes.data_as(dp), tail.ctypes.data_as(dp), inner.ctypes.data_as(dp),
                                               out.ctypes.data_as(dp)))
        return out

    # -- corridor generation (TrajPlanner::getRectangleConst on the device) ----------------------------------------------
    def set_map(self, grid: np.ndarray, origin_xy, resolution: float, occupied: int = capi.OCCUPIED):
        """grid: uint8 [size_y, size_x] (dimension 0 of the reference's GridMapND is the fastest-varying index)."""
        grid = np.ascontiguousarray(grid, dtype=np.uint8)
        m = capi.GridMap(grid.shape[1], grid.shape[0], float(origin_xy[0]), float(origin_xy[1]), float(resolution),
                         grid.ctypes.data_as(C.POINTER(C.c_uint8)), occupied)
        self._check(self._lib.dftpav_corridor_set_map(self._h, C.byref(m)))

    def corridor_rectangles(self, states: np.ndarray, cos_sin: Optional[np.ndarray] = None, keep: bool = False,
                            params: Optional[capi.CorridorParams] = None, want_host: bool = True):
        """states [n, 3] = (x, y, yaw) -> half-planes [n, 4, 4] (n_x, n_y, p_x, p_y) (dftpav_corridor_rectangles)."""
        states = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, 3)
        n = states.shape[0]
        cp = params if params is not None else capi.default_corridor_params()
        out = np.zeros((n, 4, 4)) if want_host else None
        cs = None if cos_sin is None else np.ascontiguousarray(cos_sin, dtype=np.float64).reshape(n, 2)
        dp = capi.c_double_p
        self._check(self._lib.dftpav_corridor_rectangles(
            self._h, C.byref(cp), n, states.ctypes.data_as(dp), cs.ctypes.data_as(dp) if cs is not None else None,
            out.ctypes.data_as(dp) if out is not None else None, 1 if keep else 0))
        return out

    def frontend_evaluate_pos(self, fe, path: np.ndarray, t: np.ndarray) -> np.ndarray:
        """KinoAstar::evaluatePos for queries (path index, time) on the front-end paths `fe` (synth.FrontendPaths)."""
        path = np.ascontiguousarray(path, dtype=np.int32)
        t = np.ascontiguousarray(t, dtype=np.float64)
        out = np.zeros((len(t), 3))
        c = fe.as_c()
        self._check(self._lib.dftpav_frontend_evaluate_pos(self._h, C.byref(c), len(t), path.ctypes.data_as(capi.c_int32_p),
                                                           t.ctypes.data_as(capi.c_double_p), out.ctypes.data_as(capi.c_double_p)))
        return out

    def frontend_resample(self, fe, seg_duration: np.ndarray, traj_piece_duration: float, K: int, Kd: int):
        """The resampling loop of RunMINCOParking: returns (piece_num [shots], piece_dur [shots], states [n, 3], inner [m, 2])."""
        seg_duration = np.ascontiguousarray(seg_duration, dtype=np.float64)
        SH = len(seg_duration)
        pn = np.zeros(SH, dtype=np.int32)
        pd = np.zeros(SH)
        pn_host = np.maximum((seg_duration / traj_piece_duration + 0.5).astype(np.int64), 2)
        cap = int((2 * (Kd + 1) + (pn_host - 2) * (K + 1)).sum())
        states = np.zeros((cap, 3))
        inner = np.zeros((int((pn_host - 1).sum()), 2))
        n = C.c_int64(0)
        c = fe.as_c()
        dp = capi.c_double_p
        self._check(self._lib.dftpav_frontend_resample(self._h, C.byref(c), seg_duration.ctypes.data_as(dp), traj_piece_duration,
                                                       K, Kd, pn.ctypes.data_as(capi.c_int32_p), pd.ctypes.data_as(dp), cap,
                                                       states.ctypes.data_as(dp), inner.ctypes.data_as(dp), C.byref(n)))
        return pn, pd, states[: n.value], inner

    def collision_check(self, dt: float = 0.05, vertex_spacing: float = 0.1, max_samples: int = 2048,
                        params: Optional[capi.CorridorParams] = None):
        """Collision re-check of the last solve's trajectories against the map of set_map (dftpav_collision_check):
        returns (collide [B] int32, first_hit [seg_total] int32)."""
        cp = params if params is not None else capi.default_corridor_params()
        col = np.zeros(self._dims.B, dtype=np.int32)
        first = np.zeros(self._dims.seg_total, dtype=np.int32)
        self._check(self._lib.dftpav_collision_check(self._h, C.byref(cp), dt, vertex_spacing, max_samples,
                                                     col.ctypes.data_as(capi.c_int32_p), first.ctypes.data_as(capi.c_int32_p)))
        return col, first

    # -- diagnostics (not part of the drop-in ABI) ---------------------------------------------
    def enable_trace(self, cap: int):
        self._lib.dftpav_debug_set_trace.argtypes = [C.c_void_p, C.c_int]
        self._check(self._lib.dftpav_debug_set_trace(self._h, cap))
        self._trace_cap = cap

    def get_trace(self) -> np.ndarray:
        out = np.zeros((self._dims.B, self._trace_cap, 3))
        self._lib.dftpav_debug_get_trace.argtypes = [C.c_void_p, capi.c_double_p]
        self._check(self._lib.dftpav_debug_get_trace(self._h, out.ctypes.data_as(capi.c_double_p)))
        return out

    def initial_x(self, batch: TrajBatch) -> np.ndarray:
        d = capi.Dims()
        cb = batch.as_c()
        self._check(self._lib.dftpav_batch_dims(C.byref(self.params), C.byref(cb), C.byref(d)))
        x0 = np.zeros((d.B, d.n_stride))
        self._check(self._lib.dftpav_initial_x(C.byref(self.params), C.byref(cb), x0.ctypes.data_as(capi.c_double_p)))
        return x0


class MultiGpuOptimizer:
    """dftpav_create_multi / dftpav_solve_multi: contiguous shards over several GPUs of one box, one host thread and handle
    per device inside the library, NCCL all-reduce of the convergence flag (the C++ caller's multi-GPU entry point)."""

    def __init__(self, params: Optional[capi.Params] = None, devices=(0,), variant: str = ""):
        self._lib = capi.load_library(variant)
        self.params = params if params is not None else capi.default_params()
        self._m = C.c_void_p()
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        rc = self._lib.dftpav_create_multi(C.byref(self.params), devs.ctypes.data_as(capi.c_int32_p), devs.shape[0],
                                           C.byref(self._m))
        if rc != capi.OK:
            raise DftpavError(rc, "dftpav_create_multi failed")
        self.n_dev = int(self._lib.dftpav_multi_devices(self._m))

    def uses_nccl(self) -> bool:
        return bool(self._lib.dftpav_multi_uses_nccl(self._m))

    def solve(self, batch: TrajBatch, sol: Optional[Solution] = None):
        """Returns (solution, all_accepted)."""
        cb = batch.as_c()
        if sol is None:
            d = capi.Dims()
            rc = self._lib.dftpav_batch_dims(C.byref(self.params), C.byref(cb), C.byref(d))
            if rc != capi.OK:
                raise DftpavError(rc, "invalid batch")
            sol = Solution(d.B, d.n_stride, d.row_stride, d.seg_total)
        res = sol.as_c()
        flag = C.c_int32(0)
        rc = self._lib.dftpav_solve_multi(self._m, C.byref(cb), C.byref(res), C.byref(flag))
        if rc != capi.OK:
            raise DftpavError(rc, self._lib.dftpav_multi_last_error(self._m).decode())
        return sol, bool(flag.value)

    def stats(self, dev_index: int) -> capi.Stats:
        s = capi.Stats()
        self._lib.dftpav_multi_stats(self._m, dev_index, C.byref(s))
        return s

    def close(self):
        if self._m:
            self._lib.dftpav_destroy_multi(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
