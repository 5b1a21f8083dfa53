"""Host-side container for a batch of OptimizeTrajectory problems in the flat layout of ``dftpav_batch``
(include/dftpav_b200.h).  One problem = the argument list of ``PolyTrajOptimizer::OptimizeTrajectory``
(reference traj_optimizer.h:118-120) plus the obstacle set of ``setSurroundTrajs`` (traj_optimizer.h:108).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import capi


def samples_per_segment(N: int, K: int, Kd: int) -> int:
    """Corridor polygons a segment of N pieces must carry (size check traj_optimizer.cpp:44)."""
    return (N - 2) * (K + 1) + 2 * (Kd + 1)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


@dataclass
class TrajBatch:
    """Flat arrays, all C-contiguous; see struct dftpav_batch for the meaning of every field."""

    seg_ptr: np.ndarray            # int32 [B+1]
    piece_num: np.ndarray          # int32 [G]
    singul: np.ndarray             # int32 [G]
    ini_state: np.ndarray          # f64 [G,6]  (px,py,vx,vy,ax,ay) = 2x3 col-major
    fin_state: np.ndarray          # f64 [G,6]
    init_T: np.ndarray             # f64 [G]
    inner_pts: np.ndarray          # f64 [sum 2(N-1)]
    planes: np.ndarray             # f64 [sum H, 4]
    t_now: np.ndarray              # f64 [B]
    uniform_H: int = 0
    plane_cnt: Optional[np.ndarray] = None      # int32 [Q] or None
    obs_ptr: Optional[np.ndarray] = None        # int32 [B+1] or None (no obstacles)
    obs_idx: Optional[np.ndarray] = None
    obs_start_time: Optional[np.ndarray] = None
    obs_duration: Optional[np.ndarray] = None
    obs_piece_ptr: Optional[np.ndarray] = None
    obs_piece_dur: Optional[np.ndarray] = None
    obs_coeffs: Optional[np.ndarray] = None     # f64 [pieces,12]
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        self.seg_ptr = _i32(self.seg_ptr)
        self.piece_num = _i32(self.piece_num)
        self.singul = _i32(self.singul)
        self.ini_state = _f64(self.ini_state).reshape(-1, 6)
        self.fin_state = _f64(self.fin_state).reshape(-1, 6)
        self.init_T = _f64(self.init_T)
        self.inner_pts = _f64(self.inner_pts).reshape(-1)
        if self.planes is not None:  # None: the corridor is generated on the device (BatchOptimizer.corridor_rectangles(keep=True))
            self.planes = _f64(self.planes).reshape(-1, 4)
        self.t_now = _f64(self.t_now)
        if self.plane_cnt is not None:
            self.plane_cnt = _i32(self.plane_cnt)
        if self.obs_ptr is not None:
            self.obs_ptr = _i32(self.obs_ptr)
            self.obs_idx = _i32(self.obs_idx)
            self.obs_start_time = _f64(self.obs_start_time)
            self.obs_duration = _f64(self.obs_duration)
            self.obs_piece_ptr = _i32(self.obs_piece_ptr)
            self.obs_piece_dur = _f64(self.obs_piece_dur)
            self.obs_coeffs = _f64(self.obs_coeffs).reshape(-1, 12)

    # ------------------------------------------------------------------ sizes
    @property
    def B(self) -> int:
        return int(self.seg_ptr.shape[0] - 1)

    @property
    def G(self) -> int:
        return int(self.piece_num.shape[0])

    def n_of(self, p: int) -> int:
        g0, g1 = int(self.seg_ptr[p]), int(self.seg_ptr[p + 1])
        S = g1 - g0
        return int(2 * (self.piece_num[g0:g1] - 1).sum() + S + 3 * (S - 1))

    def rows_of(self, p: int) -> int:
        g0, g1 = int(self.seg_ptr[p]), int(self.seg_ptr[p + 1])
        return int(6 * self.piece_num[g0:g1].sum())

    def n_stride(self) -> int:
        return max(self.n_of(p) for p in range(self.B)) if self.B else 0

    def row_stride(self) -> int:
        return max(self.rows_of(p) for p in range(self.B)) if self.B else 0

    def total_samples(self, K: int, Kd: int) -> int:
        return int(sum(samples_per_segment(int(N), K, Kd) for N in self.piece_num))

    def nbytes(self) -> int:
        tot = 0
        for name in ("seg_ptr", "piece_num", "singul", "ini_state", "fin_state", "init_T", "inner_pts", "planes",
                     "t_now", "plane_cnt", "obs_ptr", "obs_idx", "obs_start_time", "obs_duration", "obs_piece_ptr",
                     "obs_piece_dur", "obs_coeffs"):
            a = getattr(self, name)
            if a is not None:
                tot += a.nbytes
        return tot

    # ------------------------------------------------------------------ C view
    def as_c(self) -> capi.Batch:
        """ctypes view; the TrajBatch must outlive the returned struct."""
        b = capi.Batch()
        b.B = self.B
        b.uniform_H = int(self.uniform_H)

        def dp(a):
            return None if a is None else a.ctypes.data_as(capi.c_double_p)

        def ip(a):
            return None if a is None else a.ctypes.data_as(capi.c_int32_p)

        b.seg_ptr = ip(self.seg_ptr)
        b.piece_num = ip(self.piece_num)
        b.singul = ip(self.singul)
        b.ini_state = dp(self.ini_state)
        b.fin_state = dp(self.fin_state)
        b.init_T = dp(self.init_T)
        b.inner_pts = dp(self.inner_pts)
        b.plane_cnt = ip(self.plane_cnt)
        b.planes = dp(self.planes)
        b.t_now = dp(self.t_now)
        b.obs_ptr = ip(self.obs_ptr)
        b.obs_idx = ip(self.obs_idx)
        b.n_obs_pool = 0 if self.obs_start_time is None else int(self.obs_start_time.shape[0])
        b.obs_start_time = dp(self.obs_start_time)
        b.obs_duration = dp(self.obs_duration)
        b.obs_piece_ptr = ip(self.obs_piece_ptr)
        b.obs_piece_dur = dp(self.obs_piece_dur)
        b.obs_coeffs = dp(self.obs_coeffs)
        return b

    # ------------------------------------------------------------------ slicing (for sharding over ranks)
    def slice(self, p0: int, p1: int, K: int, Kd: int) -> "TrajBatch":
        """Problems [p0, p1) as an independent batch (contiguous shard)."""
        g0, g1 = int(self.seg_ptr[p0]), int(self.seg_ptr[p1])
        Ns = self.piece_num
        inner_pref = np.concatenate([[0], np.cumsum(2 * (Ns - 1))])
        samp = np.array([samples_per_segment(int(N), K, Kd) for N in Ns], dtype=np.int64)
        samp_pref = np.concatenate([[0], np.cumsum(samp)])
        q0, q1 = int(samp_pref[g0]), int(samp_pref[g1])
        if self.plane_cnt is not None:
            pl_pref = np.concatenate([[0], np.cumsum(self.plane_cnt.astype(np.int64))])
            pl0, pl1 = int(pl_pref[q0]), int(pl_pref[q1])
            pc = self.plane_cnt[q0:q1]
        else:
            pl0, pl1 = q0 * self.uniform_H, q1 * self.uniform_H
            pc = None
        kw = {}
        if self.obs_ptr is not None:
            o0, o1 = int(self.obs_ptr[p0]), int(self.obs_ptr[p1])
            kw = dict(obs_ptr=self.obs_ptr[p0:p1 + 1] - o0, obs_idx=self.obs_idx[o0:o1],
                      obs_start_time=self.obs_start_time, obs_duration=self.obs_duration,
                      obs_piece_ptr=self.obs_piece_ptr, obs_piece_dur=self.obs_piece_dur, obs_coeffs=self.obs_coeffs)
        return TrajBatch(seg_ptr=self.seg_ptr[p0:p1 + 1] - g0, piece_num=Ns[g0:g1], singul=self.singul[g0:g1],
                         ini_state=self.ini_state[g0:g1], fin_state=self.fin_state[g0:g1], init_T=self.init_T[g0:g1],
                         inner_pts=self.inner_pts[int(inner_pref[g0]):int(inner_pref[g1])],
                         planes=self.planes[pl0:pl1], t_now=self.t_now[p0:p1], uniform_H=self.uniform_H, plane_cnt=pc,
                         meta=dict(self.meta), **kw)


def batch_from_problems(problems: Sequence[dict]) -> TrajBatch:
    """Builds a TrajBatch from per-problem dicts that mirror the reference call:
    ``iniStates`` / ``finStates`` (list of 2x3), ``initInnerPts`` (list of 2x(N-1)), ``initTs`` (S,),
    ``hPolys`` (list over segments of list over samples of 4xH arrays), ``singuls``, ``now``,
    optional ``obstacles`` = list of dicts(start_time, duration, durs[P], coeffs[P,2,6])."""
    seg_ptr = [0]
    piece_num: List[int] = []
    singul: List[int] = []
    ini, fin, T, inner, planes, cnt, tnow = [], [], [], [], [], [], []
    obs_ptr = [0]
    obs_idx: List[int] = []
    o_start, o_dur, o_pptr, o_pdur, o_coef = [], [], [0], [], []
    any_obs = any(len(pr.get("obstacles", [])) > 0 for pr in problems)
    for pr in problems:
        S = len(pr["initInnerPts"])
        for s in range(S):
            P = np.asarray(pr["initInnerPts"][s], dtype=np.float64).reshape(2, -1)
            piece_num.append(P.shape[1] + 1)
            singul.append(int(pr["singuls"][s]))
            ini.append(np.asarray(pr["iniStates"][s], dtype=np.float64).reshape(2, 3).T.reshape(-1))
            fin.append(np.asarray(pr["finStates"][s], dtype=np.float64).reshape(2, 3).T.reshape(-1))
            T.append(float(pr["initTs"][s]))
            inner.append(P.T.reshape(-1))
            for hp in pr["hPolys"][s]:
                hp = np.asarray(hp, dtype=np.float64).reshape(4, -1)
                cnt.append(hp.shape[1])
                planes.append(hp.T.reshape(-1, 4))
        seg_ptr.append(seg_ptr[-1] + S)
        tnow.append(float(pr.get("now", 0.0)))
        for ob in pr.get("obstacles", []):
            obs_idx.append(len(o_start))
            o_start.append(float(ob["start_time"]))
            o_dur.append(float(ob["duration"]))
            durs = np.asarray(ob["durs"], dtype=np.float64)
            cf = np.asarray(ob["coeffs"], dtype=np.float64).reshape(len(durs), 2, 6)
            o_pdur.extend(durs.tolist())
            o_coef.append(cf.transpose(0, 2, 1).reshape(len(durs), 12))  # column-major 2x6
            o_pptr.append(o_pptr[-1] + len(durs))
        obs_ptr.append(len(obs_idx))
    cnt_arr = np.asarray(cnt, dtype=np.int32)
    uniform = int(cnt_arr[0]) if len(cnt_arr) and np.all(cnt_arr == cnt_arr[0]) else 0
    kw = {}
    if any_obs:
        kw = dict(obs_ptr=obs_ptr, obs_idx=obs_idx, obs_start_time=o_start, obs_duration=o_dur, obs_piece_ptr=o_pptr,
                  obs_piece_dur=o_pdur, obs_coeffs=np.concatenate(o_coef, axis=0))
    return TrajBatch(seg_ptr=seg_ptr, piece_num=piece_num, singul=singul, ini_state=np.stack(ini),
                     fin_state=np.stack(fin), init_T=T, inner_pts=np.concatenate(inner),
                     planes=np.concatenate(planes, axis=0) if planes else np.zeros((0, 4)), t_now=tnow,
                     uniform_H=uniform, plane_cnt=None if uniform else cnt_arr, **kw)
