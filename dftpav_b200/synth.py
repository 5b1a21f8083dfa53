"""Seeded synthetic problem generators for the five BASELINE.json configurations.

The reference has no batch axis and no dataset: its inputs come from a hybrid-A* front-end and an occupancy
map (traj_manager.cpp:509-641).  These generators stand in for that front-end: they emit, per problem, exactly
what ``TrajPlanner::RunMINCOParking`` hands to ``OptimizeTrajectory`` — boundary states, inner waypoints sampled
from a front-end path, per-segment durations (piece count x traj_piece_duration), one corridor polygon per
constraint sample (``getRectangleConst`` layout: 4 planes (n, p), traj_manager.cpp:1420-1466), gear signs and,
for config 5, obstacle predictions fitted with MINCO like ``ConverSurroundTrajFromPoints``
(traj_manager.cpp:743-789).

Every random parameter has its own counter-based stream keyed by (config, seed, parameter-name), so problem p is
identical for every batch size B >= p+1 and on every rank.
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import numpy as np

from . import capi
from .batch import TrajBatch, samples_per_segment

CONFIG_DESCRIPTIONS = {
    1: "single 8-piece MINCO traj, no obstacles (plumbing)",
    2: "16 pieces, 10 static convex-corridor polygons (H=6), default limits",
    3: "24 pieces, per-sample rectangles, curvature + vel/acc limits active",
    4: "parking: 3 gear segments [12,8,12] forward/reverse/forward, per-sample rectangles",
    5: "24 pieces + 8 moving obstacles (30-piece quintics), default limits",
}
DEFAULT_B = {1: 1, 2: 1024, 3: 4096, 4: 8192, 5: 16384}


class _Streams:
    def __init__(self, cfg: int, seed: int):
        self.cfg, self.seed = cfg, seed

    def rng(self, name: str) -> np.random.Generator:
        return np.random.default_rng([self.cfg, self.seed, zlib.crc32(name.encode())])

    def uniform(self, name, lo, hi, shape):
        return self.rng(name).uniform(lo, hi, size=shape)

    def normal(self, name, sigma, shape):
        return self.rng(name).normal(0.0, sigma, size=shape)


# ------------------------------------------------------------------------------------------------
# geometry helpers (vectorised over the batch axis)
# ------------------------------------------------------------------------------------------------
def _arc_param_curve(xf: np.ndarray, yf: np.ndarray):
    """Cumulative arc length of polylines xf,yf [B,F]."""
    ds = np.hypot(np.diff(xf, axis=1), np.diff(yf, axis=1))
    return np.concatenate([np.zeros((xf.shape[0], 1)), np.cumsum(ds, axis=1)], axis=1)


def _interp_rows(s_query: np.ndarray, s_grid: np.ndarray, v_grid: np.ndarray) -> np.ndarray:
    """Row-wise np.interp: s_query [B,Q], s_grid/v_grid [B,F]."""
    out = np.empty_like(s_query)
    for b in range(s_query.shape[0]):
        out[b] = np.interp(s_query[b], s_grid[b], v_grid[b])
    return out


def sample_fractions(N: int, K: int, Kd: int) -> np.ndarray:
    """Time fraction in [0,1] of every constraint sample of an N-piece segment, in pointid order
    (traj_manager.cpp:551-568: piece i, k=0..resolution, t = (i + k/resolution) * piece_dur)."""
    fr = []
    for i in range(N):
        R = Kd if (i == 0 or i == N - 1) else K
        for k in range(R + 1):
            fr.append((i + k / R) / N)
    return np.asarray(fr)


def rectangles_around(px, py, yaw, e_left, e_front, e_right, e_back, prm: capi.Params) -> np.ndarray:
    """Per-sample rectangle in getRectangleConst's plane layout (traj_manager.cpp:1420-1466):
    4 planes (n_x, n_y, p_x, p_y) around the RAW vehicle box expanded by e_* metres.  Inputs [...]; output [...,4,4]."""
    L, W, dcr = prm.veh_length, prm.veh_width, prm.veh_d_cr
    c, s = np.cos(yaw), np.sin(yaw)

    def pt(lx, ly):
        return px + c * lx - s * ly, py + s * lx + c * ly

    out = np.empty(px.shape + (4, 4))
    p1 = pt(L / 2 + dcr + e_front, W / 2 + e_left)
    p2 = pt(L / 2 + dcr + e_front, -W / 2 - e_right)
    p3 = pt(-L / 2 + dcr - e_back, -W / 2 - e_right)
    p4 = pt(-L / 2 + dcr - e_back, W / 2 + e_left)
    normals = [(-s, c), (c, s), (s, -c), (-c, -s)]
    for k, (nrm, p) in enumerate(zip(normals, (p1, p2, p3, p4))):
        out[..., k, 0] = nrm[0]
        out[..., k, 1] = nrm[1]
        out[..., k, 2] = p[0]
        out[..., k, 3] = p[1]
    return out


def minco_dense(N: int) -> np.ndarray:
    """Dense time-normalised MINCO matrix (rows as in poly_traj_utils.hpp:897-947); used only to FIT obstacle
    predictions for config 5, like ConverSurroundTrajFromPoints does with MinJerkOpt::generate."""
    n = 6 * N
    A = np.zeros((n, n))
    A[0, 0] = 1.0
    A[1, 1] = 1.0
    A[2, 2] = 2.0
    for i in range(N - 1):
        r = 6 * i
        A[r + 3, r + 3:r + 6] = [6.0, 24.0, 60.0]
        A[r + 3, r + 9] = -6.0
        A[r + 4, r + 4:r + 6] = [24.0, 120.0]
        A[r + 4, r + 10] = -24.0
        A[r + 5, r:r + 6] = 1.0
        A[r + 6, r:r + 6] = 1.0
        A[r + 6, r + 6] = -1.0
        A[r + 7, r + 1:r + 6] = [1.0, 2.0, 3.0, 4.0, 5.0]
        A[r + 7, r + 7] = -1.0
        A[r + 8, r + 2:r + 6] = [2.0, 6.0, 12.0, 20.0]
        A[r + 8, r + 8] = -2.0
    A[n - 3, n - 6:n] = 1.0
    A[n - 2, n - 5:n] = [1.0, 2.0, 3.0, 4.0, 5.0]
    A[n - 1, n - 4:n] = [2.0, 6.0, 12.0, 20.0]
    return A


def minco_fit(P: np.ndarray, dT: float, head: np.ndarray, tail: np.ndarray) -> np.ndarray:
    """MinJerkOpt::generate in numpy: P [2,N-1], head/tail [2,3] -> c [6N,2] (row 6i+k = coeff of t^k of piece i)."""
    N = P.shape[1] + 1
    A = minco_dense(N)
    b = np.zeros((6 * N, 2))
    b[0] = head[:, 0]
    b[1] = head[:, 1] * dT
    b[2] = head[:, 2] * dT ** 2
    for i in range(N - 1):
        b[6 * i + 5] = P[:, i]
    b[6 * N - 3] = tail[:, 0]
    b[6 * N - 2] = tail[:, 1] * dT
    b[6 * N - 1] = tail[:, 2] * dT ** 2
    sol = np.linalg.solve(A, b)
    scale = np.tile(dT ** -np.arange(6.0), N)[:, None]
    return sol * scale


# ------------------------------------------------------------------------------------------------
# single-segment S-curve problems (configs 1, 2, 3, 5)
# ------------------------------------------------------------------------------------------------
def _s_curve_paths(st: _Streams, B: int, length: float, f_lo: float, f_hi: float, F: int = 2048):
    """Front-end path: y = A/2 (1 - cos(2 pi f u)), u in [0,1], start (0,0) heading +x, total ARC length = `length`.
    Returns arc-length grid s [B,F] and x,y,yaw on it."""
    A = st.uniform("amp", 2.0, 6.0, B)
    f = st.uniform("freq", f_lo, f_hi, B)
    sign = np.where(st.uniform("side", 0.0, 1.0, B) < 0.5, -1.0, 1.0)
    u = np.linspace(0.0, 1.0, F)[None, :]
    # choose x-extent X so that the arc length equals `length` (fixed-point iteration on the polyline length)
    X = np.full(B, length)
    for _ in range(30):
        x = X[:, None] * u
        y = sign[:, None] * 0.5 * A[:, None] * (1.0 - np.cos(2 * np.pi * f[:, None] * u))
        s = _arc_param_curve(x, y)
        X = X * length / s[:, -1]
    x = X[:, None] * u
    y = sign[:, None] * 0.5 * A[:, None] * (1.0 - np.cos(2 * np.pi * f[:, None] * u))
    s = _arc_param_curve(x, y)
    dx = np.gradient(x, axis=1)
    dy = np.gradient(y, axis=1)
    yaw = np.arctan2(dy, dx)
    return s, x, y, yaw


def _single_segment(cfg: int, B: int, seed: int, N: int, prm: capi.Params, *, length: float, f_lo: float, f_hi: float,
                    corridor: str) -> TrajBatch:
    st = _Streams(cfg, seed)
    K, Kd = prm.traj_resolution, prm.des_traj_resolution
    M = samples_per_segment(N, K, Kd)
    s, x, y, yaw = _s_curve_paths(st, B, length, f_lo, f_hi)
    total = s[:, -1:]
    # inner waypoints: equal arc-length spacing + N(0, 0.15^2) noise
    wq = total * (np.arange(1, N)[None, :] / N)
    wx = _interp_rows(wq, s, x) + st.normal("wp_noise_x", 0.15, (B, N - 1))
    wy = _interp_rows(wq, s, y) + st.normal("wp_noise_y", 0.15, (B, N - 1))
    inner = np.stack([wx, wy], axis=2).reshape(B, -1)
    v0 = st.uniform("v0", 0.3, 3.0, B)
    ini = np.zeros((B, 6))
    ini[:, 2] = v0
    fin = np.zeros((B, 6))
    fin[:, 0] = x[:, -1]
    fin[:, 1] = y[:, -1]
    fin[:, 2] = 0.24 * np.cos(yaw[:, -1])
    fin[:, 3] = 0.24 * np.sin(yaw[:, -1])
    # corridor
    fr = sample_fractions(N, K, Kd)
    sq = total * fr[None, :]
    sx = _interp_rows(sq, s, x)
    sy = _interp_rows(sq, s, y)
    syaw = _interp_rows(sq, s, np.unwrap(yaw, axis=1))
    if corridor == "huge":
        e = np.full((B, M), 100.0)
        planes = rectangles_around(sx, sy, syaw, e, e, e, e, prm)
        H = 4
    elif corridor == "rect":
        ex = st.uniform("expand", 0.3, 1.2, (B, M, 4))
        planes = rectangles_around(sx, sy, syaw, ex[..., 0], ex[..., 1], ex[..., 2], ex[..., 3], prm)
        H = 4
    elif corridor == "poly10":
        planes = _ten_polygons(st, B, s, x, y, yaw, fr, prm)
        H = 6
    else:
        raise ValueError(corridor)
    return TrajBatch(seg_ptr=np.arange(B + 1), piece_num=np.full(B, N), singul=np.ones(B), ini_state=ini, fin_state=fin,
                     init_T=np.full(B, N * 1.0), inner_pts=inner.reshape(-1), planes=planes.reshape(-1, 4),
                     t_now=np.zeros(B), uniform_H=H,
                     meta={"config": cfg, "seed": seed, "N": [N], "K": K, "Kd": Kd, "H": H})


def _ten_polygons(st: _Streams, B, s, x, y, yaw, fr, prm: capi.Params) -> np.ndarray:
    """Config 2: 10 convex hexagons (H=6) around consecutive tenths of the front-end path, half-width U(1.5,3) m,
    assigned to samples by arc position and expanded to the per-sample list the API requires."""
    M = fr.shape[0]
    total = s[:, -1]
    half_w = st.uniform("poly_halfwidth", 1.5, 3.0, (B, 10))
    margin_f = prm.veh_d_cr + 0.5 * prm.veh_length + 2 * prm.half_margin + 1.5   # room for the front overhang
    margin_b = 0.5 * prm.veh_length - prm.veh_d_cr + 2 * prm.half_margin + 1.5
    polys = np.empty((B, 10, 6, 4))
    for k in range(10):
        a = _interp_rows((total * (k / 10.0))[:, None], s, x)[:, 0], _interp_rows((total * (k / 10.0))[:, None], s, y)[:, 0]
        b = (_interp_rows((total * ((k + 1) / 10.0))[:, None], s, x)[:, 0],
             _interp_rows((total * ((k + 1) / 10.0))[:, None], s, y)[:, 0])
        dx, dy = b[0] - a[0], b[1] - a[1]
        ln = np.hypot(dx, dy)
        ux, uy = dx / ln, dy / ln          # along
        vx, vy = -uy, ux                   # left
        # lateral sag of the path inside this tenth (so the strip contains it)
        mids = total[:, None] * np.linspace(k / 10.0, (k + 1) / 10.0, 9)[None, :]
        mx, my = _interp_rows(mids, s, x), _interp_rows(mids, s, y)
        lat = (mx - a[0][:, None]) * vx[:, None] + (my - a[1][:, None]) * vy[:, None]
        w_left = half_w[:, k] + np.maximum(lat.max(axis=1), 0.0)
        w_right = half_w[:, k] + np.maximum((-lat).max(axis=1), 0.0)
        front = ln + margin_f
        back = margin_b
        cx, cy = a[0], a[1]

        def plane(nx, ny, d):  # n . (q - c) <= d  ->  point c + n*d
            return np.stack([nx, ny, cx + nx * d, cy + ny * d], axis=1)

        polys[:, k, 0] = plane(vx, vy, w_left)
        polys[:, k, 1] = plane(ux, uy, front)
        polys[:, k, 2] = plane(-vx, -vy, w_right)
        polys[:, k, 3] = plane(-ux, -uy, back)
        r2 = np.sqrt(0.5)
        # two corner cuts at the front: clip 15 % off each front corner (keeps the polygon convex, H=6)
        polys[:, k, 4] = plane(r2 * (ux + vx), r2 * (uy + vy), 0.85 * r2 * (front + w_left) + 0.15 * r2 * np.maximum(front, w_left))
        polys[:, k, 5] = plane(r2 * (ux - vx), r2 * (uy - vy), 0.85 * r2 * (front + w_right) + 0.15 * r2 * np.maximum(front, w_right))
    idx = np.minimum((fr * 10.0).astype(np.int64), 9)          # polygon of each sample by arc position
    return polys[:, idx]                                        # [B, M, 6, 4]


# ------------------------------------------------------------------------------------------------
# config 4: parking, three gear segments
# ------------------------------------------------------------------------------------------------
def _arc_segment(p0x, p0y, psi0, kappa, dist, direction, frac):
    """Pose after travelling frac*dist along a constant-curvature arc; direction=+1 forward, -1 reverse.
    Inputs [B]; frac [Q]; returns x,y,psi [B,Q]."""
    sdist = dist[:, None] * frac[None, :]
    dpsi = kappa[:, None] * sdist * direction
    psi = psi0[:, None] + dpsi
    k = kappa[:, None] * direction
    small = np.abs(k) < 1e-9
    ksafe = np.where(small, 1.0, k)
    x = np.where(small, p0x[:, None] + direction * sdist * np.cos(psi0[:, None]),
                 p0x[:, None] + direction * (np.sin(psi) - np.sin(psi0[:, None])) / ksafe)
    y = np.where(small, p0y[:, None] + direction * sdist * np.sin(psi0[:, None]),
                 p0y[:, None] - direction * (np.cos(psi) - np.cos(psi0[:, None])) / ksafe)
    return x, y, psi


def _parking(cfg: int, B: int, seed: int, prm: capi.Params, Ns=(12, 8, 12)) -> TrajBatch:
    st = _Streams(cfg, seed)
    K, Kd = prm.traj_resolution, prm.des_traj_resolution
    S = len(Ns)
    sing = np.array([1 if i % 2 == 0 else -1 for i in range(S)])  # forward / reverse / forward (/ reverse)
    speeds = [st.uniform(f"speed{i}", 1.8, 3.0, B) if i % 2 == 0 else st.uniform(f"speed{i}", 0.6, 1.2, B) for i in range(S)]
    kap = [st.uniform(f"kappa{i}", -0.12, 0.12, B) if i % 2 == 0 else st.uniform(f"kappa{i}", -0.2, 0.2, B) for i in range(S)]
    px, py, psi = np.zeros(B), np.zeros(B), st.uniform("psi0", -np.pi, np.pi, B)
    ini = np.zeros((B, S, 6))
    fin = np.zeros((B, S, 6))
    inner_list, plane_list = [], []
    v_start = st.uniform("v0", 0.3, 2.0, B)
    for sgi, N in enumerate(Ns):
        dist = speeds[sgi] * N * 1.0
        d = float(sing[sgi])
        wfr = np.arange(1, N) / N
        wx, wy, _ = _arc_segment(px, py, psi, kap[sgi], dist, d, wfr)
        wx = wx + st.normal(f"wpx{sgi}", 0.10, (B, N - 1))
        wy = wy + st.normal(f"wpy{sgi}", 0.10, (B, N - 1))
        inner_list.append(np.stack([wx, wy], axis=2).reshape(B, -1))
        fr = sample_fractions(N, K, Kd)
        sx, sy, spsi = _arc_segment(px, py, psi, kap[sgi], dist, d, fr)
        M = fr.shape[0]
        ex = st.uniform(f"expand{sgi}", 0.3, 1.2, (B, M, 4))
        plane_list.append(rectangles_around(sx, sy, spsi, ex[..., 0], ex[..., 1], ex[..., 2], ex[..., 3], prm).reshape(B, -1))
        ex_, ey_, epsi = _arc_segment(px, py, psi, kap[sgi], dist, d, np.array([1.0]))
        ex_, ey_, epsi = ex_[:, 0], ey_[:, 0], epsi[:, 0]
        v_in = v_start if sgi == 0 else np.full(B, 0.24)
        ini[:, sgi, 0], ini[:, sgi, 1] = px, py
        ini[:, sgi, 2], ini[:, sgi, 3] = d * v_in * np.cos(psi), d * v_in * np.sin(psi)
        fin[:, sgi, 0], fin[:, sgi, 1] = ex_, ey_
        fin[:, sgi, 2], fin[:, sgi, 3] = d * 0.24 * np.cos(epsi), d * 0.24 * np.sin(epsi)
        px, py, psi = ex_, ey_, epsi
    inner = np.concatenate(inner_list, axis=1)
    planes = np.concatenate(plane_list, axis=1)
    return TrajBatch(seg_ptr=np.arange(B + 1) * S, piece_num=np.tile(np.array(Ns), B), singul=np.tile(sing, B),
                     ini_state=ini.reshape(-1, 6), fin_state=fin.reshape(-1, 6),
                     init_T=np.tile(np.array(Ns, dtype=np.float64), B), inner_pts=inner.reshape(-1),
                     planes=planes.reshape(-1, 4), t_now=np.zeros(B), uniform_H=4,
                     meta={"config": cfg, "seed": seed, "N": list(Ns), "K": K, "Kd": Kd, "H": 4})


# ------------------------------------------------------------------------------------------------
# config 5: obstacles
# ------------------------------------------------------------------------------------------------
def _circle_obstacles(st: _Streams, B: int, s, x, y, n_obs: int = 8, pieces: int = 30):
    """Each obstacle: a car on a circle of radius 12 m at 4.5 m/s (the shape dynamicObs.yaml +
    ConverSurroundTrajFromPoints produce: 30 s prediction at 1 s, fitted by MinJerkOpt::generate).
    Two of the eight circles pass through a point of the ego path near the time the ego is there; the others sit
    25-60 m to the side."""
    R, v = 12.0, 4.5
    w = v / R
    total = s[:, -1]
    T_total = 24.0
    A_inv = np.linalg.inv(minco_dense(pieces))
    tk = np.arange(pieces + 1) * 1.0
    U = B * n_obs
    frac = st.uniform("obs_meet_frac", 0.25, 0.85, (B, n_obs))
    mx = _interp_rows(total[:, None] * frac, s, x)
    my = _interp_rows(total[:, None] * frac, s, y)
    off_r = st.uniform("obs_offset", 25.0, 60.0, (B, n_obs))
    off_r[:, :2] = st.uniform("obs_near_offset", 0.0, 3.5, (B, 2))          # two near-misses per problem
    off_a = st.uniform("obs_offset_angle", -np.pi, np.pi, (B, n_obs))
    passx = mx + off_r * np.cos(off_a)
    passy = my + off_r * np.sin(off_a)
    t_meet = frac * T_total
    phase_at_meet = st.uniform("obs_phase", -np.pi, np.pi, (B, n_obs))
    ccw = np.where(st.uniform("obs_dir", 0, 1, (B, n_obs)) < 0.5, -1.0, 1.0)
    cx = passx - R * np.cos(phase_at_meet)
    cy = passy - R * np.sin(phase_at_meet)
    ph0 = phase_at_meet - ccw * w * t_meet
    ph = ph0[..., None] + ccw[..., None] * w * tk[None, None, :]           # [B,O,31]
    qx = cx[..., None] + R * np.cos(ph)
    qy = cy[..., None] + R * np.sin(ph)
    vx = -ccw[..., None] * R * w * np.sin(ph)
    vy = ccw[..., None] * R * w * np.cos(ph)
    ax = -R * w * w * np.cos(ph)
    ay = -R * w * w * np.sin(ph)
    # MINCO fit (dT = 1): b rows -> coefficients, batched
    b = np.zeros((U, 6 * pieces, 2))
    qx_, qy_, vx_, vy_, ax_, ay_ = (a.reshape(U, -1) for a in (qx, qy, vx, vy, ax, ay))
    b[:, 0, 0], b[:, 0, 1] = qx_[:, 0], qy_[:, 0]
    b[:, 1, 0], b[:, 1, 1] = vx_[:, 0], vy_[:, 0]
    b[:, 2, 0], b[:, 2, 1] = ax_[:, 0], ay_[:, 0]
    for i in range(pieces - 1):
        b[:, 6 * i + 5, 0], b[:, 6 * i + 5, 1] = qx_[:, i + 1], qy_[:, i + 1]
    n = 6 * pieces
    b[:, n - 3, 0], b[:, n - 3, 1] = qx_[:, -1], qy_[:, -1]
    b[:, n - 2, 0], b[:, n - 2, 1] = vx_[:, -1], vy_[:, -1]
    b[:, n - 1, 0], b[:, n - 1, 1] = ax_[:, -1], ay_[:, -1]
    c = np.einsum("ij,ujk->uik", A_inv, b)                                  # [U, 6P, 2], dT = 1 -> c = b~
    c = c.reshape(U, pieces, 6, 2)
    coeffs = c[:, :, ::-1, :].reshape(U * pieces, 12)                       # highest power first, (x,y) per power
    return dict(obs_ptr=np.arange(B + 1) * n_obs, obs_idx=np.arange(U), obs_start_time=np.zeros(U),
                obs_duration=np.full(U, float(pieces)), obs_piece_ptr=np.arange(U + 1) * pieces,
                obs_piece_dur=np.ones(U * pieces), obs_coeffs=coeffs)


# ------------------------------------------------------------------------------------------------
# public entry
# ------------------------------------------------------------------------------------------------
def make_config(cfg: int, B: int | None = None, seed: int = 0) -> Tuple[TrajBatch, capi.Params]:
    """Returns (batch, params) for BASELINE.json configs[cfg-1].  Parameters are the reference defaults except
    where the config says otherwise (config 3: v_max 3, a_max 2, kappa_max 0.3 so the limits are active)."""
    prm = capi.default_params()
    if B is None:
        B = DEFAULT_B[cfg]
    if cfg == 1:
        batch = _single_segment(1, B, seed, 8, prm, length=0.6 * 5.0 * 8, f_lo=0.5, f_hi=1.0, corridor="huge")
    elif cfg == 2:
        batch = _single_segment(2, B, seed, 16, prm, length=0.7 * 5.0 * 16, f_lo=0.5, f_hi=1.5, corridor="poly10")
    elif cfg == 3:
        prm.max_forward_vel = 3.0
        prm.max_forward_acc = 2.0
        prm.max_forward_cur = 0.3
        batch = _single_segment(3, B, seed, 24, prm, length=0.8 * 3.0 * 24, f_lo=1.5, f_hi=3.5, corridor="rect")
    elif cfg == 4:
        batch = _parking(4, B, seed, prm)
    elif cfg == 5:
        st = _Streams(5, seed)
        batch = _single_segment(5, B, seed, 24, prm, length=0.6 * 5.0 * 24, f_lo=1.0, f_hi=2.5, corridor="rect")
        s, x, y, _ = _s_curve_paths(st, B, 0.6 * 5.0 * 24, 1.0, 2.5)
        ob = _circle_obstacles(st, B, s, x, y)
        batch = TrajBatch(seg_ptr=batch.seg_ptr, piece_num=batch.piece_num, singul=batch.singul,
                          ini_state=batch.ini_state, fin_state=batch.fin_state, init_T=batch.init_T,
                          inner_pts=batch.inner_pts, planes=batch.planes, t_now=batch.t_now, uniform_H=4,
                          meta=dict(batch.meta, obstacles=8), **ob)
    else:
        raise ValueError(f"unknown config {cfg}")
    batch.meta["description"] = CONFIG_DESCRIPTIONS[cfg]
    return batch, prm


def make_custom(B: int, pieces, seed: int = 0) -> Tuple[TrajBatch, capi.Params]:
    """Shapes outside the five BASELINE configurations (tests of the compiled-in limits: up to 4 gear segments of up to 64
    pieces): one segment with config-3 limits and rectangles, or a forward / reverse / ... parking manoeuvre."""
    prm = capi.default_params()
    pieces = [int(n) for n in pieces]
    if len(pieces) == 1:
        N = pieces[0]
        prm.max_forward_vel = 3.0
        prm.max_forward_acc = 2.0
        prm.max_forward_cur = 0.3
        batch = _single_segment(30 + N, B, seed, N, prm, length=0.8 * 3.0 * N, f_lo=1.5 * N / 24.0, f_hi=3.5 * N / 24.0,
                                corridor="rect")
    else:
        batch = _parking(40 + len(pieces), B, seed, prm, Ns=tuple(pieces))
    batch.meta["description"] = f"custom shape {pieces}"
    return batch, prm


def make_obstacle_map(batch: TrajBatch, prm: capi.Params, resolution: float = 0.2, clutter: int = 60, seed: int = 0):
    """Occupancy grid (uint8 [size_y, size_x], OCCUPIED = 80) around the problems of a single-segment batch: random discs and
    boxes that stay clear of every problem's front-end path (the piecewise-linear path through start, waypoints and goal,
    with a margin that keeps the vehicle box free).  Returns (grid, origin_xy, resolution, states) where states [Q, 3] are
    the per-sample poses (x, y, yaw) of the path - what traj_manager.cpp:576-600 feeds getRectangleConst - in the sample
    order the optimizer expects (piece by piece, K_i + 1 samples each)."""
    rng = np.random.default_rng(1000 + seed)
    K, Kd = prm.traj_resolution, prm.des_traj_resolution
    pts_all, states = [], []
    for p in range(batch.B):
        g = int(batch.seg_ptr[p])
        N = int(batch.piece_num[g])
        off = int(sum(2 * (int(batch.piece_num[q]) - 1) for q in range(g)))
        inner = batch.inner_pts[off: off + 2 * (N - 1)].reshape(N - 1, 2)
        pts = np.vstack([batch.ini_state[g, 0:2], inner, batch.fin_state[g, 0:2]])
        pts_all.append(pts)
        for i in range(N):
            Ki = Kd if i in (0, N - 1) else K
            a, b = pts[i], pts[i + 1]
            yaw = np.arctan2(b[1] - a[1], b[0] - a[0])
            for j in range(Ki + 1):
                q = a + (b - a) * (j / Ki)
                states.append((q[0], q[1], yaw))
    states = np.array(states)
    allp = np.vstack(pts_all)
    lo = allp.min(axis=0) - 20.0
    hi = allp.max(axis=0) + 20.0
    size = np.ceil((hi - lo) / resolution).astype(int) + 1
    grid = np.full((size[1], size[0]), 127, dtype=np.uint8)  # FREE
    yy, xx = np.mgrid[0:size[1], 0:size[0]]
    cx = lo[0] + xx * resolution
    cy = lo[1] + yy * resolution
    # distance of every cell to the nearest sampled pose (coarse: poses subsampled)
    sub = states[:: max(1, len(states) // 4000), :2]
    dmin = np.full(grid.shape, 1e9)
    for k in range(0, len(sub), 256):
        blk = sub[k:k + 256]
        d = np.sqrt((cx[..., None] - blk[:, 0]) ** 2 + (cy[..., None] - blk[:, 1]) ** 2).min(axis=-1)
        dmin = np.minimum(dmin, d)
    for _ in range(clutter):
        c = lo + rng.uniform(0, 1, 2) * (hi - lo)
        if rng.uniform() < 0.5:
            r = rng.uniform(0.5, 3.0)
            mask = (cx - c[0]) ** 2 + (cy - c[1]) ** 2 <= r * r
        else:
            w, hgt = rng.uniform(0.5, 4.0, 2)
            mask = (np.abs(cx - c[0]) <= w) & (np.abs(cy - c[1]) <= hgt)
        grid[mask & (dmin > 4.5)] = capi.OCCUPIED  # the vehicle box (half diagonal ~3.6 m from the rear axle) stays free
    return grid, (float(lo[0]), float(lo[1])), resolution, states


def algorithmic_bytes_per_eval(batch: TrajBatch, prm: capi.Params, p: int = 0) -> int:
    """B_eval = 16 n + 32 sum_samples H_s + 96 S + 104 sum_obstacles P_o."""
    g0, g1 = int(batch.seg_ptr[p]), int(batch.seg_ptr[p + 1])
    S = g1 - g0
    n = batch.n_of(p)
    K, Kd = prm.traj_resolution, prm.des_traj_resolution
    samp = np.array([samples_per_segment(int(N), K, Kd) for N in batch.piece_num], dtype=np.int64)
    pref = np.concatenate([[0], np.cumsum(samp)])
    q0, q1 = int(pref[g0]), int(pref[g1])
    planes = int(batch.plane_cnt[q0:q1].sum()) if batch.plane_cnt is not None else (q1 - q0) * batch.uniform_H
    obs = 0
    if batch.obs_ptr is not None:
        for k in range(int(batch.obs_ptr[p]), int(batch.obs_ptr[p + 1])):
            u = int(batch.obs_idx[k])
            obs += int(batch.obs_piece_ptr[u + 1] - batch.obs_piece_ptr[u])
    return 16 * n + 32 * planes + 96 * S + 104 * obs


class FrontendPaths:
    """Host arrays of a dftpav_frontend (kept alive by this object) - see include/dftpav_b200.h."""

    def __init__(self, paths, shots, start_vel, end_vel, limits):
        """paths: list of [n_i, 3] pose arrays; shots: per path a list of (first_pt, last_pt, time, length, sign)."""
        self.pt_ptr = np.zeros(len(paths) + 1, dtype=np.int32)
        self.shot_ptr = np.zeros(len(paths) + 1, dtype=np.int32)
        for i, (pp, ss) in enumerate(zip(paths, shots)):
            self.pt_ptr[i + 1] = self.pt_ptr[i] + len(pp)
            self.shot_ptr[i + 1] = self.shot_ptr[i] + len(ss)
        self.pts = np.ascontiguousarray(np.concatenate(paths), dtype=np.float64)
        idx, tm, ln, sg = [], [], [], []
        for ss in shots:
            idx.append(ss[0][0])
            for (a, b, t, l, s_) in ss:
                idx.append(b)
                tm.append(t), ln.append(l), sg.append(s_)
        self.shot_index = np.array(idx, dtype=np.int32)
        self.shot_time = np.array(tm, dtype=np.float64)
        self.shot_length = np.array(ln, dtype=np.float64)
        self.shot_s = np.array(sg, dtype=np.int32)
        self.start_vel = np.ascontiguousarray(start_vel, dtype=np.float64)
        self.end_vel = np.ascontiguousarray(end_vel, dtype=np.float64)
        self.limits = tuple(float(v) for v in limits)  # max_forward_vel, max_forward_acc, max_backward_vel, max_backward_acc, non_siguav

    @property
    def n_paths(self):
        return len(self.pt_ptr) - 1

    def as_c(self) -> capi.Frontend:
        import ctypes as C
        ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        return capi.Frontend(self.n_paths, self.pt_ptr.ctypes.data_as(ip), self.pts.ctypes.data_as(dp),
                             self.shot_ptr.ctypes.data_as(ip), self.shot_index.ctypes.data_as(ip),
                             self.shot_time.ctypes.data_as(dp), self.shot_length.ctypes.data_as(dp),
                             self.shot_s.ctypes.data_as(ip), self.start_vel.ctypes.data_as(dp),
                             self.end_vel.ctypes.data_as(dp), *self.limits)


def make_frontend_paths(n_paths: int, seed: int = 0, trapezoid_duration=None, wrap_yaw: bool = True) -> FrontendPaths:
    """Synthetic hybrid-A*-like paths: arcs and straights sampled every ~0.2 m, with gear changes (a cusp reverses the
    direction of travel), divided into gear segments and timed exactly like KinoAstar::getKinoNode (kino_astar.cpp:617-668):
    the direction sign of every step, the segment lengths, evaluateDuration per segment.  `trapezoid_duration` is the
    evaluateDuration implementation to use (the library's host function or the oracle's)."""
    rng = np.random.default_rng(2000 + seed)
    vf, af, vb, ab, nsv = 3.0, 1.5, 1.5, 1.0, 0.05
    paths, shots, v0s, v1s = [], [], [], []
    for _ in range(n_paths):
        x, y, yaw = rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-np.pi, np.pi)
        pts = [(x, y, yaw)]
        direction = 1
        for leg in range(int(rng.integers(1, 4))):
            curv = rng.uniform(-0.25, 0.25)
            for _ in range(int(rng.integers(20, 120))):
                ds = 0.2 * direction
                x += ds * np.cos(yaw)
                y += ds * np.sin(yaw)
                yaw += ds * curv
                if wrap_yaw:
                    yaw = (yaw + np.pi) % (2 * np.pi) - np.pi
                pts.append((x, y, yaw))
            direction = -direction
        P = np.array(pts)
        # getKinoNode: gear sign per step, shot division
        d = P[1:, 0:2] - P[:-1, 0:2]
        sgn = np.where(d[:, 0] * np.cos(P[:-1, 2]) + d[:, 1] * np.sin(P[:-1, 2]) >= 0, 1, -1)
        lens = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])
        shot_index, shot_s, shot_len = [0], [], []
        tmpl, lastS = 0.0, int(sgn[0])
        for i in range(len(P) - 1):
            if sgn[i] * lastS >= 0:
                tmpl += lens[i]
            else:
                shot_index.append(i), shot_s.append(lastS), shot_len.append(tmpl)
                tmpl = lens[i]
            lastS = int(sgn[i])
        shot_s.append(lastS), shot_len.append(tmpl), shot_index.append(len(P) - 1)
        v0, v1 = rng.uniform(0, 1.0), rng.uniform(0, 1.0)
        ns = len(shot_s)
        ss = []
        for i in range(ns):
            a, b = (nsv, nsv)
            if i == 0:
                a = v0
            if i == ns - 1:
                b = v1
            mv, ma = (vf, af) if shot_s[i] > 0 else (vb, ab)
            ss.append((shot_index[i], shot_index[i + 1], trapezoid_duration(shot_len[i], mv, ma, a, b), shot_len[i], shot_s[i]))
        paths.append(P), shots.append(ss), v0s.append(v0), v1s.append(v1)
    return FrontendPaths(paths, shots, v0s, v1s, (vf, af, vb, ab, nsv))
