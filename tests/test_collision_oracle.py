"""The collision re-check restatement (oracle/corridor_oracle.cpp, PARITY UNPINNED: its reference lives behind the ROS /
semantic-map stack) against cases with a known answer and an independent numpy implementation of the same procedure
(traj_server_ros.cpp:385-397, semantic_map_manager.cc:639-661, shapes.cc:110-149)."""
import math

import numpy as np

from dftpav_b200 import capi


def _line_traj(x0, y0, vx, vy, N, dT):
    """N pieces of duration dT of the straight constant-velocity motion; rows = coefficient of t^k, lowest power first."""
    c = np.zeros((6 * N, 2))
    for i in range(N):
        c[6 * i] = (x0 + vx * dT * i, y0 + vy * dT * i)
        c[6 * i + 1] = (vx, vy)
    return c


def _independent(grid, origin, res, c, dT, singul, dt, spacing, occupied=capi.OCCUPIED, vp=(1.90, 4.88, 1.015)):
    W, L, dcr = vp
    N = c.shape[0] // 6
    total = 0.0
    for _ in range(N):
        total += dT

    def occ(px, py):
        cx = int(math.floor((px - origin[0]) / res + 0.5))  # = std::round away from the half-way cases
        cy = int(math.floor((py - origin[1]) / res + 0.5))
        return 0 <= cx < grid.shape[1] and 0 <= cy < grid.shape[0] and grid[cy, cx] == occupied

    t, k = 0.0, 0
    while t < total:
        i = min(int(t // dT), N - 1)
        while i > 0 and i * dT > t:
            i -= 1
        tl = t - i * dT
        pw = np.array([tl ** q for q in range(6)])
        pos = pw @ c[6 * i: 6 * i + 6]
        vel = (np.arange(1, 6) * pw[:5]) @ c[6 * i + 1: 6 * i + 6]
        yaw = math.atan2(singul * vel[1], singul * vel[0])
        ct, sn = math.cos(yaw), math.sin(yaw)
        o = pos + dcr * np.array([ct, sn])
        f, l = np.array([ct, sn]), np.array([-sn, ct])
        corners = [o + 0.5 * L * f - 0.5 * W * l, o + 0.5 * L * f + 0.5 * W * l, o - 0.5 * L * f + 0.5 * W * l,
                   o - 0.5 * L * f - 0.5 * W * l]
        verts = list(corners)
        for a, b in zip(corners, corners[1:] + corners[:1]):
            n = np.linalg.norm(b - a)
            dl = spacing
            while dl < n:
                verts.append(a + dl / n * (b - a))
                dl += spacing
        if any(occ(v[0], v[1]) for v in verts):
            return k
        t += dt
        k += 1
    return -1


def test_wall_ahead_is_hit_at_the_expected_sample(oracle):
    res = 0.2
    grid = np.full((200, 400), 127, dtype=np.uint8)
    origin = (0.0, 0.0)
    wall_cx = 100  # cells with centre x = 20.0 cover [19.9, 20.1)
    grid[:, wall_cx] = capi.OCCUPIED
    c = _line_traj(5.0, 10.0, 2.0, 0.0, N=8, dT=1.5)
    k = oracle.collision_check_segment(grid, origin, res, c, 1.5, 1)
    # front bumper = x + d_cr + length / 2 = x + 3.455 reaches 19.9 at x = 16.445, t = 5.7225 -> first sample with t >= that
    assert k == math.ceil(5.7225 / 0.05 - 1e-9) == 115
    assert oracle.collision_check_segment(np.full_like(grid, 127), origin, res, c, 1.5, 1) == -1
    grid[:] = capi.OCCUPIED
    assert oracle.collision_check_segment(grid, origin, res, c, 1.5, 1) == 0
    assert oracle.collision_check_segment(grid, origin, res, c, 1.5, 1, max_samples=1) == 0
    assert oracle.collision_check_segment(np.full_like(grid, 127), origin, res, c, 1.5, 1, max_samples=5) == -2


def test_against_an_independent_implementation(oracle):
    rng = np.random.default_rng(3)
    res = 0.25
    origin = (-10.0, -12.0)
    hits = 0
    for trial in range(40):
        grid = np.full((240, 320), 127, dtype=np.uint8)
        for _ in range(rng.integers(3, 25)):
            cx, cy = rng.integers(5, 315), rng.integers(5, 235)
            grid[cy - rng.integers(0, 3): cy + rng.integers(1, 4), cx - rng.integers(0, 3): cx + rng.integers(1, 4)] = capi.OCCUPIED
        ang = rng.uniform(-math.pi, math.pi)
        v = rng.uniform(0.5, 4.0)
        sg = int(rng.choice([-1, 1]))
        c = _line_traj(rng.uniform(10, 40), rng.uniform(10, 30), sg * v * math.cos(ang), sg * v * math.sin(ang), N=5, dT=1.2)
        c[2::6] += rng.normal(0, 0.05, size=(5, 2))  # a little curvature inside every piece
        got = oracle.collision_check_segment(grid, origin, res, c, 1.2, sg, dt=0.05, vertex_spacing=0.1)
        want = _independent(grid, origin, res, c, 1.2, sg, 0.05, 0.1)
        assert got == want, (trial, got, want)
        hits += got >= 0
    assert 5 < hits < 40
