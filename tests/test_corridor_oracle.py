"""Corridor generation: the CPU restatement of TrajPlanner::getRectangleConst
(oracle/corridor_oracle.cpp) against geometric invariants and an independent numpy implementation of the same procedure.
(The reference function itself cannot be compiled here - its translation unit needs the whole ROS / OMPL stack - so this
oracle is pinned by properties, not by the reference binary; stated in its header.)"""
import numpy as np

from dftpav_b200 import capi, synth


def _occupied(grid, origin, res, x, y):
    cx = int(np.floor(abs((x - origin[0]) / res) + 0.5) * np.sign((x - origin[0]) / res))  # std::round: half away from zero
    cy = int(np.floor(abs((y - origin[1]) / res) + 0.5) * np.sign((y - origin[1]) / res))
    if cx < 0 or cx >= grid.shape[1] or cy < 0 or cy >= grid.shape[0]:
        return False
    return grid[cy, cx] == capi.OCCUPIED


def _line(grid, origin, res, p1, p2, checkl):
    d = p2 - p1
    nrm = np.sqrt(d[0] * d[0] + d[1] * d[1])
    dl = 0.0
    while dl < nrm:
        q = d * dl / nrm + p1
        if _occupied(grid, origin, res, q[0], q[1]):
            return True
        dl += checkl
    return _occupied(grid, origin, res, p2[0], p2[1])


def _np_rectangle(grid, origin, res, state, cp):
    """Independent implementation: explicit rectangle bookkeeping in the body frame instead of the shifting source point."""
    x, y, yaw = state
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.array([[c, -s], [s, c]])
    step, limit = res, cp.limit_bound
    # body-frame extents of the current rectangle relative to the rear axle: [front, left, rear, right]
    front, rear = cp.veh_length / 2 + cp.veh_d_cr, -cp.veh_length / 2 + cp.veh_d_cr
    left, right = cp.veh_width / 2, -cp.veh_width / 2
    ex = np.zeros(4)
    fin = [False] * 4
    W = lambda a, b: np.array([x, y]) + R @ np.array([a, b])  # noqa: E731
    while not all(fin):
        for i in range(4):
            if fin[i]:
                continue
            f, l, r_, rt = front + ex[1], left + ex[0], rear - ex[3], right - ex[2]
            if i == 0:
                seg = [W(f, l), W(f, l + step), W(r_, l + step), W(r_, l)]
            elif i == 1:
                seg = [W(f, rt), W(f + step, rt), W(f + step, l), W(f, l)]
            elif i == 2:
                seg = [W(r_, rt), W(r_, rt - step), W(f, rt - step), W(f, rt)]
            else:
                seg = [W(r_, l), W(r_ - step, l), W(r_ - step, rt), W(r_, rt)]
            if any(_line(grid, origin, res, seg[k], seg[k + 1], res / 2) for k in range(3)):
                fin[i] = True
                continue
            ex[i] += step
            if ex[i] >= limit:
                fin[i] = True
    return ex


def test_against_independent_numpy_and_invariants(oracle):
    batch, prm = synth.make_config(3, B=3, seed=2)
    grid, origin, res, states = synth.make_obstacle_map(batch, prm, resolution=0.25, clutter=90, seed=1)
    cp = capi.default_corridor_params()
    sub = states[::37]
    hp, ex = oracle.corridor_rectangles(grid, origin, res, sub)
    assert (ex > 0).any() and (ex < cp.limit_bound).any(), "the map should block some directions and leave others open"
    for q, st in enumerate(sub):
        want = _np_rectangle(grid, origin, res, st, cp)
        # the independent implementation tracks the rectangle in body coordinates instead of shifting the source point: equal up
        # to rounding, i.e. the same number of steps unless a ray-march sample sits on a cell border
        assert np.all(np.abs(ex[q] - want) <= res + 1e-9), (q, ex[q], want)
        assert np.sum(np.abs(ex[q] - want) > 1e-9) <= 1
        # polygon: unit outward normals (left, front, right, rear), the vehicle box inside, consistent with the expansion lengths
        x, y, yaw = st
        c, s = np.cos(yaw), np.sin(yaw)
        n = hp[q, :, 0:2]
        assert np.allclose(n, [[-s, c], [c, s], [s, -c], [-c, -s]], atol=1e-15)
        R = np.array([[c, -s], [s, c]])
        for a in (cp.veh_length / 2 + cp.veh_d_cr, -cp.veh_length / 2 + cp.veh_d_cr):
            for b in (cp.veh_width / 2, -cp.veh_width / 2):
                v = np.array([x, y]) + R @ np.array([a, b])
                for k in range(4):
                    assert np.dot(hp[q, k, 0:2], v - hp[q, k, 2:4]) <= 1e-9
        # distances of the four faces from the rear axle = box extent + expansion
        dist = [np.dot(hp[q, k, 0:2], hp[q, k, 2:4] - np.array([x, y])) for k in range(4)]
        assert np.allclose(dist, [cp.veh_width / 2 + ex[q, 0], cp.veh_length / 2 + cp.veh_d_cr + ex[q, 1],
                                  cp.veh_width / 2 + ex[q, 2], cp.veh_length / 2 - cp.veh_d_cr + ex[q, 3]], atol=1e-9)


def test_blocked_direction_cannot_grow_and_free_map_reaches_the_bound(oracle):
    grid = np.full((200, 200), 127, dtype=np.uint8)
    hp, ex = oracle.corridor_rectangles(grid, (-20.0, -20.0), 0.2, np.array([[0.0, 0.0, 0.3]]))
    assert np.all(ex >= 10.0) and np.all(ex < 10.0 + 0.2 + 1e-9)  # every direction runs into limitBound
    grid[:, 130:] = capi.OCCUPIED  # a wall at x >= 6 m
    hp, ex = oracle.corridor_rectangles(grid, (-20.0, -20.0), 0.2, np.array([[0.0, 0.0, 0.0]]))
    front = 4.88 / 2 + 1.015 + ex[0, 1]
    assert 5.5 < front < 6.0 and ex[0, 0] >= 10.0 and ex[0, 2] >= 10.0 and ex[0, 3] >= 10.0
