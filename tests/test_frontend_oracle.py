"""Front-end resampling restatement (oracle/frontend_oracle.cpp, PARITY UNPINNED: KinoAstar lives behind OMPL / PCL / ROS):
closed-form cases of the trapezoidal speed profile (kino_astar.cpp:744-795), evaluatePos on paths with a known answer
(kino_astar.cpp:468-521), the sampling loop of RunMINCOParking (traj_manager.cpp:531-567), and bit-for-bit agreement with an
independent Python implementation of the same procedure.  The library's host function dftpav_trapezoid_duration needs no GPU."""
import math

import numpy as np

from dftpav_b200 import capi, synth


def _py_length(curt, L, mv, ma, v0, v1):
    s2, e2, m2 = v0 * v0, v1 * v1, mv * mv
    crit = (m2 - s2) / (2 * ma) + (m2 - e2) / (2 * ma)
    if L >= crit:
        t1 = (mv - v0) / ma
        t2 = t1 + (L - crit) / mv
        if curt <= t1:
            return v0 * curt + 0.5 * ma * (curt * curt)
        if curt <= t2:
            return v0 * t1 + 0.5 * ma * (t1 * t1) + (curt - t1) * mv
        u = curt - t2
        return v0 * t1 + 0.5 * ma * (t1 * t1) + (t2 - t1) * mv + mv * u - 0.5 * ma * (u * u)
    tv = math.sqrt(0.5 * (s2 + e2 + 2 * ma * L))
    tt = (tv - v0) / ma
    if curt <= tt:
        return v0 * curt + 0.5 * ma * (curt * curt)
    u = curt - tt
    return v0 * tt + 0.5 * ma * (tt * tt) + tv * u - 0.5 * ma * (u * u)


def _py_pos(fe, p, t):
    s0, s1 = fe.shot_ptr[p], fe.shot_ptr[p + 1]
    st, sl, ss = fe.shot_time[s0:s1], fe.shot_length[s0:s1], fe.shot_s[s0:s1]
    si = fe.shot_index[s0 + p: s1 + p + 1]
    pts = fe.pts[fe.pt_ptr[p]: fe.pt_ptr[p + 1]]
    vf, af, vb, ab, nsv = fe.limits
    total = 0.0
    for v in st:
        total += float(v)
    t = min(max(0.0, t), total)
    tmp, idx, cut = 0.0, -1, 0.0
    for i, v in enumerate(st):
        tmp += float(v)
        if tmp >= t:
            idx, cut = i, t - tmp + float(v)
            break
    iv = float(fe.start_vel[p]) if idx == 0 else nsv
    fv = float(fe.end_vel[p]) if idx == len(st) - 1 else nsv
    lt = pts[si[idx]: si[idx + 1] + 1]
    mv, ma = (vf, af) if ss[idx] > 0 else (vb, ab)
    arc = _py_length(cut, float(sl[idx]), mv, ma, iv, fv)
    acc = 0.0
    for i in range(len(lt) - 1):
        dx, dy = float(lt[i + 1, 0] - lt[i, 0]), float(lt[i + 1, 1] - lt[i, 1])
        l = math.sqrt(dx * dx + dy * dy)
        acc += l
        if acc >= arc:
            l1 = acc - arc
            l2 = l - l1
            out = [l1 / l * float(lt[i, q]) + l2 / l * float(lt[i + 1, q]) for q in range(3)]
            y0, y1 = float(lt[i, 2]), float(lt[i + 1, 2])
            if abs(y1 - y0) >= math.pi:
                if y1 <= 0:
                    out[2] = l1 / l * y0 + l2 / l * (y1 + 2 * math.pi)
                elif y0 <= 0:
                    out[2] = l1 / l * (y0 + 2 * math.pi) + l2 / l * y1
            return out
    return [float(v) for v in lt[-1]]


def test_trapezoid_profile_closed_forms(oracle):
    lib = capi.load_library()
    # long segment from rest to rest: 2 v/a ramps + cruise
    T = oracle.trapezoid_duration(20.0, 3.0, 1.5, 0.0, 0.0)
    assert T == 2 * 3.0 / 1.5 + (20.0 - 3.0 * 3.0 / 1.5) / 3.0
    # short segment: triangular profile, peak sqrt(a L)
    T2 = oracle.trapezoid_duration(2.0, 3.0, 1.5, 0.0, 0.0)
    assert abs(T2 - 2 * math.sqrt(1.5 * 2.0) / 1.5) < 1e-14
    rng = np.random.default_rng(0)
    for _ in range(300):
        L, mv, ma = rng.uniform(0.05, 40), rng.uniform(0.5, 5), rng.uniform(0.3, 3)
        v0, v1 = rng.uniform(0, mv), rng.uniform(0, mv)
        T = oracle.trapezoid_duration(L, mv, ma, v0, v1)
        assert T == lib.dftpav_trapezoid_duration(L, mv, ma, v0, v1)          # the library's host function, bit for bit
        ts = np.linspace(0, max(T, 0.1), 17)
        if L >= abs(v0 * v0 - v1 * v1) / (2 * ma) + 1e-9:  # the speed change fits the length (otherwise the reference's
            # formulas return what they return - restated as written - and nothing below holds)
            assert abs(oracle.trapezoid_length(T, L, mv, ma, v0, v1) - L) < 1e-9 * max(1, L)   # the profile covers the length
            ls = [oracle.trapezoid_length(t, L, mv, ma, v0, v1) for t in ts]
            assert all(b >= a - 1e-12 for a, b in zip(ls, ls[1:])) and ls[0] == 0.0
        for t in ts:
            assert oracle.trapezoid_length(t, L, mv, ma, v0, v1) == _py_length(t, L, mv, ma, v0, v1)


def test_evaluate_pos_on_a_straight_path(oracle):
    n = 101
    P = np.stack([np.linspace(0, 20, n), np.full(n, 3.0), np.zeros(n)], axis=1)
    T = oracle.trapezoid_duration(20.0, 3.0, 1.5, 0.5, 0.2)
    fe = synth.FrontendPaths([P], [[(0, n - 1, T, 20.0, 1)]], [0.5], [0.2], (3.0, 1.5, 1.5, 1.0, 0.05))
    ts = np.array([-1.0, 0.0, 0.3 * T, 0.5 * T, T, T + 5.0])
    out = oracle.frontend_evaluate_pos(fe, np.zeros(len(ts), dtype=np.int32), ts)
    want = [oracle.trapezoid_length(min(max(t, 0.0), T), 20.0, 3.0, 1.5, 0.5, 0.2) for t in ts]
    assert np.allclose(out[:, 0], want, atol=1e-12) and np.all(out[:, 1] == 3.0) and np.all(out[:, 2] == 0.0)
    assert out[0, 0] == 0.0 and abs(out[-1, 0] - 20.0) < 1e-9


def test_yaw_is_unwrapped_across_pi(oracle):
    # heading passes through +-pi between two samples: the interpolated yaw must stay next to pi, not jump to 0
    P = np.array([[0.0, 0.0, math.pi - 0.05], [-1.0, 0.05, -math.pi + 0.05], [-2.0, -0.05, -math.pi + 0.15]])
    L = float(np.sum(np.hypot(*(P[1:, :2] - P[:-1, :2]).T)))
    T = oracle.trapezoid_duration(L, 3.0, 1.5, 0.0, 0.0)
    fe = synth.FrontendPaths([P], [[(0, 2, T, L, 1)]], [0.0], [0.0], (3.0, 1.5, 1.5, 1.0, 0.05))
    ts = np.linspace(0.05 * T, 0.45 * T, 9)
    out = oracle.frontend_evaluate_pos(fe, np.zeros(9, dtype=np.int32), ts)
    assert np.all(np.abs(np.abs(out[:, 2]) - math.pi) < 0.06)


def test_against_an_independent_implementation(oracle):
    fe = synth.make_frontend_paths(24, seed=1, trapezoid_duration=oracle.trapezoid_duration)
    assert (np.diff(fe.shot_ptr) > 1).any(), "some paths must change gear"
    rng = np.random.default_rng(5)
    q = rng.integers(0, fe.n_paths, size=600).astype(np.int32)
    total = np.array([fe.shot_time[fe.shot_ptr[p]: fe.shot_ptr[p + 1]].sum() for p in range(fe.n_paths)])
    t = rng.uniform(-0.2, 1.2, size=600) * total[q]
    t[:24] = np.cumsum(fe.shot_time)[fe.shot_ptr[1:] - 1][:24] * 0 + [fe.shot_time[fe.shot_ptr[p]] for p in range(24)]  # exact shot boundaries
    q[:24] = np.arange(24)
    got = oracle.frontend_evaluate_pos(fe, q, t)
    want = np.array([_py_pos(fe, int(p), float(tt)) for p, tt in zip(q, t)])
    np.testing.assert_array_equal(got, want)
    # the sampling loop of RunMINCOParking
    seg_dur = fe.shot_time * (1 + 1e-3 * rng.standard_normal(len(fe.shot_time)))  # the FlatTrajData sums differ a little
    pn, pd, states, inner = oracle.frontend_resample(fe, seg_dur, 0.5, 6, 12)
    assert np.all(pn == np.maximum((seg_dur / 0.5 + 0.5).astype(int), 2)) and np.all(pd == seg_dur / pn)
    assert len(states) == int((2 * 13 + (pn - 2) * 7).sum()) and len(inner) == int((pn - 1).sum())
    k = 0
    ki = 0
    for p in range(fe.n_paths):
        base = 0.0
        for g in range(fe.shot_ptr[p], fe.shot_ptr[p + 1]):
            res_time = 0.0
            for i in range(pn[g]):
                r = 12 if i in (0, pn[g] - 1) else 6
                for kk in range(r + 1):
                    tt = base + res_time + 1.0 * kk / r * pd[g]
                    w = _py_pos(fe, p, tt)
                    assert list(states[k]) == w, (p, g, i, kk)
                    k += 1
                    if kk == r and i != pn[g] - 1:
                        assert list(inner[ki]) == w[:2]
                        ki += 1
                res_time += pd[g]
            base += float(seg_dur[g])
    assert k == len(states) and ki == len(inner)
