"""Independent numpy restatement of the COST (no gradients) of the hot path, written from SURVEY.md Appendix A rather
than from the oracle's code: dense MINCO solve (numpy.linalg.solve on the dense matrix, no banded LU), vectorised
sample sweep, static corridor + velocity + longitudinal-acceleration + curvature penalties, time cost, gear coupling.
Moving obstacles are not covered (config 5 is checked by finite differences and the GPU parity tests only).
Used by tests/test_oracle_np.py as a second implementation to pin the oracle's cost value; gradients of the oracle are
then checked against finite differences of THIS function."""
import numpy as np

from dftpav_b200 import synth
from dftpav_b200.batch import samples_per_segment


def virt2real(tau, mini_T):
    return np.where(tau > 0, (0.5 * tau + 1.0) * tau + 1.0, 1.0 / ((0.5 * tau - 1.0) * tau + 1.0)) + mini_T


def smoothed_l1(x, pe=1e-4):
    f3c = 1.0 / pe ** 2
    f4c = -0.5 * f3c / pe
    return np.where(x < pe, (f4c * x + f3c) * x ** 3, x - 0.5 * pe)


def cost(prm, batch, p, x):
    K, Kd = prm.traj_resolution, prm.des_traj_resolution
    g0, g1 = int(batch.seg_ptr[p]), int(batch.seg_ptr[p + 1])
    S = g1 - g0
    Ns = [int(n) for n in batch.piece_num[g0:g1]]
    off = 0
    P = []
    for N in Ns:
        P.append(x[off:off + 2 * (N - 1)].reshape(N - 1, 2).T)
        off += 2 * (N - 1)
    tau = x[off:off + S]
    off += S
    gear = x[off:off + 2 * (S - 1)].reshape(S - 1, 2)
    off += 2 * (S - 1)
    theta = x[off:off + S - 1]
    T = virt2real(tau, prm.mini_T)
    # sample / plane bookkeeping of this problem inside the flat arrays
    samp = np.array([samples_per_segment(int(n), K, Kd) for n in batch.piece_num], dtype=np.int64)
    pref = np.concatenate([[0], np.cumsum(samp)])
    H = batch.uniform_H
    assert H > 0, "np restatement handles uniform corridors"
    width = prm.veh_width + 2 * prm.half_margin
    length = prm.veh_length + 2 * prm.half_margin
    d = prm.veh_d_cr
    verts = np.array([[d + length / 2, width / 2], [d + length / 2, -width / 2], [d - length / 2, -width / 2],
                      [d - length / 2, width / 2], [d + length / 2, width / 2]])  # 5 entries (quirk Q1)
    total = 0.0
    for s in range(S):
        g = g0 + s
        N = Ns[s]
        sing = float(batch.singul[g])
        head = batch.ini_state[g].reshape(3, 2).T.copy()
        tail = batch.fin_state[g].reshape(3, 2).T.copy()
        fwd = sing > 0
        vmax = prm.max_forward_vel if fwd else prm.max_backward_vel
        amax = prm.max_forward_acc if fwd else prm.max_backward_acc
        cmax = prm.max_forward_cur if fwd else prm.max_backward_cur
        for st in (head, tail):  # clamps of OptimizeTrajectory
            for col, lim in ((1, vmax), (2, amax)):
                nv = np.linalg.norm(st[:, col])
                if nv >= lim:
                    st[:, col] = st[:, col] / nv * (lim - 1e-2)
        if s > 0:
            head[:, 0] = gear[s - 1]
            head[:, 1] = -prm.non_sinv * np.array([np.cos(theta[s - 1]), np.sin(theta[s - 1])])
        if s < S - 1:
            tail[:, 0] = gear[s]
            tail[:, 1] = prm.non_sinv * np.array([np.cos(theta[s]), np.sin(theta[s])])
        dT = T[s] / N
        c = synth.minco_fit(P[s], dT, head, tail).reshape(N, 6, 2)
        # jerk energy by Gauss-Legendre quadrature of |sigma'''|^2 (exact for degree-4 integrand with 6 nodes)
        xs, ws = np.polynomial.legendre.leggauss(6)
        ts = 0.5 * (xs + 1) * dT
        k = np.arange(3, 6)
        b3 = (k * (k - 1) * (k - 2))[None, :] * ts[:, None] ** (k - 3)[None, :]      # [q,3]
        jerk = np.einsum("qk,nkc->nqc", b3, c[:, 3:, :])
        total += float((0.5 * dT * ws[None, :, None] * jerk ** 2).sum())
        total += prm.wei_time * T[s]
        # sweep
        planes = batch.planes[pref[g] * H:pref[g + 1] * H].reshape(-1, H, 4).copy()
        nrm = np.linalg.norm(planes[:, :, :2], axis=2, keepdims=True)
        planes[:, :, :2] = planes[:, :, :2] / np.where(nrm > 0, nrm, 1.0)
        pid = 0
        for i in range(N):
            Kp = Kd if (i == 0 or i == N - 1) else K
            step = dT / Kp
            j = np.arange(Kp + 1)
            sv = j * step
            omg = np.where((j == 0) | (j == Kp), 0.5, 1.0)
            pw = sv[:, None] ** np.arange(6)[None, :]
            kk = np.arange(6)
            b0 = pw
            b1 = np.where(kk >= 1, kk * sv[:, None] ** np.maximum(kk - 1, 0), 0.0)
            b2 = np.where(kk >= 2, kk * (kk - 1) * sv[:, None] ** np.maximum(kk - 2, 0), 0.0)
            sig, dsig, ddsig = b0 @ c[i], b1 @ c[i], b2 @ c[i]
            v = np.linalg.norm(dsig, axis=1)
            skip = (v < prm.min_speed) | ((j == 0) & (i == 0)) | ((j == Kp) & (i == N - 1))
            vs = np.where(skip, 1.0, v)
            R = sing * np.stack([np.stack([dsig[:, 0], -dsig[:, 1]], 1), np.stack([dsig[:, 1], dsig[:, 0]], 1)], 1) / vs[:, None, None]
            pl = planes[pid:pid + Kp + 1]
            pid += Kp + 1
            pen = np.zeros(Kp + 1)
            for le in verts:
                bpt = sig + R @ le
                viol = (pl[:, :, :2] * (bpt[:, None, :] - pl[:, :, 2:])).sum(axis=2)
                pen += prm.wei_obs * np.where(viol > 0, smoothed_l1(np.maximum(viol, 0)), 0.0).sum(axis=1)
            z1 = (ddsig * dsig).sum(axis=1)
            z3 = ddsig[:, 1] * dsig[:, 0] - ddsig[:, 0] * dsig[:, 1]
            vv = vs ** 2
            viol_v = vv - vmax ** 2
            viol_a = z1 ** 2 / vv - amax ** 2
            cur = z3 / (vv + prm.help_eps) ** 1.5
            for viol, w in ((viol_v, prm.wei_feas), (viol_a, prm.wei_feas), (cur - cmax, prm.wei_feas * prm.cur_weight_scale),
                            (-cur - cmax, prm.wei_feas * prm.cur_weight_scale)):
                pen += w * np.where(viol > 0, smoothed_l1(np.maximum(viol, 0)), 0.0)
            total += float((np.where(skip, 0.0, omg * step * pen)).sum())
    return total
