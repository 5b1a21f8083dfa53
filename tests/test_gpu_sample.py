"""Batched trajectory evaluator (dftpav_sample / dftpav_sample_uniform / dftpav_pack_polytraj)
against the oracle's Trajectory / Piece evaluators (themselves bit-identical to the reference's header,
tests/test_traj_state_oracle.py): positions, derivatives, velocity and accelerations BIT-IDENTICAL (no libm on that path, no
FMA contraction in the kernel), angle / curvature / steer within 1e-12 (atan2, pow, atan of the CUDA math library)."""
import numpy as np
import pytest

from dftpav_b200 import capi, synth

pytestmark = pytest.mark.gpu
EXACT = [0, 1, 2, 3, 4, 5, 8, 9, 11]
LIBM = [6, 7, 10]


def _seg_coeffs(batch, sol, g):
    p = int(np.searchsorted(batch.seg_ptr, g, side="right") - 1)
    row0 = 6 * int(batch.piece_num[batch.seg_ptr[p]:g].sum())
    N = int(batch.piece_num[g])
    return sol.coeffs[p, row0:row0 + 6 * N], float(sol.piece_dT[g]), int(batch.singul[g]), N


@pytest.mark.parametrize("cfg", [3, 4])
def test_sample_against_oracle(cfg, oracle, gpu_opt_factory):
    batch, prm = synth.make_config(cfg, B=48, seed=4)
    opt = gpu_opt_factory(prm)
    sol = opt.solve(batch)
    G = int(batch.seg_ptr[-1])
    rng = np.random.default_rng(cfg)
    nq = 600
    seg = rng.integers(0, G, nq).astype(np.int32)
    dur = np.array([batch.piece_num[g] * sol.piece_dT[g] for g in range(G)])
    t = rng.uniform(-0.05, 1.15, nq) * dur[seg]
    t[:20] = 0.0
    t[20:40] = dur[seg[20:40]]
    for flags in (0, capi.SAMPLE_CLAMP, capi.SAMPLE_GETSTATE):
        got = opt.sample(seg, t, flags)
        for q in range(nq):
            c, dT, sg, N = _seg_coeffs(batch, sol, int(seg[q]))
            ref, gs = oracle.traj_state(c, dT, sg, t[q], clamp=1 if flags else 0)
            if flags == capi.SAMPLE_GETSTATE:  # what Trajectory::GetState stores (quirk Q11 thresholds)
                ref = ref.copy()
                ref[7], ref[9], ref[10] = gs[1], gs[3], gs[4]
                if abs(gs[2]) < 1.0:
                    ref[11] = 0.0
            assert np.array_equal(got[q, EXACT], ref[EXACT]), (cfg, flags, q, got[q] - ref)
            assert np.allclose(got[q, LIBM], ref[LIBM], rtol=1e-12, atol=1e-13), (cfg, flags, q, got[q, LIBM] - ref[LIBM])


def test_uniform_sampling_and_polytraj(oracle, gpu_opt_factory):
    batch, prm = synth.make_config(4, B=16, seed=9)
    opt = gpu_opt_factory(prm)
    sol = opt.solve_stepwise(batch)
    G = int(batch.seg_ptr[-1])
    dt, cap = 0.05, 700
    xyy, cnt = opt.sample_uniform(dt, cap)
    for g in range(0, G, 5):
        c, dT, sg, N = _seg_coeffs(batch, sol, g)
        total = 0.0
        for _ in range(N):
            total += dT
        k, t = 0, 0.0
        while t < total and k < cap:  # the loop of traj_server_ros.cpp:386
            ref, _ = oracle.traj_state(c, dT, sg, t)
            assert np.array_equal(xyy[g, k, :2], ref[:2]), (g, k)
            assert abs(xyy[g, k, 2] - ref[6]) <= 1e-12, (g, k)
            t += dt
            k += 1
        assert cnt[g] == k, (g, cnt[g], k)
    cx, cy, du = opt.pack_polytraj()
    P = 0
    for g in range(G):
        c, dT, sg, N = _seg_coeffs(batch, sol, g)
        for i in range(N):
            want_x = c[6 * i:6 * i + 6, 0][::-1].astype(np.float32)  # highest power first (Piece::getCoeffMat)
            want_y = c[6 * i:6 * i + 6, 1][::-1].astype(np.float32)
            assert np.array_equal(cx[P], want_x) and np.array_equal(cy[P], want_y) and du[P] == np.float32(dT)
            P += 1
    assert P == cx.shape[0]
