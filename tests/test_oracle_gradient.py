"""Oracle self-checks, cost part (traj_optimizer.cpp:206-806, 1311-1707): the analytic gradient agrees with central
finite differences for every configuration (corridor, velocity, acceleration, curvature, gear variables, moving
obstacles, time variables).  Error must be round-off dominated (~eps*|f|/h)."""
import numpy as np
import pytest

from dftpav_b200 import synth


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_fd_gradient(oracle, cfg):
    batch, prm = synth.make_config(cfg, B=3, seed=21)
    rng = np.random.default_rng(cfg)
    for p in (0, 2):
        x = oracle.initial_x(prm, batch, p)
        x = x + 0.02 * rng.normal(size=x.shape)
        f, g, _, _, terms = oracle.evaluate(prm, batch, p, x, want_terms=True)
        assert np.isfinite(f) and abs(f - terms.sum()) <= 1e-9 * abs(f)
        best = np.full(x.shape, np.inf)
        for h in (1e-5, 2e-6):
            gfd = np.zeros_like(g)
            for i in range(len(x)):
                xp, xm = x.copy(), x.copy()
                xp[i] += h
                xm[i] -= h
                gfd[i] = (oracle.evaluate(prm, batch, p, xp)[0] - oracle.evaluate(prm, batch, p, xm)[0]) / (2 * h)
            best = np.minimum(best, np.abs(gfd - g) / np.maximum(1.0, np.abs(g)))
        assert best.max() < 2e-4, (cfg, p, best.max(), int(best.argmax()))
        assert np.median(best) < 1e-6


def test_active_terms_are_exercised(oracle):
    """The configurations really activate the terms they are meant to (otherwise parity would be vacuous)."""
    act = {}
    for cfg in (2, 3, 4, 5):
        batch, prm = synth.make_config(cfg, B=6, seed=5)
        tsum = np.zeros(5)
        for p in range(batch.B):
            x0 = oracle.initial_x(prm, batch, p)
            tsum += oracle.evaluate(prm, batch, p, x0, want_terms=True)[4]
        act[cfg] = tsum
    assert act[2][2] > 0 or act[2][4] > 0          # corridor or feasibility at x0
    assert act[3][2] > 0 and act[3][4] > 0         # corridor and vel/acc/curvature penalties
    assert act[4][4] > 0                           # feasibility around the cusps
    assert act[5][3] > 0                           # moving obstacles


def test_quirk_five_vertex_list(oracle):
    """Quirk Q1: the front-left vertex is penalised twice (5-entry vec_le_).  A corridor plane violated only by that
    vertex must cost exactly twice what the mirrored plane costs for the front-right vertex."""
    from dftpav_b200.batch import batch_from_problems
    from dftpav_b200 import capi

    prm = capi.default_params()
    N, K, Kd = 2, prm.traj_resolution, prm.des_traj_resolution
    M = 2 * (Kd + 1)
    big = 1e3

    def problem(sign):
        # straight line along +x at constant speed; one plane y*sign <= 1.0 cuts the left (sign=+1) / right (-1) side
        plane = np.array([[0.0, sign, 0.0, sign * 1.0], [1, 0, big, 0], [-1, 0, -big, 0], [0, -sign, 0, -sign * big]]).T
        return dict(iniStates=[np.array([[0.0, 2, 0], [0, 0, 0]])], finStates=[np.array([[4.0, 2, 0], [0, 0, 0]])],
                    initInnerPts=[np.array([[2.0], [0.0]])], initTs=[2.0], hPolys=[[plane] * M], singuls=[1], now=0.0)

    b = batch_from_problems([problem(+1.0), problem(-1.0)])
    tl = oracle.evaluate(prm, b, 0, oracle.initial_x(prm, b, 0), want_terms=True)[4]
    tr = oracle.evaluate(prm, b, 1, oracle.initial_x(prm, b, 1), want_terms=True)[4]
    assert tl[2] > 0 and tr[2] > 0
    # left side: vertices 0 (twice) and 3; right side: vertices 1 and 2 -> ratio (2*a + b) / (a' + b') with a=a', b=b'
    assert tl[2] > tr[2] * 1.2
