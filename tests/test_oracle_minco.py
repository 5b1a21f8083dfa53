"""Oracle self-checks, MINCO part (poly_traj_utils.hpp:727-1095): banded LU vs dense solve, boundary conditions,
waypoint interpolation, C4 continuity at the knots, jerk energy vs quadrature."""
import ctypes as C

import numpy as np
import pytest

from dftpav_b200 import capi, synth

dp = capi.c_double_p


def _generate(oracle, N, P, dT, head, tail):
    c = np.zeros((6 * N, 2))
    rc = oracle.lib().oracle_minco_generate(N, np.ascontiguousarray(P.T.reshape(-1)).ctypes.data_as(dp), dT,
                                            np.ascontiguousarray(head.T.reshape(-1)).ctypes.data_as(dp),
                                            np.ascontiguousarray(tail.T.reshape(-1)).ctypes.data_as(dp), c.ctypes.data_as(dp))
    assert rc == 0
    return c


def _poly(c, i, t, der=0):
    k = np.arange(6)
    coef = np.ones(6)
    for d in range(der):
        coef = coef * (k - d)
    pw = np.where(k - der >= 0, t ** np.maximum(k - der, 0), 0.0)
    return (c[6 * i:6 * i + 6] * (coef * pw)[:, None]).sum(axis=0)


@pytest.mark.parametrize("N", [2, 3, 8, 24])
def test_banded_lu_equals_dense_solve(oracle, N):
    n = 6 * N
    dense = np.zeros((n, n))
    lu = np.zeros(13 * n)
    assert oracle.lib().oracle_minco_matrix(N, dense.ctypes.data_as(dp), lu.ctypes.data_as(dp)) == 0
    # the matrix matches the independent numpy construction used by the generators
    np.testing.assert_array_equal(dense, synth.minco_dense(N))
    # band structure (6, 6)
    i, j = np.nonzero(dense)
    assert (i - j).max() <= 6 and (j - i).max() <= 6
    rng = np.random.default_rng(N)
    P = rng.normal(size=(2, N - 1)) * 5
    head = rng.normal(size=(2, 3))
    tail = rng.normal(size=(2, 3))
    dT = 0.7
    c = _generate(oracle, N, P, dT, head, tail)
    c_ref = synth.minco_fit(P, dT, head, tail)
    np.testing.assert_allclose(c, c_ref, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("N,dT", [(2, 1.3), (8, 0.9), (24, 1.0)])
def test_minco_identities(oracle, N, dT):
    rng = np.random.default_rng(100 + N)
    P = np.cumsum(rng.normal(size=(2, N - 1)), axis=1) * 2
    head = np.stack([rng.normal(size=2), rng.normal(size=2), rng.normal(size=2)], axis=1)
    tail = np.stack([P[:, -1] + rng.normal(size=2), rng.normal(size=2), rng.normal(size=2)], axis=1)
    c = _generate(oracle, N, P, dT, head, tail)
    for d in range(3):  # boundary p, v, a
        np.testing.assert_allclose(_poly(c, 0, 0.0, d), head[:, d], atol=1e-9)
        np.testing.assert_allclose(_poly(c, N - 1, dT, d), tail[:, d], atol=1e-8)
    for i in range(N - 1):
        np.testing.assert_allclose(_poly(c, i, dT, 0), P[:, i], atol=1e-9)       # waypoint interpolation
        for d in range(5):                                                        # C4 continuity
            np.testing.assert_allclose(_poly(c, i, dT, d), _poly(c, i + 1, 0.0, d), atol=1e-7 * (1 + abs(_poly(c, i, dT, d)).max()))


def test_jerk_energy_matches_quadrature(oracle):
    """Smoothness term of the cost = integral of |jerk|^2 (getTrajJerkCost, poly_traj_utils.hpp:998-1009)."""
    batch, prm = synth.make_config(1, B=1, seed=4)
    x0 = oracle.initial_x(prm, batch, 0)
    f, g, c, dT, terms = oracle.evaluate(prm, batch, 0, x0, want_terms=True)
    N = int(batch.piece_num[0])
    ts, w = np.polynomial.legendre.leggauss(12)
    ts = 0.5 * (ts + 1) * dT[0]
    w = 0.5 * w * dT[0]
    energy = sum(w[q] * (_poly(c, i, ts[q], 3) ** 2).sum() for i in range(N) for q in range(len(ts)))
    assert abs(energy - terms[0]) <= 1e-9 * max(1.0, abs(terms[0]))
    assert abs(terms[1] - prm.wei_time * N * dT[0]) <= 1e-9 * terms[1]
