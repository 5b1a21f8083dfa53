"""Second, independently written implementation (tests/np_restatement.py: dense MINCO solve, vectorised numpy sweep,
jerk energy by quadrature) pins the oracle's cost on configs 1-4; the oracle's analytic gradient is then checked against
central finite differences of the numpy cost."""
import numpy as np
import pytest

from dftpav_b200 import synth

import np_restatement as NP


@pytest.mark.parametrize("cfg", [1, 2, 3, 4])
def test_cost_matches_independent_numpy(oracle, cfg):
    batch, prm = synth.make_config(cfg, B=3, seed=31)
    rng = np.random.default_rng(cfg)
    for p in range(batch.B):
        x0 = oracle.initial_x(prm, batch, p)
        for scale in (0.0, 0.05):
            x = x0 + scale * rng.normal(size=x0.shape)
            f = oracle.evaluate(prm, batch, p, x)[0]
            fn = NP.cost(prm, batch, p, x)
            assert abs(f - fn) <= 1e-9 * max(1.0, abs(f)), (cfg, p, scale, f, fn)


@pytest.mark.parametrize("cfg", [2, 3, 4])
def test_gradient_matches_fd_of_numpy_cost(oracle, cfg):
    batch, prm = synth.make_config(cfg, B=1, seed=32)
    x = oracle.initial_x(prm, batch, 0)
    x = x + 0.02 * np.random.default_rng(cfg).normal(size=x.shape)
    _, g, _, _ = oracle.evaluate(prm, batch, 0, x)
    idx = np.random.default_rng(1).choice(len(x), size=min(12, len(x)), replace=False)
    h = 1e-5
    for i in idx:
        xp, xm = x.copy(), x.copy()
        xp[i] += h
        xm[i] -= h
        fd = (NP.cost(prm, batch, 0, xp) - NP.cost(prm, batch, 0, xm)) / (2 * h)
        assert abs(fd - g[i]) <= 2e-4 * max(1.0, abs(g[i])), (cfg, int(i), fd, g[i])
