"""The oracle's restatement of the reference's MINCO classes against the reference's OWN header.

`oracle/_ref/libref_minco.so` is plan_utils/poly_traj_utils.hpp (BandedSystem, MinJerkOpt, Piece, Trajectory) compiled
where it lies, unmodified, against a test-only Eigen-compatible layer (oracle/eigen_shim2 + oracle/ref_stubs; recipe in
oracle/Makefile).  The call sequence of one cost evaluation - generate, getTrajJerkCost, initSmGradCost, [constraint
gradient added to gdC / gdT], calGrads_PT - and the trajectory evaluators used for moving obstacles are run on both sides
with identical inputs.  Live when the library is present (this container and, prebuilt, the GPU box), and always against
tests/golden/minco_reference_header.npz recorded from it by tests/make_golden.py.  Agreement is bit-for-bit: both sides
evaluate every expression element by element in the order the reference writes it."""
import os

import numpy as np
import pytest

import make_golden as MG  # tests/make_golden.py (same directory)

GOLD = os.path.join(os.path.dirname(__file__), "golden", "minco_reference_header.npz")
KEYS = ("c", "energy", "gdC0", "gdT0", "gdP", "gdHead", "gdTail", "gdT")


def _check(a, b, what):
    for k in KEYS:
        np.testing.assert_array_equal(a[k], b[k], err_msg=f"{what}: {k}")


def test_oracle_matches_recorded_reference_header():
    g = np.load(GOLD)
    from oracle import oracle_py as O
    L = O.lib()
    for i, case in enumerate(MG.minco_cases()):
        got = O.minco_run(L.oracle_minco_run, *case)
        _check(got, {k: g[f"run{i}_{k}"] for k in KEYS}, f"case {i} (N={case[0]}, dT={case[2]})")
        for j, t in enumerate(MG.traj_times(case)):
            out = O.traj_eval(L.oracle_traj_eval, case[0], case[1], case[2], case[3], case[4], t)
            ref = g[f"traj{i}_{j}"]
            sel = [k for k in range(17) if k != 14]  # 14 = heading angle, not on the hot path
            np.testing.assert_array_equal(out[sel], ref[sel], err_msg=f"case {i} t={t}")


def test_oracle_matches_reference_header_live():
    from oracle import oracle_py as O
    R = O.ref_minco_lib()
    if R is None:
        pytest.skip("oracle/_ref/libref_minco.so not available (needs /root/reference or the prebuilt copy)")
    L = O.lib()
    rng = np.random.default_rng(123)
    for N in (2, 3, 5, 16, 24, 40):
        for dT in (0.25, 1.0, 3.5):
            P = np.cumsum(rng.normal(size=(N - 1, 2)) * 2.0, axis=0)
            head = rng.normal(size=(3, 2)) * np.array([[5.0], [3.0], [1.0]])
            tail = rng.normal(size=(3, 2)) * np.array([[5.0], [3.0], [1.0]]) + np.array([[P[-1, 0], P[-1, 1]], [0, 0], [0, 0]])
            ex = rng.normal(size=(6 * N, 2)) * 10.0
            a = O.minco_run(R.ref_minco_run, N, P, dT, head, tail, ex, 1.7)
            b = O.minco_run(L.oracle_minco_run, N, P, dT, head, tail, ex, 1.7)
            _check(b, a, f"N={N} dT={dT}")
            for t in (0.0, 0.3 * dT, dT, 2.6 * dT, N * dT, N * dT + 0.7):
                ra = O.traj_eval(R.ref_traj_eval, N, P, dT, head, tail, t)
                rb = O.traj_eval(L.oracle_traj_eval, N, P, dT, head, tail, t)
                sel = [k for k in range(17) if k != 14]
                np.testing.assert_array_equal(rb[sel], ra[sel], err_msg=f"N={N} dT={dT} t={t}")
