"""Batched MINCO fits (dftpav_minco_fit) against the oracle's MinJerkOpt::generate restatement (itself
bit-identical to the reference header, tests/test_oracle_vs_reference_minco.py): coefficients BIT-IDENTICAL - the kernel runs
the reference's banded substitutions with explicitly rounded operations."""
import numpy as np
import pytest

from dftpav_b200 import capi, synth

pytestmark = pytest.mark.gpu


def test_fits_bit_identical_to_the_oracle(oracle, gpu_opt_factory):
    rng = np.random.default_rng(5)
    Ns = [2, 3, 8, 30, 30, 17, 64, 30, 5, 100]
    pp = np.concatenate([[0], np.cumsum(Ns)]).astype(np.int32)
    dT = rng.uniform(0.3, 2.0, len(Ns))
    head = rng.normal(size=(len(Ns), 3, 2)) * [[5.0], [2.0], [1.0]]
    tail = rng.normal(size=(len(Ns), 3, 2)) * [[5.0], [2.0], [1.0]] + [[40.0], [0.0], [0.0]]
    inner = [np.cumsum(rng.normal(size=(N - 1, 2)) + [1.5, 0.2], axis=0) for N in Ns]
    opt = gpu_opt_factory(capi.default_params())
    got = opt.minco_fit(pp, dT, head.reshape(len(Ns), 6), tail.reshape(len(Ns), 6), np.concatenate([a.reshape(-1) for a in inner]))
    L = oracle.lib()
    for k, N in enumerate(Ns):
        ref = oracle.minco_run(L.oracle_minco_run, N, inner[k], dT[k], head[k], tail[k], np.zeros((6 * N, 2)), 0.0)["c"]
        want = ref.reshape(N, 6, 2)[:, ::-1, :].reshape(N, 12)  # row 6i+k = coefficient of t^k  ->  highest power first
        assert np.array_equal(got[pp[k]:pp[k + 1]], want), (k, N, np.abs(got[pp[k]:pp[k + 1]] - want).max())


def test_obstacle_pool_from_predicted_states(oracle, gpu_opt_factory):
    """ConverSurroundTrajFromPoints (traj_manager.cpp:743-789): predicted states of the moving obstacles -> fitted pool -> the same
    evaluation as with the pool the synthetic generator fitted with numpy (independent dense solve): agreement to 1e-9."""
    batch, prm = synth.make_config(5, B=6, seed=3)
    U = int(batch.obs_piece_ptr.shape[0] - 1)
    Np = int(batch.obs_piece_ptr[1] - batch.obs_piece_ptr[0])
    lib = capi.load_library()
    # predicted states sampled from the generator's own trajectories at the piece boundaries
    heads, tails, inners, dTs = [], [], [], []
    for u in range(U):
        p0 = int(batch.obs_piece_ptr[u])
        cm = batch.obs_coeffs[p0:p0 + Np].reshape(Np, 6, 2)  # highest power first
        dur = batch.obs_piece_dur[p0:p0 + Np]
        pos0 = cm[:, 5, :]
        T = dur[-1]
        end = sum(cm[-1, 5 - k, :] * T ** k for k in range(6))
        vel = lambda i, t: sum(k * cm[i, 5 - k, :] * t ** (k - 1) for k in range(1, 6))  # noqa: E731
        acc = lambda i, t: sum(k * (k - 1) * cm[i, 5 - k, :] * t ** (k - 2) for k in range(2, 6))  # noqa: E731
        heads.append(np.concatenate([pos0[0], vel(0, 0.0), acc(0, 0.0)]))
        tails.append(np.concatenate([end, vel(Np - 1, T), acc(Np - 1, T)]))
        inners.append(pos0[1:].reshape(-1))
        dTs.append(dur.sum() / Np)
    opt = gpu_opt_factory(prm)
    pp = (np.arange(U + 1) * Np).astype(np.int32)
    fit = opt.minco_fit(pp, np.array(dTs), np.array(heads), np.array(tails), np.concatenate(inners))
    assert np.abs(fit - batch.obs_coeffs).max() <= 1e-7 * max(1.0, np.abs(batch.obs_coeffs).max())
    import copy

    b2 = copy.copy(batch)
    b2.obs_coeffs = fit
    opt.upload(batch)
    x0 = opt.initial_x(batch)
    f1, _ = opt.evaluate(x0)
    opt.upload(b2)
    f2, _ = opt.evaluate(x0)
    assert np.abs(f1 - f2).max() <= 1e-6 * np.abs(f1).max()
    # state_to_flat_output (host helper): a state moving along its heading with a curvature
    st = np.array([1.0, 2.0, 0.3, 4.5, 0.2, 0.05])
    flat = np.zeros(6)
    lib.dftpav_state_to_flat_output(st.ctypes.data_as(capi.c_double_p), flat.ctypes.data_as(capi.c_double_p))
    c, s = np.cos(0.3), np.sin(0.3)
    assert np.allclose(flat, [1.0, 2.0, c * 4.5, s * 4.5, c * 0.2 - s * 0.05 * 4.5 ** 2, s * 0.2 + c * 0.05 * 4.5 ** 2], rtol=1e-15)
