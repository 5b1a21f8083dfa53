"""Front-end resampling on the GPU (dftpav_frontend_evaluate_pos, dftpav_frontend_resample)
against the CPU restatement: only + - * / sqrt on the path, so every output is bit-identical; and the resampled states feed
the corridor kernel, the inner waypoints / piece counts / durations the optimizer - the device pipeline from a searched path
to an optimised trajectory."""
import numpy as np
import pytest

from dftpav_b200 import capi, synth
from dftpav_b200.batch import TrajBatch

pytestmark = pytest.mark.gpu


def _paths(n, seed):
    lib = capi.load_library()
    return synth.make_frontend_paths(n, seed=seed, trapezoid_duration=lib.dftpav_trapezoid_duration)


def test_evaluate_pos_bit_identical(oracle, gpu_opt_factory):
    fe = _paths(96, 3)
    opt = gpu_opt_factory(capi.default_params())
    rng = np.random.default_rng(1)
    nq = 20000
    q = rng.integers(0, fe.n_paths, size=nq).astype(np.int32)
    total = np.array([fe.shot_time[fe.shot_ptr[p]: fe.shot_ptr[p + 1]].sum() for p in range(fe.n_paths)])
    t = rng.uniform(-0.1, 1.1, size=nq) * total[q]
    q[: fe.n_paths] = np.arange(fe.n_paths)
    t[: fe.n_paths] = [fe.shot_time[fe.shot_ptr[p]] for p in range(fe.n_paths)]   # exactly on the first gear change
    got = opt.frontend_evaluate_pos(fe, q, t)
    want = oracle.frontend_evaluate_pos(fe, q, t)
    bad = np.nonzero(np.any(got != want, axis=1))[0]
    assert bad.size == 0, (bad[:5], got[bad[:2]], want[bad[:2]])


@pytest.mark.parametrize("K,Kd", [(8, 16), (6, 12)])
def test_resample_bit_identical(K, Kd, oracle, gpu_opt_factory):
    fe = _paths(200, 7)
    opt = gpu_opt_factory(capi.default_params())
    rng = np.random.default_rng(2)
    seg_dur = fe.shot_time * (1 + 1e-3 * rng.standard_normal(len(fe.shot_time)))
    pn, pd, states, inner = opt.frontend_resample(fe, seg_dur, 0.5, K, Kd)
    ms = opt.stats().solve_ms
    pn_o, pd_o, st_o, in_o = oracle.frontend_resample(fe, seg_dur, 0.5, K, Kd)
    np.testing.assert_array_equal(pn, pn_o)
    np.testing.assert_array_equal(pd, pd_o)
    np.testing.assert_array_equal(states, st_o)
    np.testing.assert_array_equal(inner, in_o)
    print(f"{fe.n_paths} paths, {len(seg_dur)} gear segments, {len(states)} states resampled in {ms:.3f} ms on the device")


def test_path_to_trajectory_pipeline(gpu_opt_factory):
    """searched path -> resampled states + waypoints (GPU) -> corridor (GPU, stays on the device) -> optimizer."""
    fe = _paths(40, 11)
    prm = capi.default_params()
    K, Kd = prm.traj_resolution, prm.des_traj_resolution
    opt = gpu_opt_factory(prm)
    seg_dur = fe.shot_time.copy()
    pn, pd, states, inner = opt.frontend_resample(fe, seg_dur, 0.5, K, Kd)
    # an empty map of the right size: the corridor becomes the 10 m box around every state
    lo = fe.pts[:, 0:2].min(axis=0) - 25.0
    size = np.ceil((fe.pts[:, 0:2].max(axis=0) + 25.0 - lo) / 0.2).astype(int) + 1
    opt.set_map(np.full((size[1], size[0]), 127, dtype=np.uint8), lo, 0.2)
    opt.corridor_rectangles(states, keep=True, want_host=False)
    # one optimizer problem per path: its gear segments, boundary states from the path ends, waypoints and durations from above
    S = np.diff(fe.shot_ptr)
    G = int(S.sum())
    ini = np.zeros((G, 6))
    fin = np.zeros((G, 6))
    for p in range(fe.n_paths):
        for g in range(fe.shot_ptr[p], fe.shot_ptr[p + 1]):
            a, b = fe.shot_index[g + p], fe.shot_index[g + p + 1]
            pa, pb = fe.pts[fe.pt_ptr[p] + a], fe.pts[fe.pt_ptr[p] + b]
            sg = fe.shot_s[g]
            va = fe.start_vel[p] if g == fe.shot_ptr[p] else fe.limits[4]
            vb = fe.end_vel[p] if g == fe.shot_ptr[p + 1] - 1 else fe.limits[4]
            ini[g] = (pa[0], pa[1], sg * max(va, 0.05) * np.cos(pa[2]), sg * max(va, 0.05) * np.sin(pa[2]), 0.0, 0.0)
            fin[g] = (pb[0], pb[1], sg * max(vb, 0.05) * np.cos(pb[2]), sg * max(vb, 0.05) * np.sin(pb[2]), 0.0, 0.0)
    batch = TrajBatch(seg_ptr=fe.shot_ptr.astype(np.int32), piece_num=pn.astype(np.int32), singul=fe.shot_s.astype(np.int32),
                      ini_state=ini, fin_state=fin, inner_pts=inner.reshape(-1), init_T=pd * pn, planes=None, uniform_H=4,
                      plane_cnt=None, t_now=np.zeros(fe.n_paths))
    sol = opt.solve(batch)
    assert opt.stats().h2d_bytes < 2_000_000, "the corridor must not be uploaded"
    assert np.all(np.isfinite(sol.cost)) and (sol.status != capi.PROBLEM_REJECTED).all()
    assert sol.success.mean() > 0.5


def test_malformed_front_end_is_refused(gpu_opt_factory):
    import ctypes as C

    fe = _paths(5, 2)
    opt = gpu_opt_factory(capi.default_params())
    lib = capi.load_library()
    dp = capi.c_double_p
    # a capacity that is too small: refused, but the needed number of states is reported
    seg = fe.shot_time.copy()
    pn = np.zeros(len(seg), dtype=np.int32)
    pd = np.zeros(len(seg))
    st = np.zeros((10, 3))
    inn = np.zeros((1000, 2))
    n = C.c_int64(0)
    c = fe.as_c()
    rc = lib.dftpav_frontend_resample(opt._h, C.byref(c), seg.ctypes.data_as(dp), 0.5, 8, 16, pn.ctypes.data_as(capi.c_int32_p),
                                      pd.ctypes.data_as(dp), 10, st.ctypes.data_as(dp), inn.ctypes.data_as(dp), C.byref(n))
    assert rc == capi.ERR_INVALID_ARG and n.value == int((2 * 17 + (np.maximum((seg / 0.5 + 0.5).astype(int), 2) - 2) * 9).sum())
    # a decreasing shot index, a query on a path that does not exist, a non-positive segment duration
    bad = _paths(5, 2)
    bad.shot_index[1] = -3
    with pytest.raises(Exception):
        opt.frontend_evaluate_pos(bad, np.zeros(1, dtype=np.int32), np.zeros(1))
    with pytest.raises(Exception):
        opt.frontend_evaluate_pos(fe, np.array([99], dtype=np.int32), np.zeros(1))
    seg[0] = 0.0
    with pytest.raises(Exception):
        opt.frontend_resample(fe, seg, 0.5, 8, 16)
