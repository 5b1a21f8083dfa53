"""Pins the oracle's L-BFGS / Lewis-Overton restatement (lbfgs.hpp:276-390, 440-751) against the UNMODIFIED reference
header: (1) live, when oracle/_ref/libref_lbfgs.so is available (built in place from /root/reference), (2) always,
against tests/golden/lbfgs_reference_header.npz recorded from that build.  Iterates must be bit-identical."""
import ctypes as C
import os

import numpy as np
import pytest

from dftpav_b200 import capi, synth

import make_golden as MG  # tests/make_golden.py (same directory)

GOLD = os.path.join(os.path.dirname(__file__), "golden", "lbfgs_reference_header.npz")
dp = capi.c_double_p


def _run_oracle(oracle, n, x0, fgptr, user, prm):
    x = x0.copy()
    f = C.c_double()
    it, ev = C.c_int(), C.c_int()
    tr = np.zeros((MG.CAP, 3))
    ret = oracle.lib().oracle_lbfgs(n, x.ctypes.data_as(dp), C.cast(C.byref(f), dp), fgptr, user, C.byref(prm), C.byref(it),
                                    C.byref(ev), tr.ctypes.data_as(dp), MG.CAP)
    return ret, f.value, x, ev.value, it.value, tr


def _check(gold, key, res):
    ret, f, x, ev, it, tr = res
    assert ret == int(gold[key + "_ret"])
    assert f == float(gold[key + "_f"])
    np.testing.assert_array_equal(x, gold[key + "_x"])
    assert ev == int(gold[key + "_ev"])
    k = max(it - 1, 0) if ret < 0 else it      # the reference's progress callback is not invoked on a failed search
    np.testing.assert_array_equal(tr[: min(k, MG.CAP)], gold[key + "_tr"][: min(k, MG.CAP)])


def test_analytic_functions_match_reference_header(oracle):
    gold = np.load(GOLD)
    L = oracle.lib()
    tf = C.cast(L.oracle_testfunc_fg, C.c_void_p)
    for fid, n in MG.ANALYTIC:
        for si, (mem, past, delta, geps) in enumerate(MG.LB_SETTINGS):
            prm = capi.default_params()
            prm.lbfgs_mem_size, prm.lbfgs_past, prm.lbfgs_delta, prm.lbfgs_g_epsilon = mem, past, delta, geps
            idv = C.c_int(fid)
            res = _run_oracle(oracle, n, MG.analytic_x0(n), tf, C.cast(C.byref(idv), C.c_void_p), prm)
            _check(gold, f"an_{fid}_{n}_{si}", res)
    # sanity on the optimiser itself: Rosenbrock and the quadratic reach their minima with the tight setting
    prm = capi.default_params()
    prm.lbfgs_past, prm.lbfgs_delta, prm.lbfgs_g_epsilon = 0, 0.0, 1e-10
    idv = C.c_int(0)
    ret, f, x, *_ = _run_oracle(oracle, 10, MG.analytic_x0(10), tf, C.cast(C.byref(idv), C.c_void_p), prm)
    assert ret == 0 and f < 1e-15 and np.allclose(x, 1.0, atol=1e-6)


def test_trajectory_problems_match_reference_header(oracle):
    """The reference's own lbfgs_optimize loop driving the oracle's cost function == the oracle's loop (all configs)."""
    gold = np.load(GOLD)
    L = oracle.lib()
    for cfg in (1, 2, 3, 4, 5):
        batch, prm = synth.make_config(cfg, B=3, seed=7)
        cb = batch.as_c()
        for p in range(batch.B):
            h = L.oracle_problem_create(C.byref(prm), C.byref(cb), p)
            n = L.oracle_problem_n(h)
            x0 = np.zeros(n)
            L.oracle_problem_initial_x(h, x0.ctypes.data_as(dp))
            res = _run_oracle(oracle, n, x0, C.cast(L.oracle_problem_fg, C.c_void_p), h, prm)
            L.oracle_problem_free(h)
            _check(gold, f"pr_{cfg}_{p}", res)


def test_live_reference_header_build(oracle):
    """Where oracle/_ref exists (this container, or shipped prebuilt) re-run the reference header itself."""
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref/libref_lbfgs.so not available")
    gold = np.load(GOLD)
    L = oracle.lib()
    tf = C.cast(L.oracle_testfunc_fg, C.c_void_p)
    for fid, n in MG.ANALYTIC[:2]:
        mem, past, delta, geps = MG.LB_SETTINGS[0]
        idv = C.c_int(fid)
        ret, f, x, ev, tr = MG.run_ref(n, MG.analytic_x0(n), tf, C.cast(C.byref(idv), C.c_void_p), mem, past, delta, geps)
        key = f"an_{fid}_{n}_0"
        assert ret == int(gold[key + "_ret"]) and f == float(gold[key + "_f"])
        np.testing.assert_array_equal(x, gold[key + "_x"])
