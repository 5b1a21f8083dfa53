"""The C-ABI library loads and exports every symbol include/dftpav_b200.h declares; host-only entry points work
without a GPU; with no sm_100 device the solver refuses to run (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dftpav_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return capi.load_library()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "dftpav_b200.h")).read()
    declared = set(re.findall(r"\b(dftpav_[a-z_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_params_struct_layout_matches(lib):
    p = capi.Params()
    lib.dftpav_default_params(C.byref(p))
    q = capi.default_params()
    for name, _ in capi.Params._fields_:
        assert getattr(p, name) == getattr(q, name), name
    assert p.lbfgs_mem_size == 256 and p.non_sinv == 0.24 and p.traj_resolution == 16


def test_dims_and_initial_x_host_only(lib, oracle):
    for cfg in (1, 2, 3, 4, 5):
        batch, prm = synth.make_config(cfg, B=3, seed=1)
        cb = batch.as_c()
        d = capi.Dims()
        assert lib.dftpav_batch_dims(C.byref(prm), C.byref(cb), C.byref(d)) == capi.OK
        assert d.B == 3 and d.n_stride == batch.n_stride() and d.row_stride == batch.row_stride()
        assert d.sample_total == batch.total_samples(prm.traj_resolution, prm.des_traj_resolution)
        x0 = np.zeros((d.B, d.n_stride))
        assert lib.dftpav_initial_x(C.byref(prm), C.byref(cb), x0.ctypes.data_as(capi.c_double_p)) == capi.OK
        for p in range(3):
            np.testing.assert_array_equal(x0[p, : batch.n_of(p)], oracle.initial_x(prm, batch, p))


def test_preconditions_rejected(lib):
    batch, prm = synth.make_config(3, B=2, seed=0)
    d = capi.Dims()
    batch.init_T[1] = 0.05  # < mini_T  (traj_optimizer.cpp:30-33)
    cb = batch.as_c()
    # a problem the reference would refuse is flagged, not an error of the batch (per-problem status at solve time)
    assert lib.dftpav_batch_dims(C.byref(prm), C.byref(cb), C.byref(d)) == capi.OK
    rej = np.zeros(2, dtype=np.int32)
    assert lib.dftpav_batch_check(C.byref(prm), C.byref(cb), rej.ctypes.data_as(capi.c_int32_p)) == capi.OK
    assert rej.tolist() == [0, 1]
    batch.init_T[1] = 24.0
    prm.abi_version = 99
    cb = batch.as_c()
    assert lib.dftpav_batch_dims(C.byref(prm), C.byref(cb), C.byref(d)) == capi.ERR_INVALID_ARG
    prm.abi_version = capi.ABI_VERSION
    prm.lbfgs_mem_size = 1000
    assert lib.dftpav_batch_dims(C.byref(prm), C.byref(cb), C.byref(d)) == capi.ERR_UNSUPPORTED


def test_no_cpu_fallback(lib):
    """Without an sm_100 GPU dftpav_create must fail: the product path never runs on the CPU."""
    import torch

    prm = capi.default_params()
    h = C.c_void_p()
    rc = lib.dftpav_create(C.byref(prm), -1, C.byref(h))
    if torch.cuda.is_available():
        assert rc == capi.OK
        lib.dftpav_destroy(h)
    else:
        assert rc == capi.ERR_NO_DEVICE and not h.value


def test_product_does_not_import_oracle():
    """The product package must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "dftpav_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(base, f)).read()
                assert "oracle_py" not in txt and "dftpav_oracle" not in txt and "from oracle" not in txt, (base, f)
