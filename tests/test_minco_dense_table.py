"""Host logic of the dense MINCO path (no GPU): the table M = A^-1[:, R] the kernels multiply with, computed by the
library in extended precision, against numpy's dense inverse of the time-normalised collocation matrix
(synth.minco_dense builds the same matrix the oracle factorises), and the two identities the kernels rely on:
b~ = M b_R for a right-hand side supported on R, and lambda_R = M^T g for the adjoint solve."""
import ctypes as C

import numpy as np
import pytest

from dftpav_b200 import capi, synth


def _table(N):
    lib = capi.load_library()
    lib.dftpav_debug_minco_dense.restype = C.c_longlong
    lib.dftpav_debug_minco_dense.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    cnt = lib.dftpav_debug_minco_dense(N, None, None)
    assert cnt >= (N + 5) * (6 * N + 1)
    buf = np.zeros(cnt)
    rows = np.zeros(N + 5, dtype=np.int32)
    assert lib.dftpav_debug_minco_dense(N, buf.ctypes.data_as(C.POINTER(C.c_double)), rows.ctypes.data_as(C.POINTER(C.c_int))) == cnt
    ld = 6 * N + 1
    Mt = buf[: (N + 5) * ld].reshape(N + 5, ld)[:, : 6 * N]
    return Mt.T.copy(), rows  # M [6N][N+5]


@pytest.mark.parametrize("N", [1, 2, 3, 8, 24, 40])
def test_dense_table_is_the_inverse_restricted_to_rhs_rows(N):
    M, rows = _table(N)
    A = synth.minco_dense(N)
    expect_rows = [0, 1, 2] + [6 * i + 5 for i in range(N - 1)] + [6 * N - 3, 6 * N - 2, 6 * N - 1]
    assert rows.tolist() == expect_rows
    # A M = E_R  (columns of the identity at R), to a few ulps of cond(A) * eps
    E = np.zeros((6 * N, N + 5))
    E[rows, np.arange(N + 5)] = 1.0
    res = np.abs(A @ M - E).max()
    assert res < 1e-11, res
    rng = np.random.default_rng(N)
    bR = rng.normal(size=(N + 5, 2)) * 10
    b = np.zeros((6 * N, 2))
    b[rows] = bR
    c = np.linalg.solve(A, b)
    assert np.abs(M @ bR - c).max() <= 1e-11 * max(1.0, np.abs(c).max())
    g = rng.normal(size=(6 * N, 2))
    lam = np.linalg.solve(A.T, g)
    assert np.abs(M.T @ g - lam[rows]).max() <= 1e-11 * max(1.0, np.abs(lam).max())


def test_bad_piece_count_is_rejected():
    lib = capi.load_library()
    lib.dftpav_debug_minco_dense.restype = C.c_longlong
    lib.dftpav_debug_minco_dense.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    assert lib.dftpav_debug_minco_dense(0, None, None) == -1
    assert lib.dftpav_debug_minco_dense(65, None, None) == -1


def test_band_lu_factor_is_bit_identical_to_the_oracle():
    """The host factorises the collocation matrix row by row (dftpav_api.cu, BandLU); every entry must carry the same bits
    as BandedSystem::factorizeLU's column-oriented loop nest (poly_traj_utils.hpp:776-800, restated in the oracle), which
    the strict verification build and the large-N fallback rely on."""
    import ctypes as C

    from dftpav_b200 import capi
    from oracle import oracle_py

    lib = capi.load_library()
    lib.dftpav_debug_minco_band.argtypes = [C.c_int, capi.c_double_p]
    lib.dftpav_debug_minco_band.restype = C.c_longlong
    O = oracle_py.lib()
    O.oracle_minco_matrix.argtypes = [C.c_int, capi.c_double_p, capi.c_double_p]
    for N in (2, 3, 8, 16, 24, 33, 64):
        n = 6 * N
        packed = np.zeros((n, 16))
        assert lib.dftpav_debug_minco_band(N, packed.ctypes.data_as(capi.c_double_p)) == n * 16
        dense = np.zeros((n, n))
        band = np.zeros((13, n))
        assert O.oracle_minco_matrix(N, dense.ctypes.data_as(capi.c_double_p), band.ctypes.data_as(capi.c_double_p)) == 0
        for j in range(n):
            for r in range(13):
                i = j + r - 6
                if 0 <= i < n:
                    assert packed[j, r] == band[r, j], (N, i, j, packed[j, r], band[r, j])
