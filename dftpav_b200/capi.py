"""ctypes mirror of include/dftpav_b200.h (the C-ABI drop-in boundary) and the loader of the CUDA library.

The product library is ``dftpav_b200/csrc/libdftpav_b200.so`` (built in-tree by ``__graft_entry__.build()``).
There is NO CPU fallback: if the library is missing or no sm_100 GPU is present, calls fail loudly.

Reference interface replaced: ``plan_manage::PolyTrajOptimizer``
(src/Plan/traj_planner/include/plan_manage/traj_optimizer.h:24-248).
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 1

OK = 0
ERR_INVALID_ARG = -1
ERR_BAD_INPUT = -2
ERR_CUDA = -3
ERR_NO_DEVICE = -4
ERR_STATE = -5
ERR_UNSUPPORTED = -6

# LBFGS.hpp:135-184
LBFGS_CONVERGENCE = 0
LBFGS_STOP = 1
LBFGS_CANCELED = 2
LBFGSERR_UNKNOWNERROR = -1024
LBFGSERR_INVALID_N = -1023
LBFGSERR_INVALID_MEMSIZE = -1022
LBFGSERR_INVALID_GEPSILON = -1021
LBFGSERR_INVALID_TESTPERIOD = -1020
LBFGSERR_INVALID_DELTA = -1019
LBFGSERR_INVALID_MINSTEP = -1018
LBFGSERR_INVALID_MAXSTEP = -1017
LBFGSERR_INVALID_FDECCOEFF = -1016
LBFGSERR_INVALID_SCURVCOEFF = -1015
LBFGSERR_INVALID_MACHINEPREC = -1014
LBFGSERR_INVALID_MAXLINESEARCH = -1013
LBFGSERR_INVALID_FUNCVAL = -1012
LBFGSERR_MINIMUMSTEP = -1011
LBFGSERR_MAXIMUMSTEP = -1010
LBFGSERR_MAXIMUMLINESEARCH = -1009
LBFGSERR_MAXIMUMITERATION = -1008
LBFGSERR_WIDTHTOOSMALL = -1007
LBFGSERR_INVALIDPARAMETERS = -1006
LBFGSERR_INCREASEGRADIENT = -1005
STATE_DOUBLES = 12
SAMPLE_CLAMP = 1
SAMPLE_GETSTATE = 2
WHEEL_BASE = 2.85  # VehicleParam::wheel_base (semantics.h:68)
PROBLEM_REJECTED = -2000  # precondition of OptimizeTrajectory violated (traj_optimizer.cpp:26-48), per problem

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class Params(C.Structure):
    """struct dftpav_params (include/dftpav_b200.h)."""

    _fields_ = [
        ("abi_version", C.c_int32),
        ("traj_resolution", C.c_int32),
        ("des_traj_resolution", C.c_int32),
        ("gear_opt", C.c_int32),
        ("wei_obs", C.c_double),
        ("wei_surround", C.c_double),
        ("wei_feas", C.c_double),
        ("wei_time", C.c_double),
        ("surround_clearance", C.c_double),
        ("half_margin", C.c_double),
        ("veh_width", C.c_double),
        ("veh_length", C.c_double),
        ("veh_d_cr", C.c_double),
        ("max_forward_vel", C.c_double),
        ("max_backward_vel", C.c_double),
        ("max_forward_cur", C.c_double),
        ("max_backward_cur", C.c_double),
        ("max_forward_acc", C.c_double),
        ("max_backward_acc", C.c_double),
        ("non_sinv", C.c_double),
        ("mini_T", C.c_double),
        ("help_eps", C.c_double),
        ("min_speed", C.c_double),
        ("smooth_eps", C.c_double),
        ("cur_weight_scale", C.c_double),
        ("lse_alpha", C.c_double),
        ("obs_cull_factor", C.c_double),
        ("cost_limit", C.c_double),
        ("lbfgs_mem_size", C.c_int32),
        ("lbfgs_past", C.c_int32),
        ("lbfgs_max_iterations", C.c_int32),
        ("lbfgs_max_linesearch", C.c_int32),
        ("lbfgs_delta", C.c_double),
        ("lbfgs_g_epsilon", C.c_double),
        ("lbfgs_min_step", C.c_double),
        ("lbfgs_max_step", C.c_double),
        ("lbfgs_f_dec_coeff", C.c_double),
        ("lbfgs_s_curv_coeff", C.c_double),
        ("lbfgs_cautious_factor", C.c_double),
        ("lbfgs_machine_prec", C.c_double),
    ]


def default_params() -> Params:
    """Reference defaults: minco_config.pb.txt:65-100, traj_optimizer.cpp:127-134, traj_optimizer.h:68,
    VehicleParam defaults (Sim/core/common/inc/common/basics/semantics.h:66-76).  Pure-Python twin of
    ``dftpav_default_params`` so that CPU-only tests do not need the CUDA library."""
    p = Params()
    p.abi_version = ABI_VERSION
    p.traj_resolution = 16
    p.des_traj_resolution = 32
    p.gear_opt = 1
    p.wei_obs = 1000.0
    p.wei_surround = 5000.0
    p.wei_feas = 2500.0
    p.wei_time = 500.0
    p.surround_clearance = 0.4
    p.half_margin = 0.15
    p.veh_width = 1.90
    p.veh_length = 4.88
    p.veh_d_cr = 1.015
    p.max_forward_vel = 5.0
    p.max_backward_vel = 2.0
    p.max_forward_cur = 1.0
    p.max_backward_cur = 1.0
    p.max_forward_acc = 8.0
    p.max_backward_acc = 4.0
    p.non_sinv = 0.24
    p.mini_T = 0.1
    p.help_eps = 0.0
    p.min_speed = 1.0e-4
    p.smooth_eps = 1.0e-4
    p.cur_weight_scale = 10.0
    p.lse_alpha = 100.0
    p.obs_cull_factor = 1.5
    p.cost_limit = 50000.0
    p.lbfgs_mem_size = 256
    p.lbfgs_past = 3
    p.lbfgs_max_iterations = 12000
    p.lbfgs_max_linesearch = 64
    p.lbfgs_delta = 1.0e-4
    p.lbfgs_g_epsilon = 1.0e-16
    p.lbfgs_min_step = 1.0e-32
    p.lbfgs_max_step = 1.0e20
    p.lbfgs_f_dec_coeff = 1.0e-4
    p.lbfgs_s_curv_coeff = 0.9
    p.lbfgs_cautious_factor = 1.0e-6
    p.lbfgs_machine_prec = 1.0e-16
    return p


class Batch(C.Structure):
    """struct dftpav_batch."""

    _fields_ = [
        ("B", C.c_int32),
        ("uniform_H", C.c_int32),
        ("seg_ptr", c_int32_p),
        ("piece_num", c_int32_p),
        ("singul", c_int32_p),
        ("ini_state", c_double_p),
        ("fin_state", c_double_p),
        ("init_T", c_double_p),
        ("inner_pts", c_double_p),
        ("plane_cnt", c_int32_p),
        ("planes", c_double_p),
        ("t_now", c_double_p),
        ("obs_ptr", c_int32_p),
        ("obs_idx", c_int32_p),
        ("n_obs_pool", C.c_int32),
        ("obs_start_time", c_double_p),
        ("obs_duration", c_double_p),
        ("obs_piece_ptr", c_int32_p),
        ("obs_piece_dur", c_double_p),
        ("obs_coeffs", c_double_p),
    ]


class GridMap(C.Structure):
    """struct dftpav_gridmap (mirror of common::GridMapND<uint8_t, 2>)."""

    _fields_ = [("size_x", C.c_int32), ("size_y", C.c_int32), ("origin_x", C.c_double), ("origin_y", C.c_double),
                ("resolution", C.c_double), ("data", C.POINTER(C.c_uint8)), ("occupied_value", C.c_uint8)]


class Frontend(C.Structure):
    """struct dftpav_frontend: the searched paths of KinoAstar after getKinoNode (front-end resampling, row N3)."""

    _fields_ = [("n_paths", C.c_int32), ("pt_ptr", C.POINTER(C.c_int32)), ("pts", C.POINTER(C.c_double)),
                ("shot_ptr", C.POINTER(C.c_int32)), ("shot_index", C.POINTER(C.c_int32)), ("shot_time", C.POINTER(C.c_double)),
                ("shot_length", C.POINTER(C.c_double)), ("shot_s", C.POINTER(C.c_int32)), ("start_vel", C.POINTER(C.c_double)),
                ("end_vel", C.POINTER(C.c_double)), ("max_forward_vel", C.c_double), ("max_forward_acc", C.c_double),
                ("max_backward_vel", C.c_double), ("max_backward_acc", C.c_double), ("non_siguav", C.c_double)]


class CorridorParams(C.Structure):
    _fields_ = [("veh_width", C.c_double), ("veh_length", C.c_double), ("veh_d_cr", C.c_double), ("limit_bound", C.c_double)]


def default_corridor_params() -> "CorridorParams":
    """The default common::VehicleParam getRectangleConst expands from (semantics.h:66-76) and its 10 m bound."""
    return CorridorParams(1.90, 4.88, 1.015, 10.0)


OCCUPIED = 80  # GridMap2D::OCCUPIED (semantics.h:354)


class Dims(C.Structure):
    _fields_ = [
        ("B", C.c_int32),
        ("n_stride", C.c_int32),
        ("row_stride", C.c_int32),
        ("seg_total", C.c_int32),
        ("sample_total", C.c_int64),
    ]


class Result(C.Structure):
    _fields_ = [
        ("x", c_double_p),
        ("cost", c_double_p),
        ("status", c_int32_p),
        ("success", c_int32_p),
        ("iters", c_int32_p),
        ("evals", c_int32_p),
        ("coeffs", c_double_p),
        ("piece_dT", c_double_p),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("upload_ms", C.c_double),
        ("solve_ms", C.c_double),
        ("download_ms", C.c_double),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("total_iters", C.c_int64),
        ("total_evals", C.c_int64),
        ("algorithmic_bytes", C.c_int64),
        ("kernel_launches", C.c_int32),
        ("n_shapes", C.c_int32),
    ]


# every symbol include/dftpav_b200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "dftpav_default_params",
    "dftpav_batch_dims",
    "dftpav_batch_check",
    "dftpav_initial_x",
    "dftpav_create",
    "dftpav_destroy",
    "dftpav_set_stream",
    "dftpav_upload_batch",
    "dftpav_get_dims",
    "dftpav_eval",
    "dftpav_optimize",
    "dftpav_download",
    "dftpav_solve",
    "dftpav_success_count",
    "dftpav_sample",
    "dftpav_sample_uniform",
    "dftpav_pack_polytraj",
    "dftpav_minco_fit",
    "dftpav_state_to_flat_output",
    "dftpav_corridor_set_map",
    "dftpav_corridor_rectangles",
    "dftpav_collision_check",
    "dftpav_frontend_evaluate_pos",
    "dftpav_frontend_resample",
    "dftpav_trapezoid_duration",
    "dftpav_create_multi",
    "dftpav_destroy_multi",
    "dftpav_solve_multi",
    "dftpav_multi_uses_nccl",
    "dftpav_multi_devices",
    "dftpav_multi_stats",
    "dftpav_multi_last_error",
    "dftpav_last_stats",
    "dftpav_last_error",
    "dftpav_version",
]

_LIBS = {}
_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(_CSRC, "libdftpav_b200.so")
# the strict-order VERIFICATION build of the same sources (-DDFTPAV_STRICT=1 -fmad=false, dftpav_strict.cuh): complete solves
# bit-identical to the CPU oracle; used by tests/test_gpu_strict.py only
STRICT_LIB_PATH = os.path.join(_CSRC, "libdftpav_b200_strict.so")


def load_library(variant: str = "") -> C.CDLL:
    """Loads the in-tree CUDA library (variant "" = the product, "strict" = the verification build, or a path);
    raises (never falls back) if it has not been built."""
    if variant in _LIBS:
        return _LIBS[variant]
    path = LIB_PATH if variant == "" else (STRICT_LIB_PATH if variant == "strict" else variant)
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the dftpav_b200 hot path."
        )
    lib = C.CDLL(path)
    H = C.c_void_p
    lib.dftpav_default_params.argtypes = [C.POINTER(Params)]
    lib.dftpav_default_params.restype = None
    lib.dftpav_batch_dims.argtypes = [C.POINTER(Params), C.POINTER(Batch), C.POINTER(Dims)]
    lib.dftpav_batch_check.argtypes = [C.POINTER(Params), C.POINTER(Batch), c_int32_p]
    lib.dftpav_initial_x.argtypes = [C.POINTER(Params), C.POINTER(Batch), c_double_p]
    lib.dftpav_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(H)]
    lib.dftpav_destroy.argtypes = [H]
    lib.dftpav_set_stream.argtypes = [H, C.c_void_p]
    lib.dftpav_upload_batch.argtypes = [H, C.POINTER(Batch)]
    lib.dftpav_get_dims.argtypes = [H, C.POINTER(Dims)]
    lib.dftpav_eval.argtypes = [H, c_double_p, c_double_p, c_double_p]
    lib.dftpav_optimize.argtypes = [H]
    lib.dftpav_download.argtypes = [H, C.POINTER(Result)]
    lib.dftpav_solve.argtypes = [H, C.POINTER(Batch), C.POINTER(Result)]
    lib.dftpav_success_count.argtypes = [H, c_int32_p]
    lib.dftpav_sample.argtypes = [H, C.c_int64, c_int32_p, c_double_p, C.c_int32, C.c_double, c_double_p]
    lib.dftpav_sample_uniform.argtypes = [H, C.c_double, C.c_int32, c_double_p, c_int32_p]
    lib.dftpav_pack_polytraj.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.dftpav_minco_fit.argtypes = [H, C.c_int32, c_int32_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]
    lib.dftpav_state_to_flat_output.argtypes = [c_double_p, c_double_p]
    lib.dftpav_state_to_flat_output.restype = None
    lib.dftpav_corridor_set_map.argtypes = [H, C.POINTER(GridMap)]
    lib.dftpav_corridor_rectangles.argtypes = [H, C.POINTER(CorridorParams), C.c_int64, c_double_p, c_double_p, c_double_p,
                                               C.c_int32]
    lib.dftpav_frontend_evaluate_pos.argtypes = [H, C.POINTER(Frontend), C.c_int64, c_int32_p, c_double_p, c_double_p]
    lib.dftpav_frontend_resample.argtypes = [H, C.POINTER(Frontend), c_double_p, C.c_double, C.c_int32, C.c_int32, c_int32_p,
                                             c_double_p, C.c_int64, c_double_p, c_double_p, C.POINTER(C.c_int64)]
    lib.dftpav_trapezoid_duration.argtypes = [C.c_double] * 5
    lib.dftpav_trapezoid_duration.restype = C.c_double
    lib.dftpav_collision_check.argtypes = [H, C.POINTER(CorridorParams), C.c_double, C.c_double, C.c_int32, c_int32_p, c_int32_p]
    lib.dftpav_create_multi.argtypes = [C.POINTER(Params), c_int32_p, C.c_int32, C.POINTER(H)]
    lib.dftpav_destroy_multi.argtypes = [H]
    lib.dftpav_solve_multi.argtypes = [H, C.POINTER(Batch), C.POINTER(Result), c_int32_p]
    lib.dftpav_multi_uses_nccl.argtypes = [H]
    lib.dftpav_multi_devices.argtypes = [H]
    lib.dftpav_multi_stats.argtypes = [H, C.c_int32, C.POINTER(Stats)]
    lib.dftpav_multi_last_error.argtypes = [H]
    lib.dftpav_multi_last_error.restype = C.c_char_p
    lib.dftpav_last_stats.argtypes = [H, C.POINTER(Stats)]
    lib.dftpav_last_error.argtypes = [H]
    lib.dftpav_last_error.restype = C.c_char_p
    lib.dftpav_version.argtypes = []
    lib.dftpav_version.restype = C.c_char_p
    lib.dftpav_debug_force_banded.argtypes = [H, C.c_int]
    lib.dftpav_debug_build_flags.argtypes = []
    for name in EXPORTED_SYMBOLS:
        getattr(lib, name)
    _LIBS[variant] = lib
    return lib
