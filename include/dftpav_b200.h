/*
 * dftpav_b200.h — C-ABI of the B200-native batched MINCO / L-BFGS trajectory optimizer.
 *
 * This is the drop-in boundary for ONE hot path of ZJU-FAST-Lab/Dftpav: the back-end
 * optimizer `plan_manage::PolyTrajOptimizer` (reference
 * src/Plan/traj_planner/include/plan_manage/traj_optimizer.h:24-248, implementation
 * src/Plan/traj_planner/src/traj_optimizer.cpp).  Every entry point below states the reference
 * interface it replaces.  All arrays are plain caller-owned host buffers of IEEE fp64 / int32 /
 * int64; there are no C++ or torch types in any signature.
 *
 * The same structs are consumed by the CPU oracle (oracle/dftpav_oracle.h), which is test
 * infrastructure only.
 *
 * Shorthand in the comments:  OPT.cpp = src/Plan/traj_planner/src/traj_optimizer.cpp,
 * OPT.h = include/plan_manage/traj_optimizer.h, MINCO.hpp = include/plan_utils/poly_traj_utils.hpp,
 * LBFGS.hpp = include/geo_utils2d/lbfgs.hpp, CFG = config/minco_config.pb.txt (all under
 * src/Plan/traj_planner/ of the reference).
 */
#ifndef DFTPAV_B200_H
#define DFTPAV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFTPAV_ABI_VERSION 1

/* ---- return codes of the API functions themselves (not of the solver) ---- */
#define DFTPAV_OK 0
#define DFTPAV_ERR_INVALID_ARG (-1)   /* NULL pointer, negative size, inconsistent CSR arrays       */
#define DFTPAV_ERR_BAD_INPUT (-2)     /* reserved (round 1: a problem violating OPT.cpp:26-48 failed the whole batch;
                                         now reported per problem as status DFTPAV_PROBLEM_REJECTED)  */
#define DFTPAV_ERR_CUDA (-3)          /* CUDA runtime failure; see dftpav_last_error()               */
#define DFTPAV_ERR_NO_DEVICE (-4)     /* no usable sm_100 device: there is NO CPU fallback           */
#define DFTPAV_ERR_STATE (-5)         /* call order violated (e.g. optimize before upload)           */
#define DFTPAV_ERR_UNSUPPORTED (-6)   /* shape exceeds a compiled-in limit (see dftpav_limits)       */

/* ---- per-problem solver status: the enum of LBFGS.hpp:135-184, same numeric values ---- */
#define DFTPAV_LBFGS_CONVERGENCE 0
#define DFTPAV_LBFGS_STOP 1
#define DFTPAV_LBFGS_CANCELED 2
#define DFTPAV_LBFGSERR_UNKNOWNERROR (-1024)
#define DFTPAV_LBFGSERR_INVALID_N (-1023)
#define DFTPAV_LBFGSERR_INVALID_MEMSIZE (-1022)
#define DFTPAV_LBFGSERR_INVALID_GEPSILON (-1021)
#define DFTPAV_LBFGSERR_INVALID_TESTPERIOD (-1020)
#define DFTPAV_LBFGSERR_INVALID_DELTA (-1019)
#define DFTPAV_LBFGSERR_INVALID_MINSTEP (-1018)
#define DFTPAV_LBFGSERR_INVALID_MAXSTEP (-1017)
#define DFTPAV_LBFGSERR_INVALID_FDECCOEFF (-1016)
#define DFTPAV_LBFGSERR_INVALID_SCURVCOEFF (-1015)
#define DFTPAV_LBFGSERR_INVALID_MACHINEPREC (-1014)
#define DFTPAV_LBFGSERR_INVALID_MAXLINESEARCH (-1013)
#define DFTPAV_LBFGSERR_INVALID_FUNCVAL (-1012)
#define DFTPAV_LBFGSERR_MINIMUMSTEP (-1011)
#define DFTPAV_LBFGSERR_MAXIMUMSTEP (-1010)
#define DFTPAV_LBFGSERR_MAXIMUMLINESEARCH (-1009)
#define DFTPAV_LBFGSERR_MAXIMUMITERATION (-1008)
#define DFTPAV_LBFGSERR_WIDTHTOOSMALL (-1007)
#define DFTPAV_LBFGSERR_INVALIDPARAMETERS (-1006)
#define DFTPAV_LBFGSERR_INCREASEGRADIENT (-1005)
/* Not an L-BFGS code: the problem violates a precondition of OptimizeTrajectory itself (a segment with a single piece,
 * initTs < mini_T: OPT.cpp:26-48 `return false`).  It is not solved; success = 0, every other output of it is zero. The
 * other problems of the batch are unaffected. */
#define DFTPAV_PROBLEM_REJECTED (-2000)

/*
 * dftpav_params — every number that reaches the arithmetic, explicit (no hidden constants).
 * Replaces: PolyTrajOptimizer::setParam (OPT.cpp:1713-1781, fields of planning::minco::Config /
 * CFG:65-100), the hard-coded members of OPT.h:68-74 and the L-BFGS parameter block of
 * OPT.cpp:127-134 / LBFGS.hpp:15-129.  dftpav_default_params() fills the reference defaults.
 */
typedef struct dftpav_params {
  int32_t abi_version;          /* must be DFTPAV_ABI_VERSION                                        */
  int32_t traj_resolution;      /* K   constraint samples per inner piece      CFG:66  (16)          */
  int32_t des_traj_resolution;  /* Kd  samples of first/last piece             CFG:67  (32)          */
  int32_t gear_opt;             /* GearOpt                                      CFG:96  (1)          */
  double wei_obs;               /* wei_sta_obs   1000                           OPT.cpp:1717          */
  double wei_surround;          /* wei_dyn_obs   5000                           OPT.cpp:1718          */
  double wei_feas;              /* 2500                                         OPT.cpp:1719          */
  double wei_time;              /* 500                                          OPT.cpp:1721          */
  double surround_clearance;    /* dyn_obs_clearance 0.4                        OPT.cpp:1722          */
  double half_margin;           /* 0.15: vehicle box inflated by 2*half_margin  OPT.cpp:1749-1750     */
  double veh_width;             /* raw VehicleParam 1.90  (Sim/core/common semantics.h:66)           */
  double veh_length;            /* 4.88                                                              */
  double veh_d_cr;              /* 1.015 rear axle -> geometric centre                                */
  double max_forward_vel, max_backward_vel;   /* 5, 2      OPT.cpp:1725-1726                          */
  double max_forward_cur, max_backward_cur;   /* 1, 1      OPT.cpp:1727-1728                          */
  double max_forward_acc, max_backward_acc;   /* 8, 4      OPT.cpp:1729-1730                          */
  double non_sinv;              /* 0.24 speed at gear-switch points             OPT.h:68              */
  double mini_T;                /* 0.1                                          CFG:100               */
  double help_eps;              /* curvature denominator epsilon; the manager passes 0.0
                                   (traj_manager.cpp:610), header default 1e-4 (OPT.h:120)           */
  double min_speed;             /* 1e-4: samples with |sigma'| below are skipped OPT.cpp:550          */
  double smooth_eps;            /* 1e-4: width of positiveSmoothedL1             OPT.cpp:785          */
  double cur_weight_scale;      /* 10.0: curvature penalty weight factor         OPT.cpp:690-692      */
  double lse_alpha;             /* 100.0: log-sum-exp sharpness                  OPT.cpp:1336         */
  double obs_cull_factor;       /* 1.5: cull obstacle if farther than factor*length OPT.cpp:1393     */
  double cost_limit;            /* 50000: solve is a failure if final cost >=    OPT.cpp:197          */
  /* L-BFGS (LBFGS.hpp:15-129 with the overrides of OPT.cpp:127-134) */
  int32_t lbfgs_mem_size;       /* 256   CFG:97                                                      */
  int32_t lbfgs_past;           /* 3     CFG:98                                                      */
  int32_t lbfgs_max_iterations; /* 12000 OPT.cpp:134                                                 */
  int32_t lbfgs_max_linesearch; /* 64    LBFGS.hpp:76                                                */
  double lbfgs_delta;           /* 1e-4  CFG:99                                                      */
  double lbfgs_g_epsilon;       /* 1e-16 OPT.cpp:130                                                 */
  double lbfgs_min_step;        /* 1e-32 OPT.cpp:132                                                 */
  double lbfgs_max_step;        /* 1e20  LBFGS.hpp:94                                                */
  double lbfgs_f_dec_coeff;     /* 1e-4  LBFGS.hpp:101                                               */
  double lbfgs_s_curv_coeff;    /* 0.9   LBFGS.hpp:112                                               */
  double lbfgs_cautious_factor; /* 1e-6  LBFGS.hpp:121                                               */
  double lbfgs_machine_prec;    /* 1e-16 LBFGS.hpp:128                                               */
} dftpav_params;

/*
 * dftpav_batch — B independent OptimizeTrajectory problems, flattened.
 * One problem = the argument list of PolyTrajOptimizer::OptimizeTrajectory (OPT.h:118-120,
 * OPT.cpp:7-10) plus the obstacle set given to setSurroundTrajs (OPT.h:108).
 *
 * Segments ("gear segments", trajnum in the reference) of all problems are concatenated:
 * problem p owns segments seg_ptr[p] .. seg_ptr[p+1]-1.  For segment g with N = piece_num[g]:
 *   ini_state/fin_state : 6 doubles, the 2x3 column-major matrix [p v a] of iniStates[i]/finStates[i]
 *   init_T              : initTs(i), total duration of the segment (>= mini_T)
 *   inner_pts           : 2*(N-1) doubles, initInnerPts[i] column-major (x0,y0,x1,y1,...),
 *                         all segments concatenated in order
 *   samples             : M = (N-2)*(K+1) + 2*(Kd+1) corridor polygons (hPoly_container[i][k],
 *                         size check OPT.cpp:44), all segments concatenated in order.
 *                         Sample q has plane_cnt[q] half-planes (or uniform_H if plane_cnt==NULL);
 *                         each half-plane is 4 doubles (n_x, n_y, p_x, p_y) = one column of the
 *                         reference's 4xH hPoly matrix; normals need not be unit (they are
 *                         normalised like OPT.cpp:51).  planes holds them concatenated.
 * Obstacles: problem p uses obstacle indices obs_idx[obs_ptr[p] .. obs_ptr[p+1]-1] into a pool
 * (several problems may share pool entries).  Pool entry u = one plan_utils::LocalTrajData
 * (traj_container.hpp:29-39): start_time, duration and a piecewise quintic with pieces
 * obs_piece_ptr[u] .. obs_piece_ptr[u+1]-1; each piece has a duration and 12 coefficients =
 * the 2x6 column-major Piece::coeffMat, HIGHEST power first (MINCO.hpp:19,77-87).
 * obs_ptr == NULL means "no obstacles" (surround_trajs_ == NULL, OPT.cpp:636).
 */
typedef struct dftpav_batch {
  int32_t B;
  int32_t uniform_H;            /* used when plane_cnt == NULL                                       */
  const int32_t* seg_ptr;       /* [B+1]                                                             */
  const int32_t* piece_num;     /* [G]   G = seg_ptr[B]                                              */
  const int32_t* singul;        /* [G]   +1 forward / -1 reverse                                     */
  const double* ini_state;      /* [G][6]                                                            */
  const double* fin_state;      /* [G][6]                                                            */
  const double* init_T;         /* [G]                                                               */
  const double* inner_pts;      /* [sum 2*(N_g-1)]                                                   */
  const int32_t* plane_cnt;     /* [Q] or NULL;  Q = total samples                                   */
  const double* planes;         /* [sum_q H_q][4]                                                    */
  const double* t_now;          /* [B]  `now` argument of OptimizeTrajectory                         */
  const int32_t* obs_ptr;       /* [B+1] or NULL                                                     */
  const int32_t* obs_idx;       /* [obs_ptr[B]]                                                      */
  int32_t n_obs_pool;
  const double* obs_start_time; /* [n_obs_pool]                                                      */
  const double* obs_duration;   /* [n_obs_pool]                                                      */
  const int32_t* obs_piece_ptr; /* [n_obs_pool+1]                                                    */
  const double* obs_piece_dur;  /* [obs_piece_ptr[n_obs_pool]]                                       */
  const double* obs_coeffs;     /* [obs_piece_ptr[n_obs_pool]][12]                                   */
} dftpav_batch;

/*
 * Output strides.  Decision vector of a problem (OPT.cpp:96-115):
 *   x = [ inner points of seg 0 .. S-1 | tau_0..tau_{S-1} | Gear_0(2)..Gear_{S-2} | theta_0..theta_{S-2} ]
 *   n = sum 2(N_i-1) + S + 3(S-1).
 * Coefficient rows of a problem: sum_i 6*N_i rows, segment after segment; row 6j+k of a segment is
 * the (x,y) coefficient of t^k of piece j (MINCO.hpp:981-984).
 */
typedef struct dftpav_dims {
  int32_t B;
  int32_t n_stride;     /* max n over the batch: stride of x / g arrays                              */
  int32_t row_stride;   /* max coefficient rows over the batch                                       */
  int32_t seg_total;    /* G                                                                         */
  int64_t sample_total; /* Q                                                                         */
} dftpav_dims;

/*
 * dftpav_result — caller-provided output buffers; any pointer may be NULL (then it is skipped).
 * Replaces the read-back through getMinJerkOptPtr() (OPT.h:112) / MinJerkOpt::getTraj,
 * getCoeffs, getDt (MINCO.hpp:987-997,1069-1074) and the bool returned by OptimizeTrajectory.
 */
typedef struct dftpav_result {
  double* x;          /* [B][n_stride]   final decision vector (reverted point on line-search error) */
  double* cost;       /* [B]             final_cost (LBFGS.hpp:748)                                  */
  int32_t* status;    /* [B]             lbfgs return code                                           */
  int32_t* success;   /* [B]             flag_success of OPT.cpp:176-201                             */
  int32_t* iters;     /* [B]             L-BFGS iteration counter k at exit                          */
  int32_t* evals;     /* [B]             cost+gradient evaluations (iter_num_, OPT.cpp:334)          */
  double* coeffs;     /* [B][row_stride][2]  coefficients of the LAST EVALUATED point (quirk Q8)     */
  double* piece_dT;   /* [G]             per-piece duration T_i/N_i of the last evaluated point      */
} dftpav_result;

/* Timings and counters of the last call, device-side (CUDA events on the handle's stream). */
typedef struct dftpav_stats {
  double upload_ms;        /* H2D copies + layout/normalisation kernel                               */
  double solve_ms;         /* the solver (or eval) kernel(s) only                                    */
  double download_ms;      /* D2H copies                                                             */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int64_t total_iters;     /* sum over problems                                                      */
  int64_t total_evals;
  int64_t algorithmic_bytes; /* BASELINE.md section 4 formula with the actual counts                 */
  int32_t kernel_launches; /* kernels of this library launched by the last call                      */
  int32_t n_shapes;        /* distinct problem shapes in the uploaded batch                          */
} dftpav_stats;

typedef struct dftpav_handle dftpav_handle;

/* Fills *p with the reference defaults (CFG:65-100, OPT.cpp:127-134, OPT.h:68). */
void dftpav_default_params(dftpav_params* p);

/* Validates the arrays of a batch (consistent CSR pointers, compiled-in limits) and fills *dims. No device needed. */
int dftpav_batch_dims(const dftpav_params* prm, const dftpav_batch* batch, dftpav_dims* dims);

/* Which problems OptimizeTrajectory itself would refuse (OPT.cpp:26-48: a segment with one piece, initTs < mini_T):
 * rejected[p] = 1 for those, 0 otherwise ([B], may be NULL).  Returns the same codes as dftpav_batch_dims; a rejected problem
 * is NOT an error of the batch (it gets status DFTPAV_PROBLEM_REJECTED when the batch is solved).  No device needed. */
int dftpav_batch_check(const dftpav_params* prm, const dftpav_batch* batch, int32_t* rejected);

/* Packs the initial decision vector x0 of every problem exactly as OPT.cpp:96-115 (including the
 * ini/fin clamps of OPT.cpp:55-76 that feed Gear/theta).  x0 is [B][n_stride].  Host-only helper. */
int dftpav_initial_x(const dftpav_params* prm, const dftpav_batch* batch, double* x0);

/* [setup]  replaces `new PolyTrajOptimizer` + setParam (traj_manager.cpp:28-29).
 * device < 0 selects the current CUDA device.  Fails with DFTPAV_ERR_NO_DEVICE if there is no
 * sm_100 GPU — there is no CPU path behind this ABI. */
int dftpav_create(const dftpav_params* prm, int device, dftpav_handle** out);
int dftpav_destroy(dftpav_handle* h);

/* Optional: run on a caller-owned CUDA stream (cudaStream_t passed as void*). NULL = own stream. */
int dftpav_set_stream(dftpav_handle* h, void* cuda_stream);

/* [setup]  replaces the argument marshalling of OptimizeTrajectory (deep copies at OPT.cpp:15-18),
 * the plane normalisation (OPT.cpp:51), the boundary clamps (OPT.cpp:55-76), MinJerkOpt::reset
 * (OPT.cpp:79) and setSurroundTrajs (OPT.h:108).  Host buffers are copied; they may be freed after
 * the call returns. */
int dftpav_upload_batch(dftpav_handle* h, const dftpav_batch* batch);
int dftpav_get_dims(const dftpav_handle* h, dftpav_dims* dims);

/* Parity hook = one call of costFunctionCallback per problem (OPT.cpp:206-350):
 * x is [B][n_stride] (host), f is [B], g is [B][n_stride]. */
int dftpav_eval(dftpav_handle* h, const double* x, double* f, double* g);

/* [optimize]  replaces lbfgs::lbfgs_optimize driven by OptimizeTrajectory (OPT.cpp:159-166) for the
 * whole uploaded batch; starts from the uploaded initial guess every time it is called.
 * Results stay on the device until dftpav_download. */
int dftpav_optimize(dftpav_handle* h);

/* [getTraj]  copies results of the last dftpav_optimize to the host. */
int dftpav_download(dftpav_handle* h, const dftpav_result* out);

/* Number of problems of the last solve with success == 1 (the local half of the global convergence flag that multi-GPU
 * callers all-reduce: every shard finished and was accepted). */
int dftpav_success_count(dftpav_handle* h, int32_t* n_success);

/* upload + optimize + download in one call (the end-to-end path a planner would use). */
int dftpav_solve(dftpav_handle* h, const dftpav_batch* batch, const dftpav_result* out);

/* ---- read-back: batched evaluation of the solved trajectories, from the device-resident coefficients ----
 * Replaces the per-sample calls consumers make on plan_utils::Trajectory (MINCO.hpp:554-645: getPos, getdSigma, getddSigma,
 * getAngle, getCurv, getVel, getAcc, getLatAcc, getSteer; GetState 378-406) after getMinJerkOptPtr()->at(i).getTraj(singul).
 * A gear segment g (global index: seg_ptr[p] + s) is one Trajectory of piece_num[g] pieces with its singul; times are
 * local to the segment.  Query q evaluates segment seg[q] at t[q]:
 *   out[q] = { x, y, dsigma_x, dsigma_y, ddsigma_x, ddsigma_y, angle, curvature, velocity, acceleration, steer, lat_acc }
 * Without flags the getters' semantics (time not clamped: past the end the last piece is extrapolated, MINCO.hpp:510-528;
 * curvature / acceleration zero when |sigma'| < 1e-6).  DFTPAV_SAMPLE_CLAMP clamps t to the segment duration.
 * DFTPAV_SAMPLE_GETSTATE gives what Trajectory::GetState stores: clamped time and getStateExpPos' threshold, which tests
 * `abs(velocity) < 1e-6` with the unqualified (integer) abs, i.e. curvature, acceleration and steer are zero whenever
 * |velocity| < 1 (quirk Q11).  wheel_base: VehicleParam::wheel_base (2.85, semantics.h:68) for steer = atan(L * curvature).
 * Positions, derivatives, velocity and accelerations are bit-identical to the reference's evaluators; angle, curvature and
 * steer within 1e-12 relative.  Call after dftpav_optimize / dftpav_solve. */
#define DFTPAV_STATE_DOUBLES 12
#define DFTPAV_SAMPLE_CLAMP 1
#define DFTPAV_SAMPLE_GETSTATE 2
int dftpav_sample(dftpav_handle* h, int64_t nq, const int32_t* seg, const double* t, int32_t flags, double wheel_base,
                  double* out);

/* The sampling loop of the collision re-check of the executing trajectory (traj_server_ros.cpp:385-397): every segment at
 * t = 0, dt, 2 dt, ... while t < duration (t accumulated like the reference's `t += 0.05`), at most max_per_seg samples:
 * xyyaw[g][k] = { x, y, yaw } (getPos, getAngle), count[g] = samples written. */
int dftpav_sample_uniform(dftpav_handle* h, double dt, int32_t max_per_seg, double* xyyaw, int32_t* count);

/* The fields of msg/PolyTraj.msg (uint8 order = 5, float32[] coef_x, coef_y, duration) for all segments concatenated in
 * segment order: per piece six coefficients, highest power first like Piece::getCoeffMat (MINCO.hpp:19, 987-997).
 * coef_x, coef_y: [6 * sum piece_num], duration: [sum piece_num]. */
int dftpav_pack_polytraj(dftpav_handle* h, float* coef_x, float* coef_y, float* duration);

/* ---- batched MINCO fits: predicted obstacle trajectories and initial guesses ----
 * Replaces MinJerkOpt::reset + generate + getTraj (MINCO.hpp:880-997) as TrajPlanner::ConverSurroundTrajFromPoints runs them
 * for every predicted obstacle trajectory (traj_manager.cpp:743-789: N = states - 1 pieces of the mean duration, inner points =
 * the intermediate positions, head / tail = state_to_flat_output of the first / last state) and for the initial guess
 * (traj_manager.cpp:579-581).  Trajectory k has piece_ptr[k+1] - piece_ptr[k] >= 2 pieces of duration dT[k], flat boundary
 * states head[k][6] / tail[k][6] (2x3 column-major p v a) and 2*(N_k - 1) inner points (concatenated).  coeffs receives
 * Piece::coeffMat for every piece: [sum N_k][12], 2x6 column-major, HIGHEST power first - directly usable as
 * dftpav_batch.obs_coeffs (with obs_piece_dur = dT per piece, obs_duration = the SUM of the piece durations).
 * Bit-identical to the reference's banded solve (explicitly rounded operations, no libm on the path). */
int dftpav_minco_fit(dftpav_handle* h, int32_t n_traj, const int32_t* piece_ptr, const double* dT, const double* head,
                     const double* tail, const double* inner_pts, double* coeffs);
/* TrajPlanner::state_to_flat_output (traj_manager.cpp:139-158), host only (cos / sin / pow of the C library, like the
 * reference): state = (x, y, angle, velocity, acceleration, curvature) -> flat[6] = 2x3 column-major (p, v, a). */
void dftpav_state_to_flat_output(const double* state, double* flat);

/* ---- corridor generation (the step immediately before the optimizer) ----
 * Replaces TrajPlanner::getRectangleConst (src/traj_manager.cpp:1211-1469): for every sampled state (x, y, yaw) of the
 * front-end path a safe rectangle is grown cell by cell on the occupancy grid - round-robin over +dy, +dx, -dy, -dx, each
 * step ray-marching the three sides of the strip it would add at resolution/2 (TrajPlannerAdapter::CheckIfCollisionUsingLine,
 * src/map_adapter.cpp:117-129) until blocked or limit_bound metres - and returned as four half-planes (n_x, n_y, p_x, p_y):
 * exactly the per-sample polygon layout dftpav_batch.planes takes (uniform_H = 4).
 * dftpav_gridmap mirrors common::GridMapND<uint8_t, 2> (Sim/core/common semantics.h:376-560): cell (cx, cy) of a position p is
 * round((p - origin) / resolution) per axis (std::round), data index cx + cy * size_x (dimension 0 fastest), a cell is an
 * obstacle iff data == occupied_value (OCCUPIED = 80); positions outside the map are free. */
typedef struct dftpav_gridmap {
  int32_t size_x, size_y;
  double origin_x, origin_y, resolution;
  const uint8_t* data;      /* [size_y][size_x] */
  uint8_t occupied_value;   /* 80 */
} dftpav_gridmap;

typedef struct dftpav_corridor_params {
  double veh_width, veh_length, veh_d_cr;  /* the default common::VehicleParam the reference expands from: 1.90, 4.88, 1.015 */
  double limit_bound;                       /* 10.0 (traj_manager.cpp:1218) */
} dftpav_corridor_params;

/* Copies the occupancy grid to the device (replaces the GridMap2D copy of traj_manager.cpp:1213-1214). */
int dftpav_corridor_set_map(dftpav_handle* h, const dftpav_gridmap* map);
/* Rectangles for n states.  states: [n][3] = (x, y, yaw).  cos_sin (may be NULL): [n][2] = (cos yaw, sin yaw) computed by the
 * caller - with it the result is bit-identical to a CPU evaluation that uses the same values (the CUDA and glibc sin / cos
 * differ by an ulp now and then, which can move a ray-march sample across a cell border).  hpolys (may be NULL): [n][4][4]
 * host output.  If keep != 0 the polygons also stay on the device as the corridor of the NEXT dftpav_upload_batch /
 * dftpav_solve whose batch has planes == NULL and uniform_H == 4 (n must then equal the batch's sample total): the
 * 56 KB-per-problem plane upload disappears from the planner cycle. */
int dftpav_corridor_rectangles(dftpav_handle* h, const dftpav_corridor_params* cp, int64_t n, const double* states,
                               const double* cos_sin, double* hpolys, int32_t keep);

/* ---- front-end path -> optimizer inputs (the resampling between the hybrid-A* search and the
 * optimizer) ----
 * dftpav_frontend mirrors what path_searching::KinoAstar keeps after getKinoNode (src/kino_astar.cpp:554-743) for every
 * searched path p: SampleTraj = pts[pt_ptr[p] .. pt_ptr[p+1]) (x, y, yaw); its gear segments ("shots") shot_ptr[p] ..
 * shot_ptr[p+1]: shot_time = shot_timeList, shot_length = shot_lengthList, shot_s = shot_SList (+1 forward, -1 backward),
 * shot_index = shotindex (ns + 1 point indices per path, stored at shot_index[shot_ptr[p] + p ...]); start_vel / end_vel =
 * fabs(start_state_[3]) / fabs(end_state_[3]); the limits are the search's max_forward_vel / _acc, max_backward_vel / _acc
 * and non_siguav. */
typedef struct dftpav_frontend {
  int32_t n_paths;
  const int32_t* pt_ptr;      /* [n_paths + 1] */
  const double* pts;          /* [pt_total][3] */
  const int32_t* shot_ptr;    /* [n_paths + 1] */
  const int32_t* shot_index;  /* [shot_total + n_paths] */
  const double* shot_time;    /* [shot_total] */
  const double* shot_length;  /* [shot_total] */
  const int32_t* shot_s;      /* [shot_total] */
  const double* start_vel;    /* [n_paths] */
  const double* end_vel;      /* [n_paths] */
  double max_forward_vel, max_forward_acc, max_backward_vel, max_backward_acc, non_siguav;
} dftpav_frontend;

/* KinoAstar::evaluatePos (src/kino_astar.cpp:468-521) for nq queries: out[q] = (x, y, yaw) of path path[q] at time t[q]
 * (clamped to [0, total time]): the shot that contains t, the arc length its trapezoidal speed profile has covered
 * (evaluateLength, kino_astar.cpp:763-795), linear interpolation between the path samples around that arc length.
 * Only + - * / sqrt on the path: bit-identical to the CPU restatement. */
int dftpav_frontend_evaluate_pos(dftpav_handle* h, const dftpav_frontend* fe, int64_t nq, const int32_t* path, const double* t,
                                 double* out);
/* The resampling loop of TrajPlanner::RunMINCOParking (src/traj_manager.cpp:531-567) for every path: per gear segment g,
 * piece_num[g] = max(int(seg_duration[g] / traj_piece_duration + 0.5), 2) pieces of piece_dur[g] = seg_duration[g] /
 * piece_num[g] (seg_duration = the summed time steps of the segment's FlatTrajData), the states evaluatePos gives at
 * K + 1 (inner pieces) or Kd + 1 (first / last piece) times per piece - [n_states][3], the statelist getRectangleConst
 * consumes, directly usable by dftpav_corridor_rectangles - and the piece_num - 1 inner waypoints per segment (inner,
 * [.][2], = dftpav_batch.inner_pts).  states_cap = capacity of states in states (inner needs sum (piece_num - 1) pairs);
 * *n_states receives the number of states; DFTPAV_ERR_INVALID_ARG if states_cap is too small (*n_states is still set). */
int dftpav_frontend_resample(dftpav_handle* h, const dftpav_frontend* fe, const double* seg_duration,
                             double traj_piece_duration, int32_t K, int32_t Kd, int32_t* piece_num, double* piece_dur,
                             int64_t states_cap, double* states, double* inner, int64_t* n_states);
/* KinoAstar::evaluateDuration (src/kino_astar.cpp:744-762): duration of the trapezoidal speed profile over `length` (host). */
double dftpav_trapezoid_duration(double length, double max_vel, double max_acc, double start_vel, double end_vel);

/* Collision re-check of the trajectories of the last solve against the occupancy grid of dftpav_corridor_set_map - the loop the
 * replanning server runs over the executing trajectory every cycle (src/traj_server_ros.cpp:385-397): every gear segment is
 * sampled at t = 0, dt, dt + dt, ... while t < duration (dt = 0.05 there), the footprint at (getPos(t), getAngle(t)) is the
 * vertex set of SemanticMapManager::CheckCollisionUsingPosAndYaw (Sim/core/semantic_map_manager/src/semantic_map_manager.cc:
 * 639-661; box centre = pos + veh_d_cr (cos yaw, sin yaw); corners and edge points every vertex_spacing metres,
 * ShapeUtils::GetDenseVerticesOfOrientedBoundingBox, Sim/core/common/src/common/basics/shapes.cc:110-149, default 0.1) and a
 * sample collides if any vertex lies in an occupied cell.  collide[B]: 1 if any sample of any segment of the problem collides
 * (the reference returns true at the first one).  first_hit (may be NULL) [seg_total]: index of the first colliding sample
 * of the segment, -1 if none.  max_samples bounds the samples per segment; DFTPAV_ERR_INVALID_ARG if a segment is longer than
 * max_samples * dt.  vp->limit_bound is not used. */
int dftpav_collision_check(dftpav_handle* h, const dftpav_corridor_params* vp, double dt, double vertex_spacing,
                           int32_t max_samples, int32_t* collide, int32_t* first_hit);

/* ---- multi-GPU: the batch shards embarrassingly over the GPUs of one box ----
 * dftpav_create_multi: one handle, host thread and stream per device of device_ids[0..n_dev) and - for n_dev > 1 - an NCCL
 * communicator over them (libnccl.so.2 is loaded at run time; without it the flag below is reduced on the host and
 * dftpav_multi_uses_nccl() returns 0).
 * dftpav_solve_multi: problems are split into n_dev contiguous slices, every slice goes through dftpav_solve on its device
 * concurrently, results are written into the caller's arrays (layout and strides of dftpav_batch_dims of the WHOLE batch;
 * bit-identical to a single-device solve), and the path's single collective runs: an all-reduce (MIN) over the devices of
 * "every problem of my slice ended with success == 1" -> *all_accepted (may be NULL).  No data-path collective exists. */
typedef struct dftpav_multi dftpav_multi;
int dftpav_create_multi(const dftpav_params* prm, const int32_t* device_ids, int32_t n_dev, dftpav_multi** out);
int dftpav_destroy_multi(dftpav_multi* m);
int dftpav_solve_multi(dftpav_multi* m, const dftpav_batch* batch, const dftpav_result* out, int32_t* all_accepted);
int dftpav_multi_uses_nccl(const dftpav_multi* m);
int dftpav_multi_devices(const dftpav_multi* m);
int dftpav_multi_stats(const dftpav_multi* m, int32_t dev_index, dftpav_stats* stats);
const char* dftpav_multi_last_error(const dftpav_multi* m);

int dftpav_last_stats(const dftpav_handle* h, dftpav_stats* stats);
const char* dftpav_last_error(const dftpav_handle* h);
const char* dftpav_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DFTPAV_B200_H */
