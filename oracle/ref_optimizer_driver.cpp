// TEST INFRASTRUCTURE ONLY.  Compiles the reference's OWN optimizer (src/traj_optimizer.cpp, with plan_manage/
// traj_optimizer.h, plan_utils/poly_traj_utils.hpp, plan_utils/traj_container.hpp and geo_utils2d/lbfgs.hpp), included
// from where it lies under /root/reference and unmodified, against the test-only Eigen-compatible layer
// oracle/eigen_shim2 and the stub headers oracle/ref_stubs (ROS, message types, the protobuf-generated config, RViz /
// DecompUtil helpers: all inert) into oracle/_ref/libref_optimizer.so (recipe: oracle/Makefile, target `ref`).
// The driver feeds one problem of a dftpav_batch to PolyTrajOptimizer::setParam / setSurroundTrajs /
// OptimizeTrajectory - the reference's public API - and reads the result back through getMinJerkOptPtr().
// Two preprocessor devices, neither of which changes the reference's source text:
//   * `#define private public`  : the driver sets the vehicle box and non_sinv and calls costFunctionCallback directly;
//   * `#define lbfgs_optimize lbfgs_optimize_spy` after lbfgs.hpp has been included: OptimizeTrajectory's single call of
//     the solver goes through a forwarding function that records x, the final cost, the return code and the iteration /
//     evaluation counts (the reference keeps them in locals), or - in evaluation mode - only evaluates the reference's
//     cost callback at a caller-supplied point.
// No reference source is copied into the repository.
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/dftpav_b200.h"

#define private public
#include <Eigen/Eigen>
#include "geo_utils2d/lbfgs.hpp"

namespace lbfgs {
struct Spy {
  int mode = 0;  // 0: run the solver, 1: evaluate at x_eval only
  const double* x_eval = nullptr;
  Eigen::VectorXd x, g;
  double cost = 0.0;
  int ret = 0, evals = 0, iters = 0;
  lbfgs_evaluate_t eval = nullptr;
  void* inst = nullptr;
};
static thread_local Spy g_spy;  // one per host thread: ref_opt_solve_batch runs problems concurrently
static double spy_evaluate(void* /*inst*/, const Eigen::VectorXd& x, Eigen::VectorXd& g) {
  g_spy.evals++;
  return g_spy.eval(g_spy.inst, x, g);
}
static int spy_progress(void*, const Eigen::VectorXd&, const Eigen::VectorXd&, const double, const double, const int k,
                        const int) {
  g_spy.iters = k;
  return 0;
}
inline int lbfgs_optimize_spy(Eigen::VectorXd& x, double& f, lbfgs_evaluate_t proc_evaluate, lbfgs_stepbound_t proc_stepbound,
                              lbfgs_progress_t /*proc_progress (NULL in the reference)*/, void* instance,
                              const lbfgs_parameter_t& param) {
  g_spy.eval = proc_evaluate;
  g_spy.inst = instance;
  g_spy.evals = 0;
  g_spy.iters = 0;
  int ret;
  if (g_spy.mode == 1) {
    for (int i = 0; i < x.size(); i++) x(i) = g_spy.x_eval[i];
    g_spy.g = Eigen::VectorXd::Zero(x.size());
    f = proc_evaluate(instance, x, g_spy.g);
    ret = LBFGS_CONVERGENCE;
  } else {
    ret = lbfgs_optimize(x, f, spy_evaluate, proc_stepbound, spy_progress, instance, param);
    // the solver's iteration counter k is a local of lbfgs_optimize; the progress callback sees it after every
    // completed iteration, so when a line search fails (negative return other than the iteration limit) the counter at
    // exit is one more than the last value seen
    if (ret < 0 && ret != LBFGSERR_MAXIMUMITERATION) g_spy.iters += 1;
  }
  g_spy.x = x;
  g_spy.cost = f;
  g_spy.ret = ret;
  return ret;
}
}  // namespace lbfgs
#define lbfgs_optimize lbfgs_optimize_spy
#include "traj_optimizer.cpp"  // found through -I $(REF_SRC): the reference source, where it lies
#undef lbfgs_optimize
#undef private

namespace {

struct Setup {
  plan_manage::PolyTrajOptimizer opt;
  plan_utils::SurroundTrajData sur;
  std::vector<Eigen::MatrixXd> ini, fin, inner;
  Eigen::VectorXd Ts;
  std::vector<std::vector<Eigen::MatrixXd>> hpolys;
  std::vector<int> singuls;
  double now = 0.0, help_eps = 0.0;
  int n = 0, rows = 0;
};

// the constants the reference hard-codes must match the parameter block, otherwise this build cannot represent it
bool representable(const dftpav_params* p) {
  return p->min_speed == 1e-4 && p->smooth_eps == 1e-4 && p->cur_weight_scale == 10.0 && p->lse_alpha == 100.0 &&
         p->obs_cull_factor == 1.5 && p->cost_limit == 50000.0 && p->lbfgs_max_iterations == 12000 &&
         p->lbfgs_max_linesearch == 64 && p->lbfgs_g_epsilon == 1e-16 && p->lbfgs_min_step == 1e-32 &&
         p->lbfgs_max_step == 1e20 && p->lbfgs_f_dec_coeff == 1e-4 && p->lbfgs_s_curv_coeff == 0.9 &&
         p->lbfgs_cautious_factor == 1e-6 && p->lbfgs_machine_prec == 1e-16;
}

int build(const dftpav_params* prm, const dftpav_batch* b, int p, Setup& S) {
  if (!prm || !b || p < 0 || p >= b->B) return DFTPAV_ERR_INVALID_ARG;
  if (!representable(prm)) return DFTPAV_ERR_UNSUPPORTED;
  planning::minco::Config cfg;
  planning::minco::OptCfg& o = cfg.opt;
  o.traj_resolution_ = prm->traj_resolution, o.des_traj_resolution_ = prm->des_traj_resolution;
  o.wei_sta_obs_ = prm->wei_obs, o.wei_dyn_obs_ = prm->wei_surround, o.wei_feas_ = prm->wei_feas, o.wei_time_ = prm->wei_time;
  o.dyn_obs_clearance_ = prm->surround_clearance, o.half_margin_ = prm->half_margin;
  o.max_forward_vel_ = prm->max_forward_vel, o.max_backward_vel_ = prm->max_backward_vel;
  o.max_forward_cur_ = prm->max_forward_cur, o.max_backward_cur_ = prm->max_backward_cur;
  o.max_forward_acc_ = prm->max_forward_acc, o.max_backward_acc_ = prm->max_backward_acc;
  o.gearopt_ = prm->gear_opt != 0;
  o.lbfgs_memsize_ = prm->lbfgs_mem_size, o.lbfgs_past_ = prm->lbfgs_past, o.lbfgs_delta_ = prm->lbfgs_delta;
  o.mini_t_ = prm->mini_T;
  S.opt.veh_param_.set_width(prm->veh_width);
  S.opt.veh_param_.set_length(prm->veh_length);
  S.opt.veh_param_.set_d_cr(prm->veh_d_cr);
  S.opt.non_sinv = prm->non_sinv;
  S.opt.setParam(ros::NodeHandle(), cfg);

  const int g0 = b->seg_ptr[p], g1 = b->seg_ptr[p + 1], K = prm->traj_resolution, Kd = prm->des_traj_resolution;
  // offsets of this problem inside the flat arrays
  long long inner_off = 0, samp_off = 0;
  for (int g = 0; g < g0; g++) {
    inner_off += 2 * (b->piece_num[g] - 1);
    samp_off += (long long)(b->piece_num[g] - 2) * (K + 1) + 2 * (Kd + 1);
  }
  long long plane_off = 0;
  if (b->plane_cnt) for (long long q = 0; q < samp_off; q++) plane_off += b->plane_cnt[q];
  else plane_off = samp_off * b->uniform_H;
  const int Sg = g1 - g0;
  S.Ts.resize(Sg);
  S.n = 0, S.rows = 0;
  for (int g = g0; g < g1; g++) {
    const int N = b->piece_num[g];
    Eigen::MatrixXd is(2, 3), fs(2, 3), ip(2, N - 1);
    for (int k = 0; k < 3; k++) {
      is(0, k) = b->ini_state[6 * g + 2 * k], is(1, k) = b->ini_state[6 * g + 2 * k + 1];
      fs(0, k) = b->fin_state[6 * g + 2 * k], fs(1, k) = b->fin_state[6 * g + 2 * k + 1];
    }
    for (int i = 0; i < N - 1; i++) ip(0, i) = b->inner_pts[inner_off + 2 * i], ip(1, i) = b->inner_pts[inner_off + 2 * i + 1];
    inner_off += 2 * (N - 1);
    S.ini.push_back(is), S.fin.push_back(fs), S.inner.push_back(ip);
    S.Ts(g - g0) = b->init_T[g];
    S.singuls.push_back(b->singul[g]);
    const long long M = (long long)(N - 2) * (K + 1) + 2 * (Kd + 1);
    std::vector<Eigen::MatrixXd> hp;
    for (long long q = 0; q < M; q++) {
      const int H = b->plane_cnt ? b->plane_cnt[samp_off + q] : b->uniform_H;
      Eigen::MatrixXd m(4, H);
      for (int h = 0; h < H; h++)
        for (int r = 0; r < 4; r++) m(r, h) = b->planes[4 * (plane_off + h) + r];
      plane_off += H;
      hp.push_back(m);
    }
    samp_off += M;
    S.hpolys.push_back(hp);
    S.n += 2 * (N - 1);
    S.rows += 6 * N;
  }
  S.n += Sg + 3 * (Sg - 1);
  S.now = b->t_now ? b->t_now[p] : 0.0;
  S.help_eps = prm->help_eps;
  if (b->obs_ptr) {
    for (int e = b->obs_ptr[p]; e < b->obs_ptr[p + 1]; e++) {
      const int u = b->obs_idx[e];
      std::vector<double> durs;
      std::vector<plan_utils::CoefficientMat> cms;
      for (int q = b->obs_piece_ptr[u]; q < b->obs_piece_ptr[u + 1]; q++) {
        durs.push_back(b->obs_piece_dur[q]);
        plan_utils::CoefficientMat cm;
        for (int k = 0; k < 6; k++) cm(0, k) = b->obs_coeffs[12 * q + 2 * k], cm(1, k) = b->obs_coeffs[12 * q + 2 * k + 1];
        cms.push_back(cm);
      }
      plan_utils::LocalTrajData ld;
      ld.traj = plan_utils::Trajectory(durs, cms, 1);
      ld.drone_id = u, ld.traj_id = u;
      ld.duration = b->obs_duration[u];
      ld.start_time = b->obs_start_time[u];
      ld.end_time = ld.start_time + ld.duration;
      ld.start_pos = ld.traj.getJuncPos(0);
      ld.init_angle = 0.0;
      S.sur.push_back(ld);
    }
    S.opt.setSurroundTrajs(&S.sur);
  }
  return DFTPAV_OK;
}

// The reference prints progress to std::cout from inside OptimizeTrajectory (traj_optimizer.cpp:34,148,158,171).
// ref_opt_solve_batch runs several solves concurrently, so the sink must hold NO state: a streambuf without a put area
// whose overflow()/xsputn() discard their input.  (Round 1 redirected into one std::ostringstream shared by all OpenMP
// threads: unsynchronised writes into the same stringbuf corrupted the heap and killed bench.py --impl reference.)
struct NullBuf : std::streambuf {
  int_type overflow(int_type c) override { return traits_type::not_eof(c); }
  std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
};
NullBuf g_null_buf;
struct Quiet {
  std::streambuf* old;
  Quiet() : old(std::cout.rdbuf(&g_null_buf)) {}
  ~Quiet() { std::cout.rdbuf(old); }
};

}  // namespace

extern "C" {

// One full OptimizeTrajectory of problem p.  x [n], coeffs [rows][2] row-major, piece_dT [segments].
int ref_opt_solve(const dftpav_params* prm, const dftpav_batch* b, int p, double* x, double* cost, int* status,
                  int* success, int* iters, int* evals, double* coeffs, double* piece_dT) {
  auto S = std::make_unique<Setup>();
  int rc = build(prm, b, p, *S);
  if (rc != DFTPAV_OK) return rc;
  Quiet q;
  lbfgs::g_spy.mode = 0;
  const bool ok = S->opt.OptimizeTrajectory(S->ini, S->fin, S->inner, S->Ts, S->hpolys, S->singuls, S->now, S->help_eps);
  for (int i = 0; i < S->n; i++) x[i] = lbfgs::g_spy.x(i);
  *cost = lbfgs::g_spy.cost, *status = lbfgs::g_spy.ret, *success = ok ? 1 : 0;
  *iters = lbfgs::g_spy.iters, *evals = lbfgs::g_spy.evals;
  const std::vector<plan_utils::MinJerkOpt>* J = S->opt.getMinJerkOptPtr();
  int row = 0;
  for (size_t s = 0; s < J->size(); s++) {
    const Eigen::MatrixXd& c = (*J)[s].getCoeffs();
    for (int i = 0; i < c.rows(); i++, row++) coeffs[2 * row] = c(i, 0), coeffs[2 * row + 1] = c(i, 1);
    piece_dT[s] = (*J)[s].getDt();
  }
  return DFTPAV_OK;
}

// The reference's costFunctionCallback at a caller-supplied x (after the set-up part of OptimizeTrajectory).
// x0_out (may be null) receives the initial decision vector the reference packs.
int ref_opt_eval(const dftpav_params* prm, const dftpav_batch* b, int p, const double* x, double* f, double* g,
                 double* coeffs, double* piece_dT) {
  auto S = std::make_unique<Setup>();
  int rc = build(prm, b, p, *S);
  if (rc != DFTPAV_OK) return rc;
  Quiet q;
  lbfgs::g_spy.mode = 1;
  lbfgs::g_spy.x_eval = x;
  S->opt.OptimizeTrajectory(S->ini, S->fin, S->inner, S->Ts, S->hpolys, S->singuls, S->now, S->help_eps);
  lbfgs::g_spy.mode = 0;
  *f = lbfgs::g_spy.cost;
  for (int i = 0; i < S->n; i++) g[i] = lbfgs::g_spy.g(i);
  const std::vector<plan_utils::MinJerkOpt>* J = S->opt.getMinJerkOptPtr();
  int row = 0;
  for (size_t s = 0; s < J->size(); s++) {
    const Eigen::MatrixXd& c = (*J)[s].getCoeffs();
    for (int i = 0; i < c.rows(); i++, row++)
      if (coeffs) coeffs[2 * row] = c(i, 0), coeffs[2 * row + 1] = c(i, 1);
    if (piece_dT) piece_dT[s] = (*J)[s].getDt();
  }
  return DFTPAV_OK;
}

// The initial decision vector the reference packs (OPT.cpp:96-115), obtained by evaluating nothing: mode 1 with x_eval
// pointing at a scratch copy would overwrite it, so this runs the packing and returns before the first evaluation.
int ref_opt_dims(const dftpav_params* prm, const dftpav_batch* b, int p, int* n, int* rows) {
  auto S = std::make_unique<Setup>();
  int rc = build(prm, b, p, *S);
  if (rc != DFTPAV_OK) return rc;
  *n = S->n, *rows = S->rows;
  return DFTPAV_OK;
}

// Problems p0..p1-1, one complete OptimizeTrajectory each, OpenMP over problems (the reference itself is single-
// threaded per solve).  Used by bench.py --impl reference to time the reference build.  Outputs are [p1-p0] arrays.
int ref_opt_solve_batch(const dftpav_params* prm, const dftpav_batch* b, int p0, int p1, int threads, double* cost,
                        int* status, int* success, int* iters, int* evals) {
  if (!prm || !b || p0 < 0 || p1 > b->B || p0 > p1) return DFTPAV_ERR_INVALID_ARG;
  if (!representable(prm)) return DFTPAV_ERR_UNSUPPORTED;
  Quiet q;
  int bad = 0;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
  for (int p = p0; p < p1; p++) {
    auto S = std::make_unique<Setup>();
    if (build(prm, b, p, *S) != DFTPAV_OK) {
#pragma omp atomic
      bad++;
      continue;
    }
    lbfgs::g_spy.mode = 0;
    const bool ok = S->opt.OptimizeTrajectory(S->ini, S->fin, S->inner, S->Ts, S->hpolys, S->singuls, S->now, S->help_eps);
    const int k = p - p0;
    cost[k] = lbfgs::g_spy.cost, status[k] = lbfgs::g_spy.ret, success[k] = ok ? 1 : 0;
    iters[k] = lbfgs::g_spy.iters, evals[k] = lbfgs::g_spy.evals;
  }
  return bad ? DFTPAV_ERR_INVALID_ARG : DFTPAV_OK;
}

}  // extern "C"
