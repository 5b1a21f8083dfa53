// poly_traj_optimizer.hpp — source-compatible C++ adapter: the reference's optimizer API on top of the C-ABI.
//
// Mirrors plan_manage::PolyTrajOptimizer (reference src/Plan/traj_planner/include/plan_manage/traj_optimizer.h:24-248):
//   setParam / setSurroundTrajs / OptimizeTrajectory / getMinJerkOptPtr()->at(i).getTraj(singul), getTrajJerkCost().
// The plan manager (traj_manager.cpp:28-29, 604-623) calls exactly these.  All numerics run in libdftpav_b200.so on the
// GPU (batch of one problem per call; planners that evaluate many candidates should use the batch C-ABI directly).
//
// Two build modes:
//   * inside the reference's catkin workspace: define DFTPAV_WITH_REFERENCE_TYPES before including; Eigen, ROS and
//     plan_utils::Trajectory / SurroundTrajData come from the reference tree and the signatures are character-for-
//     character those of traj_optimizer.h (setParam(ros::NodeHandle, planning::minco::Config) included);
//   * standalone (this repository's tests; Eigen is not installed in the build image): any matrix type offering
//     rows(), cols(), data() (column-major) and operator()(i,j) — e.g. tests/mini_matrix.hpp — is accepted.
#pragma once
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dftpav_b200.h"

#ifdef DFTPAV_WITH_REFERENCE_TYPES
#include <Eigen/Eigen>
#include <ros/ros.h>
#include <plan_utils/traj_container.hpp>
#include "minco_config.pb.h"
namespace dftpav_adapter {
typedef Eigen::MatrixXd MatrixXd;
typedef Eigen::VectorXd VectorXd;
}
#else
#ifndef DFTPAV_ADAPTER_MATRIX_HEADER
#error "standalone build: define DFTPAV_ADAPTER_MATRIX_HEADER to a header providing dftpav_adapter::MatrixXd / VectorXd"
#endif
#include DFTPAV_ADAPTER_MATRIX_HEADER
#endif

namespace plan_manage {

using dftpav_adapter::MatrixXd;
using dftpav_adapter::VectorXd;

// One obstacle prediction = plan_utils::LocalTrajData (traj_container.hpp:29-39) reduced to what the optimizer reads.
struct SurroundTrajLite {
  double start_time = 0.0, duration = 0.0;
  std::vector<double> piece_durations;
  std::vector<double> coeffs;  // per piece 12 doubles: Piece::coeffMat 2x6 column-major, highest power first
};

// Result view of one gear segment = the subset of plan_utils::MinJerkOpt the manager reads (traj_manager.cpp:619-623).
class MinJerkOptView {
 public:
  int N = 0;
  double dT = 0.0;
  std::vector<double> c;  // 6N x 2 row-major pairs, row 6i+k = coefficient of t^k of piece i (poly_traj_utils.hpp:981-984)
  inline const double& getDt() const { return dT; }
  inline double getTrajJerkCost() const {  // poly_traj_utils.hpp:998-1009
    const double t1 = dT, t2 = t1 * t1, t3 = t2 * t1, t4 = t2 * t2, t5 = t4 * t1;
    double energy = 0.0;
    auto d = [&](int a, int b) { return c[2 * a] * c[2 * b] + c[2 * a + 1] * c[2 * b + 1]; };
    for (int i = 0; i < N; i++) {
      const int r3 = 6 * i + 3, r4 = 6 * i + 4, r5 = 6 * i + 5;
      energy += 36.0 * d(r3, r3) * t1 + 144.0 * d(r4, r3) * t2 + 192.0 * d(r4, r4) * t3 + 240.0 * d(r5, r3) * t3 +
                720.0 * d(r5, r4) * t4 + 720.0 * d(r5, r5) * t5;
    }
    return energy;
  }
  // Piece coefficient matrix in the reference's Piece layout (2x6 column-major, highest power first):
  // what MinJerkOpt::getTraj builds with c.block<6,2>(6i,0).transpose().rowwise().reverse() (poly_traj_utils.hpp:993)
  inline void pieceCoeffMat(int i, double out12[12]) const {
    for (int k = 0; k < 6; k++) {
      out12[2 * k] = c[2 * (6 * i + 5 - k)];
      out12[2 * k + 1] = c[2 * (6 * i + 5 - k) + 1];
    }
  }
#ifdef DFTPAV_WITH_REFERENCE_TYPES
  inline plan_utils::Trajectory getTraj(int s) const {  // poly_traj_utils.hpp:987-997
    plan_utils::Trajectory traj;
    traj.reserve(N);
    for (int i = 0; i < N; i++) {
      plan_utils::CoefficientMat cm;
      double buf[12];
      pieceCoeffMat(i, buf);
      std::memcpy(cm.data(), buf, sizeof(buf));
      traj.emplace_back(dT, cm, s);
    }
    return traj;
  }
#endif
};

class PolyTrajOptimizer {
 public:
  typedef std::unique_ptr<PolyTrajOptimizer> Ptr;
  PolyTrajOptimizer() { dftpav_default_params(&prm_); }
  ~PolyTrajOptimizer() {
    if (h_) dftpav_destroy(h_);
  }
  PolyTrajOptimizer(const PolyTrajOptimizer&) = delete;
  PolyTrajOptimizer& operator=(const PolyTrajOptimizer&) = delete;

  /* set variables */
  void setParam(const dftpav_params& p) {
    prm_ = p;
    reset_handle();
  }
#ifdef DFTPAV_WITH_REFERENCE_TYPES
  // field-by-field copy of traj_optimizer.cpp:1713-1736
  void setParam(ros::NodeHandle, planning::minco::Config cfg_) {
    const auto& o = cfg_.opt_cfg();
    prm_.traj_resolution = o.traj_resolution();
    prm_.des_traj_resolution = o.des_traj_resolution();
    prm_.wei_obs = o.wei_sta_obs();
    prm_.wei_surround = o.wei_dyn_obs();
    prm_.wei_feas = o.wei_feas();
    prm_.wei_time = o.wei_time();
    prm_.surround_clearance = o.dyn_obs_clearance();
    prm_.half_margin = o.half_margin();
    prm_.max_forward_vel = o.max_forward_vel();
    prm_.max_backward_vel = o.max_backward_vel();
    prm_.max_forward_cur = o.max_forward_cur();
    prm_.max_backward_cur = o.max_backward_cur();
    prm_.max_forward_acc = o.max_forward_acc();
    prm_.max_backward_acc = o.max_backward_acc();
    prm_.gear_opt = o.gearopt() ? 1 : 0;
    prm_.lbfgs_mem_size = o.lbfgs_memsize();
    prm_.lbfgs_past = o.lbfgs_past();
    prm_.lbfgs_delta = o.lbfgs_delta();
    prm_.mini_T = o.mini_t();
    reset_handle();
  }
  void setSurroundTrajs(plan_utils::SurroundTrajData* p) {  // borrowed pointer, traj_optimizer.h:30,108
    ref_surround_ = p;
  }
#endif
  void setSurroundTrajs(const std::vector<SurroundTrajLite>* p) { surround_ = p; }

  /* helper functions */
  inline const std::vector<MinJerkOptView>* getMinJerkOptPtr(void) { return &jerkOpt_container; }
  inline int get_traj_resolution_() { return prm_.traj_resolution; }
  inline int get_destraj_resolution_() { return prm_.des_traj_resolution; }
  inline double getsurroundClearance(void) { return prm_.surround_clearance; }
  inline double lastCost() const { return final_cost_; }
  inline int lastStatus() const { return status_; }

  /* main planning API (traj_optimizer.h:118-120) */
  bool OptimizeTrajectory(const std::vector<MatrixXd>& iniStates, const std::vector<MatrixXd>& finStates,
                          std::vector<MatrixXd>& initInnerPts, const VectorXd& initTs,
                          std::vector<std::vector<MatrixXd>>& hPoly_container, std::vector<int> singuls, double now,
                          double help_eps = 1.0e-4) {
    const int S = (int)initInnerPts.size();
    jerkOpt_container.clear();
    if ((int)initTs.size() != S || (int)iniStates.size() != S || (int)finStates.size() != S ||
        (int)hPoly_container.size() != S || (int)singuls.size() != S)
      return false;  // traj_optimizer.cpp:26-29
    if (prm_.help_eps != help_eps) {
      prm_.help_eps = help_eps;
      reset_handle();
    }
    if (!h_ && dftpav_create(&prm_, -1, &h_) != DFTPAV_OK) throw std::runtime_error("dftpav_create failed: no sm_100 GPU");
    // ---- flatten into a batch of one
    std::vector<int32_t> seg_ptr = {0, S}, piece_num(S), sing(singuls.begin(), singuls.end()), plane_cnt;
    std::vector<double> ini(6 * S), fin(6 * S), T(S), inner, planes, tnow = {now};
    for (int s = 0; s < S; s++) {
      if (initInnerPts[s].cols() == 0) return false;  // "There is only a piece?" traj_optimizer.cpp:38-41
      piece_num[s] = (int)initInnerPts[s].cols() + 1;
      std::memcpy(&ini[6 * s], iniStates[s].data(), 6 * sizeof(double));  // 2x3 column-major
      std::memcpy(&fin[6 * s], finStates[s].data(), 6 * sizeof(double));
      T[s] = initTs(s);
      inner.insert(inner.end(), initInnerPts[s].data(), initInnerPts[s].data() + 2 * (piece_num[s] - 1));
      const int M = (piece_num[s] - 2) * (prm_.traj_resolution + 1) + 2 * (prm_.des_traj_resolution + 1);
      if ((int)hPoly_container[s].size() != M) return false;  // "cfgHs size error!" traj_optimizer.cpp:44-48
      for (const MatrixXd& hp : hPoly_container[s]) {         // 4 x H, one column per half-plane
        plane_cnt.push_back((int)hp.cols());
        planes.insert(planes.end(), hp.data(), hp.data() + 4 * hp.cols());
      }
    }
    dftpav_batch b;
    std::memset(&b, 0, sizeof(b));
    b.B = 1;
    b.seg_ptr = seg_ptr.data();
    b.piece_num = piece_num.data();
    b.singul = sing.data();
    b.ini_state = ini.data();
    b.fin_state = fin.data();
    b.init_T = T.data();
    b.inner_pts = inner.data();
    b.plane_cnt = plane_cnt.data();
    b.planes = planes.data();
    b.t_now = tnow.data();
    std::vector<int32_t> obs_ptr, obs_idx, obs_piece_ptr;
    std::vector<double> o_start, o_dur, o_pdur, o_coef;
    std::vector<SurroundTrajLite> converted;
    const std::vector<SurroundTrajLite>* sur = surround_;
#ifdef DFTPAV_WITH_REFERENCE_TYPES
    if (ref_surround_) {
      for (const auto& lt : *ref_surround_) {
        SurroundTrajLite o;
        o.start_time = lt.start_time;
        o.duration = lt.duration;
        for (int i = 0; i < lt.traj.getPieceNum(); i++) {
          o.piece_durations.push_back(lt.traj[i].getDuration());
          const double* cm = lt.traj[i].getCoeffMat().data();
          o.coeffs.insert(o.coeffs.end(), cm, cm + 12);
        }
        converted.push_back(o);
      }
      sur = &converted;
    }
#endif
    if (sur && !sur->empty()) {  // surround_trajs_ != NULL, traj_optimizer.cpp:636
      obs_ptr = {0, (int32_t)sur->size()};
      obs_piece_ptr.push_back(0);
      for (size_t u = 0; u < sur->size(); u++) {
        const SurroundTrajLite& o = (*sur)[u];
        obs_idx.push_back((int32_t)u);
        o_start.push_back(o.start_time);
        o_dur.push_back(o.duration);
        o_pdur.insert(o_pdur.end(), o.piece_durations.begin(), o.piece_durations.end());
        o_coef.insert(o_coef.end(), o.coeffs.begin(), o.coeffs.end());
        obs_piece_ptr.push_back((int32_t)o_pdur.size());
      }
      b.obs_ptr = obs_ptr.data();
      b.obs_idx = obs_idx.data();
      b.n_obs_pool = (int32_t)sur->size();
      b.obs_start_time = o_start.data();
      b.obs_duration = o_dur.data();
      b.obs_piece_ptr = obs_piece_ptr.data();
      b.obs_piece_dur = o_pdur.data();
      b.obs_coeffs = o_coef.data();
    }
    dftpav_dims dims;
    int rc = dftpav_batch_dims(&prm_, &b, &dims);
    if (rc == DFTPAV_ERR_BAD_INPUT) return false;  // e.g. "mini segment T < mini_T!" traj_optimizer.cpp:30-33
    if (rc != DFTPAV_OK) throw std::runtime_error("dftpav_batch_dims failed");
    std::vector<double> x(dims.n_stride), coeffs(2 * (size_t)dims.row_stride), dT(S);
    int32_t success = 0;
    dftpav_result r;
    std::memset(&r, 0, sizeof(r));
    r.x = x.data();
    r.cost = &final_cost_;
    r.status = &status_;
    r.success = &success;
    r.coeffs = coeffs.data();
    r.piece_dT = dT.data();
    rc = dftpav_solve(h_, &b, &r);
    if (rc == DFTPAV_ERR_BAD_INPUT) return false;
    if (rc != DFTPAV_OK) throw std::runtime_error(std::string("dftpav_solve: ") + dftpav_last_error(h_));
    // ---- results stay in the optimizer until the next call (traj_optimizer.cpp:20-23)
    jerkOpt_container.resize(S);
    size_t row = 0;
    for (int s = 0; s < S; s++) {
      MinJerkOptView& v = jerkOpt_container[s];
      v.N = piece_num[s];
      v.dT = dT[s];
      v.c.assign(coeffs.begin() + 2 * row, coeffs.begin() + 2 * (row + 6 * (size_t)v.N));
      row += 6 * (size_t)v.N;
    }
    last_x_ = x;
    return success != 0;  // flag_success of traj_optimizer.cpp:176-201
  }

  const std::vector<double>& lastDecisionVector() const { return last_x_; }

 private:
  void reset_handle() {
    if (h_) dftpav_destroy(h_);
    h_ = nullptr;
  }
  dftpav_params prm_;
  dftpav_handle* h_ = nullptr;
  const std::vector<SurroundTrajLite>* surround_ = nullptr;
#ifdef DFTPAV_WITH_REFERENCE_TYPES
  plan_utils::SurroundTrajData* ref_surround_ = nullptr;
#endif
  std::vector<MinJerkOptView> jerkOpt_container;
  std::vector<double> last_x_;
  double final_cost_ = 0.0;
  int32_t status_ = 0;
};

}  // namespace plan_manage
