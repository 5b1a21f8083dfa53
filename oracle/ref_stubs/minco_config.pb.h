// TEST INFRASTRUCTURE ONLY: stands in for the protobuf-generated minco_config.pb.h (generated code, not in the source
// tree).  Plain structs with the accessor names PolyTrajOptimizer::setParam reads; the test driver fills the values.
#pragma once
namespace planning {
namespace minco {
struct OptCfg {
  int traj_resolution_ = 16, des_traj_resolution_ = 32, lbfgs_memsize_ = 256, lbfgs_past_ = 3;
  double wei_sta_obs_ = 0, wei_dyn_obs_ = 0, wei_feas_ = 0, wei_sqrvar_ = 0, wei_time_ = 0, dyn_obs_clearance_ = 0,
         half_margin_ = 0, max_phidot_ = 0, max_forward_vel_ = 0, max_backward_vel_ = 0, max_forward_cur_ = 0,
         max_backward_cur_ = 0, max_forward_acc_ = 0, max_backward_acc_ = 0, max_latacc_ = 0, lbfgs_delta_ = 0,
         mini_t_ = 0;
  bool gearopt_ = true;
  int traj_resolution() const { return traj_resolution_; }
  int des_traj_resolution() const { return des_traj_resolution_; }
  double wei_sta_obs() const { return wei_sta_obs_; }
  double wei_dyn_obs() const { return wei_dyn_obs_; }
  double wei_feas() const { return wei_feas_; }
  double wei_sqrvar() const { return wei_sqrvar_; }
  double wei_time() const { return wei_time_; }
  double dyn_obs_clearance() const { return dyn_obs_clearance_; }
  double half_margin() const { return half_margin_; }
  double max_phidot() const { return max_phidot_; }
  double max_forward_vel() const { return max_forward_vel_; }
  double max_backward_vel() const { return max_backward_vel_; }
  double max_forward_cur() const { return max_forward_cur_; }
  double max_backward_cur() const { return max_backward_cur_; }
  double max_forward_acc() const { return max_forward_acc_; }
  double max_backward_acc() const { return max_backward_acc_; }
  double max_latacc() const { return max_latacc_; }
  bool gearopt() const { return gearopt_; }
  int lbfgs_memsize() const { return lbfgs_memsize_; }
  int lbfgs_past() const { return lbfgs_past_; }
  double lbfgs_delta() const { return lbfgs_delta_; }
  double mini_t() const { return mini_t_; }
};
struct Config {
  OptCfg opt;
  const OptCfg& opt_cfg() const { return opt; }
};
}  // namespace minco
}  // namespace planning
