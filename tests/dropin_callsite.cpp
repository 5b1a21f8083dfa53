// The optimizer call sequence of the reference's plan manager (TrajPlanner::init and TrajPlanner::RunMINCOParking,
// traj_manager.cpp:29, 604-623), written once and compiled twice by tests/test_adapter.py:
//   -DUSE_REFERENCE_OPTIMIZER : against the reference's own plan_manage/traj_optimizer.h
//   (default)                 : against this repository's drop-in dftpav_b200/host/poly_traj_optimizer.hpp
// Both must compile from the same text: that is what "drop-in behind the reference's optimizer API" means at the source
// level.  Eigen / ROS / protobuf headers come from the test-only layers under oracle/ (they are absent from the image).
#ifdef USE_REFERENCE_OPTIMIZER
#include "plan_manage/traj_optimizer.h"
#else
#define DFTPAV_WITH_REFERENCE_TYPES
#include "../dftpav_b200/host/poly_traj_optimizer.hpp"
#endif
#include <iostream>
#include <memory>

int plan_once(ros::NodeHandle nh_, planning::minco::Config cfg_, std::vector<Eigen::MatrixXd>& iniState_container,
              std::vector<Eigen::MatrixXd>& finState_container, std::vector<Eigen::MatrixXd>& initInnerPts_container,
              Eigen::VectorXd& duration_container, std::vector<std::vector<Eigen::MatrixXd>>& sfc_container,
              std::vector<int>& singul_container, plan_utils::SurroundTrajData& surround_trajs,
              plan_utils::TrajContainer& traj_container_, double worldtime, int ego_id_) {
  std::unique_ptr<plan_manage::PolyTrajOptimizer> ploy_traj_opt_(new plan_manage::PolyTrajOptimizer);
  ploy_traj_opt_->setParam(nh_, cfg_);                                                       // traj_manager.cpp:29
  ploy_traj_opt_->setSurroundTrajs(&surround_trajs);                                         // :604
  int flag_success = ploy_traj_opt_->OptimizeTrajectory(iniState_container, finState_container, initInnerPts_container,
                                                        duration_container, sfc_container, singul_container, worldtime,
                                                        0.0);                                 // :608-610
  if (flag_success) {
    for (unsigned int i = 0; i < iniState_container.size(); i++) {
      traj_container_.addSingulTraj((*ploy_traj_opt_->getMinJerkOptPtr())[i].getTraj(singul_container[i]), worldtime,
                                    ego_id_);                                                 // :619
      std::cout << (*ploy_traj_opt_->getMinJerkOptPtr())[i].getTraj(1).getTotalDuration()
                << (*ploy_traj_opt_->getMinJerkOptPtr())[i].getTrajJerkCost() << std::endl;  // :622-623
    }
  }
  return flag_success + ploy_traj_opt_->get_traj_resolution_() + ploy_traj_opt_->get_destraj_resolution_() +
         (int)ploy_traj_opt_->getsurroundClearance();
}
