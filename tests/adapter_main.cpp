// Drives the C++ adapter like TrajPlanner::RunMINCOParking does (traj_manager.cpp:509-625): one problem read from a
// flat text dump, OptimizeTrajectory, read-back through getMinJerkOptPtr().  Prints cost / status / first coefficients.
#define DFTPAV_ADAPTER_MATRIX_HEADER "../../tests/mini_matrix.hpp"
#include <cstdio>
#include <fstream>
#include <iostream>

#include "../dftpav_b200/host/poly_traj_optimizer.hpp"

using namespace plan_manage;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1]);
  int S, K, Kd;
  double vmax, amax, cmax;
  in >> S >> K >> Kd >> vmax >> amax >> cmax;
  std::vector<MatrixXd> ini(S), fin(S), inner(S);
  VectorXd T(S);
  std::vector<std::vector<MatrixXd>> hp(S);
  std::vector<int> sing(S);
  for (int s = 0; s < S; s++) {
    int N, M;
    in >> N >> sing[s] >> T(s) >> M;
    ini[s] = MatrixXd(2, 3);
    fin[s] = MatrixXd(2, 3);
    for (int k = 0; k < 6; k++) in >> ini[s].data()[k];
    for (int k = 0; k < 6; k++) in >> fin[s].data()[k];
    inner[s] = MatrixXd(2, N - 1);
    for (int k = 0; k < 2 * (N - 1); k++) in >> inner[s].data()[k];
    hp[s].resize(M);
    for (int q = 0; q < M; q++) {
      int H;
      in >> H;
      hp[s][q] = MatrixXd(4, H);
      for (int k = 0; k < 4 * H; k++) in >> hp[s][q].data()[k];
    }
  }
  PolyTrajOptimizer::Ptr opt(new PolyTrajOptimizer());
  dftpav_params prm;
  dftpav_default_params(&prm);
  prm.traj_resolution = K;
  prm.des_traj_resolution = Kd;
  prm.max_forward_vel = vmax;
  prm.max_forward_acc = amax;
  prm.max_forward_cur = cmax;
  opt->setParam(prm);
  bool ok = opt->OptimizeTrajectory(ini, fin, inner, T, hp, sing, 0.0, 0.0);
  std::printf("success %d status %d cost %.17g\n", ok ? 1 : 0, opt->lastStatus(), opt->lastCost());
  const auto* mj = opt->getMinJerkOptPtr();
  for (size_t s = 0; s < mj->size(); s++) {
    double cm[12];
    (*mj)[s].pieceCoeffMat(0, cm);
    std::printf("segment %zu pieces %d dT %.17g jerk %.17g c0 %.17g %.17g\n", s, (*mj)[s].N, (*mj)[s].getDt(),
                (*mj)[s].getTrajJerkCost(), cm[10], cm[11]);
  }
  return ok ? 0 : 1;
}
