// TEST INFRASTRUCTURE ONLY.  Compiles the reference's OWN MINCO header (plan_utils/poly_traj_utils.hpp: BandedSystem,
// MinJerkOpt, Piece, Trajectory), included from where it lies under /root/reference and unmodified, against the
// test-only Eigen-compatible layer oracle/eigen_shim2 and the stub headers oracle/ref_stubs, into
// oracle/_ref/libref_minco.so (recipe: oracle/Makefile, target `ref`).  tests/test_oracle_vs_reference_minco.py runs
// this next to the oracle's restatement on the same inputs.  No reference source is copied into the repository.
#include "plan_utils/poly_traj_utils.hpp"

extern "C" {

// generate() -> getTrajJerkCost() -> initSmGradCost() -> [gdC += extra_gdC, gdT += extra_gdT: what addPVAGradCost2CT
// does between the two calls] -> calGrads_PT().   Layouts: P 2x(N-1) column-major, head/tail 2x3 column-major (p v a),
// matrices 6N x 2 row-major.
int ref_minco_run(int N, const double* P, double dT, const double* head, const double* tail, const double* extra_gdC,
                  double extra_gdT, double* c, double* energy, double* gdC0, double* gdT0, double* gdP, double* gdHead,
                  double* gdTail, double* gdT) {
  plan_utils::MinJerkOpt opt;
  opt.reset(N);
  Eigen::MatrixXd inPs(2, N - 1), hs(2, 3), ts(2, 3);
  for (int i = 0; i < N - 1; i++) inPs(0, i) = P[2 * i], inPs(1, i) = P[2 * i + 1];
  for (int k = 0; k < 3; k++) {
    hs(0, k) = head[2 * k], hs(1, k) = head[2 * k + 1];
    ts(0, k) = tail[2 * k], ts(1, k) = tail[2 * k + 1];
  }
  opt.generate(inPs, dT, hs, ts);
  const Eigen::MatrixXd& cc = opt.getCoeffs();
  for (int i = 0; i < 6 * N; i++) c[2 * i] = cc(i, 0), c[2 * i + 1] = cc(i, 1);
  *energy = opt.getTrajJerkCost();
  opt.initSmGradCost();
  Eigen::MatrixXd& g = opt.get_gdC();
  for (int i = 0; i < 6 * N; i++) gdC0[2 * i] = g(i, 0), gdC0[2 * i + 1] = g(i, 1);
  *gdT0 = opt.get_gdT();
  if (extra_gdC)
    for (int i = 0; i < 6 * N; i++) g(i, 0) += extra_gdC[2 * i], g(i, 1) += extra_gdC[2 * i + 1];
  opt.get_gdT() += extra_gdT;
  opt.calGrads_PT();
  const Eigen::MatrixXd p = opt.get_gdP(), gh = opt.get_gdHead(), gt = opt.get_gdTail();
  for (int i = 0; i < N - 1; i++) gdP[2 * i] = p(0, i), gdP[2 * i + 1] = p(1, i);
  for (int k = 0; k < 3; k++) {
    gdHead[2 * k] = gh(0, k), gdHead[2 * k + 1] = gh(1, k);
    gdTail[2 * k] = gt(0, k), gdTail[2 * k + 1] = gt(1, k);
  }
  *gdT = opt.get_gdT();
  return 0;
}

// Trajectory of a generate() call sampled at time t (Trajectory::locatePieceIdx + Piece evaluators):
// out = pos(2) dsigma(2) ddsigma(2) R(4, row-major) Rdot(4, row-major) angle piece_index local_t
int ref_traj_eval(int N, const double* P, double dT, const double* head, const double* tail, int singul, double t,
                  double* out) {
  plan_utils::MinJerkOpt opt;
  opt.reset(N);
  Eigen::MatrixXd inPs(2, N - 1), hs(2, 3), ts(2, 3);
  for (int i = 0; i < N - 1; i++) inPs(0, i) = P[2 * i], inPs(1, i) = P[2 * i + 1];
  for (int k = 0; k < 3; k++) {
    hs(0, k) = head[2 * k], hs(1, k) = head[2 * k + 1];
    ts(0, k) = tail[2 * k], ts(1, k) = tail[2 * k + 1];
  }
  opt.generate(inPs, dT, hs, ts);
  const plan_utils::Trajectory traj = opt.getTraj(singul);
  const Eigen::Vector2d pos = traj.getPos(t), ds = traj.getdSigma(t), dds = traj.getddSigma(t);
  const Eigen::Matrix2d R = traj.getR(t), Rd = traj.getRdot(t);
  out[0] = pos(0), out[1] = pos(1), out[2] = ds(0), out[3] = ds(1), out[4] = dds(0), out[5] = dds(1);
  out[6] = R(0, 0), out[7] = R(0, 1), out[8] = R(1, 0), out[9] = R(1, 1);
  out[10] = Rd(0, 0), out[11] = Rd(0, 1), out[12] = Rd(1, 0), out[13] = Rd(1, 1);
  out[14] = traj.getAngle(t);
  double tl = t;
  out[15] = traj.locatePieceIdx(tl);
  out[16] = tl;
  return 0;
}

}  // extern "C"
