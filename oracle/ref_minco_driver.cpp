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

// The reference's Trajectory / Piece evaluators on caller-supplied coefficients (the read-back side of the hot path:
// MinJerkOpt::getTraj, poly_traj_utils.hpp:987-997, then Trajectory::GetState / getPos / getdSigma / getddSigma / getAngle /
// getCurv / getVel / getAcc / getSteer / getLatAcc, 378-406 and 554-645).  c = 6N x 2 row-major (row 6i+k = coefficient of
// t^k of piece i), all pieces of duration dT.  clamp != 0: the time is clamped to the total duration like GetState does.
// out[12] = x y dsigma(2) ddsigma(2) angle curvature velocity acceleration steer lateral-acceleration;
// gs[6] (optional) = what GetState writes: angle curvature velocity acceleration steer + the clamped time.
int ref_traj_state(int N, const double* c, double dT, int singul, double t, int clamp, double* out, double* gs) {
  plan_utils::Trajectory traj;
  traj.reserve(N);
  for (int i = 0; i < N; i++) {
    plan_utils::CoefficientMat cm;
    for (int k = 0; k < 6; k++) cm(0, 5 - k) = c[2 * (6 * i + k)], cm(1, 5 - k) = c[2 * (6 * i + k) + 1];
    traj.emplace_back(dT, cm, singul);
  }
  double tt = t;
  if (clamp && tt > traj.getTotalDuration()) tt = traj.getTotalDuration();
  const Eigen::Vector2d pos = traj.getPos(tt), ds = traj.getdSigma(tt), dds = traj.getddSigma(tt);
  out[0] = pos(0), out[1] = pos(1), out[2] = ds(0), out[3] = ds(1), out[4] = dds(0), out[5] = dds(1);
  out[6] = traj.getAngle(tt), out[7] = traj.getCurv(tt), out[8] = traj.getVel(tt), out[9] = traj.getAcc(tt);
  out[10] = traj.getSteer(tt), out[11] = traj.getLatAcc(tt);
  if (gs) {
    common::State st;
    traj.GetState(t, &st);
    gs[0] = st.angle, gs[1] = st.curvature, gs[2] = st.velocity, gs[3] = st.acceleration, gs[4] = st.steer;
    gs[5] = st.vec_position(0);
  }
  return 0;
}

}  // extern "C"
