// frontend_oracle.cpp — CPU restatement of the front-end resampling that feeds the optimizer.
//
// TEST INFRASTRUCTURE ONLY (like everything under oracle/).  PARITY UNPINNED: the reference functions are members of
// path_searching::KinoAstar (src/Plan/traj_planner/src/kino_astar.cpp, needs OMPL + PCL + ROS) and of TrajPlanner
// (src/traj_manager.cpp, ROS); neither translation unit compiles here.  The restatement follows the sources statement by
// statement; what checks it instead: closed-form cases (tests/test_frontend_oracle.py) and an independent numpy
// implementation of the same procedure.
//
//   KinoAstar::evaluatePos(t)        kino_astar.cpp:468-521   pose on the searched path at time t: the gear segment ("shot")
//                                                              that contains t, the arc length the trapezoidal speed profile
//                                                              has covered (evaluateLength), linear interpolation between the
//                                                              path samples around that arc length (yaw unwrapped across +-pi)
//   KinoAstar::evaluateDuration      kino_astar.cpp:744-762   duration of the profile over a length
//   KinoAstar::evaluateLength        kino_astar.cpp:763-795   arc length covered at time curt
//   TrajPlanner::RunMINCOParking     traj_manager.cpp:531-567 per gear segment: piece count, piece duration, the sampled states
//                                                              (corridor input) and the inner waypoints (optimizer input)
#include <algorithm>
#include <cmath>
#include <cstdint>

#include "../include/dftpav_b200.h"

namespace {

// kino_astar.cpp:763-795 (`localtime` is an unused parameter there)
double evaluate_length(double curt, double locallength, double max_vel, double max_acc, double startV, double endV) {
  const double startv2 = std::pow(startV, 2), endv2 = std::pow(endV, 2), maxv2 = std::pow(max_vel, 2);
  const double critical_len = (maxv2 - startv2) / (2 * max_acc) + (maxv2 - endv2) / (2 * max_acc);
  if (locallength >= critical_len) {
    const double t1 = (max_vel - startV) / max_acc;
    const double t2 = t1 + (locallength - critical_len) / max_vel;
    if (curt <= t1) return startV * curt + 0.5 * max_acc * std::pow(curt, 2);
    if (curt <= t2) return startV * t1 + 0.5 * max_acc * std::pow(t1, 2) + (curt - t1) * max_vel;
    return startV * t1 + 0.5 * max_acc * std::pow(t1, 2) + (t2 - t1) * max_vel + max_vel * (curt - t2) -
           0.5 * max_acc * std::pow(curt - t2, 2);
  }
  const double tmpv = std::sqrt(0.5 * (startv2 + endv2 + 2 * max_acc * locallength));
  const double tmpt = (tmpv - startV) / max_acc;
  if (curt <= tmpt) return startV * curt + 0.5 * max_acc * std::pow(curt, 2);
  return startV * tmpt + 0.5 * max_acc * std::pow(tmpt, 2) + tmpv * (curt - tmpt) - 0.5 * max_acc * std::pow(curt - tmpt, 2);
}

// kino_astar.cpp:468-521 for path p of the front end
void evaluate_pos(const dftpav_frontend* fe, int p, double t, double* out) {
  const int s0 = fe->shot_ptr[p], ns = fe->shot_ptr[p + 1] - s0;
  const double* st = fe->shot_time + s0;
  const double* sl = fe->shot_length + s0;
  const int32_t* ss = fe->shot_s + s0;
  const int32_t* si = fe->shot_index + s0 + p;   // ns + 1 entries
  const double* pts = fe->pts + 3 * (size_t)fe->pt_ptr[p];
  double total = 0.0;  // totalTrajTime (kino_astar.cpp:739-742)
  for (int i = 0; i < ns; i++) total += st[i];
  t = std::min<double>(std::max<double>(0, t), total);
  const double startvel = fe->start_vel[p], endvel = fe->end_vel[p];
  int index = -1;
  double tmpT = 0, CutTime = 0;
  for (int i = 0; i < ns; i++) {
    tmpT += st[i];
    if (tmpT >= t) {
      index = i;
      CutTime = t - tmpT + st[i];
      break;
    }
  }
  double initv = fe->non_siguav, finv = fe->non_siguav;
  if (index == 0) initv = startvel;
  if (index == ns - 1) finv = endvel;
  const double locallength = sl[index];
  const int front = si[index], back = si[index + 1];
  const double* lt = pts + 3 * (size_t)front;  // localTraj
  const int nl = back - front + 1;
  const double arclength = ss[index] > 0
                               ? evaluate_length(CutTime, locallength, fe->max_forward_vel, fe->max_forward_acc, initv, finv)
                               : evaluate_length(CutTime, locallength, fe->max_backward_vel, fe->max_backward_acc, initv, finv);
  double tmparc = 0;
  for (int i = 0; i < nl - 1; i++) {
    const double dx = lt[3 * (i + 1)] - lt[3 * i], dy = lt[3 * (i + 1) + 1] - lt[3 * i + 1];
    const double l = std::sqrt(dx * dx + dy * dy);
    tmparc += l;
    if (tmparc >= arclength) {
      const double l1 = tmparc - arclength;
      const double l2 = l - l1;
      for (int q = 0; q < 3; q++) out[q] = l1 / l * lt[3 * i + q] + l2 / l * lt[3 * (i + 1) + q];
      const double y0 = lt[3 * i + 2], y1 = lt[3 * (i + 1) + 2];
      if (std::fabs(y1 - y0) >= M_PI) {
        double normalize_yaw = out[2];  // (the reference leaves it uninitialised when neither branch fires: cannot happen for yaws in (-pi, pi])
        if (y1 <= 0) normalize_yaw = l1 / l * y0 + l2 / l * (y1 + 2 * M_PI);
        else if (y0 <= 0) normalize_yaw = l1 / l * (y0 + 2 * M_PI) + l2 / l * y1;
        out[2] = normalize_yaw;
      }
      return;
    }
  }
  for (int q = 0; q < 3; q++) out[q] = lt[3 * (nl - 1) + q];
}

}  // namespace

extern "C" {

double oracle_trapezoid_duration(double length, double max_vel, double max_acc, double startV, double endV) {
  // kino_astar.cpp:744-762
  const double startv2 = std::pow(startV, 2), endv2 = std::pow(endV, 2), maxv2 = std::pow(max_vel, 2);
  const double critical_len = (maxv2 - startv2) / (2 * max_acc) + (maxv2 - endv2) / (2 * max_acc);
  if (length >= critical_len) return (max_vel - startV) / max_acc + (max_vel - endV) / max_acc + (length - critical_len) / max_vel;
  const double tmpv = std::sqrt(0.5 * (startv2 + endv2 + 2 * max_acc * length));
  return (tmpv - startV) / max_acc + (tmpv - endV) / max_acc;
}

double oracle_trapezoid_length(double curt, double locallength, double max_vel, double max_acc, double startV, double endV) {
  return evaluate_length(curt, locallength, max_vel, max_acc, startV, endV);
}

int oracle_frontend_evaluate_pos(const dftpav_frontend* fe, int64_t nq, const int32_t* path, const double* t, double* out) {
  for (int64_t q = 0; q < nq; q++) evaluate_pos(fe, path[q], t[q], out + 3 * q);
  return 0;
}

// traj_manager.cpp:531-567 for every path: seg_duration[shot] = the sum of the FlatTrajData time steps of that gear segment
// (initTotalduration).  Outputs per gear segment: piece_num, piece_dur; states [.][3] (K + 1 or Kd + 1 per piece, the
// order getRectangleConst consumes) and inner [.][2] (piece_num - 1 per segment), both concatenated in segment order.
// Returns the number of states written (or needed, when states is null).
int64_t oracle_frontend_resample(const dftpav_frontend* fe, const double* seg_duration, double traj_piece_duration, int K, int Kd,
                                 int32_t* piece_num, double* piece_dur, double* states, double* inner) {
  int64_t ns = 0, ni = 0;
  for (int p = 0; p < fe->n_paths; p++) {
    double basetime = 0.0;
    for (int g = fe->shot_ptr[p]; g < fe->shot_ptr[p + 1]; g++) {
      const double initTotalduration = seg_duration[g];
      const int pn = std::max(int(initTotalduration / traj_piece_duration + 0.5), 2);
      const double timePerPiece = initTotalduration / pn;
      if (piece_num) piece_num[g] = pn;
      if (piece_dur) piece_dur[g] = timePerPiece;
      double res_time = 0;
      for (int i = 0; i < pn; i++) {
        const int resolution = (i == 0 || i == pn - 1) ? Kd : K;
        for (int k = 0; k <= resolution; k++) {
          const double t = basetime + res_time + 1.0 * k / resolution * timePerPiece;
          double pos[3];
          evaluate_pos(fe, p, t, pos);
          if (states) states[3 * ns] = pos[0], states[3 * ns + 1] = pos[1], states[3 * ns + 2] = pos[2];
          ns++;
          if (k == resolution && i != pn - 1) {
            if (inner) inner[2 * ni] = pos[0], inner[2 * ni + 1] = pos[1];
            ni++;
          }
        }
        res_time += timePerPiece;
      }
      basetime += initTotalduration;
    }
  }
  return ns;
}

}  // extern "C"
