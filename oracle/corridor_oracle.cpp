// corridor_oracle.cpp — CPU restatement of the reference's per-sample safe-rectangle generation.
//
// TEST INFRASTRUCTURE ONLY (like everything under oracle/).  PARITY UNPINNED: the reference function is a member of
// TrajPlanner (src/Plan/traj_planner/src/traj_manager.cpp:1211-1469) whose translation unit pulls in the whole ROS / OMPL /
// semantic-map stack and cannot be compiled here; the restatement below follows its source statement by statement, and
// the two leaf routines it calls - TrajPlannerAdapter::CheckIfCollisionUsingLine (src/map_adapter.cpp:117-129) and
// GridMapND::CheckIfEqualUsingGlobalPosition / GetCoordUsingGlobalPosition (Sim/core/common/src/common/basics/
// semantics.cc:169-179, 214-221, 264-288; OCCUPIED = 80, semantics.h:354) - likewise.  What pins it instead: geometric
// invariants checked by tests/test_corridor_oracle.py (every returned rectangle contains the vehicle box, is collision-free
// on its boundary at the sampled points, cannot grow by one more cell in any finished direction) and an independent numpy
// implementation of the same procedure.
//
// Reference procedure (MGR.cpp = traj_manager.cpp):
//   for each state (x, y, yaw):  sourcePt = (x, y), sourceVp = default VehicleParam (width 1.90, length 4.88, d_cr 1.015),
//   expandLength = 0; round-robin over the four directions +dy, +dx, -dy, -dx while any is unfinished (MGR.cpp:1306-1432):
//   the three segments  point1 -> newpoint1 -> newpoint2 -> point2  of the strip that one more `step` (= map resolution)
//   would add are ray-marched at resolution/2 (CheckIfCollisionUsingLine); blocked -> the direction is finished; else
//   expandLength[i] += step (finished when >= 10 m) and the source rectangle grows (centre shifts by step/2, width or length
//   grows by step).  The polygon is then built from the RAW pose and the four expansion lengths (MGR.cpp:1433-1460):
//   four columns (n_x, n_y, p_x, p_y) - the layout the optimizer's corridor takes.
#include <cmath>
#include <cstdint>

#include "../include/dftpav_b200.h"

namespace {

struct Grid {
  const dftpav_gridmap* g;
  // GridMapND::CheckIfEqualUsingGlobalPosition(p, OCCUPIED): false outside the map
  bool occupied(double px, double py) const {
    const int cx = (int)std::round((px - g->origin_x) / g->resolution);
    const int cy = (int)std::round((py - g->origin_y) / g->resolution);
    if (cx < 0 || cx >= g->size_x || cy < 0 || cy >= g->size_y) return false;
    return g->data[(size_t)cx + (size_t)cy * g->size_x] == g->occupied_value;
  }
  // TrajPlannerAdapter::CheckIfCollisionUsingLine (map_adapter.cpp:117-129)
  bool line(double p1x, double p1y, double p2x, double p2y, double checkl) const {
    const double dx = p2x - p1x, dy = p2y - p1y;
    const double nrm = std::sqrt(dx * dx + dy * dy);
    for (double dl = 0.0; dl < nrm; dl += checkl) {
      const double px = dx * dl / nrm + p1x, py = dy * dl / nrm + p1y;
      if (occupied(px, py)) return true;
    }
    return occupied(p2x, p2y);
  }
};

struct Vp {
  double width, length, d_cr;
};

}  // namespace

extern "C" int oracle_traj_state(int N, const double* c, double dT, int singul, double t, int clamp, double wheel_base,
                                 double* out, double* gs);  // dftpav_oracle.cpp (pinned to the reference's evaluators)

extern "C" {

// Collision re-check of ONE gear segment (N pieces of duration dT, coefficients c = 6N x 2 row-major, lowest power first),
// PARITY UNPINNED like the corridor: the loop of traj_server_ros.cpp:385-397
//     for (t = 0; t < duration; t += 0.05) { pos = getPos(t); yaw = getAngle(t); CheckIfCollisionUsingPosAndYaw(vp, state) }
// with SemanticMapManager::CheckCollisionUsingPosAndYaw (Sim/core/semantic_map_manager/src/semantic_map_manager.cc:639-661)
// and ShapeUtils::GetDenseVerticesOfOrientedBoundingBox (Sim/core/common/src/common/basics/shapes.cc:110-149, res = 0.1 by
// default, shapes.h:200-201) restated statement by statement; the trajectory evaluators are the pinned ones.
// Returns the index of the first colliding sample, -1 if none, -2 if max_samples was reached before the end.
int oracle_collision_check_segment(const dftpav_gridmap* map, const dftpav_corridor_params* vp, int N, const double* c,
                                   double dT, int singul, double dt, double res, int max_samples) {
  const Grid G{map};
  double total = 0.0;  // Trajectory::getTotalDuration
  for (int i = 0; i < N; i++) total += dT;
  int k = 0;
  for (double t = 0.0; t < total; t += dt, k++) {
    if (k >= max_samples) return -2;
    double st[12];
    oracle_traj_state(N, c, dT, singul, t, 0, 2.85, st, nullptr);
    const double sx = st[0], sy = st[1], yaw = st[6];
    // obb_ego (semantic_map_manager.cc:645-649)
    const double ox = sx + vp->veh_d_cr * std::cos(yaw), oy = sy + vp->veh_d_cr * std::sin(yaw);
    const double W = vp->veh_width, L = vp->veh_length;
    // shapes.cc:114-127
    const double ct = std::cos(yaw), sn = std::sin(yaw);
    const double c1x = ox + 0.5 * L * ct + 0.5 * W * sn, c1y = oy + 0.5 * L * sn - 0.5 * W * ct;
    const double c2x = ox + 0.5 * L * ct - 0.5 * W * sn, c2y = oy + 0.5 * L * sn + 0.5 * W * ct;
    const double c3x = ox - 0.5 * L * ct - 0.5 * W * sn, c3y = oy - 0.5 * L * sn + 0.5 * W * ct;
    const double c4x = ox - 0.5 * L * ct + 0.5 * W * sn, c4y = oy - 0.5 * L * sn - 0.5 * W * ct;
    auto edge = [&](double ax, double ay, double bx, double by) {  // shapes.cc:128-143
      const double dx = bx - ax, dy = by - ay;
      const double nrm = std::sqrt(dx * dx + dy * dy);
      for (double dl = res; dl < nrm; dl += res) {
        const double f = dl / nrm;
        if (G.occupied(f * dx + ax, f * dy + ay)) return true;
      }
      return false;
    };
    // the vertex order of the reference (edges 12, 23, 34, 41, then the corners) does not matter: any hit returns true
    if (edge(c1x, c1y, c2x, c2y) || edge(c2x, c2y, c3x, c3y) || edge(c3x, c3y, c4x, c4y) || edge(c4x, c4y, c1x, c1y) ||
        G.occupied(c1x, c1y) || G.occupied(c2x, c2y) || G.occupied(c3x, c3y) || G.occupied(c4x, c4y))
      return k;
  }
  return -1;
}

// states [n][3] = (x, y, yaw); cs (optional) [n][2] = (cos yaw, sin yaw) to use instead of libm; hpolys [n][4][4]:
// plane k = (n_x, n_y, p_x, p_y); expand (optional) [n][4] = the four expansion lengths (+dy, +dx, -dy, -dx).
int oracle_corridor_rectangles(const dftpav_gridmap* map, const dftpav_corridor_params* cp, int64_t n, const double* states,
                               const double* cs, double* hpolys, double* expand) {
  if (!map || !cp || !states || !hpolys || n < 0) return DFTPAV_ERR_INVALID_ARG;
  const Grid G{map};
  const double resolution = map->resolution;
  const double step = resolution * 1.0;  // MGR.cpp:1217
  const double limitBound = cp->limit_bound;
  for (int64_t q = 0; q < n; q++) {
    const double x = states[3 * q], y = states[3 * q + 1], yaw = states[3 * q + 2];
    const double c = cs ? cs[2 * q] : std::cos(yaw), s = cs ? cs[2 * q + 1] : std::sin(yaw);
    auto R = [&](double a, double b, double& ox, double& oy) {  // egoR * (a, b)
      ox = c * a + (-s) * b;
      oy = s * a + c * b;
    };
    double spx = x, spy = y;  // sourcePt
    Vp vp{cp->veh_width, cp->veh_length, cp->veh_d_cr};
    const Vp raw = vp;
    int notfin[4] = {1, 1, 1, 1};
    double ex[4] = {0.0, 0.0, 0.0, 0.0};
    while (notfin[0] || notfin[1] || notfin[2] || notfin[3]) {
      for (int i = 0; i < 4; i++) {
        if (!notfin[i]) continue;
        double a1, b1, a2, b2, na1, nb1, na2, nb2;  // body-frame corners of the strip: point1, point2, newpoint1, newpoint2
        const double hl = vp.length / 2.0, hw = vp.width / 2.0;
        switch (i) {
          case 0:  // +dy (MGR.cpp:1320-1349)
            a1 = hl + vp.d_cr, b1 = hw, a2 = -hl + vp.d_cr, b2 = hw;
            na1 = hl + vp.d_cr, nb1 = hw + step, na2 = -hl + vp.d_cr, nb2 = hw + step;
            break;
          case 1:  // +dx (1351-1380)
            a1 = hl + vp.d_cr, b1 = -hw, a2 = hl + vp.d_cr, b2 = hw;
            na1 = step + hl + vp.d_cr, nb1 = -hw, na2 = step + hl + vp.d_cr, nb2 = hw;
            break;
          case 2:  // -dy (1382-1411)
            a1 = -hl + vp.d_cr, b1 = -hw, a2 = hl + vp.d_cr, b2 = -hw;
            na1 = -hl + vp.d_cr, nb1 = -hw - step, na2 = hl + vp.d_cr, nb2 = -hw - step;
            break;
          default:  // -dx (1413-1429)
            a1 = -hl + vp.d_cr, b1 = hw, a2 = -hl + vp.d_cr, b2 = -hw;
            na1 = -hl + vp.d_cr - step, nb1 = hw, na2 = -hl + vp.d_cr - step, nb2 = -hw;
            break;
        }
        double p1x, p1y, p2x, p2y, n1x, n1y, n2x, n2y, tx, ty;
        R(a1, b1, tx, ty), p1x = spx + tx, p1y = spy + ty;
        R(a2, b2, tx, ty), p2x = spx + tx, p2y = spy + ty;
        R(na1, nb1, tx, ty), n1x = spx + tx, n1y = spy + ty;
        R(na2, nb2, tx, ty), n2x = spx + tx, n2y = spy + ty;
        if (G.line(p1x, p1y, n1x, n1y, resolution / 2.0) || G.line(n1x, n1y, n2x, n2y, resolution / 2.0) ||
            G.line(n2x, n2y, p2x, p2y, resolution / 2.0)) {
          notfin[i] = 0;
          continue;
        }
        ex[i] += step;
        if (ex[i] >= limitBound) {
          notfin[i] = 0;
          continue;
        }
        // the source rectangle grows by one step in direction i
        switch (i) {
          case 0: R(0.0, step / 2.0, tx, ty), vp.width = vp.width + step; break;
          case 1: R(step / 2.0, 0.0, tx, ty), vp.length = vp.length + step; break;
          case 2: R(0.0, -step / 2.0, tx, ty), vp.width = vp.width + step; break;
          default: R(-step / 2.0, 0.0, tx, ty), vp.length = vp.length + step; break;
        }
        spx = spx + tx, spy = spy + ty;
      }
    }
    // the polygon from the raw pose and the expansion lengths (MGR.cpp:1433-1460)
    double* h = hpolys + 16 * q;
    double tx, ty;
    R(raw.length / 2.0 + raw.d_cr + ex[1], raw.width / 2.0 + ex[0], tx, ty);
    h[0] = -s, h[1] = c, h[2] = x + tx, h[3] = y + ty;
    R(raw.length / 2.0 + raw.d_cr + ex[1], -raw.width / 2.0 - ex[2], tx, ty);
    h[4] = c, h[5] = s, h[6] = x + tx, h[7] = y + ty;
    R(-raw.length / 2.0 + raw.d_cr - ex[3], -raw.width / 2.0 - ex[2], tx, ty);
    h[8] = s, h[9] = -c, h[10] = x + tx, h[11] = y + ty;
    R(-raw.length / 2.0 + raw.d_cr - ex[3], raw.width / 2.0 + ex[0], tx, ty);
    h[12] = -c, h[13] = -s, h[14] = x + tx, h[15] = y + ty;
    if (expand)
      for (int k = 0; k < 4; k++) expand[4 * q + k] = ex[k];
  }
  return DFTPAV_OK;
}

}  // extern "C"
