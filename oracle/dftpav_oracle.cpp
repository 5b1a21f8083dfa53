// dftpav_oracle.cpp — CPU restatement of the Dftpav MINCO / L-BFGS hot path.
//
// TEST INFRASTRUCTURE ONLY (see dftpav_oracle.h).  Parity PINNED: bit-identical to the reference's own sources
// compiled in place into oracle/_ref (see dftpav_oracle.h).
//
// Plain C++17, no Eigen.  Every function cites the reference lines it restates; evaluation order of
// every floating-point expression follows the reference source (left-to-right association of the
// Eigen expressions, sequential sample loop, sequential accumulation into gdC/gdT/cost) so that the
// only differences to the real Eigen build are Eigen's internal reduction orders for dynamic-size
// dot products.  Compile WITHOUT -ffast-math and with -ffp-contract=off (see Makefile).
//
// Shorthand: OPT.cpp = src/Plan/traj_planner/src/traj_optimizer.cpp
//            MINCO.hpp = src/Plan/traj_planner/include/plan_utils/poly_traj_utils.hpp
//            LBFGS.hpp = src/Plan/traj_planner/include/geo_utils2d/lbfgs.hpp
#include "dftpav_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V2 {
  double x, y;
};
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator-(V2 a) { return {-a.x, -a.y}; }
inline V2 operator*(double s, V2 a) { return {s * a.x, s * a.y}; }
inline V2 operator*(V2 a, double s) { return {a.x * s, a.y * s}; }
inline V2 operator/(V2 a, double s) { return {a.x / s, a.y / s}; }
inline double dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
inline double norm(V2 a) { return std::sqrt(a.x * a.x + a.y * a.y); }

// 2x2 matrix, row-major fields m00 m01 / m10 m11
struct M2 {
  double a, b, c, d;
};
inline M2 operator*(double s, M2 m) { return {s * m.a, s * m.b, s * m.c, s * m.d}; }
inline M2 operator*(M2 m, double s) { return {m.a * s, m.b * s, m.c * s, m.d * s}; }
inline M2 operator/(M2 m, double s) { return {m.a / s, m.b / s, m.c / s, m.d / s}; }
inline M2 operator-(M2 p, M2 q) { return {p.a - q.a, p.b - q.b, p.c - q.c, p.d - q.d}; }
inline M2 operator*(M2 p, M2 q) {
  return {p.a * q.a + p.b * q.c, p.a * q.b + p.b * q.d, p.c * q.a + p.d * q.c, p.c * q.b + p.d * q.d};
}
inline V2 operator*(M2 m, V2 v) { return {m.a * v.x + m.b * v.y, m.c * v.x + m.d * v.y}; }
inline M2 transpose(M2 m) { return {m.a, m.c, m.b, m.d}; }
// row-vector * matrix
inline V2 rowmul(V2 r, M2 m) { return {r.x * m.a + r.y * m.c, r.x * m.b + r.y * m.d}; }
// outer product u v^T
inline M2 outer(V2 u, V2 v) { return {u.x * v.x, u.x * v.y, u.y * v.x, u.y * v.y}; }
// [l, Bl] with B = [[0,-1],[1,0]]  (OPT.cpp:596-598, 1426-1428)
inline M2 l_Bl(V2 l) { return {l.x, -l.y, l.y, l.x}; }
const M2 kBh = {0.0, -1.0, 1.0, 0.0};  // OPT.cpp:1741-1742

// ---------------------------------------------------------------------------------------------
// BandedSystem  (MINCO.hpp:727-853)
// ---------------------------------------------------------------------------------------------
struct Banded {
  int n = 0, lo = 0, up = 0;
  std::vector<double> d;
  void create(int n_, int p, int q) {  // MINCO.hpp:731-741
    n = n_;
    lo = p;
    up = q;
    d.assign((size_t)n * (lo + up + 1), 0.0);
  }
  inline double& at(int i, int j) { return d[(size_t)(i - j + up) * n + j]; }  // MINCO.hpp:766-772
  inline double at(int i, int j) const { return d[(size_t)(i - j + up) * n + j]; }

  void factorizeLU() {  // MINCO.hpp:776-800 (no pivoting, exact zeros skipped)
    for (int k = 0; k <= n - 2; k++) {
      int iM = std::min(k + lo, n - 1);
      double cVl = at(k, k);
      for (int i = k + 1; i <= iM; i++) {
        if (at(i, k) != 0.0) at(i, k) /= cVl;
      }
      int jM = std::min(k + up, n - 1);
      for (int j = k + 1; j <= jM; j++) {
        cVl = at(k, j);
        if (cVl != 0.0) {
          for (int i = k + 1; i <= iM; i++) {
            if (at(i, k) != 0.0) at(i, j) -= at(i, k) * cVl;
          }
        }
      }
    }
  }
  // b is n x 2, stored b[2*row + col]
  void solve(double* b) const {  // MINCO.hpp:805-826
    for (int j = 0; j <= n - 1; j++) {
      int iM = std::min(j + lo, n - 1);
      for (int i = j + 1; i <= iM; i++) {
        double a = at(i, j);
        if (a != 0.0) {
          b[2 * i] -= a * b[2 * j];
          b[2 * i + 1] -= a * b[2 * j + 1];
        }
      }
    }
    for (int j = n - 1; j >= 0; j--) {
      double dj = at(j, j);
      b[2 * j] /= dj;
      b[2 * j + 1] /= dj;
      int iM = std::max(0, j - up);
      for (int i = iM; i <= j - 1; i++) {
        double a = at(i, j);
        if (a != 0.0) {
          b[2 * i] -= a * b[2 * j];
          b[2 * i + 1] -= a * b[2 * j + 1];
        }
      }
    }
  }
  void solveAdj(double* b) const {  // MINCO.hpp:831-852
    for (int j = 0; j <= n - 1; j++) {
      double dj = at(j, j);
      b[2 * j] /= dj;
      b[2 * j + 1] /= dj;
      int iM = std::min(j + up, n - 1);
      for (int i = j + 1; i <= iM; i++) {
        double a = at(j, i);
        if (a != 0.0) {
          b[2 * i] -= a * b[2 * j];
          b[2 * i + 1] -= a * b[2 * j + 1];
        }
      }
    }
    for (int j = n - 1; j >= 0; j--) {
      int iM = std::max(0, j - lo);
      for (int i = iM; i <= j - 1; i++) {
        double a = at(j, i);
        if (a != 0.0) {
          b[2 * i] -= a * b[2 * j];
          b[2 * i + 1] -= a * b[2 * j + 1];
        }
      }
    }
  }
};

// MinJerkOpt::reset matrix fill (MINCO.hpp:897-947).  `set(i,j,v)` abstracts dense/banded storage.
template <class F>
void fill_minco_matrix(int N, F set) {
  set(0, 0, 1.0);
  set(1, 1, 1.0);
  set(2, 2, 2.0);
  for (int i = 0; i < N - 1; i++) {
    set(6 * i + 3, 6 * i + 3, 6.0);
    set(6 * i + 3, 6 * i + 4, 24.0);
    set(6 * i + 3, 6 * i + 5, 60.0);
    set(6 * i + 3, 6 * i + 9, -6.0);
    set(6 * i + 4, 6 * i + 4, 24.0);
    set(6 * i + 4, 6 * i + 5, 120.0);
    set(6 * i + 4, 6 * i + 10, -24.0);
    for (int k = 0; k < 6; k++) set(6 * i + 5, 6 * i + k, 1.0);
    for (int k = 0; k < 6; k++) set(6 * i + 6, 6 * i + k, 1.0);
    set(6 * i + 6, 6 * i + 6, -1.0);
    set(6 * i + 7, 6 * i + 1, 1.0);
    set(6 * i + 7, 6 * i + 2, 2.0);
    set(6 * i + 7, 6 * i + 3, 3.0);
    set(6 * i + 7, 6 * i + 4, 4.0);
    set(6 * i + 7, 6 * i + 5, 5.0);
    set(6 * i + 7, 6 * i + 7, -1.0);
    set(6 * i + 8, 6 * i + 2, 2.0);
    set(6 * i + 8, 6 * i + 3, 6.0);
    set(6 * i + 8, 6 * i + 4, 12.0);
    set(6 * i + 8, 6 * i + 5, 20.0);
    set(6 * i + 8, 6 * i + 8, -2.0);
  }
  for (int k = 0; k < 6; k++) set(6 * N - 3, 6 * N - 6 + k, 1.0);
  set(6 * N - 2, 6 * N - 5, 1.0);
  set(6 * N - 2, 6 * N - 4, 2.0);
  set(6 * N - 2, 6 * N - 3, 3.0);
  set(6 * N - 2, 6 * N - 2, 4.0);
  set(6 * N - 2, 6 * N - 1, 5.0);
  set(6 * N - 1, 6 * N - 4, 2.0);
  set(6 * N - 1, 6 * N - 3, 6.0);
  set(6 * N - 1, 6 * N - 2, 12.0);
  set(6 * N - 1, 6 * N - 1, 20.0);
}

// ---------------------------------------------------------------------------------------------
// MinJerkOpt  (MINCO.hpp:855-1095).  Matrices 6N x 2 stored [2*row + col].
// ---------------------------------------------------------------------------------------------
struct MinJerk {
  int N = 0;
  double head[6], tail[6];  // 2x3 column-major [p v a]
  std::vector<double> b, c, adj, gdC;
  Banded A;
  double t[6], tInv[6];
  double gdT = 0.0;
  double gdHead[6], gdTail[6];  // 2x3 column-major
  std::vector<double> gdP;      // 2 x (N-1) column-major

  void reset(int pieceNum) {  // MINCO.hpp:880-951
    N = pieceNum;
    A.create(6 * N, 6, 6);
    b.assign(12 * N, 0.0);
    c.assign(12 * N, 0.0);
    adj.assign(12 * N, 0.0);
    gdC.assign(12 * N, 0.0);
    gdP.assign(2 * (N - 1), 0.0);
    t[0] = 1.0;
    fill_minco_matrix(N, [&](int i, int j, double v) { A.at(i, j) = v; });
    A.factorizeLU();
  }

  void generate(const double* inPs, double dT, const double* headState, const double* tailState) {
    // MINCO.hpp:953-986
    std::memcpy(head, headState, sizeof(head));
    std::memcpy(tail, tailState, sizeof(tail));
    t[1] = dT;
    t[2] = t[1] * t[1];
    t[3] = t[2] * t[1];
    t[4] = t[2] * t[2];
    t[5] = t[4] * t[1];
    for (int k = 0; k < 6; k++) tInv[k] = 1.0 / t[k];
    std::fill(b.begin(), b.end(), 0.0);
    for (int col = 0; col < 2; col++) {
      b[2 * 0 + col] = head[0 + col];
      b[2 * 1 + col] = head[2 + col] * t[1];
      b[2 * 2 + col] = head[4 + col] * t[2];
      for (int i = 0; i < N - 1; i++) b[2 * (6 * i + 5) + col] = inPs[2 * i + col];
      b[2 * (6 * N - 3) + col] = tail[0 + col];
      b[2 * (6 * N - 2) + col] = tail[2 + col] * t[1];
      b[2 * (6 * N - 1) + col] = tail[4 + col] * t[2];
    }
    A.solve(b.data());
    for (int i = 0; i < N; i++)
      for (int k = 0; k < 6; k++)
        for (int col = 0; col < 2; col++) c[2 * (6 * i + k) + col] = b[2 * (6 * i + k) + col] * tInv[k];
  }

  inline double rdot(int r1, int r2) const { return c[2 * r1] * c[2 * r2] + c[2 * r1 + 1] * c[2 * r2 + 1]; }

  double getTrajJerkCost() const {  // MINCO.hpp:998-1009
    double energy = 0.0;
    for (int i = 0; i < N; i++) {
      int r3 = 6 * i + 3, r4 = 6 * i + 4, r5 = 6 * i + 5;
      energy += 36.0 * rdot(r3, r3) * t[1] + 144.0 * rdot(r4, r3) * t[2] + 192.0 * rdot(r4, r4) * t[3] +
                240.0 * rdot(r5, r3) * t[3] + 720.0 * rdot(r5, r4) * t[4] + 720.0 * rdot(r5, r5) * t[5];
    }
    return energy;
  }

  void initSmGradCost() {  // MINCO.hpp:1012-1035
    for (int i = 0; i < N; i++) {
      for (int col = 0; col < 2; col++) {
        double c3 = c[2 * (6 * i + 3) + col], c4 = c[2 * (6 * i + 4) + col], c5 = c[2 * (6 * i + 5) + col];
        gdC[2 * (6 * i + 5) + col] = 240.0 * c3 * t[3] + 720.0 * c4 * t[4] + 1440.0 * c5 * t[5];
        gdC[2 * (6 * i + 4) + col] = 144.0 * c3 * t[2] + 384.0 * c4 * t[3] + 720.0 * c5 * t[4];
        gdC[2 * (6 * i + 3) + col] = 72.0 * c3 * t[1] + 144.0 * c4 * t[2] + 240.0 * c5 * t[3];
        gdC[2 * (6 * i + 0) + col] = 0.0;
        gdC[2 * (6 * i + 1) + col] = 0.0;
        gdC[2 * (6 * i + 2) + col] = 0.0;
      }
    }
    gdT = 0.0;
    for (int i = 0; i < N; i++) {
      int r3 = 6 * i + 3, r4 = 6 * i + 4, r5 = 6 * i + 5;
      gdT += 36.0 * rdot(r3, r3) + 288.0 * rdot(r4, r3) * t[1] + 576.0 * rdot(r4, r4) * t[2] +
             720.0 * rdot(r5, r3) * t[2] + 2880.0 * rdot(r5, r4) * t[3] + 3600.0 * rdot(r5, r5) * t[4];
    }
  }

  void calGrads_PT() {  // MINCO.hpp:1037-1066
    for (int i = 0; i < N; i++)
      for (int k = 0; k < 6; k++)
        for (int col = 0; col < 2; col++)
          adj[2 * (6 * i + k) + col] = gdC[2 * (6 * i + k) + col] * tInv[k];
    A.solveAdj(adj.data());
    for (int i = 0; i < N - 1; i++) {
      gdP[2 * i] = adj[2 * (6 * i + 5)];
      gdP[2 * i + 1] = adj[2 * (6 * i + 5) + 1];
    }
    for (int k = 0; k < 3; k++)
      for (int col = 0; col < 2; col++) {
        gdHead[2 * k + col] = adj[2 * k + col] * t[k];
        gdTail[2 * k + col] = adj[2 * (6 * N - 3 + k) + col] * t[k];
      }
    gdT += head[2] * adj[2 * 1] + head[3] * adj[2 * 1 + 1];
    gdT += (head[4] * adj[2 * 2] + head[5] * adj[2 * 2 + 1]) * 2.0 * t[1];
    gdT += tail[2] * adj[2 * (6 * N - 2)] + tail[3] * adj[2 * (6 * N - 2) + 1];
    gdT += (tail[4] * adj[2 * (6 * N - 1)] + tail[5] * adj[2 * (6 * N - 1) + 1]) * 2.0 * t[1];
    double gdtInv[6];
    gdtInv[0] = 0.0;
    gdtInv[1] = -1.0 * tInv[2];
    gdtInv[2] = -2.0 * tInv[3];
    gdtInv[3] = -3.0 * tInv[4];
    gdtInv[4] = -4.0 * tInv[5];
    gdtInv[5] = -5.0 * tInv[5] * tInv[1];
    for (int i = 0; i < N; i++) {
      double s = 0.0;
      for (int k = 0; k < 6; k++) {
        int r = 6 * i + k;
        double gdcol = gdC[2 * r] * b[2 * r] + gdC[2 * r + 1] * b[2 * r + 1];
        s += gdtInv[k] * gdcol;
      }
      gdT += s;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Obstacle trajectory evaluation  (Piece / Trajectory, MINCO.hpp:77-112, 179-211, 510-603)
// ---------------------------------------------------------------------------------------------
struct ObsTraj {
  int P;
  const double* dur;     // [P]
  const double* coeffs;  // [P][12]  2x6 column-major, highest power first
  double start_time, duration;

  int locate(double& t) const {  // Trajectory::locatePieceIdx MINCO.hpp:510-528
    int idx;
    double d;
    for (idx = 0; idx < P && t > (d = dur[idx]); idx++) t -= d;
    if (idx == P) {
      idx--;
      t += dur[idx];
    }
    return idx;
  }
  static V2 col(const double* cm, int i) { return {cm[2 * i], cm[2 * i + 1]}; }
  static V2 piecePos(const double* cm, double t) {  // Piece::getPos MINCO.hpp:77-87
    V2 pos = {0.0, 0.0};
    double tn = 1.0;
    for (int i = 5; i >= 0; i--) {
      pos = pos + tn * col(cm, i);
      tn *= t;
    }
    return pos;
  }
  static V2 pieceVel(const double* cm, double t) {  // Piece::getdSigma MINCO.hpp:179-194
    V2 ds = {0.0, 0.0};
    double tn = 1.0;
    int n = 1;
    for (int i = 4; i >= 0; i--) {
      ds = ds + (n * tn) * col(cm, i);
      tn *= t;
      n++;
    }
    return ds;
  }
  static V2 pieceAcc(const double* cm, double t) {  // Piece::getddSigma MINCO.hpp:196-211
    V2 dd = {0.0, 0.0};
    double tn = 1.0;
    int m = 1, n = 2;
    for (int i = 3; i >= 0; i--) {
      dd = dd + ((m * n) * tn) * col(cm, i);
      tn *= t;
      m++;
      n++;
    }
    return dd;
  }
  V2 getPos(double t) const {
    int i = locate(t);
    return piecePos(coeffs + 12 * i, t);
  }
  V2 getdSigma(double t) const {
    int i = locate(t);
    return pieceVel(coeffs + 12 * i, t);
  }
  V2 getddSigma(double t) const {
    int i = locate(t);
    return pieceAcc(coeffs + 12 * i, t);
  }
  M2 getR(double t) const {  // Piece::getR MINCO.hpp:89-98, obstacle singul = +1 (traj_manager.cpp:773)
    int i = locate(t);
    V2 v = pieceVel(coeffs + 12 * i, t);
    M2 rot = {v.x, -v.y, v.y, v.x};
    return (1.0 * rot) / norm(v);
  }
  M2 getRdot(double t) const {  // Piece::getRdot MINCO.hpp:100-112
    int i = locate(t);
    V2 v = pieceVel(coeffs + 12 * i, t);
    V2 a = pieceAcc(coeffs + 12 * i, t);
    M2 ta = {a.x, -a.y, a.y, a.x};
    M2 tv = {v.x, -v.y, v.y, v.x};
    double nv = norm(v);
    return 1.0 * (ta / nv - (tv / std::pow(nv, 3)) * dot(v, a));
  }
};

// positiveSmoothedL1  (OPT.cpp:783-806)
inline void smoothedL1(double pe, double x, double& f, double& df) {
  const double half = 0.5 * pe;
  const double f3c = 1.0 / (pe * pe);
  const double f4c = -0.5 * f3c / pe;
  const double d2c = 3.0 * f3c;
  const double d3c = 4.0 * f4c;
  if (x < pe) {
    f = (f4c * x + f3c) * x * x * x;
    df = (d3c * x + d2c) * x * x;
  } else {
    f = x - half;
    df = 1.0;
  }
}

// log_sum_exp  (OPT.cpp:1686-1707): overwrites v with the exponentials (quirk Q7).
inline double log_sum_exp(double alpha, double* v, int n, double& exp_sum) {
  double d0 = v[0];
  if (alpha > 0) {
    for (int j = 1; j < n; j++) d0 = std::max(d0, v[j]);
  } else {
    for (int j = 1; j < n; j++) d0 = std::min(d0, v[j]);
  }
  exp_sum = 0;
  for (int j = 0; j < n; j++) {
    v[j] = std::exp(alpha * (v[j] - d0));
    exp_sum += v[j];
  }
  return std::log(exp_sum) / alpha + d0;
}

// ---------------------------------------------------------------------------------------------
// One problem = one PolyTrajOptimizer instance after OptimizeTrajectory's prelude (OPT.cpp:13-115)
// ---------------------------------------------------------------------------------------------
struct Problem {
  const dftpav_params* prm;
  int S = 0, n = 0, rows = 0;
  std::vector<int> N, singul;
  std::vector<std::vector<double>> ini, fin;  // clamped copies, 6 each (OPT.cpp:55-76)
  std::vector<double> initT;
  std::vector<const double*> innerPts;
  // corridor: per segment, per sample: H and normalised planes (copy; OPT.cpp:15,49-52)
  std::vector<std::vector<int>> H;
  std::vector<std::vector<std::vector<double>>> planes;
  std::vector<ObsTraj> obs;
  bool have_obs = false;
  double t_now = 0.0;
  std::vector<MinJerk> mj;
  V2 vec_le[5];
  double veh_length = 0.0;
  long evals = 0;
  double last_terms[5] = {0, 0, 0, 0, 0};

  int samplesOf(int N_) const { return (N_ - 2) * (prm->traj_resolution + 1) + 2 * (prm->des_traj_resolution + 1); }

  // returns 0, or DFTPAV_ERR_BAD_INPUT on the reference's `return false` preconditions
  int setup(const dftpav_params* prm_, const dftpav_batch* bt, int p) {
    prm = prm_;
    const int g0 = bt->seg_ptr[p], g1 = bt->seg_ptr[p + 1];
    S = g1 - g0;
    if (S <= 0) return DFTPAV_ERR_BAD_INPUT;
    // offsets of this problem's segments in the concatenated arrays
    long inner_off = 0, sample_off = 0;
    for (int g = 0; g < g0; g++) {
      inner_off += 2L * (bt->piece_num[g] - 1);
      sample_off += samplesOf(bt->piece_num[g]);
    }
    long plane_off = 0;
    if (bt->plane_cnt) {
      for (long q = 0; q < sample_off; q++) plane_off += bt->plane_cnt[q];
    } else {
      plane_off = sample_off * bt->uniform_H;
    }
    // vehicle box (OPT.cpp:1749-1775)
    double width = prm->veh_width + 2 * prm->half_margin;
    double length = prm->veh_length + 2 * prm->half_margin;
    veh_length = length;
    double dcr = prm->veh_d_cr;
    vec_le[0] = {dcr + length / 2.0, width / 2.0};
    vec_le[1] = {dcr + length / 2.0, -width / 2.0};
    vec_le[2] = {dcr - length / 2.0, -width / 2.0};
    vec_le[3] = {dcr - length / 2.0, width / 2.0};
    vec_le[4] = vec_le[0];  // quirk Q1: first vertex pushed twice (OPT.cpp:1773)

    N.resize(S);
    singul.resize(S);
    ini.resize(S);
    fin.resize(S);
    initT.resize(S);
    innerPts.resize(S);
    H.resize(S);
    planes.resize(S);
    mj.resize(S);
    n = 0;
    rows = 0;
    int rc = 0;
    for (int s = 0; s < S; s++) {
      int g = g0 + s;
      N[s] = bt->piece_num[g];
      if (N[s] < 2) rc = DFTPAV_ERR_BAD_INPUT;  // OPT.cpp:38-41 "There is only a piece?"
      singul[s] = bt->singul[g];
      initT[s] = bt->init_T[g];
      if (initT[s] < prm->mini_T) rc = DFTPAV_ERR_BAD_INPUT;  // OPT.cpp:30-33
      innerPts[s] = bt->inner_pts + inner_off;
      inner_off += 2L * (N[s] - 1);
      int M = samplesOf(N[s]);
      H[s].resize(M);
      planes[s].resize(M);
      for (int q = 0; q < M; q++) {
        int h = bt->plane_cnt ? bt->plane_cnt[sample_off + q] : bt->uniform_H;
        H[s][q] = h;
        planes[s][q].assign(bt->planes + 4 * plane_off, bt->planes + 4 * (plane_off + h));
        plane_off += h;
        for (int k = 0; k < h; k++) {  // colwise().normalize() of the top two rows, OPT.cpp:51
          double* pl = &planes[s][q][4 * k];
          double z = pl[0] * pl[0] + pl[1] * pl[1];
          if (z > 0.0) {
            double nz = std::sqrt(z);
            pl[0] /= nz;
            pl[1] /= nz;
          }
        }
      }
      sample_off += M;
      ini[s].assign(bt->ini_state + 6 * g, bt->ini_state + 6 * g + 6);
      fin[s].assign(bt->fin_state + 6 * g, bt->fin_state + 6 * g + 6);
      // clamps OPT.cpp:55-76
      double max_vel, max_acc;
      if (singul[s] > 0) {
        max_vel = prm->max_forward_vel;
        max_acc = prm->max_forward_acc;
      } else {
        max_vel = prm->max_backward_vel;
        max_acc = prm->max_backward_acc;
      }
      auto clampcol = [](double* v, double lim) {
        double nv = std::sqrt(v[0] * v[0] + v[1] * v[1]);
        if (nv >= lim) {
          // Eigen normalized(): v / sqrt(squaredNorm) if squaredNorm > 0
          double z = v[0] * v[0] + v[1] * v[1];
          double nx = v[0], ny = v[1];
          if (z > 0.0) {
            double sz = std::sqrt(z);
            nx = v[0] / sz;
            ny = v[1] / sz;
          }
          v[0] = nx * (lim - 1.0e-2);
          v[1] = ny * (lim - 1.0e-2);
        }
      };
      clampcol(&ini[s][2], max_vel);
      clampcol(&fin[s][2], max_vel);
      clampcol(&ini[s][4], max_acc);
      clampcol(&fin[s][4], max_acc);
      if (rc == 0) mj[s].reset(N[s]);  // OPT.cpp:79
      n += 2 * (N[s] - 1);
      rows += 6 * N[s];
    }
    n += S;            // OPT.cpp:84
    n += 2 * (S - 1);  // OPT.cpp:85
    n += 1 * (S - 1);  // OPT.cpp:86
    t_now = bt->t_now ? bt->t_now[p] : 0.0;
    have_obs = (bt->obs_ptr != nullptr);
    obs.clear();
    if (have_obs) {
      for (int k = bt->obs_ptr[p]; k < bt->obs_ptr[p + 1]; k++) {
        int u = bt->obs_idx[k];
        ObsTraj o;
        o.P = bt->obs_piece_ptr[u + 1] - bt->obs_piece_ptr[u];
        o.dur = bt->obs_piece_dur + bt->obs_piece_ptr[u];
        o.coeffs = bt->obs_coeffs + 12L * bt->obs_piece_ptr[u];
        o.start_time = bt->obs_start_time[u];
        o.duration = bt->obs_duration[u];
        obs.push_back(o);
      }
    }
    return rc;
  }

  // RealT2VirtualT  OPT.cpp:361-369
  double real2virt(double RT) const {
    double mini_T = prm->mini_T;
    return RT > 1.0 + mini_T ? (std::sqrt(2.0 * RT - 1.0 - 2 * mini_T) - 1.0)
                             : (1.0 - std::sqrt(2.0 / (RT - mini_T) - 1.0));
  }
  // VirtualT2RealT  OPT.cpp:372-379
  double virt2real(double VT) const {
    double mini_T = prm->mini_T;
    return VT > 0.0 ? ((0.5 * VT + 1.0) * VT + 1.0) + mini_T : 1.0 / ((0.5 * VT - 1.0) * VT + 1.0) + mini_T;
  }

  void initialX(double* x) const {  // OPT.cpp:96-115
    int off = 0;
    for (int s = 0; s < S; s++) {
      std::memcpy(x + off, innerPts[s], sizeof(double) * 2 * (N[s] - 1));
      off += 2 * (N[s] - 1);
    }
    for (int s = 0; s < S; s++) x[off + s] = real2virt(initT[s]);
    off += S;
    for (int s = 0; s < S - 1; s++) {
      x[off] = fin[s][0];
      x[off + 1] = fin[s][1];
      off += 2;
    }
    for (int s = 0; s < S - 1; s++) x[off + s] = std::atan2(fin[s][3], fin[s][2]);
  }

  // dynamicObsGradCostP  (OPT.cpp:1311-1684)
  double dynamicObs(MinJerk& J, double omg, double step, double t, const double* beta0, const double* beta1,
                    double gama, int pieceid, int trajres, V2 sigma, V2 dsigma, V2 ddsigma, M2 ego_R,
                    int trajid, double trajtime) {
    const int singul_ = singul[trajid];
    if (obs.size() < 1) return 0.0;
    const double alpha = prm->lse_alpha;
    const double d_min = prm->surround_clearance + std::log(8.0) / alpha;  // OPT.cpp:1336
    double temp0 = norm(dsigma);
    double temp0_reci = (temp0 != 0.0) ? 1.0 / temp0 : 0.0;
    double temp3 = temp0_reci * temp0_reci;
    const V2* vec_lo = vec_le;  // lo_k = le_k (OPT.cpp:1769)
    double totalPenalty = 0.0;
    const double wei_surround = prm->wei_surround;

    for (size_t sur_id = 0; sur_id < obs.size(); sur_id++) {
      const ObsTraj& ob = obs[sur_id];
      double offsettime = t_now - ob.start_time + trajtime;
      double pt_time = offsettime + t;
      V2 sp, sv, sa;
      if (pt_time < ob.duration) {  // OPT.cpp:1372-1378
        sp = ob.getPos(pt_time);
        sv = ob.getdSigma(pt_time);
        sa = ob.getddSigma(pt_time);
      } else {  // OPT.cpp:1379-1389
        sa = ob.getddSigma(ob.duration);
        double exceed_time = pt_time - ob.duration;
        sv = ob.getdSigma(ob.duration) + exceed_time * sa;
        sp = ob.getPos(ob.duration) + exceed_time * ob.getdSigma(ob.duration) + (0.5 * sa) * exceed_time * exceed_time;
      }
      if (norm(sp - sigma) > veh_length * prm->obs_cull_factor) continue;  // OPT.cpp:1393

      M2 sur_R = ob.getR(pt_time);  // OPT.cpp:1410 (quirk Q6: evaluated past the duration too)
      double sum_exp_U[4], d_U[4], sum_exp_E[4], d_E[4];
      V2 ego_normal[4], sur_normal[4];
      double dUoe[4][4], dEeo[4][4];
      M2 F_dle[4], F_le[4];
      for (int e = 0; e < 4; e++) {  // OPT.cpp:1417-1461
        V2 le = vec_le[e];
        V2 delta_le = vec_le[e + 1] - vec_le[e];
        double dn = norm(delta_le);
        double dn_inv = 1 / dn;
        M2 t1 = l_Bl(delta_le);
        F_dle[e] = (double(singul_) * transpose(t1)) * temp0_reci - outer(dsigma, ego_R * delta_le) * temp3;
        M2 t2 = l_Bl(le);
        F_le[e] = (double(singul_) * transpose(t2)) * temp0_reci - outer(dsigma, ego_R * le) * temp3;
        V2 H_tilde = ((kBh * ego_R) * delta_le) * dn_inv;
        ego_normal[e] = H_tilde;
        double d_U_e_tilde = dot(H_tilde, (sp - sigma) - ego_R * le);
        for (int o = 0; o < 4; o++) dUoe[e][o] = dot(rowmul(H_tilde, sur_R), vec_lo[o]);  // (H^T R_u) l_o, OPT.cpp:1449
        double es;
        d_U[e] = log_sum_exp(-alpha, dUoe[e], 4, es) + d_U_e_tilde;
        sum_exp_U[e] = es;
      }
      for (int o = 0; o < 4; o++) {  // OPT.cpp:1468-1496
        V2 lo = vec_lo[o];
        V2 delta_lo = vec_lo[o + 1] - vec_lo[o];
        double dn = norm(delta_lo);
        double dn_inv = 1 / dn;
        V2 H_tilde = ((kBh * sur_R) * delta_lo) * dn_inv;
        sur_normal[o] = H_tilde;
        double d_E_o_tilde = dot(H_tilde, (sigma - sp) - sur_R * lo);
        for (int e = 0; e < 4; e++) dEeo[o][e] = dot(rowmul(H_tilde, ego_R), vec_le[e]);  // OPT.cpp:1485
        double es;
        d_E[o] = log_sum_exp(-alpha, dEeo[o], 4, es) + d_E_o_tilde;
        sum_exp_E[o] = es;
      }
      double d_test[8];
      for (int k = 0; k < 4; k++) {
        d_test[k] = d_U[k];
        d_test[4 + k] = d_E[k];
      }
      double exp_sum_d = 0;
      double costp = d_min - log_sum_exp(alpha, d_test, 8, exp_sum_d);  // OPT.cpp:1502
      if (costp <= 0) continue;
      double pena, penaD;
      smoothedL1(prm->smooth_eps, costp, pena, penaD);
      totalPenalty += omg * step * wei_surround * pena;  // OPT.cpp:1508

      V2 pG_sigma = {0.0, 0.0};  // OPT.cpp:1511-1523
      for (int e = 0; e < 4; e++) pG_sigma = pG_sigma - (d_test[e] / exp_sum_d) * (-ego_normal[e]);
      for (int o = 0; o < 4; o++) pG_sigma = pG_sigma - (d_test[o + 4] / exp_sum_d) * sur_normal[o];

      V2 pG_dsigma = {0.0, 0.0};  // OPT.cpp:1528-1573
      for (int e = 0; e < 4; e++) {
        V2 le = vec_le[e];
        V2 delta_le = vec_le[e + 1] - vec_le[e];
        double dnrm = norm(delta_le);
        V2 a1 = (F_dle[e] * kBh) * ((-sp + sigma) + ego_R * le);
        V2 a2 = ((F_le[e] * kBh) * ego_R) * delta_le;
        V2 pdU = (a1 - a2) / dnrm;
        for (int o = 0; o < 4; o++) {
          V2 lo = vec_lo[o];
          V2 pd = ((F_dle[e] * transpose(kBh)) * (sur_R * lo)) / dnrm;
          pdU = pdU + (dUoe[e][o] / sum_exp_U[e]) * pd;
        }
        pG_dsigma = pG_dsigma - (d_test[e] / exp_sum_d) * pdU;
      }
      for (int o = 0; o < 4; o++) {
        V2 delta_lo = vec_lo[o + 1] - vec_lo[o];
        double dnrm = norm(delta_lo);
        V2 pdE = {0.0, 0.0};
        for (int e = 0; e < 4; e++) {
          V2 pd = (((F_le[e] * kBh) * sur_R) * delta_lo) / dnrm;
          pdE = pdE + (dEeo[o][e] / sum_exp_E[o]) * pd;
        }
        pG_dsigma = pG_dsigma - (d_test[o + 4] / exp_sum_d) * pdE;
      }
      double pG_tbar = dot(pG_sigma, dsigma) + dot(pG_dsigma, ddsigma);  // OPT.cpp:1578-1580

      double pG_that = 0.0;  // OPT.cpp:1586-1646
      M2 Rud = ob.getRdot(pt_time);
      for (int e = 0; e < 4; e++) {
        double pdU = dot(ego_normal[e], sv);
        for (int o = 0; o < 4; o++) {
          double pd = dot(rowmul(ego_normal[e], Rud), vec_lo[o]);  // OPT.cpp:1614-1615
          pdU += dUoe[e][o] / sum_exp_U[e] * pd;
        }
        pG_that -= d_test[e] / exp_sum_d * pdU;
      }
      for (int o = 0; o < 4; o++) {
        V2 lo = vec_lo[o];
        V2 delta_lo = vec_lo[o + 1] - vec_lo[o];
        double dnrm = norm(delta_lo);
        V2 r1 = ((kBh * Rud) * delta_lo) / dnrm;
        V2 r2 = ((kBh * sur_R) * delta_lo) / dnrm;
        double pdE = dot(r1, (sigma - sp) - sur_R * lo) + dot(r2, -sv - Rud * lo);
        for (int e = 0; e < 4; e++) {
          V2 row = rowmul(rowmul(ego_R * vec_le[e], kBh), Rud);
          double pd = dot(row, delta_lo) / dnrm;
          pdE += dEeo[o][e] / sum_exp_E[o] * pd;
        }
        pG_that -= d_test[o + 4] / exp_sum_d * pdE;
      }

      // scatter  OPT.cpp:1650-1676
      double gradViolaPt = gama * pG_tbar;
      double wc = omg * step * wei_surround * penaD;
      double* gd = &J.gdC[2 * (pieceid * 6)];
      for (int k = 0; k < 6; k++) {
        gd[2 * k] += wc * (beta0[k] * pG_sigma.x + beta1[k] * pG_dsigma.x);
        gd[2 * k + 1] += wc * (beta0[k] * pG_sigma.y + beta1[k] * pG_dsigma.y);
      }
      J.gdT += omg * wei_surround * (pena / trajres + penaD * gradViolaPt * step);
      J.gdT += omg * step * wei_surround * pG_that * penaD * pieceid;
      J.gdT += omg * step * wei_surround * gama * pG_that * penaD;
      for (int idx = 0; idx < trajid; idx++)  // quirk Q5
        J.gdT += omg * step * wei_surround * pG_that * penaD * N[trajid];
    }
    return totalPenalty;
  }

  // addPVAGradCost2CT  (OPT.cpp:422-779)
  void addPVAGradCost2CT(double* costs, int trajid, double trajtime) {
    MinJerk& J = mj[trajid];
    const int Np = N[trajid];
    const int singul_ = singul[trajid];
    double max_vel, max_cur, max_acc;
    if (singul_ > 0) {
      max_vel = prm->max_forward_vel;
      max_cur = prm->max_forward_cur;
      max_acc = prm->max_forward_acc;
    } else {
      max_vel = prm->max_backward_vel;
      max_cur = prm->max_backward_cur;
      max_acc = prm->max_backward_acc;
    }
    const double wei_obs = prm->wei_obs, wei_feas = prm->wei_feas, epis = prm->help_eps;
    const double pe = prm->smooth_eps, cw = prm->cur_weight_scale;
    costs[0] = costs[1] = costs[2] = 0.0;
    double t = 0;
    int pointid = -1;
    double beta0[6], beta1[6], beta2[6], beta3[6];
    for (int i = 0; i < Np; ++i) {
      int K = (i == 0 || i == Np - 1) ? prm->des_traj_resolution : prm->traj_resolution;
      const double* c = &J.c[2 * (i * 6)];
      double step = J.t[1] / K;
      double s1 = 0.0;
      for (int j = 0; j <= K; ++j) {
        double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
        beta0[0] = 1.0, beta0[1] = s1, beta0[2] = s2, beta0[3] = s3, beta0[4] = s4, beta0[5] = s5;
        beta1[0] = 0.0, beta1[1] = 1.0, beta1[2] = 2.0 * s1, beta1[3] = 3.0 * s2, beta1[4] = 4.0 * s3, beta1[5] = 5.0 * s4;
        beta2[0] = 0.0, beta2[1] = 0.0, beta2[2] = 2.0, beta2[3] = 6.0 * s1, beta2[4] = 12.0 * s2, beta2[5] = 20.0 * s3;
        beta3[0] = 0.0, beta3[1] = 0.0, beta3[2] = 0.0, beta3[3] = 6.0, beta3[4] = 24.0 * s1, beta3[5] = 60.0 * s2;
        double alpha = 1.0 / K * j;
        s1 += step;  // quirk Q3: accumulated
        pointid++;
        V2 sigma = {0, 0}, dsigma = {0, 0}, ddsigma = {0, 0}, dddsigma = {0, 0};
        for (int k = 0; k < 6; k++) {
          sigma.x += c[2 * k] * beta0[k];
          sigma.y += c[2 * k + 1] * beta0[k];
          dsigma.x += c[2 * k] * beta1[k];
          dsigma.y += c[2 * k + 1] * beta1[k];
          ddsigma.x += c[2 * k] * beta2[k];
          ddsigma.y += c[2 * k + 1] * beta2[k];
          dddsigma.x += c[2 * k] * beta3[k];
          dddsigma.y += c[2 * k + 1] * beta3[k];
        }
        double omg = (j == 0 || j == K) ? 0.5 : 1.0;
        double z_h0 = norm(dsigma);
        double z_h1 = dot(ddsigma, dsigma);
        double z_h2 = dot(dddsigma, dsigma);
        double z_h3 = ddsigma.y * dsigma.x + (-ddsigma.x) * dsigma.y;  // ddsigma^T B_h dsigma
        if (z_h0 < prm->min_speed || (j == 0 && i == 0) || (i == Np - 1 && j == K)) continue;  // OPT.cpp:550
        double vel2_reci = 1.0 / (z_h0 * z_h0);
        double vel2_reci_e = 1.0 / (z_h0 * z_h0 + epis);
        double vel3_2_reci_e = vel2_reci_e * std::sqrt(vel2_reci_e);
        z_h0 = 1.0 / z_h0;
        double z_h4 = z_h1 * vel2_reci;
        double violaVel = 1.0 / vel2_reci - max_vel * max_vel;
        double acc2 = z_h1 * z_h1 * vel2_reci;
        double cur = z_h3 * vel3_2_reci_e;
        double violaAcc = acc2 - max_acc * max_acc;
        double violaCurL = cur - max_cur;
        double violaCurR = -cur - max_cur;
        M2 rot = {dsigma.x, -dsigma.y, dsigma.y, dsigma.x};
        M2 ego_R = (double(singul_) * rot) * z_h0;
        M2 temp_a = {ddsigma.x, -ddsigma.y, ddsigma.y, ddsigma.x};
        M2 R_dot = double(singul_) * (temp_a * z_h0 - ((rot * vel2_reci) * z_h0) * z_h1);
        double* gd = &J.gdC[2 * (i * 6)];

        const int corr_k = H[trajid][pointid];
        const double* hp = planes[trajid][pointid].data();
        for (int v = 0; v < 5; v++) {  // quirk Q1: 5 entries
          V2 le = vec_le[v];
          V2 bpt = sigma + ego_R * le;
          M2 tl = l_Bl(le);
          for (int k = 0; k < corr_k; k++) {
            V2 nrm = {hp[4 * k], hp[4 * k + 1]};
            V2 pnt = {hp[4 * k + 2], hp[4 * k + 3]};
            double violaPos = dot(nrm, bpt - pnt);
            if (violaPos > 0) {
              double pena, penaD;
              smoothedL1(pe, violaPos, pena, penaD);
              M2 Mx = (double(singul_) * tl) * z_h0 - outer(ego_R * le, dsigma) * vel2_reci;
              double gradViolaPt = dot(alpha * nrm, dsigma + R_dot * le);
              double wc = omg * step * wei_obs * penaD;
              for (int r = 0; r < 6; r++) {
                V2 b1n = beta1[r] * nrm;  // (beta1 * n^T) row r
                V2 pc = {beta0[r] * nrm.x + (b1n.x * Mx.a + b1n.y * Mx.c), beta0[r] * nrm.y + (b1n.x * Mx.b + b1n.y * Mx.d)};
                gd[2 * r] += wc * pc.x;
                gd[2 * r + 1] += wc * pc.y;
              }
              J.gdT += omg * wei_obs * (penaD * gradViolaPt * step + pena / K);
              costs[0] += omg * step * wei_obs * pena;
            }
          }
        }
        if (have_obs) {  // OPT.cpp:636-638
          costs[1] += dynamicObs(J, omg, step, t + step * j, beta0, beta1, alpha, i, K, sigma, dsigma, ddsigma, ego_R,
                                 trajid, trajtime);
        }
        if (violaVel > 0.0) {  // OPT.cpp:642-653
          double pena, penaD;
          smoothedL1(pe, violaVel, pena, penaD);
          double gradViolaVt = 2.0 * alpha * z_h1;
          double wc = omg * step * wei_feas * penaD;
          for (int r = 0; r < 6; r++) {
            double b = 2.0 * beta1[r];
            gd[2 * r] += wc * (b * dsigma.x);
            gd[2 * r + 1] += wc * (b * dsigma.y);
          }
          J.gdT += omg * wei_feas * (penaD * gradViolaVt * step + pena / K);
          costs[2] += omg * step * wei_feas * pena;
        }
        if (violaAcc > 0.0) {  // OPT.cpp:655-665
          double pena, penaD;
          smoothedL1(pe, violaAcc, pena, penaD);
          V2 w1 = z_h4 * ddsigma - (z_h4 * z_h4) * dsigma;
          V2 w2 = dsigma;
          double gradViolaAt = 2.0 * alpha * (z_h4 * (dot(ddsigma, ddsigma) + z_h2) - z_h4 * z_h4 * z_h1);
          double wc = omg * step * wei_feas * penaD;
          for (int r = 0; r < 6; r++) {
            double b1 = 2.0 * beta1[r], b2 = 2.0 * beta2[r] * z_h4;
            gd[2 * r] += wc * (b1 * w1.x + b2 * w2.x);
            gd[2 * r + 1] += wc * (b1 * w1.y + b2 * w2.y);
          }
          J.gdT += omg * wei_feas * (penaD * gradViolaAt * step + pena / K);
          costs[2] += omg * step * wei_feas * pena;
        }
        // curvature  OPT.cpp:684-705.  ddsigma^T B_h = (ddy, -ddx);  dsigma^T B_h^T = (-dy, dx)
        V2 aB = {ddsigma.y, -ddsigma.x};
        V2 Bv = {-dsigma.y, dsigma.x};
        double z_d3 = dddsigma.y * dsigma.x + (-dddsigma.x) * dsigma.y;  // dddsigma^T B_h dsigma
        if (violaCurL > 0.0) {
          double pena, penaD;
          smoothedL1(pe, violaCurL, pena, penaD);
          V2 w1 = vel3_2_reci_e * aB - (3 * vel3_2_reci_e * vel2_reci_e * z_h3) * dsigma;
          double gradT = alpha * vel3_2_reci_e * (z_d3 - 3 * vel2_reci_e * z_h3 * z_h1);
          double wc = omg * step * wei_feas * cw * penaD;
          for (int r = 0; r < 6; r++) {
            double b2 = beta2[r] * vel3_2_reci_e;
            gd[2 * r] += wc * (beta1[r] * w1.x + b2 * Bv.x);
            gd[2 * r + 1] += wc * (beta1[r] * w1.y + b2 * Bv.y);
          }
          J.gdT += omg * wei_feas * cw * (penaD * gradT * step + pena / K);
          costs[2] += omg * step * wei_feas * cw * pena;
        }
        if (violaCurR > 0.0) {
          double pena, penaD;
          smoothedL1(pe, violaCurR, pena, penaD);
          V2 w1 = vel3_2_reci_e * aB - (3 * vel3_2_reci_e * vel2_reci_e * z_h3) * dsigma;
          double gradT = -(alpha * vel3_2_reci_e * (z_d3 - 3 * vel2_reci_e * z_h3 * z_h1));
          double wc = omg * step * wei_feas * cw * penaD;
          for (int r = 0; r < 6; r++) {
            double b2 = beta2[r] * vel3_2_reci_e;
            gd[2 * r] += wc * (-(beta1[r] * w1.x + b2 * Bv.x));
            gd[2 * r + 1] += wc * (-(beta1[r] * w1.y + b2 * Bv.y));
          }
          J.gdT += omg * wei_feas * cw * (penaD * gradT * step + pena / K);
          costs[2] += omg * step * wei_feas * cw * pena;
        }
      }
      t += J.t[1];
    }
  }

  // costFunctionCallback  (OPT.cpp:206-350)
  double eval(const double* x, double* grad) {
    double total_smcost = 0.0, total_timecost = 0.0, penalty_cost = 0.0;
    std::vector<int> poff(S);
    int off = 0;
    for (int s = 0; s < S; s++) {
      poff[s] = off;
      off += 2 * (N[s] - 1);
    }
    for (int k = 0; k < off; k++) grad[k] = 0.0;
    const double* tv = x + off;
    double* gradt = grad + off;
    off += S;
    std::vector<double> T(S);
    for (int s = 0; s < S; s++) T[s] = virt2real(tv[s]);
    std::vector<double> trajtimes;
    trajtimes.push_back(0.0);
    for (int s = 0; s < S; s++) trajtimes.push_back(T[s]);  // quirk Q5: not cumulative
    const double* gear = x + off;
    double* gradGear = grad + off;
    for (int k = 0; k < 2 * (S - 1); k++) gradGear[k] = 0.0;
    off += 2 * (S - 1);
    const double* angles = x + off;
    double* gradAngles = grad + off;
    for (int k = 0; k < S - 1; k++) gradAngles[k] = 0.0;
    const double nsv = prm->non_sinv;
    double tc[5] = {0, 0, 0, 0, 0};

    for (int s = 0; s < S; s++) {
      double costs[3];
      double IniS[6], FinS[6];
      std::memcpy(IniS, ini[s].data(), sizeof(IniS));
      std::memcpy(FinS, fin[s].data(), sizeof(FinS));
      if (s > 0) {  // OPT.cpp:273-277
        double theta = angles[s - 1];
        IniS[0] = gear[2 * (s - 1)];
        IniS[1] = gear[2 * (s - 1) + 1];
        IniS[2] = -nsv * std::cos(theta);
        IniS[3] = -nsv * std::sin(theta);
      }
      if (s < S - 1) {  // OPT.cpp:278-282
        double theta = angles[s];
        FinS[0] = gear[2 * s];
        FinS[1] = gear[2 * s + 1];
        FinS[2] = nsv * std::cos(theta);
        FinS[3] = nsv * std::sin(theta);
      }
      mj[s].generate(x + poff[s], T[s] / N[s], IniS, FinS);
      mj[s].initSmGradCost();
      double smoo_cost = mj[s].getTrajJerkCost();
      addPVAGradCost2CT(costs, s, trajtimes[s]);
      total_smcost += smoo_cost;
      penalty_cost += (costs[0] + costs[1]) + costs[2];
      tc[2] += costs[0];
      tc[3] += costs[1];
      tc[4] += costs[2];
    }
    for (int s = 0; s < S; s++) {
      mj[s].calGrads_PT();
      std::memcpy(grad + poff[s], mj[s].gdP.data(), sizeof(double) * 2 * (N[s] - 1));
      const double* gIni = mj[s].gdHead;
      const double* gFin = mj[s].gdTail;
      if (prm->gear_opt) {  // OPT.cpp:308-319
        if (s > 0) {
          double theta = angles[s - 1];
          gradGear[2 * (s - 1)] += gIni[0];
          gradGear[2 * (s - 1) + 1] += gIni[1];
          gradAngles[s - 1] += gIni[2] * (nsv * std::sin(theta)) + gIni[3] * (-nsv * std::cos(theta));
        }
        if (s < S - 1) {
          double theta = angles[s];
          gradGear[2 * s] += gFin[0];
          gradGear[2 * s + 1] += gFin[1];
          gradAngles[s] += gFin[2] * (-nsv * std::sin(theta)) + gFin[3] * (nsv * std::cos(theta));
        }
      }
      // VirtualTGradCost scalar overload  OPT.cpp:405-419
      double RT = T[s], VT = tv[s], gdRT = mj[s].gdT / N[s];
      double gdVT2Rt;
      if (VT > 0) {
        gdVT2Rt = VT + 1.0;
      } else {
        double denSqrt = (0.5 * VT - 1.0) * VT + 1.0;
        gdVT2Rt = (1.0 - VT) / (denSqrt * denSqrt);
      }
      gradt[s] = (gdRT + prm->wei_time) * gdVT2Rt;
      double time_cost = RT * prm->wei_time;
      total_timecost += time_cost;
    }
    evals++;
    tc[0] = total_smcost;
    tc[1] = total_timecost;
    std::memcpy(last_terms, tc, sizeof(tc));
    return total_smcost + total_timecost + penalty_cost;
  }
};

// ---------------------------------------------------------------------------------------------
// L-BFGS  (LBFGS.hpp:276-390 line_search_lewisoverton, 440-751 lbfgs_optimize)
// ---------------------------------------------------------------------------------------------
struct LbfgsParam {
  int mem_size, past, max_iterations, max_linesearch;
  double g_epsilon, delta, min_step, max_step, f_dec_coeff, s_curv_coeff, cautious_factor, machine_prec;
};
LbfgsParam lbfgs_param_of(const dftpav_params* p) {
  LbfgsParam q;
  q.mem_size = p->lbfgs_mem_size;
  q.past = p->lbfgs_past;
  q.max_iterations = p->lbfgs_max_iterations;
  q.max_linesearch = p->lbfgs_max_linesearch;
  q.g_epsilon = p->lbfgs_g_epsilon;
  q.delta = p->lbfgs_delta;
  q.min_step = p->lbfgs_min_step;
  q.max_step = p->lbfgs_max_step;
  q.f_dec_coeff = p->lbfgs_f_dec_coeff;
  q.s_curv_coeff = p->lbfgs_s_curv_coeff;
  q.cautious_factor = p->lbfgs_cautious_factor;
  q.machine_prec = p->lbfgs_machine_prec;
  return q;
}

inline double vdot(const double* a, const double* b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}
inline double vmaxabs(const double* a, int n) {
  double m = std::fabs(a[0]);
  for (int i = 1; i < n; i++) m = std::max(m, std::fabs(a[i]));
  return m;
}

template <class Eval>
int line_search_lewisoverton(int n, double* x, double& f, double* g, double& stp, const double* s, const double* xp,
                             const double* gp, double stpmin, double stpmax, Eval& eval, const LbfgsParam& param) {
  int count = 0;
  bool brackt = false, touched = false;
  double finit, dginit, dgtest, dstest;
  double mu = 0.0, nu = stpmax;
  if (!(stp > 0.0)) return DFTPAV_LBFGSERR_INVALIDPARAMETERS;
  dginit = vdot(gp, s, n);
  if (0.0 < dginit) return DFTPAV_LBFGSERR_INCREASEGRADIENT;
  finit = f;
  dgtest = param.f_dec_coeff * dginit;
  dstest = param.s_curv_coeff * dginit;
  while (true) {
    for (int i = 0; i < n; i++) x[i] = xp[i] + stp * s[i];
    f = eval(x, g);
    ++count;
    if (std::isinf(f) || std::isnan(f)) return DFTPAV_LBFGSERR_INVALID_FUNCVAL;
    if (param.past > 0 && std::fabs(finit - f) / (std::fabs(finit) + 1.0) < param.delta / param.past)
      return count;  // LBFGS.hpp:326-329
    if (f > finit + stp * dgtest) {
      nu = stp;
      brackt = true;
    } else {
      if (vdot(g, s, n) < dstest) {
        mu = stp;
      } else {
        return count;
      }
    }
    if (param.max_linesearch <= count) return DFTPAV_LBFGSERR_MAXIMUMLINESEARCH;
    if (brackt && (nu - mu) < param.machine_prec * nu) return DFTPAV_LBFGSERR_WIDTHTOOSMALL;
    if (brackt) {
      stp = 0.5 * (mu + nu);
    } else {
      stp *= 2.0;
    }
    if (stp < stpmin) return DFTPAV_LBFGSERR_MINIMUMSTEP;
    if (stp > stpmax) {
      if (touched) {
        return DFTPAV_LBFGSERR_MAXIMUMSTEP;
      } else {
        touched = true;
        stp = stpmax;
      }
    }
  }
}

// ---- design study for the GPU kernel (profiles/README.md, "Next"): the two-loop recursion in blocks of BS history
// entries.  Inside a block the alphas follow from BS independent dot products with the direction at the START of the block
// plus the Gram products s_i . y_k of the block (alpha_1 = (s_1.q - alpha_0 s_1.y_0) / ys_1, ...), so the BS reductions no
// longer depend on each other.  Mathematically the reference's recursion; the rounding differs.  Enabled only by the
// environment variable DFTPAV_ORACLE_BLOCKED_TWO_LOOP=BS (tests/test_oracle_lbfgs.py measures what it does to the results).
inline int blocked_two_loop_size() {
  static const int bs = [] { const char* e = std::getenv("DFTPAV_ORACLE_BLOCKED_TWO_LOOP"); return e ? std::atoi(e) : 0; }();
  return bs;
}
inline void blocked_two_loop(double* d, const double* S, const double* Y, const double* ysv, double* alpha, int n, int m,
                             int end, int bound, double scale, int BS) {
  std::vector<int> idx(BS);
  std::vector<double> u(BS), a(BS);
  // first loop: newest to oldest
  int j = end;
  for (int i0 = 0; i0 < bound; i0 += BS) {
    const int cnt = std::min(BS, bound - i0);
    for (int k = 0; k < cnt; k++) {
      j = (j + m - 1) % m;
      idx[k] = j;
      u[k] = vdot(&S[(size_t)j * n], d, n);
    }
    for (int k = 0; k < cnt; k++) {
      double v = u[k];
      for (int l = 0; l < k; l++) v -= a[l] * vdot(&S[(size_t)idx[k] * n], &Y[(size_t)idx[l] * n], n);
      a[k] = v / ysv[idx[k]];
      alpha[idx[k]] = a[k];
    }
    for (int k = 0; k < cnt; k++) {
      const double* y = &Y[(size_t)idx[k] * n];
      for (int q = 0; q < n; q++) d[q] -= a[k] * y[q];
    }
  }
  for (int q = 0; q < n; q++) d[q] *= scale;
  // second loop: oldest to newest, starting where the first loop stopped
  for (int i0 = 0; i0 < bound; i0 += BS) {
    const int cnt = std::min(BS, bound - i0);
    for (int k = 0; k < cnt; k++) {
      idx[k] = j;
      u[k] = vdot(&Y[(size_t)j * n], d, n);
      j = (j + 1) % m;
    }
    for (int k = 0; k < cnt; k++) {
      double v = u[k];
      for (int l = 0; l < k; l++) v += a[l] * vdot(&Y[(size_t)idx[k] * n], &S[(size_t)idx[l] * n], n);
      const double beta = v / ysv[idx[k]];
      a[k] = alpha[idx[k]] - beta;
    }
    for (int k = 0; k < cnt; k++) {
      const double* sv = &S[(size_t)idx[k] * n];
      for (int q = 0; q < n; q++) d[q] += a[k] * sv[q];
    }
  }
}

template <class Eval>
int lbfgs_optimize(int n, double* x, double& f, Eval& eval, const LbfgsParam& param, int* iters_out, double* trace,
                   int trace_cap) {
  int ret, i, j, k, ls, end, bound;
  double step, step_min, step_max, fx, ys, yy;
  double gnorm_inf, xnorm_inf, beta, rate, cau;
  const int m = param.mem_size;
  if (iters_out) *iters_out = 0;
  if (n <= 0) return DFTPAV_LBFGSERR_INVALID_N;
  if (m <= 0) return DFTPAV_LBFGSERR_INVALID_MEMSIZE;
  if (param.g_epsilon < 0.0) return DFTPAV_LBFGSERR_INVALID_GEPSILON;
  if (param.past < 0) return DFTPAV_LBFGSERR_INVALID_TESTPERIOD;
  if (param.delta < 0.0) return DFTPAV_LBFGSERR_INVALID_DELTA;
  if (param.min_step < 0.0) return DFTPAV_LBFGSERR_INVALID_MINSTEP;
  if (param.max_step < param.min_step) return DFTPAV_LBFGSERR_INVALID_MAXSTEP;
  if (!(param.f_dec_coeff > 0.0 && param.f_dec_coeff < 1.0)) return DFTPAV_LBFGSERR_INVALID_FDECCOEFF;
  if (!(param.s_curv_coeff < 1.0 && param.s_curv_coeff > param.f_dec_coeff)) return DFTPAV_LBFGSERR_INVALID_SCURVCOEFF;
  if (!(param.machine_prec > 0.0)) return DFTPAV_LBFGSERR_INVALID_MACHINEPREC;
  if (param.max_linesearch <= 0) return DFTPAV_LBFGSERR_INVALID_MAXLINESEARCH;

  std::vector<double> xp(n), g(n), gp(n), d(n), pf(std::max(1, param.past));
  std::vector<double> lm_alpha(m, 0.0), lm_s((size_t)n * m, 0.0), lm_y((size_t)n * m, 0.0), lm_ys(m, 0.0);

  fx = eval(x, g.data());
  pf[0] = fx;
  for (i = 0; i < n; i++) d[i] = -g[i];
  gnorm_inf = vmaxabs(g.data(), n);
  xnorm_inf = vmaxabs(x, n);
  k = 0;
  if (gnorm_inf / std::max(1.0, xnorm_inf) < param.g_epsilon) {
    ret = DFTPAV_LBFGS_CONVERGENCE;
  } else {
    step = 1.0 / std::sqrt(vdot(d.data(), d.data(), n));
    k = 1;
    end = 0;
    bound = 0;
    while (true) {
      std::copy(x, x + n, xp.begin());
      gp = g;
      step_min = param.min_step;
      step_max = param.max_step;
      ls = line_search_lewisoverton(n, x, fx, g.data(), step, d.data(), xp.data(), gp.data(), step_min, step_max, eval,
                                    param);
      if (trace && k - 1 < trace_cap) {
        trace[3 * (k - 1)] = fx;
        trace[3 * (k - 1) + 1] = step;
        trace[3 * (k - 1) + 2] = (double)ls;
      }
      if (ls < 0) {
        std::copy(xp.begin(), xp.end(), x);
        g = gp;
        ret = ls;
        break;
      }
      gnorm_inf = vmaxabs(g.data(), n);
      xnorm_inf = vmaxabs(x, n);
      if (gnorm_inf / std::max(1.0, xnorm_inf) < param.g_epsilon) {
        ret = DFTPAV_LBFGS_CONVERGENCE;
        break;
      }
      if (0 < param.past) {
        if (param.past <= k) {
          rate = std::fabs(pf[k % param.past] - fx) / std::max(1.0, std::fabs(fx));
          if (rate < param.delta) {
            ret = DFTPAV_LBFGS_STOP;
            break;
          }
        }
        pf[k % param.past] = fx;
      }
      if (param.max_iterations != 0 && param.max_iterations <= k) {
        ret = DFTPAV_LBFGSERR_MAXIMUMITERATION;
        break;
      }
      ++k;
      double* se = &lm_s[(size_t)end * n];
      double* ye = &lm_y[(size_t)end * n];
      for (i = 0; i < n; i++) {
        se[i] = x[i] - xp[i];
        ye[i] = g[i] - gp[i];
      }
      ys = vdot(ye, se, n);
      yy = vdot(ye, ye, n);
      lm_ys[end] = ys;
      for (i = 0; i < n; i++) d[i] = -g[i];
      cau = vdot(se, se, n) * std::sqrt(vdot(gp.data(), gp.data(), n)) * param.cautious_factor;
      if (ys > cau) {
        ++bound;
        bound = m < bound ? m : bound;
        end = (end + 1) % m;
        j = end;
        if (blocked_two_loop_size() > 1) {
          // EXPERIMENT (round-2 design study, never used by the tests' oracle calls): blocked form of the recursion
          blocked_two_loop(d.data(), lm_s.data(), lm_y.data(), lm_ys.data(), lm_alpha.data(), n, m, end, bound, ys / yy,
                           blocked_two_loop_size());
          step = 1.0;
          continue;
        }
        for (i = 0; i < bound; ++i) {
          j = (j + m - 1) % m;
          lm_alpha[j] = vdot(&lm_s[(size_t)j * n], d.data(), n) / lm_ys[j];
          double a = -lm_alpha[j];
          const double* yj = &lm_y[(size_t)j * n];
          for (int q = 0; q < n; q++) d[q] += a * yj[q];
        }
        double sc = ys / yy;
        for (int q = 0; q < n; q++) d[q] *= sc;
        for (i = 0; i < bound; ++i) {
          beta = vdot(&lm_y[(size_t)j * n], d.data(), n) / lm_ys[j];
          double a = lm_alpha[j] - beta;
          const double* sj = &lm_s[(size_t)j * n];
          for (int q = 0; q < n; q++) d[q] += a * sj[q];
          j = (j + 1) % m;
        }
      }
      step = 1.0;
    }
  }
  f = fx;
  if (iters_out) *iters_out = k;
  return ret;
}

struct CbEval {
  oracle_fg_t fg;
  void* user;
  int n;
  long evals = 0;
  double operator()(const double* x, double* g) {
    evals++;
    return fg(user, x, g, n);
  }
};
struct ProbEval {
  Problem* P;
  double operator()(const double* x, double* g) { return P->eval(x, g); }
};

long sample_prefix(const dftpav_params* prm, const dftpav_batch* bt, int g) {
  long s = 0;
  for (int q = 0; q < g; q++)
    s += (bt->piece_num[q] - 2) * (prm->traj_resolution + 1) + 2 * (prm->des_traj_resolution + 1);
  return s;
}

void batch_strides(const dftpav_batch* bt, int* n_stride, int* row_stride) {
  int ns = 0, rs = 0;
  for (int p = 0; p < bt->B; p++) {
    int n = 0, rows = 0, S = bt->seg_ptr[p + 1] - bt->seg_ptr[p];
    for (int g = bt->seg_ptr[p]; g < bt->seg_ptr[p + 1]; g++) {
      n += 2 * (bt->piece_num[g] - 1);
      rows += 6 * bt->piece_num[g];
    }
    n += S + 3 * (S - 1);
    ns = std::max(ns, n);
    rs = std::max(rs, rows);
  }
  *n_stride = ns;
  *row_stride = rs;
}

}  // namespace

extern "C" {

int oracle_problem_dims(const dftpav_params* prm, const dftpav_batch* batch, int p, int* n, int* rows, int* samples) {
  if (!prm || !batch || p < 0 || p >= batch->B) return DFTPAV_ERR_INVALID_ARG;
  int S = batch->seg_ptr[p + 1] - batch->seg_ptr[p];
  int nn = 0, rr = 0, ss = 0;
  for (int g = batch->seg_ptr[p]; g < batch->seg_ptr[p + 1]; g++) {
    nn += 2 * (batch->piece_num[g] - 1);
    rr += 6 * batch->piece_num[g];
    ss += (batch->piece_num[g] - 2) * (prm->traj_resolution + 1) + 2 * (prm->des_traj_resolution + 1);
  }
  nn += S + 3 * (S - 1);
  if (n) *n = nn;
  if (rows) *rows = rr;
  if (samples) *samples = ss;
  return DFTPAV_OK;
}

int oracle_eval(const dftpav_params* prm, const dftpav_batch* batch, int p, const double* x, double* f, double* g,
                double* coeffs, double* piece_dT, double* terms) {
  if (!prm || !batch || !x || !f || !g || p < 0 || p >= batch->B) return DFTPAV_ERR_INVALID_ARG;
  Problem P;
  int rc = P.setup(prm, batch, p);
  if (rc) return rc;
  *f = P.eval(x, g);
  if (coeffs) {
    int off = 0;
    for (int s = 0; s < P.S; s++) {
      std::memcpy(coeffs + off, P.mj[s].c.data(), sizeof(double) * 12 * P.N[s]);
      off += 12 * P.N[s];
    }
  }
  if (piece_dT)
    for (int s = 0; s < P.S; s++) piece_dT[s] = P.mj[s].t[1];
  if (terms) std::memcpy(terms, P.last_terms, sizeof(double) * 5);
  return DFTPAV_OK;
}

int oracle_optimize(const dftpav_params* prm, const dftpav_batch* batch, int p0, int p1, int threads,
                    const dftpav_result* out, double* trace, int trace_cap) {
  if (!prm || !batch || !out || p0 < 0 || p1 > batch->B || p0 > p1) return DFTPAV_ERR_INVALID_ARG;
  int n_stride, row_stride;
  batch_strides(batch, &n_stride, &row_stride);
  const LbfgsParam lp = lbfgs_param_of(prm);
  (void)sample_prefix;
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int p = p0; p < p1; p++) {
    Problem P;
    int rc = P.setup(prm, batch, p);
    std::vector<double> x(std::max(1, P.n), 0.0);
    double final_cost = 0.0;
    // a problem the reference rejects before solving (OPT.cpp:26-48 `return false`) is reported like the C-ABI does
    int result = DFTPAV_PROBLEM_REJECTED, iters = 0;
    bool success = false;
    if (rc == 0) {
      P.initialX(x.data());
      ProbEval ev{&P};
      result = lbfgs_optimize(P.n, x.data(), final_cost, ev, lp, &iters,
                              trace ? trace + (size_t)p * trace_cap * 3 : nullptr, trace_cap);
      // OPT.cpp:176-201
      if (result == DFTPAV_LBFGS_CONVERGENCE || result == DFTPAV_LBFGS_CANCELED || result == DFTPAV_LBFGS_STOP ||
          result == DFTPAV_LBFGSERR_MAXIMUMITERATION) {
        success = true;
      } else if (result == DFTPAV_LBFGSERR_MAXIMUMLINESEARCH) {
        success = true;
      }
      if (final_cost >= prm->cost_limit) success = false;
    }
    if (out->x) {
      double* dst = out->x + (size_t)p * n_stride;
      for (int i = 0; i < n_stride; i++) dst[i] = i < P.n ? x[i] : 0.0;
    }
    if (out->cost) out->cost[p] = final_cost;
    if (out->status) out->status[p] = result;
    if (out->success) out->success[p] = success ? 1 : 0;
    if (out->iters) out->iters[p] = iters;
    if (out->evals) out->evals[p] = (int)P.evals;
    if (out->coeffs && rc == 0) {
      double* dst = out->coeffs + (size_t)p * row_stride * 2;
      int off = 0;
      for (int s = 0; s < P.S; s++) {
        std::memcpy(dst + off, P.mj[s].c.data(), sizeof(double) * 12 * P.N[s]);
        off += 12 * P.N[s];
      }
      for (int i = off; i < row_stride * 2; i++) dst[i] = 0.0;
    }
    if (out->piece_dT && rc == 0)
      for (int s = 0; s < P.S; s++) out->piece_dT[batch->seg_ptr[p] + s] = P.mj[s].t[1];
  }
  return DFTPAV_OK;
}

int oracle_lbfgs(int n, double* x, double* f, oracle_fg_t fg, void* user, const dftpav_params* prm, int* iters,
                 int* evals, double* trace, int trace_cap) {
  if (!x || !f || !fg || !prm) return DFTPAV_ERR_INVALID_ARG;
  CbEval ev{fg, user, n};
  LbfgsParam lp = lbfgs_param_of(prm);
  int ret = lbfgs_optimize(n, x, *f, ev, lp, iters, trace, trace_cap);
  if (evals) *evals = (int)ev.evals;
  return ret;
}

int oracle_minco_matrix(int N, double* dense, double* lu_band) {
  if (N < 2) return DFTPAV_ERR_INVALID_ARG;
  int n = 6 * N;
  if (dense) {
    std::fill(dense, dense + (size_t)n * n, 0.0);
    fill_minco_matrix(N, [&](int i, int j, double v) { dense[(size_t)i * n + j] = v; });
  }
  if (lu_band) {
    MinJerk J;
    J.reset(N);
    std::memcpy(lu_band, J.A.d.data(), sizeof(double) * J.A.d.size());
  }
  return DFTPAV_OK;
}

int oracle_minco_generate(int N, const double* P, double dT, const double* head, const double* tail, double* c) {
  if (N < 2 || !P || !head || !tail || !c) return DFTPAV_ERR_INVALID_ARG;
  MinJerk J;
  J.reset(N);
  J.generate(P, dT, head, tail);
  std::memcpy(c, J.c.data(), sizeof(double) * 12 * N);
  return DFTPAV_OK;
}


int oracle_minco_run(int N, const double* P, double dT, const double* head, const double* tail, const double* extra_gdC,
                     double extra_gdT, double* c, double* energy, double* gdC0, double* gdT0, double* gdP, double* gdHead,
                     double* gdTail, double* gdT) {
  if (N < 2 || !P || !head || !tail) return DFTPAV_ERR_INVALID_ARG;
  MinJerk J;
  J.reset(N);
  J.generate(P, dT, head, tail);
  std::memcpy(c, J.c.data(), sizeof(double) * 12 * N);
  *energy = J.getTrajJerkCost();
  J.initSmGradCost();
  std::memcpy(gdC0, J.gdC.data(), sizeof(double) * 12 * N);
  *gdT0 = J.gdT;
  if (extra_gdC)
    for (int i = 0; i < 12 * N; i++) J.gdC[i] += extra_gdC[i];
  J.gdT += extra_gdT;
  J.calGrads_PT();
  std::memcpy(gdP, J.gdP.data(), sizeof(double) * 2 * (N - 1));
  std::memcpy(gdHead, J.gdHead, sizeof(double) * 6);
  std::memcpy(gdTail, J.gdTail, sizeof(double) * 6);
  *gdT = J.gdT;
  return DFTPAV_OK;
}

int oracle_traj_eval(int N, const double* P, double dT, const double* head, const double* tail, int singul, double t,
                     double* out) {
  if (N < 2 || singul != 1) return DFTPAV_ERR_INVALID_ARG;  // the hot path evaluates obstacle trajectories with singul = +1
  MinJerk J;
  J.reset(N);
  J.generate(P, dT, head, tail);
  // Piece coefficient matrix = c.block<6,2>(6i,0)^T with the columns reversed (MINCO.hpp:991): highest power first
  std::vector<double> dur(N, dT), cm(12 * (size_t)N);
  for (int i = 0; i < N; i++)
    for (int k = 0; k < 6; k++)
      for (int col = 0; col < 2; col++) cm[12 * i + 2 * k + col] = J.c[2 * (6 * i + 5 - k) + col];
  ObsTraj tr{N, dur.data(), cm.data(), 0.0, dT * N};
  const V2 p = tr.getPos(t), v = tr.getdSigma(t), a = tr.getddSigma(t);
  const M2 R = tr.getR(t), Rd = tr.getRdot(t);
  out[0] = p.x, out[1] = p.y, out[2] = v.x, out[3] = v.y, out[4] = a.x, out[5] = a.y;
  out[6] = R.a, out[7] = R.b, out[8] = R.c, out[9] = R.d;
  out[10] = Rd.a, out[11] = Rd.b, out[12] = Rd.c, out[13] = Rd.d;
  out[14] = std::nan("");  // Piece::getAngle is not on the hot path
  double tl = t;
  out[15] = tr.locate(tl);
  out[16] = tl;
  return DFTPAV_OK;
}

// The read-back side of the hot path: MinJerkOpt::getTraj (MINCO.hpp:987-997) followed by the Trajectory / Piece evaluators
// consumers sample the result with (GetState 378-406; getPos, getdSigma, getddSigma, getAngle, getCurv, getVel, getAcc,
// getLatAcc, getSteer 77-87, 179-211, 233-296, 554-645).  Same signature and output layout as ref_traj_state in
// oracle/ref_minco_driver.cpp, which runs the reference's own header.
int oracle_traj_state(int N, const double* c, double dT, int singul, double t, int clamp, double wheel_base, double* out,
                      double* gs) {
  if (N < 1 || !c || !out) return DFTPAV_ERR_INVALID_ARG;
  std::vector<double> dur(N, dT), cm(12 * (size_t)N);
  for (int i = 0; i < N; i++)
    for (int k = 0; k < 6; k++)
      for (int col = 0; col < 2; col++) cm[12 * i + 2 * k + col] = c[2 * (6 * i + 5 - k) + col];
  ObsTraj tr{N, dur.data(), cm.data(), 0.0, 0.0};
  double total = 0.0;  // Trajectory::getTotalDuration (MINCO.hpp:425-434)
  for (int i = 0; i < N; i++) total += dT;
  double tt = t;
  if (clamp && tt > total) tt = total;
  const double sg = (double)singul;
  auto state_at = [&](double tq, double* o) {
    const V2 p = tr.getPos(tq), v = tr.getdSigma(tq), a = tr.getddSigma(tq);
    o[0] = p.x, o[1] = p.y, o[2] = v.x, o[3] = v.y, o[4] = a.x, o[5] = a.y;
    const double nv = norm(v);
    o[6] = std::atan2(sg * v.y, sg * v.x);                                        // getAngle, MINCO.hpp:233-241
    o[7] = nv < 1e-6 ? 0.0 : sg * (v.x * a.y - v.y * a.x) / std::pow(nv, 3);      // getCurv, 244-256
    o[8] = sg * nv;                                                               // getVel, 260-265
    o[9] = nv < 1e-6 ? 0.0 : sg * (v.x * a.x + v.y * a.y) / nv;                   // getAcc, 268-278
    o[10] = std::atan(wheel_base * o[7]);                                         // getSteer, 293-296
    o[11] = nv < 1e-6 ? 0.0 : sg * (v.x * a.y - v.y * a.x) / nv;                  // getLatAcc, 280-290
  };
  state_at(tt, out);
  if (gs) {  // Trajectory::GetState + Piece::getStateExpPos (MINCO.hpp:299-338, 378-406): always clamps
    double ti = t;
    if (ti > total) ti = total;
    const V2 p = tr.getPos(ti), v = tr.getdSigma(ti), a = tr.getddSigma(ti);
    gs[0] = std::atan2(sg * v.y, sg * v.x);
    gs[2] = sg * norm(v);
    // quirk Q11: getStateExpPos tests `abs(otherstate[2]) < 1e-6` with an UNQUALIFIED abs (MINCO.hpp:316); compiled as the
    // reference writes it (oracle/_ref, g++ 13) that is ::abs(int): the speed is truncated to an integer first, so curvature,
    // acceleration and steer are zeroed whenever |v| < 1 m/s.  (The getters getCurv / getAcc / getSteer test norm() < 1e-6.)
    if (std::abs((int)gs[2]) < 1e-6) {
      gs[1] = gs[3] = gs[4] = 0.0;
    } else {
      gs[1] = (v.x * a.y - v.y * a.x) / std::pow(gs[2], 3);
      gs[3] = (v.x * a.x + v.y * a.y) / gs[2];
      gs[4] = std::atan(wheel_base * gs[1]);
    }
    gs[5] = p.x;
  }
  return DFTPAV_OK;
}

// ---- analytic test functions (shared, via C callback, with the reference-header build) ----
double oracle_testfunc(int id, const double* x, double* g, int n) {
  double f = 0.0;
  if (id == 0) {  // Rosenbrock chain
    for (int i = 0; i < n; i++) g[i] = 0.0;
    for (int i = 0; i + 1 < n; i++) {
      double a = x[i + 1] - x[i] * x[i], b = 1.0 - x[i];
      f += 100.0 * a * a + b * b;
      g[i] += -400.0 * a * x[i] - 2.0 * b;
      g[i + 1] += 200.0 * a;
    }
  } else if (id == 1) {  // ill-conditioned quadratic with a shifted minimum
    for (int i = 0; i < n; i++) {
      double w = std::pow(10.0, 4.0 * i / std::max(1, n - 1));
      double d = x[i] - 0.1 * (i + 1);
      f += 0.5 * w * d * d;
      g[i] = w * d;
    }
  } else {  // piecewise-smooth: quadratic + |x| kinks (exercises the weak-Wolfe bracketing)
    for (int i = 0; i < n; i++) {
      double d = x[i] - 1.0;
      f += 0.5 * (i + 1) * d * d + 0.3 * std::fabs(x[i]);
      g[i] = (i + 1) * d + (x[i] > 0 ? 0.3 : (x[i] < 0 ? -0.3 : 0.0));
    }
  }
  return f;
}
double oracle_testfunc_fg(void* user, const double* x, double* g, int n) {
  return oracle_testfunc(*reinterpret_cast<int*>(user), x, g, n);
}

// ---- problem handles: let an external optimizer loop (the reference header) drive the oracle's cost function ----
void* oracle_problem_create(const dftpav_params* prm, const dftpav_batch* batch, int p) {
  if (!prm || !batch || p < 0 || p >= batch->B) return nullptr;
  Problem* P = new Problem();
  if (P->setup(prm, batch, p) != 0) {
    delete P;
    return nullptr;
  }
  return P;
}
void oracle_problem_free(void* h) { delete reinterpret_cast<Problem*>(h); }
int oracle_problem_n(void* h) { return reinterpret_cast<Problem*>(h)->n; }
void oracle_problem_initial_x(void* h, double* x) { reinterpret_cast<Problem*>(h)->initialX(x); }
double oracle_problem_fg(void* h, const double* x, double* g, int n) {
  (void)n;
  return reinterpret_cast<Problem*>(h)->eval(x, g);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
