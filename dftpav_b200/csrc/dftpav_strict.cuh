// dftpav_strict.cuh — strict-order evaluation for the VERIFICATION build of the solver (-DDFTPAV_STRICT=1 -fmad=false).
//
// The production kernels change the arithmetic of the reference in a handful of rounding-level ways (FMA contraction,
// parallel reductions, products with A^-1[:, R] instead of banded substitutions, 1/ys products, j*step, rsqrt, vertex 0
// weighted 2, per-vertex corridor sums).  The reference algorithm amplifies such differences
// (the penalty has curvature 1e8 inside its smoothing zone), so a production solve cannot be compared bit by bit with the
// reference.  This build removes every one of them: ONE lane per trajectory evaluates costFunctionCallback exactly as the
// reference writes it -
//   * sequential sample loop, s1 += step (quirk Q3), every sample's terms added to gdC / gdT / cost in source order,
//     the front-left vertex visited twice (quirk Q1)                                  traj_optimizer.cpp:422-779
//   * sqrt and divisions where the reference has them                                  traj_optimizer.cpp:527-559
//   * BandedSystem::solve / solveAdj on the LU factor with divisions by U(j,j)        poly_traj_utils.hpp:805-852
//   * MinJerkOpt::generate / initSmGradCost / getTrajJerkCost / calGrads_PT           poly_traj_utils.hpp:953-1066
//   * left-to-right association of every expression, no FMA (-fmad=false)
// and the L-BFGS vector operations use sequential dot products and divisions by ys (two_loop_generic<true>).  What is
// shared with the production build is everything else: the solver state machine of solve_kernel (fetch, first
// evaluation, Lewis-Overton trials, convergence tests, history ring indices, revert on failure, result write-back), the
// upload path, the plane layout and the batch plumbing.  sqrt and division are correctly rounded on both sides and no
// libm function is on the path of configurations without gear segments and moving obstacles, so complete solves of this
// build are BIT-IDENTICAL to the CPU oracle (tests/test_gpu_strict.py) - which proves the state machine to the last bit
// and leaves the rounding-level differences listed above as the only thing between a production solve and the reference.
// Moving obstacles (exp / log) are not supported by this build; gear segments (sin / cos) are compared as a fraction.
#pragma once

namespace dftpav {

// sequential banded substitutions on the packed LU columns: A(i,j) = C[j][i - j + 6]
__device__ __forceinline__ double lu_at(const double* C, int i, int j) { return C[(size_t)j * kLuLd + (i - j + 6)]; }

__device__ __noinline__ void strict_band_solve(const double* C, double* b, int n) {  // poly_traj_utils.hpp:805-826
  for (int j = 0; j <= n - 1; j++) {
    const int iM = min(j + 6, n - 1);
    for (int i = j + 1; i <= iM; i++) {
      const double a = lu_at(C, i, j);
      if (a != 0.0) {
        b[2 * i] -= a * b[2 * j];
        b[2 * i + 1] -= a * b[2 * j + 1];
      }
    }
  }
  for (int j = n - 1; j >= 0; j--) {
    const double dj = lu_at(C, j, j);
    b[2 * j] /= dj;
    b[2 * j + 1] /= dj;
    const int iM = max(0, j - 6);
    for (int i = iM; i <= j - 1; i++) {
      const double a = lu_at(C, i, j);
      if (a != 0.0) {
        b[2 * i] -= a * b[2 * j];
        b[2 * i + 1] -= a * b[2 * j + 1];
      }
    }
  }
}

__device__ __noinline__ void strict_band_solve_adj(const double* C, double* b, int n) {  // poly_traj_utils.hpp:831-852
  for (int j = 0; j <= n - 1; j++) {
    const double dj = lu_at(C, j, j);
    b[2 * j] /= dj;
    b[2 * j + 1] /= dj;
    const int iM = min(j + 6, n - 1);
    for (int i = j + 1; i <= iM; i++) {
      const double a = lu_at(C, j, i);
      if (a != 0.0) {
        b[2 * i] -= a * b[2 * j];
        b[2 * i + 1] -= a * b[2 * j + 1];
      }
    }
  }
  for (int j = n - 1; j >= 0; j--) {
    const int iM = max(0, j - 6);
    for (int i = iM; i <= j - 1; i++) {
      const double a = lu_at(C, j, i);
      if (a != 0.0) {
        b[2 * i] -= a * b[2 * j];
        b[2 * i + 1] -= a * b[2 * j + 1];
      }
    }
  }
}

// positiveSmoothedL1 exactly as traj_optimizer.cpp:783-806 (constants recomputed from pe like the reference does)
__device__ __forceinline__ void strict_l1(double pe, double x, double& f, double& df) {
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

// address of component cc of plane k of the sample (piece, j) of segment s in the pass-major plane storage
__device__ __forceinline__ const double* strict_plane_ptr(const ShapeDev& sh, const double* planes, int s, int piece, int j) {
  const int N = sh.N[s], K = sh.K, Kd = sh.Kd;
  const int td = (Kd + kTile - 1) / kTile, tk = (K + kTile - 1) / kTile;
  const int Kp = (piece == 0 || piece == N - 1) ? Kd : K;
  const int npass_reg = (sh.ntile[s] + 1) >> 1;
  int pass, ln;
  if (j < Kp) {
    int t;
    if (piece == 0) t = j / kTile;
    else if (piece == N - 1) t = td + (N - 2) * tk + sh.tpad[s] + j / kTile;
    else t = td + (piece - 1) * tk + j / kTile;
    pass = t >> 1;
    ln = (t & 1) * 16 + (j % kTile);
  } else {
    pass = npass_reg + piece / 32;
    ln = piece % 32;
  }
  const int nchunk = sh.H / 4;
  return planes + ((size_t)(sh.samp0[s] / 32) + pass) * nchunk * 512 + ln;
}

// addPVAGradCost2CT (traj_optimizer.cpp:422-779), static terms only; sequential, one lane
__device__ __noinline__ void strict_sweep(const ShapeDev& sh, const ParamsDev& P, const double* planes, WarpWS& ws,
                                          int s, double* costs) {
  const int Np = sh.N[s];
  const int singul_ = (int)ws.seg[s].singul;
  const int sidx = singul_ > 0 ? 0 : 1;
  const double max_vel = P.max_vel[sidx], max_cur = P.max_cur[sidx], max_acc = P.max_acc[sidx];
  const double wei_obs = P.wei_obs, wei_feas = P.wei_feas, epis = P.help_eps;
  const double pe = P.smooth_eps, cw = P.cur_weight_scale;
  costs[0] = costs[1] = costs[2] = 0.0;
  const double t1 = ws.seg[s].dT;
  double gdT = ws.seg[s].gdT;
  double beta0[6], beta1[6], beta2[6], beta3[6];
  V2 vle[5];
  for (int k = 0; k < 4; k++) vle[k] = mk(P.le[k][0], P.le[k][1]);
  vle[4] = vle[0];  // quirk Q1
  const int H = sh.H;
  for (int i = 0; i < Np; ++i) {
    const int K = (i == 0 || i == Np - 1) ? sh.Kd : sh.K;
    const double* c = ws.c + 2 * (sh.row0[s] + i * 6);
    double* gd = ws.gd + 2 * (sh.row0[s] + i * 6);
    const double step = t1 / K;
    double s1 = 0.0;
    for (int j = 0; j <= K; ++j) {
      const double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
      beta0[0] = 1.0, beta0[1] = s1, beta0[2] = s2, beta0[3] = s3, beta0[4] = s4, beta0[5] = s5;
      beta1[0] = 0.0, beta1[1] = 1.0, beta1[2] = 2.0 * s1, beta1[3] = 3.0 * s2, beta1[4] = 4.0 * s3, beta1[5] = 5.0 * s4;
      beta2[0] = 0.0, beta2[1] = 0.0, beta2[2] = 2.0, beta2[3] = 6.0 * s1, beta2[4] = 12.0 * s2, beta2[5] = 20.0 * s3;
      beta3[0] = 0.0, beta3[1] = 0.0, beta3[2] = 0.0, beta3[3] = 6.0, beta3[4] = 24.0 * s1, beta3[5] = 60.0 * s2;
      const double alpha = 1.0 / K * j;
      s1 += step;  // quirk Q3
      V2 sigma = mk(0.0, 0.0), dsigma = mk(0.0, 0.0), ddsigma = mk(0.0, 0.0), dddsigma = mk(0.0, 0.0);
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
      const double omg = (j == 0 || j == K) ? 0.5 : 1.0;
      double z_h0 = sqrt(dsigma.x * dsigma.x + dsigma.y * dsigma.y);
      const double z_h1 = ddsigma.x * dsigma.x + ddsigma.y * dsigma.y;
      const double z_h2 = dddsigma.x * dsigma.x + dddsigma.y * dsigma.y;
      const double z_h3 = ddsigma.y * dsigma.x + (-ddsigma.x) * dsigma.y;
      if (z_h0 < P.min_speed || (j == 0 && i == 0) || (i == Np - 1 && j == K)) continue;  // traj_optimizer.cpp:550
      const double vel2_reci = 1.0 / (z_h0 * z_h0);
      const double vel2_reci_e = 1.0 / (z_h0 * z_h0 + epis);
      const double vel3_2_reci_e = vel2_reci_e * sqrt(vel2_reci_e);
      z_h0 = 1.0 / z_h0;
      const double z_h4 = z_h1 * vel2_reci;
      const double violaVel = 1.0 / vel2_reci - max_vel * max_vel;
      const double acc2 = z_h1 * z_h1 * vel2_reci;
      const double cur = z_h3 * vel3_2_reci_e;
      const double violaAcc = acc2 - max_acc * max_acc;
      const double violaCurL = cur - max_cur;
      const double violaCurR = -cur - max_cur;
      const M2 rot = M2{dsigma.x, -dsigma.y, dsigma.y, dsigma.x};
      const double sgd = (double)singul_;
      const M2 ego_R = M2{(sgd * rot.a) * z_h0, (sgd * rot.b) * z_h0, (sgd * rot.c) * z_h0, (sgd * rot.d) * z_h0};
      const M2 temp_a = M2{ddsigma.x, -ddsigma.y, ddsigma.y, ddsigma.x};
      // R_dot = singul * (temp_a * z_h0 - ((rot * vel2_reci) * z_h0) * z_h1)
      const M2 R_dot = M2{sgd * (temp_a.a * z_h0 - ((rot.a * vel2_reci) * z_h0) * z_h1),
                          sgd * (temp_a.b * z_h0 - ((rot.b * vel2_reci) * z_h0) * z_h1),
                          sgd * (temp_a.c * z_h0 - ((rot.c * vel2_reci) * z_h0) * z_h1),
                          sgd * (temp_a.d * z_h0 - ((rot.d * vel2_reci) * z_h0) * z_h1)};
      const double* hp = strict_plane_ptr(sh, planes, s, i, j);
      for (int v = 0; v < 5; v++) {
        const V2 le = vle[v];
        const V2 Rl = mk(ego_R.a * le.x + ego_R.b * le.y, ego_R.c * le.x + ego_R.d * le.y);
        const V2 bpt = mk(sigma.x + Rl.x, sigma.y + Rl.y);
        const M2 tl = M2{le.x, -le.y, le.y, le.x};
        for (int k = 0; k < H; k++) {
          const double* pk = hp + (size_t)(k >> 2) * 512 + (k & 3) * 128;
          const V2 nrm = mk(pk[0], pk[32]);
          const V2 pnt = mk(pk[64], pk[96]);
          const double violaPos = nrm.x * (bpt.x - pnt.x) + nrm.y * (bpt.y - pnt.y);
          if (violaPos > 0) {
            double pena, penaD;
            strict_l1(pe, violaPos, pena, penaD);
            // Mx = (singul * tl) * z_h0 - outer(ego_R * le, dsigma) * vel2_reci
            const M2 Mx = M2{(sgd * tl.a) * z_h0 - (Rl.x * dsigma.x) * vel2_reci, (sgd * tl.b) * z_h0 - (Rl.x * dsigma.y) * vel2_reci,
                             (sgd * tl.c) * z_h0 - (Rl.y * dsigma.x) * vel2_reci, (sgd * tl.d) * z_h0 - (Rl.y * dsigma.y) * vel2_reci};
            const V2 Rdl = mk(R_dot.a * le.x + R_dot.b * le.y, R_dot.c * le.x + R_dot.d * le.y);
            const double gradViolaPt = (alpha * nrm.x) * (dsigma.x + Rdl.x) + (alpha * nrm.y) * (dsigma.y + Rdl.y);
            const double wc = omg * step * wei_obs * penaD;
            for (int r = 0; r < 6; r++) {
              const V2 b1n = mk(beta1[r] * nrm.x, beta1[r] * nrm.y);
              const double pcx = beta0[r] * nrm.x + (b1n.x * Mx.a + b1n.y * Mx.c);
              const double pcy = beta0[r] * nrm.y + (b1n.x * Mx.b + b1n.y * Mx.d);
              gd[2 * r] += wc * pcx;
              gd[2 * r + 1] += wc * pcy;
            }
            gdT += omg * wei_obs * (penaD * gradViolaPt * step + pena / K);
            costs[0] += omg * step * wei_obs * pena;
          }
        }
      }
      if (violaVel > 0.0) {  // traj_optimizer.cpp:642-653
        double pena, penaD;
        strict_l1(pe, violaVel, pena, penaD);
        const double gradViolaVt = 2.0 * alpha * z_h1;
        const double wc = omg * step * wei_feas * penaD;
        for (int r = 0; r < 6; r++) {
          const double b = 2.0 * beta1[r];
          gd[2 * r] += wc * (b * dsigma.x);
          gd[2 * r + 1] += wc * (b * dsigma.y);
        }
        gdT += omg * wei_feas * (penaD * gradViolaVt * step + pena / K);
        costs[2] += omg * step * wei_feas * pena;
      }
      if (violaAcc > 0.0) {  // traj_optimizer.cpp:655-665
        double pena, penaD;
        strict_l1(pe, violaAcc, pena, penaD);
        const V2 w1 = mk(z_h4 * ddsigma.x - (z_h4 * z_h4) * dsigma.x, z_h4 * ddsigma.y - (z_h4 * z_h4) * dsigma.y);
        const V2 w2 = dsigma;
        const double gradViolaAt =
            2.0 * alpha * (z_h4 * ((ddsigma.x * ddsigma.x + ddsigma.y * ddsigma.y) + z_h2) - z_h4 * z_h4 * z_h1);
        const double wc = omg * step * wei_feas * penaD;
        for (int r = 0; r < 6; r++) {
          const double b1 = 2.0 * beta1[r], b2 = 2.0 * beta2[r] * z_h4;
          gd[2 * r] += wc * (b1 * w1.x + b2 * w2.x);
          gd[2 * r + 1] += wc * (b1 * w1.y + b2 * w2.y);
        }
        gdT += omg * wei_feas * (penaD * gradViolaAt * step + pena / K);
        costs[2] += omg * step * wei_feas * pena;
      }
      // curvature (traj_optimizer.cpp:684-705)
      const V2 aB = mk(ddsigma.y, -ddsigma.x);
      const V2 Bv = mk(-dsigma.y, dsigma.x);
      const double z_d3 = dddsigma.y * dsigma.x + (-dddsigma.x) * dsigma.y;
      if (violaCurL > 0.0) {
        double pena, penaD;
        strict_l1(pe, violaCurL, pena, penaD);
        const double k3 = 3 * vel3_2_reci_e * vel2_reci_e * z_h3;
        const V2 w1 = mk(vel3_2_reci_e * aB.x - k3 * dsigma.x, vel3_2_reci_e * aB.y - k3 * dsigma.y);
        const double gradT = alpha * vel3_2_reci_e * (z_d3 - 3 * vel2_reci_e * z_h3 * z_h1);
        const double wc = omg * step * wei_feas * cw * penaD;
        for (int r = 0; r < 6; r++) {
          const double b2 = beta2[r] * vel3_2_reci_e;
          gd[2 * r] += wc * (beta1[r] * w1.x + b2 * Bv.x);
          gd[2 * r + 1] += wc * (beta1[r] * w1.y + b2 * Bv.y);
        }
        gdT += omg * wei_feas * cw * (penaD * gradT * step + pena / K);
        costs[2] += omg * step * wei_feas * cw * pena;
      }
      if (violaCurR > 0.0) {
        double pena, penaD;
        strict_l1(pe, violaCurR, pena, penaD);
        const double k3 = 3 * vel3_2_reci_e * vel2_reci_e * z_h3;
        const V2 w1 = mk(vel3_2_reci_e * aB.x - k3 * dsigma.x, vel3_2_reci_e * aB.y - k3 * dsigma.y);
        const double gradT = -(alpha * vel3_2_reci_e * (z_d3 - 3 * vel2_reci_e * z_h3 * z_h1));
        const double wc = omg * step * wei_feas * cw * penaD;
        for (int r = 0; r < 6; r++) {
          const double b2 = beta2[r] * vel3_2_reci_e;
          gd[2 * r] += wc * (-(beta1[r] * w1.x + b2 * Bv.x));
          gd[2 * r + 1] += wc * (-(beta1[r] * w1.y + b2 * Bv.y));
        }
        gdT += omg * wei_feas * cw * (penaD * gradT * step + pena / K);
        costs[2] += omg * step * wei_feas * cw * pena;
      }
    }
  }
  ws.seg[s].gdT = gdT;
}

// costFunctionCallback (traj_optimizer.cpp:206-350) by lane 0 of the warp; x, g in shared memory.  Leaves c, b~ and the
// segment scalars in the workspace like the production evaluation.  `adj` = scratch array of 2*rows doubles.
__device__ __noinline__ double strict_eval(const ShapeDev& sh, const ParamsDev& P, const double* planes,
                                           const double* __restrict__ lu, WarpWS& ws, double* adj, const double* x,
                                           double* g, int lane) {
  double result = 0.0;
  if (lane == 0) {
    const int S = sh.S;
    const int tv0 = sh.var0[S - 1] + 2 * (sh.N[S - 1] - 1);
    const int gb = tv0 + S, ab = gb + 2 * (S - 1);
    for (int k = 0; k < sh.n; k++) g[k] = 0.0;
    const double nsv = P.non_sinv;
    double total_smcost = 0.0, total_timecost = 0.0, penalty_cost = 0.0;
    for (int s = 0; s < S; s++) {
      const double VT = x[tv0 + s];
      ws.seg[s].T = VT > 0.0 ? ((0.5 * VT + 1.0) * VT + 1.0) + P.mini_T : 1.0 / ((0.5 * VT - 1.0) * VT + 1.0) + P.mini_T;
    }
    for (int s = 0; s < S; s++) ws.seg[s].trajtime = s == 0 ? 0.0 : ws.seg[s - 1].T;  // quirk Q5
    for (int s = 0; s < S; s++) {
      SegScratch& sc = ws.seg[s];
      const int N = sh.N[s], n6 = 6 * N;
      for (int k = 0; k < 6; k++) sc.head[k] = sc.ini[k], sc.tail[k] = sc.fin[k];
      if (s > 0) {  // traj_optimizer.cpp:273-277
        const double theta = x[ab + s - 1];
        sc.head[0] = x[gb + 2 * (s - 1)];
        sc.head[1] = x[gb + 2 * (s - 1) + 1];
        sc.head[2] = -nsv * cos(theta);
        sc.head[3] = -nsv * sin(theta);
      }
      if (s < S - 1) {  // traj_optimizer.cpp:278-282
        const double theta = x[ab + s];
        sc.tail[0] = x[gb + 2 * s];
        sc.tail[1] = x[gb + 2 * s + 1];
        sc.tail[2] = nsv * cos(theta);
        sc.tail[3] = nsv * sin(theta);
      }
      // ---- generate (poly_traj_utils.hpp:953-986)
      sc.dT = sc.T / N;
      double t[6], tInv[6];
      t[0] = 1.0;
      t[1] = sc.dT;
      t[2] = t[1] * t[1];
      t[3] = t[2] * t[1];
      t[4] = t[2] * t[2];
      t[5] = t[4] * t[1];
      for (int k = 0; k < 6; k++) tInv[k] = 1.0 / t[k];
      double* b = ws.bt + 2 * sh.row0[s];
      double* c = ws.c + 2 * sh.row0[s];
      double* gd = ws.gd + 2 * sh.row0[s];
      const double* Pin = x + sh.var0[s];
      for (int k = 0; k < 2 * n6; k++) b[k] = 0.0;
      for (int col = 0; col < 2; col++) {
        b[2 * 0 + col] = sc.head[0 + col];
        b[2 * 1 + col] = sc.head[2 + col] * t[1];
        b[2 * 2 + col] = sc.head[4 + col] * t[2];
        for (int i = 0; i < N - 1; i++) b[2 * (6 * i + 5) + col] = Pin[2 * i + col];
        b[2 * (n6 - 3) + col] = sc.tail[0 + col];
        b[2 * (n6 - 2) + col] = sc.tail[2 + col] * t[1];
        b[2 * (n6 - 1) + col] = sc.tail[4 + col] * t[2];
      }
      strict_band_solve(lu + sh.lu_off[s], b, n6);
      for (int i = 0; i < N; i++)
        for (int k = 0; k < 6; k++)
          for (int col = 0; col < 2; col++) c[2 * (6 * i + k) + col] = b[2 * (6 * i + k) + col] * tInv[k];
      // ---- initSmGradCost (poly_traj_utils.hpp:1012-1035)
      auto rdot = [&](int r1, int r2) { return c[2 * r1] * c[2 * r2] + c[2 * r1 + 1] * c[2 * r2 + 1]; };
      for (int i = 0; i < N; i++) {
        for (int col = 0; col < 2; col++) {
          const double c3 = c[2 * (6 * i + 3) + col], c4 = c[2 * (6 * i + 4) + col], c5 = c[2 * (6 * i + 5) + col];
          gd[2 * (6 * i + 5) + col] = 240.0 * c3 * t[3] + 720.0 * c4 * t[4] + 1440.0 * c5 * t[5];
          gd[2 * (6 * i + 4) + col] = 144.0 * c3 * t[2] + 384.0 * c4 * t[3] + 720.0 * c5 * t[4];
          gd[2 * (6 * i + 3) + col] = 72.0 * c3 * t[1] + 144.0 * c4 * t[2] + 240.0 * c5 * t[3];
          gd[2 * (6 * i + 0) + col] = 0.0;
          gd[2 * (6 * i + 1) + col] = 0.0;
          gd[2 * (6 * i + 2) + col] = 0.0;
        }
      }
      double gdT = 0.0;
      for (int i = 0; i < N; i++) {
        const int r3 = 6 * i + 3, r4 = 6 * i + 4, r5 = 6 * i + 5;
        gdT += 36.0 * rdot(r3, r3) + 288.0 * rdot(r4, r3) * t[1] + 576.0 * rdot(r4, r4) * t[2] +
               720.0 * rdot(r5, r3) * t[2] + 2880.0 * rdot(r5, r4) * t[3] + 3600.0 * rdot(r5, r5) * t[4];
      }
      sc.gdT = gdT;
      // ---- getTrajJerkCost (poly_traj_utils.hpp:998-1009)
      double energy = 0.0;
      for (int i = 0; i < N; i++) {
        const int r3 = 6 * i + 3, r4 = 6 * i + 4, r5 = 6 * i + 5;
        energy += 36.0 * rdot(r3, r3) * t[1] + 144.0 * rdot(r4, r3) * t[2] + 192.0 * rdot(r4, r4) * t[3] +
                  240.0 * rdot(r5, r3) * t[3] + 720.0 * rdot(r5, r4) * t[4] + 720.0 * rdot(r5, r5) * t[5];
      }
      sc.jerk = energy;
      double costs[3];
      strict_sweep(sh, P, planes, ws, s, costs);
      sc.pen[0] = costs[0], sc.pen[1] = costs[1], sc.pen[2] = costs[2];
      total_smcost += energy;
      penalty_cost += (costs[0] + costs[1]) + costs[2];
    }
    for (int s = 0; s < S; s++) {
      SegScratch& sc = ws.seg[s];
      const int N = sh.N[s], n6 = 6 * N;
      double t[6], tInv[6];
      t[0] = 1.0;
      t[1] = sc.dT;
      t[2] = t[1] * t[1];
      t[3] = t[2] * t[1];
      t[4] = t[2] * t[2];
      t[5] = t[4] * t[1];
      for (int k = 0; k < 6; k++) tInv[k] = 1.0 / t[k];
      const double* b = ws.bt + 2 * sh.row0[s];
      const double* gd = ws.gd + 2 * sh.row0[s];
      double* aj = adj + 2 * sh.row0[s];
      // ---- calGrads_PT (poly_traj_utils.hpp:1037-1066)
      for (int i = 0; i < N; i++)
        for (int k = 0; k < 6; k++)
          for (int col = 0; col < 2; col++) aj[2 * (6 * i + k) + col] = gd[2 * (6 * i + k) + col] * tInv[k];
      strict_band_solve_adj(lu + sh.lu_off[s], aj, n6);
      double* gP = g + sh.var0[s];
      for (int i = 0; i < N - 1; i++) {
        gP[2 * i] = aj[2 * (6 * i + 5)];
        gP[2 * i + 1] = aj[2 * (6 * i + 5) + 1];
      }
      double gdHead[6], gdTail[6];
      for (int k = 0; k < 3; k++)
        for (int col = 0; col < 2; col++) {
          gdHead[2 * k + col] = aj[2 * k + col] * t[k];
          gdTail[2 * k + col] = aj[2 * (n6 - 3 + k) + col] * t[k];
        }
      double gdT = sc.gdT;
      gdT += sc.head[2] * aj[2 * 1] + sc.head[3] * aj[2 * 1 + 1];
      gdT += (sc.head[4] * aj[2 * 2] + sc.head[5] * aj[2 * 2 + 1]) * 2.0 * t[1];
      gdT += sc.tail[2] * aj[2 * (n6 - 2)] + sc.tail[3] * aj[2 * (n6 - 2) + 1];
      gdT += (sc.tail[4] * aj[2 * (n6 - 1)] + sc.tail[5] * aj[2 * (n6 - 1) + 1]) * 2.0 * t[1];
      double gdtInv[6];
      gdtInv[0] = 0.0;
      gdtInv[1] = -1.0 * tInv[2];
      gdtInv[2] = -2.0 * tInv[3];
      gdtInv[3] = -3.0 * tInv[4];
      gdtInv[4] = -4.0 * tInv[5];
      gdtInv[5] = -5.0 * tInv[5] * tInv[1];
      for (int i = 0; i < N; i++) {
        double sacc = 0.0;
        for (int k = 0; k < 6; k++) {
          const int r = 6 * i + k;
          const double gdcol = gd[2 * r] * b[2 * r] + gd[2 * r + 1] * b[2 * r + 1];
          sacc += gdtInv[k] * gdcol;
        }
        gdT += sacc;
      }
      sc.gdT = gdT;
      if (P.gear_opt) {  // traj_optimizer.cpp:308-319
        if (s > 0) {
          const double theta = x[ab + s - 1];
          g[gb + 2 * (s - 1)] += gdHead[0];
          g[gb + 2 * (s - 1) + 1] += gdHead[1];
          g[ab + s - 1] += gdHead[2] * (nsv * sin(theta)) + gdHead[3] * (-nsv * cos(theta));
        }
        if (s < S - 1) {
          const double theta = x[ab + s];
          g[gb + 2 * s] += gdTail[0];
          g[gb + 2 * s + 1] += gdTail[1];
          g[ab + s] += gdTail[2] * (-nsv * sin(theta)) + gdTail[3] * (nsv * cos(theta));
        }
      }
      // VirtualTGradCost (traj_optimizer.cpp:405-419)
      const double RT = sc.T, VT = x[tv0 + s], gdRT = gdT / N;
      double gdVT2Rt;
      if (VT > 0) {
        gdVT2Rt = VT + 1.0;
      } else {
        const double denSqrt = (0.5 * VT - 1.0) * VT + 1.0;
        gdVT2Rt = (1.0 - VT) / (denSqrt * denSqrt);
      }
      g[tv0 + s] = (gdRT + P.wei_time) * gdVT2Rt;
      total_timecost += RT * P.wei_time;
    }
    result = total_smcost + total_timecost + penalty_cost;
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, result, 0);
}

}  // namespace dftpav
