// dftpav_kernels.cu — sm_100a kernels of the batched MINCO / L-BFGS trajectory optimizer.
//
// Reference path restated here (file:line of ZJU-FAST-Lab/Dftpav):
//   costFunctionCallback            src/Plan/traj_planner/src/traj_optimizer.cpp:206-350      -> eval_pre() / sweep_ranks() / eval_post()
//   addPVAGradCost2CT               traj_optimizer.cpp:422-779                                -> sweep_segment()
//   dynamicObsGradCostP/log_sum_exp traj_optimizer.cpp:1311-1707                              -> dyn_pair() / apply_pair() / lse_n()
//   MinJerkOpt::generate/calGrads   include/plan_utils/poly_traj_utils.hpp:953-1066           -> eval_pre() / eval_post(), dense_forward() / dense_adjoint()
//   BandedSystem::solve/solveAdj    poly_traj_utils.hpp:805-852                               -> banded_solve() / banded_solve_adj() (large shapes)
//   lbfgs_optimize / lewisoverton   include/geo_utils2d/lbfgs.hpp:276-390, 440-751            -> solve_kernel() state machine, two_loop()
//   plane normalisation + copy      traj_optimizer.cpp:15, 49-52                              -> layout_problem()
//   Trajectory evaluators           poly_traj_utils.hpp:378-406, 554-645                      -> sample_kernel() ...
//   getRectangleConst               src/traj_manager.cpp:1211-1469, src/map_adapter.cpp:117-129 -> corridor_kernel()
//   ConverSurroundTrajFromPoints    src/traj_manager.cpp:743-789 (MinJerkOpt::generate)        -> minco_fit_kernel()
// All arithmetic is IEEE fp64; no fast-math.  One warp per trajectory; nothing here touches tensor cores
// (there is no dense contraction on this path).
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>

#include "dftpav_device.cuh"
#include "../../include/dftpav_b200.h"

namespace dftpav {

#define FULL 0xffffffffu

// warps per CTA (one CTA per SM)
#ifndef DFTPAV_MAX_WARPS
#define DFTPAV_MAX_WARPS 12  // measured on B200, batch 4096: 8 warps 153 k, 10: 153 k, 12: 161 k, 14: 138 k, 16: 139 k solves/s (168 vs 128 registers: no spills)
#endif

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// four sums at once: the butterflies are interleaved (the latency of one)
__device__ __noinline__ void warp_sum4(double (&v)[4]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] += __shfl_xor_sync(FULL, v[q], o);
  }
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __noinline__ double warp_dot(const double* a, const double* b, int n, int lane) {
#if DFTPAV_STRICT
  // the reference's sequential sum (Eigen dot of the element-by-element layer the oracle is pinned to)
  double s = 0.0;
  if (lane == 0)
    for (int i = 0; i < n; i++) s += a[i] * b[i];
  return __shfl_sync(FULL, s, 0);
#else
  double s = 0.0;
#pragma unroll 1
  for (int i = lane; i < n; i += 32) s += a[i] * b[i];
  return warp_sum(s);
#endif
}
__device__ __noinline__ double warp_maxabs(const double* a, int n, int lane) {
  double m = 0.0;
#pragma unroll 1
  for (int i = lane; i < n; i += 32) m = fmax(m, fabs(a[i]));
  return warp_max(m);
}

// positiveSmoothedL1 (traj_optimizer.cpp:783-806); the five constants are precomputed on the host (ParamsDev)
__device__ __forceinline__ void smoothed_l1(const ParamsDev& P, double x, double& f, double& df) {
  if (x < P.smooth_eps) {
    f = (P.l1_f4c * x + P.l1_f3c) * x * x * x;
    df = (P.l1_d3c * x + P.l1_d2c) * x * x;
  } else {
    f = x - P.l1_half;
    df = 1.0;
  }
}

struct V2 {
  double x, y;
};
struct M2 {
  double a, b, c, d;  // row-major [[a,b],[c,d]]
};
__device__ __forceinline__ V2 mk(double x, double y) { return V2{x, y}; }
__device__ __forceinline__ V2 operator+(V2 a, V2 b) { return V2{a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ V2 operator-(V2 a, V2 b) { return V2{a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ V2 operator-(V2 a) { return V2{-a.x, -a.y}; }
__device__ __forceinline__ V2 operator*(double s, V2 a) { return V2{s * a.x, s * a.y}; }
__device__ __forceinline__ V2 operator/(V2 a, double s) { return V2{a.x / s, a.y / s}; }
__device__ __forceinline__ double dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ double norm(V2 a) { return sqrt(a.x * a.x + a.y * a.y); }
__device__ __forceinline__ M2 operator*(double s, M2 m) { return M2{s * m.a, s * m.b, s * m.c, s * m.d}; }
__device__ __forceinline__ M2 operator-(M2 p, M2 q) { return M2{p.a - q.a, p.b - q.b, p.c - q.c, p.d - q.d}; }
__device__ __forceinline__ M2 operator*(M2 p, M2 q) {
  return M2{p.a * q.a + p.b * q.c, p.a * q.b + p.b * q.d, p.c * q.a + p.d * q.c, p.c * q.b + p.d * q.d};
}
__device__ __forceinline__ V2 operator*(M2 m, V2 v) { return V2{m.a * v.x + m.b * v.y, m.c * v.x + m.d * v.y}; }
__device__ __forceinline__ M2 transpose(M2 m) { return M2{m.a, m.c, m.b, m.d}; }
__device__ __forceinline__ V2 rowmul(V2 r, M2 m) { return V2{r.x * m.a + r.y * m.c, r.x * m.b + r.y * m.d}; }
__device__ __forceinline__ M2 outer(V2 u, V2 v) { return M2{u.x * v.x, u.x * v.y, u.y * v.x, u.y * v.y}; }
__device__ __forceinline__ M2 l_Bl(V2 l) { return M2{l.x, -l.y, l.y, l.x}; }
// B_h * M for B_h = [[0,-1],[1,0]]
__device__ __forceinline__ M2 Bh_mul(M2 m) { return M2{-m.c, -m.d, m.a, m.b}; }
// M * B_h
__device__ __forceinline__ M2 mul_Bh(M2 m) { return M2{m.b, -m.a, m.d, -m.c}; }
// M * B_h^T
__device__ __forceinline__ M2 mul_BhT(M2 m) { return M2{-m.b, m.a, -m.d, m.c}; }

// ------------------------------------------------------------------------------------------------
// per-warp workspace in shared memory
// ------------------------------------------------------------------------------------------------
struct SegScratch {   // 38 doubles per segment
  double ini[6], fin[6];  // clamped boundary states of the problem, staged once per problem (stage_problem)
  double singul;          // +1 forward / -1 reverse, staged once per problem
  double pad1;
  double head[6];
  double tail[6];
  double T, dT;
  double gdT;         // dJ/d(dT) accumulated (MinJerkOpt::gdT)
  double jerk;        // smoothness cost of the segment
  double pen[3];      // corridor / dynamic / feasibility penalty of the segment
  double trajtime;    // quirk Q5: T of the previous segment (0 for the first)
  double pad[4];
};

// ------------------------------------------------------------------------------------------------
// Drain-phase helpers.  The batch time at the reference batch sizes is set by the last problems leaving the queue
// (profiles/README.md): once the work queue is empty, a warp without a problem becomes a HELPER of a warp of its CTA
// that is still solving (its OWNER) and takes a share of the regular sweep passes of each of the owner's evaluations.
//
// Determinism without partial buffers.  With the reference resolutions (K <= 16 < Kd <= 32) every regular pass adds into
// the gdC rows of pieces that no other regular pass touches (one 16-sample tile per inner piece; the two tiles of the
// first / last piece share one pass, see ShapeDev::tpad), so owner and helpers accumulate straight into the owner's gdC in
// any interleaving; the piece-end samples (j = K_i, "end passes") are added by the owner after the helpers have
// finished.  Every gdC entry therefore receives its terms in ONE fixed order: jerk gradient, tile sums of the piece
// in tile order, end sample.  The scalar sums (gdT and the three penalty sums) are reduced per VIRTUAL RANK r = pass % kVR
// (lane accumulators over the passes of the rank in pass order, one butterfly per rank) and merged by the owner in rank
// order, whoever computed them.  A solve is bit-identical with 0 ... kVR-1 helpers (tests/test_gpu_parity.py).
// Shapes with other resolutions (ShapeDev::helpable == 0) run without helpers.
//
// Protocol (shared-memory flags, no named barriers: a CTA has 16 of them, not enough for 12 owners x teams).  Owner w
// publishes every evaluation as ctl[w].go = ++seq after writing ctl[w].h = number of helpers taking part (its snapshot
// of ctl[w].nhelp) and the problem; helper i (attach index, from atomicAdd on nhelp) serves evaluation seq iff h > i,
// takes virtual ranks chunk(i + 1 of h + 1), and reports with atomicAdd on ctl[w].done.  h = -1 retires the owner and
// releases its helpers, which then look for another owner of the CTA or exit.
// ------------------------------------------------------------------------------------------------
#ifndef DFTPAV_VRANKS
#define DFTPAV_VRANKS 4  // measured at config 3, batch 4096 (k solves/s): 2: 182, 3: 150, 4: 188-193, 6: 192, 8: 180
#endif
#ifndef DFTPAV_STRICT
#define DFTPAV_STRICT 0
#endif
constexpr bool kStrict = DFTPAV_STRICT != 0;          // strict-order verification build (dftpav_strict.cuh)
constexpr int kVR = kStrict ? 1 : DFTPAV_VRANKS;      // virtual ranks of the sweep = max team size (owner + kVR-1 helpers)
constexpr int kMaxHelp = kVR - 1;
static_assert(kVR >= 1 && kVR <= 15, "virtual ranks");

struct TeamCtl {
  int go;       // sequence number of the owner's current evaluation
  int h;        // helpers taking part in evaluation `go`; -1: the owner has retired
  int nhelp;    // helpers attached to this owner
  int done;     // cumulative number of helper completions
  int alive;    // 1 until the warp finds the work queue empty
  int prob;     // problem of evaluation `go`
  unsigned long long planes;  // plane pointer of that problem
};

// where a sweep adds its results (the owner's workspace; generic pointers: a helper writes into its owner's arrays)
struct SweepOut {
  double* gd;     // [2*rows] gdC of the owner
  double* rscal;  // [S][kVR+1][4] per-segment, per-virtual-rank (gdT, pen0, pen1, pen2); slot kVR = the end passes
};

struct WarpWS {
  double* c;      // [2*rows]  MINCO coefficients, row-major pairs
  double* bt;     // [2*rows]  solved rhs before 1/dT^k scaling (b of MinJerkOpt after A.solve)
  double* gd;     // [2*rows]  gdC, then the adjoint solution in place
  double* x;      // [npad] current point
  double* xp;     // [npad]
  double* g;      // [npad]
  double* gp;     // [npad]
  double* d;      // [npad]
  double* stage;  // [2][12][kStageLd]; doubles as the L-BFGS alpha array between evaluations
  double* pf;     // [8] past cost values (lbfgs.hpp:508, 658)
  double* obs;    // [kObsCap][4] obstacle metadata of the current problem: start, duration, uniform piece duration, (P<<32|piece_ptr)
  SegScratch* seg;  // [S]
  double* rscal;  // [S][kVR+1][4] scalar sums of the sweep per virtual rank
  double* adj;    // [2*rows] strict build only: the adjoint right-hand side (the reference keeps gdC beside it)
};

// workspace size in doubles; obstacle metadata only when the batch has obstacles
__host__ __device__ inline int ws_doubles_for(int rows, int npad, int S, int has_obs) {
  return 6 * rows + 5 * npad + 2 * 12 * kStageLd + 8 + (has_obs ? 4 * kObsCap : 0) +
         S * (int)(sizeof(SegScratch) / sizeof(double)) + S * (kVR + 1) * 4 + (kStrict ? 2 * rows : 0);
}

__device__ __forceinline__ WarpWS carve_ws(double* base, const ShapeDev& sh) {
  WarpWS w;
  w.c = base;
  w.bt = w.c + 2 * sh.rows;
  w.gd = w.bt + 2 * sh.rows;
  w.x = w.gd + 2 * sh.rows;
  w.xp = w.x + sh.npad;
  w.g = w.xp + sh.npad;
  w.gp = w.g + sh.npad;
  w.d = w.gp + sh.npad;
  w.stage = w.d + sh.npad;
  w.pf = w.stage + 2 * 12 * kStageLd;
  w.obs = w.pf + 8;
  double* q = w.obs + (sh.has_obs ? 4 * kObsCap : 0);
  w.seg = reinterpret_cast<SegScratch*>(q);
  w.rscal = q + sh.S * (int)(sizeof(SegScratch) / sizeof(double));
  w.adj = w.rscal + sh.S * (kVR + 1) * 4;
  return w;
}

// virtual ranks [r0, r1) of member `i` (0 = the owner) of a team of t: contiguous chunks, the larger ones last (the
// owner also has the serial parts of the evaluation to do)
__device__ __forceinline__ void rank_chunk(int i, int t, int& r0, int& r1) {
  const int base = kVR / t, rem = kVR % t;
  const int big0 = t - rem;  // members big0 .. t-1 get base + 1 ranks
  r0 = i * base + max(0, i - big0);
  r1 = r0 + base + (i >= big0 ? 1 : 0);
}

// Per-problem constants that every evaluation needs (boundary states, direction of each gear segment): read from global
// memory ONCE when the warp takes the problem instead of at every evaluation (the dependent seg_base -> singul load
// alone stalled the start of every sweep for a DRAM round trip: 6 % of all stall samples, profiles/README.md).
__device__ __forceinline__ void stage_problem(const ShapeDev& sh, const BatchDev& bd, int prob, WarpWS& ws, int lane) {
  if (lane < sh.S) {
    const int g = bd.seg_base[prob] + lane;
    SegScratch& sc = ws.seg[lane];
#pragma unroll
    for (int k = 0; k < 6; k++) {
      sc.ini[k] = bd.ini[6 * g + k];
      sc.fin[k] = bd.fin[6 * g + k];
    }
    sc.singul = (double)bd.singul[g];
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// banded substitutions (poly_traj_utils.hpp:805-852) on the packed LU table
//   packed column j: C[j][0..5] = U(j-6..j-1, j), C[j][6] = U(j,j), C[j][7..12] = L(j+1..j+6, j), C[j][13] = 1/U(j,j)
// Each calling lane owns one right-hand-side column `col` of the row-major n x 2 array b.
// The 1/U(j,j) product replaces the reference's division (<= 1 ulp per row).
// ------------------------------------------------------------------------------------------------
// Shared-memory load the compiler may schedule freely (a pure function of its address for the compiler: no memory
// clobber, not volatile).  ONLY for data that is not written between the producing barrier and the consuming loop
// iteration: the read-only LU table, and right-hand-side rows that the running substitution pass has not reached
// yet.  In the banded passes a plain C++ load sits behind the previous row's store in program order (the compiler
// cannot prove the two shared arrays distinct after inlining) and the in-order issue then exposes the full LDS
// latency on every row of the dependent chain.
__device__ __forceinline__ double lds_pure(unsigned addr) {
  double v;
  asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// Opaque zero created by a volatile asm with a memory clobber: the compiler keeps it behind every earlier store and
// barrier.  Adding it to the base address of lds_pure loads pins those loads behind that point as well.
__device__ __forceinline__ unsigned order_token() {
  unsigned t;
  asm volatile("mov.u32 %0, 0;" : "=r"(t)::"memory");
  return t;
}

// ------------------------------------------------------------------------------------------------
// Banded substitutions (poly_traj_utils.hpp:805-852) — pattern-specialised, software-pipelined.
//
// All four passes are in update ("column") form, i.e. exactly the reference loops: when row j is final its contribution is
// pushed into the pending rows with independent multiply-adds, so the dependent chain per row is one FMA (+ one multiply
// by the stored 1/U(j,j)); every row receives its terms in the reference order.
// The LU factors of the MINCO matrix have a sparsity pattern of period 6 (one piece): 20 of 36 sub-diagonal and 11 of
// 36 super-diagonal entries per piece are structurally non-zero.  The reference skips exact zeros at run time
// (`if (A(i,j) != 0.0)`); here the pattern is a compile-time mask per position in the piece, so the zero terms cost
// nothing — same arithmetic, fewer instructions.  The masks are supersets of the true pattern for every piece except the
// last one in the two L passes (its columns hold the tail conditions), which runs with the full mask; the host verifies
// the masks against the factorised table at upload (ShapeDev::band_masks_ok) and the kernel falls back to full masks
// otherwise.  The loops are software-pipelined by hand: table entries and the right-hand-side row of step j+1 are
// loaded at the top of step j (non-aliasing `lds_pure`), otherwise the in-order issue exposes the shared-memory latency
// on the dependent chain of every row.
// ------------------------------------------------------------------------------------------------
enum { PASS_FWD = 0, PASS_BWD = 1, PASS_ADJ_U = 2, PASS_ADJ_L = 3 };
// bit k (1..6) set <=> the k-th neighbour entry is structurally non-zero; index = (row or column) mod 6
__device__ constexpr int kBandMask[4][6] = {
    {0x60, 0x70, 0x78, 0x3c, 0x1e, 0x0e},  // FWD  : L(j+k, j)   by column type
    {0x00, 0x00, 0x00, 0x56, 0x6e, 0x06},  // BWD  : U(j-k, j)   by column type
    {0x00, 0x0c, 0x06, 0x46, 0x42, 0x30},  // ADJ_U: U(j, j+k)   by row type
    {0x7e, 0x7c, 0x78, 0x00, 0x00, 0x3e},  // ADJ_L: L(j, j-k)   by row type
};
constexpr unsigned kColBytes = kLuLd * 8u;

template <int PASS>
__device__ __forceinline__ unsigned band_coef_addr(unsigned cj, int k) {
  if (PASS == PASS_FWD) return cj + (6 + k) * 8u;
  if (PASS == PASS_BWD) return cj + (6 - k) * 8u;
  if (PASS == PASS_ADJ_U) return cj + k * kColBytes + (6 - k) * 8u;
  return cj - k * kColBytes + (6 + k) * 8u;
}
template <int PASS, int MASK>
__device__ __forceinline__ void band_load(double (&c)[7], unsigned cj) {
#pragma unroll
  for (int k = 1; k <= 6; k++)
    if ((MASK >> k) & 1) c[k] = lds_pure(band_coef_addr<PASS>(cj, k));
  if (PASS == PASS_BWD || PASS == PASS_ADJ_U) c[0] = lds_pure(cj + 13 * 8u);  // 1/U(j,j)
}
// one row: finalise row j from the window, push its contribution into the pending rows, shift the window
template <int PASS, int MASK>
__device__ __forceinline__ void band_step(double (&w)[7], const double (&c)[7], unsigned b_addr, double next_in) {
  const double out = (PASS == PASS_BWD || PASS == PASS_ADJ_U) ? w[0] * c[0] : w[0];
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(b_addr), "d"(out) : "memory");
#pragma unroll
  for (int k = 1; k <= 6; k++) w[k - 1] = ((MASK >> k) & 1) ? w[k] - c[k] * out : w[k];
  w[6] = next_in;
}

// one row of a piece: T = position inside the piece in processing order
template <int PASS, bool ALLK, int T>
__device__ __forceinline__ void band_row(double (&w)[7], double (&c)[7], unsigned Ca, unsigned ba, int j0, int n) {
  constexpr bool ASC = (PASS == PASS_FWD || PASS == PASS_ADJ_U);
  constexpr int type = ASC ? T : 5 - T;
  constexpr int mask = ALLK ? 0x7e : kBandMask[PASS][type];
  constexpr int ntype = ASC ? (T + 1) % 6 : (10 - T) % 6;
  // the next row may belong to a piece that runs with the full mask: prefetch everything at piece borders
  constexpr int nmask = (T == 5 || ALLK) ? 0x7e : kBandMask[PASS][ntype];
  const int j = ASC ? j0 + T : j0 - T;
  const int jn = ASC ? j + 1 : j - 1;
  const int jin = ASC ? j + 7 : j - 7;
  double cn[7];
  band_load<PASS, nmask>(cn, Ca + jn * kColBytes);
  // rows outside [0, n) belong to the neighbouring arrays of the warp workspace (c | bt | gd | x ...): readable, and
  // they only ever feed accumulators of rows that do not exist
  const double nin = lds_pure(ba + 16u * jin);
  band_step<PASS, mask>(w, c, ba + 16u * j, nin);
#pragma unroll
  for (int k = 0; k <= 6; k++) c[k] = cn[k];
}

// one piece = 6 rows
template <int PASS, bool ALLK>
__device__ __forceinline__ void band_piece(double (&w)[7], double (&c)[7], unsigned Ca, unsigned ba, int j0, int n) {
  band_row<PASS, ALLK, 0>(w, c, Ca, ba, j0, n);
  band_row<PASS, ALLK, 1>(w, c, Ca, ba, j0, n);
  band_row<PASS, ALLK, 2>(w, c, Ca, ba, j0, n);
  band_row<PASS, ALLK, 3>(w, c, Ca, ba, j0, n);
  band_row<PASS, ALLK, 4>(w, c, Ca, ba, j0, n);
  band_row<PASS, ALLK, 5>(w, c, Ca, ba, j0, n);
}

// a full pass over the n = 6*N rows of one right-hand-side column (`col`) of the row-major n x 2 array b
template <int PASS>
__device__ __forceinline__ void band_pass(const double* C, double* b, int n, int col, bool masks_ok) {
  constexpr bool ASC = (PASS == PASS_FWD || PASS == PASS_ADJ_U);
  constexpr bool LPASS = (PASS == PASS_FWD || PASS == PASS_ADJ_L);
  const unsigned tok = order_token();  // keeps every lds_pure below behind the stores / barriers above
  const unsigned Ca = smem_addr(C) + tok;
  const unsigned ba = smem_addr(b) + 8u * col + tok;
  const int NP = n / 6;
  double w[7], c[7];
  const int jstart = ASC ? 0 : n - 1;
#pragma unroll
  for (int k = 0; k <= 6; k++) {
    const int r = ASC ? k : n - 1 - k;
    w[k] = (r >= 0 && r < n) ? lds_pure(ba + 16u * r) : 0.0;
  }
  band_load<PASS, 0x7e>(c, Ca + jstart * kColBytes);
  if (!masks_ok) {
    for (int p = 0; p < NP; p++) band_piece<PASS, true>(w, c, Ca, ba, ASC ? 6 * p : n - 1 - 6 * p, n);
    return;
  }
  if (ASC) {
    const int nspec = LPASS ? NP - 1 : NP;  // the last piece of an L pass runs with the full mask
    for (int p = 0; p < nspec; p++) band_piece<PASS, false>(w, c, Ca, ba, 6 * p, n);
    if (LPASS) band_piece<PASS, true>(w, c, Ca, ba, 6 * (NP - 1), n);
  } else {
    int p = 0;
    if (LPASS) {
      band_piece<PASS, true>(w, c, Ca, ba, n - 1, n);
      p = 1;
    }
    for (; p < NP; p++) band_piece<PASS, false>(w, c, Ca, ba, n - 1 - 6 * p, n);
  }
}

__device__ __forceinline__ void banded_solve(const double* C, double* b, int n, int col, bool masks_ok) {
  band_pass<PASS_FWD>(C, b, n, col, masks_ok);  // L y = b      (poly_traj_utils.hpp:808-815)
  band_pass<PASS_BWD>(C, b, n, col, masks_ok);  // U x = y      (poly_traj_utils.hpp:816-824)
}
__device__ __forceinline__ void banded_solve_adj(const double* C, double* b, int n, int col, bool masks_ok) {
  band_pass<PASS_ADJ_U>(C, b, n, col, masks_ok);  // U^T z = b  (poly_traj_utils.hpp:834-842)
  band_pass<PASS_ADJ_L>(C, b, n, col, masks_ok);  // L^T w = z  (poly_traj_utils.hpp:843-850)
}

// ------------------------------------------------------------------------------------------------
// obstacle trajectory evaluation (Piece/Trajectory, poly_traj_utils.hpp:77-112,179-211,510-603)
// ------------------------------------------------------------------------------------------------
struct ObsView {
  double uni;  // common piece duration or NaN
  int P;
  const double* dur;
  const double* cf;  // [P][12]
  double start_time, duration;
};
// Trajectory::locatePieceIdx (poly_traj_utils.hpp:510-528): sequential subtraction of the piece durations.  When all
// durations of the obstacle are bit-identical (`uni` is that value, else NaN) the loop needs no loads and is still exact.
__device__ __forceinline__ int locate_piece(double uni, const double* __restrict__ dur, int P, double& t) {
  int idx = 0;
  if (uni == uni) {
    // Power-of-two piece duration (the 1 s prediction step of the reference's obstacle trajectories): every subtraction
    // t -= uni of the reference's loop is EXACT (the result lies on t's own floating-point grid and is smaller), so the loop
    // has the closed form idx = ceil(t / uni) - 1, t -= idx * uni, bit for bit - without a dependent chain of 20-30
    // subtractions per (sample, obstacle).
    const long long ub = __double_as_longlong(uni);
    if ((ub & 0x000fffffffffffffLL) == 0 && uni > 0.0 && t > uni && t < 4.0e15 * uni) {
      const double k = fmin(ceil(t / uni) - 1.0, (double)P);  // steps the loop takes (it stops at idx == P)
      idx = (int)k;
      t -= k * uni;
      if (idx == P) {
        idx--;
        t += uni;
      }
      return idx;
    }
    while (idx < P && t > uni) {
      t -= uni;
      idx++;
    }
    if (idx == P) {
      idx--;
      t += uni;
    }
  } else {
    double d = 0.0;
    for (idx = 0; idx < P && t > (d = dur[idx]); idx++) t -= d;
    if (idx == P) {
      idx--;
      t += dur[idx];
    }
  }
  return idx;
}
__device__ __forceinline__ int obs_locate(const ObsView& o, double& t) { return locate_piece(o.uni, o.dur, o.P, t); }
__device__ __forceinline__ V2 piece_pos(const double* cm, double t) {
  V2 pos = mk(0.0, 0.0);
  double tn = 1.0;
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    pos = pos + tn * mk(cm[2 * i], cm[2 * i + 1]);
    tn *= t;
  }
  return pos;
}
__device__ __forceinline__ V2 piece_vel(const double* cm, double t) {
  V2 ds = mk(0.0, 0.0);
  double tn = 1.0;
  int n = 1;
#pragma unroll
  for (int i = 4; i >= 0; i--) {
    ds = ds + (n * tn) * mk(cm[2 * i], cm[2 * i + 1]);
    tn *= t;
    n++;
  }
  return ds;
}
__device__ __forceinline__ V2 piece_acc(const double* cm, double t) {
  V2 dd = mk(0.0, 0.0);
  double tn = 1.0;
  int m = 1, n = 2;
#pragma unroll
  for (int i = 3; i >= 0; i--) {
    dd = dd + ((m * n) * tn) * mk(cm[2 * i], cm[2 * i + 1]);
    tn *= t;
    m++;
    n++;
  }
  return dd;
}

// log_sum_exp (traj_optimizer.cpp:1686-1707); v is overwritten with the exponentials (quirk Q7).
// OUT OF LINE with rolled loops, and so is exp(): the moving-obstacle term calls this 17 times per colliding pair and the
// inlined form (48 copies of the ~50-instruction exp sequence) made dyn_obstacles ~100 KB of SASS - every call streamed
// through the 32 KB instruction cache (config 5: `no_instruction` was 60-90 % of all stall samples, profiles/README.md).
__device__ __noinline__ double exp_ool(double x) { return exp(x); }
__device__ __noinline__ double lse_n(double alpha, double* v, int nv, double& exp_sum) {
  double d0 = v[0];
  if (alpha > 0) {
#pragma unroll 1
    for (int j = 1; j < nv; j++) d0 = fmax(d0, v[j]);
  } else {
#pragma unroll 1
    for (int j = 1; j < nv; j++) d0 = fmin(d0, v[j]);
  }
  exp_sum = 0.0;
#pragma unroll 1
  for (int j = 0; j < nv; j++) {
    v[j] = exp_ool(alpha * (v[j] - d0));
    exp_sum += v[j];
  }
  return log(exp_sum) / alpha + d0;
}
template <int NV>
__device__ __forceinline__ double lse(double alpha, double* v, double& exp_sum) {
  return lse_n(alpha, v, NV, exp_sum);
}

// cold path (help_eps != 0, the header default of traj_optimizer.h:120): kept out of line to keep the sweep loop small.
// Returns by value: reference parameters would force the caller's four scalars into local memory on the HOT path too.
struct SpeedHelpers {
  double z_h0, vel2_reci, vel2_reci_e, vel3_2_reci_e;
};
__device__ __noinline__ SpeedHelpers speed_helpers_eps(double v2, double eps) {
  SpeedHelpers r;
  r.vel2_reci = 1.0 / v2;
  r.vel2_reci_e = 1.0 / (v2 + eps);
  r.vel3_2_reci_e = r.vel2_reci_e * sqrt(r.vel2_reci_e);
  r.z_h0 = 1.0 / sqrt(v2);
  return r;
}

constexpr int kPlaneChunk = 4;
constexpr int kObsRound = 16;  // obstacles per compaction round of the moving-obstacle term (pair list: 32 x 16 entries)
// Plane storage of one problem is pass-major: [pass][chunk of 4 planes][16 components][32 lanes], so the 16 values a
// lane needs for one chunk sit at compile-time offsets k*32 doubles from one base pointer (immediate-offset LDG.64,
// each warp-wide load is one contiguous 256 B row).  The host pads the plane count to a multiple of 4 with null planes.
// -DDFTPAV_PLANES_STREAM=1: planes are read with the streaming (evict-first) cache operator so that the once-per-
// evaluation plane stream does not push the L-BFGS history of the resident problems out of L2.
// Planes are read with the streaming (evict-first) cache operator: the once-per-evaluation plane stream must not push
// the L-BFGS history of the resident problems out of L2 (measured 197-199 k vs 192-193 k solves/s).  Not ld.global.nc:
// the tiled copy is written by this kernel (layout_problem).
#define DFTPAV_PLANE_LD(p) __ldcs(p)
__device__ __forceinline__ void load_plane_chunk(double (&pn)[kPlaneChunk][4], const double* __restrict__ lanep) {
#pragma unroll
  for (int hh = 0; hh < kPlaneChunk; hh++) {
#pragma unroll
    for (int cc = 0; cc < 4; cc++) pn[hh][cc] = DFTPAV_PLANE_LD(lanep + (hh * 4 + cc) * 32);
  }
}

// ------------------------------------------------------------------------------------------------
// Plane layout on first use: raw (n, p) planes [sample][h][4] of ONE problem -> its pass-major SoA
// [pass][chunk][16 components][32 lanes] with unit normals (replaces the deep copy + colwise().normalize() of
// traj_optimizer.cpp:15, 49-52).  Slot u = pass*32 + lane holds the sample that lane handles in that pass of the sweep
// (two 16-sample tiles per regular pass, then the piece-end samples).  Done by the warp that takes the problem, so the
// host->device copy of the planes (the dominant bytes of an upload) can still be running while the first problems are
// being solved: the warp first waits for the arrival flag of the problem's upload chunk.
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ void layout_problem(const ShapeDev& sh, const BatchDev& bd, int prob, double* __restrict__ dst,
                                            int lane) {
  if (lane == 0) {
    const volatile int* rd = bd.ready + prob / bd.chunk;
    while (*rd != bd.epoch) __nanosleep(500);
    __threadfence();
  }
  __syncwarp();
  if (*reinterpret_cast<const volatile int*>(bd.laid + prob) == bd.epoch) return;
  const int nchunk = sh.H / 4;
  const int K = sh.K, Kd = sh.Kd;
  const int td = (Kd + kTile - 1) / kTile, tk = (K + kTile - 1) / kTile;
  long long pid0 = bd.samp_base[prob];
  for (int s = 0; s < sh.S; s++) {
    const int N = sh.N[s];
    const int ntile = sh.ntile[s];
    const int tlast0 = td + (N - 2) * tk + sh.tpad[s];
    const int npass_reg = (ntile + 1) >> 1, npass = npass_reg + ((N + 31) >> 5);
    for (int pass = 0; pass < npass; pass++) {
      int piece = -1, j = 0;
      if (pass < npass_reg) {
        const int t = 2 * pass + (lane >> 4), r = lane & 15;
        if (t < ntile) {
          int j0;
          if (t < td) {
            piece = 0;
            j0 = kTile * t;
          } else if (t < td + (N - 2) * tk) {
            const int uu = t - td;
            piece = 1 + uu / tk;
            j0 = kTile * (uu % tk);
          } else {
            const int uu = t - tlast0;  // -1: the padding tile
            piece = N - 1;
            j0 = uu < 0 ? Kd : kTile * uu;
          }
          j = j0 + r;
          const int Kp = (piece == 0 || piece == N - 1) ? Kd : K;
          if (j >= Kp) piece = -1;
        }
      } else {
        piece = (pass - npass_reg) * 32 + lane;
        if (piece >= N) piece = -1;
        else j = (piece == 0 || piece == N - 1) ? Kd : K;
      }
      int Hs = 0;
      const double* src = nullptr;
      if (piece >= 0) {
        // pointid inside the segment (traj_optimizer.cpp:514): pieces before contribute K_i + 1 samples each
        long long pid = pid0 + j;
        if (piece > 0) pid += (Kd + 1) + (long long)(piece - 1) * (K + 1);
        if (bd.plane_ptr) {
          Hs = (int)(bd.plane_ptr[pid + 1] - bd.plane_ptr[pid]);
          src = bd.planes_raw + 4 * bd.plane_ptr[pid];
        } else {
          Hs = bd.uniform_H;
          src = bd.planes_raw + 4 * pid * bd.uniform_H;
        }
      }
      double* d0 = dst + ((size_t)(sh.samp0[s] / 32) + pass) * nchunk * 512 + lane;
      for (int h = 0; h < sh.H; h++) {
        double nx = 0.0, ny = 0.0, px = 0.0, py = 0.0;
        if (h < Hs) {
          const double2 a = *reinterpret_cast<const double2*>(src + 4 * h);
          const double2 b = *reinterpret_cast<const double2*>(src + 4 * h + 2);
          nx = a.x, ny = a.y, px = b.x, py = b.y;
          const double z = nx * nx + ny * ny;
          if (z > 0.0) {
            const double nz = sqrt(z);
            nx /= nz;
            ny /= nz;
          }
        }
        double* d4 = d0 + (size_t)(h >> 2) * 512 + (h & 3) * 128;  // chunk h/4, plane h%4 -> components 4*(h%4)..+3
        d4[0] = nx;
        d4[32] = ny;
        d4[64] = px;
        d4[96] = py;
      }
    }
    pid0 += (long long)(N - 2) * (K + 1) + 2 * (Kd + 1);
  }
  __syncwarp();
  if (lane == 0) {
    __threadfence();
    *reinterpret_cast<volatile int*>(bd.laid + prob) = bd.epoch;
  }
  __syncwarp();
}

// Stages the obstacle metadata of a problem into the warp workspace (called once per problem).
__device__ __forceinline__ void stage_obstacles(const BatchDev& bd, int prob, double* obs_sm, int lane) {
  if (!bd.obs_ptr) return;
  const int o0 = bd.obs_ptr[prob], o1 = bd.obs_ptr[prob + 1];
  const int cnt = min(o1 - o0, kObsCap);
  if (lane < cnt) {
    const int u = bd.obs_idx[o0 + lane];
    const int p0 = bd.obs_piece_ptr[u], p1 = bd.obs_piece_ptr[u + 1];
    obs_sm[4 * lane + 0] = bd.obs_start[u];
    obs_sm[4 * lane + 1] = bd.obs_dur[u];
    obs_sm[4 * lane + 2] = bd.obs_uniform[u];
    obs_sm[4 * lane + 3] = __longlong_as_double(((long long)(p1 - p0) << 32) | (long long)(unsigned)p0);
  }
  __syncwarp();
}

// Accumulators of one sample: G0 = d/d sigma, G1 = d/d sigma', G2 = d/d sigma'' (already weighted), plus the
// sample's contributions to gdT and to the three penalty sums.
struct SampleAcc {
  V2 G0, G1, G2;
  double gdT;
  double pen0, pen1, pen2;
};

// dynamicObsGradCostP (traj_optimizer.cpp:1311-1684) for ONE (sample, obstacle) pair; `t` = time of the sample inside the
// segment.  Returns false where the reference takes `continue` (distance cull, no penetration); otherwise the ingredients of
// the pair's contributions, which the lane that owns the sample adds to its accumulators (apply_pair) - the pairs of a
// sweep pass are compacted and evaluated one pair per lane (sweep_segment, stage D), so this 30 KB function runs with
// (nearly) full warps instead of once per obstacle with the few lanes whose sample is near that obstacle.
struct PairOut {
  double pena, penaD, tbar, that;
  V2 gs, gd;  // dG/dsigma, dG/dsigma'
};
__device__ __noinline__ bool dyn_pair(const ParamsDev& P, const BatchDev& bd, int prob, double t, V2 sigma, V2 dsigma,
                                      V2 ddsigma, M2 ego_R, int singul_, double trajtime, int oi, PairOut& res) {
  const double alpha = P.lse_alpha, d_min = P.d_min;
  const double temp0 = norm(dsigma);
  const double temp0_reci = (temp0 != 0.0) ? 1.0 / temp0 : 0.0;
  const double temp3 = temp0_reci * temp0_reci;
  const double t_now = bd.t_now[prob];
  V2 vle[5];
#pragma unroll
  for (int k = 0; k < 4; k++) vle[k] = mk(P.le[k][0], P.le[k][1]);
  vle[4] = vle[0];

  {
    const int u = bd.obs_idx[oi];
    ObsView ob;
    ob.P = bd.obs_piece_ptr[u + 1] - bd.obs_piece_ptr[u];
    ob.dur = bd.obs_piece_dur + bd.obs_piece_ptr[u];
    ob.cf = bd.obs_coeffs + 12L * bd.obs_piece_ptr[u];
    ob.start_time = bd.obs_start[u];
    ob.duration = bd.obs_dur[u];
    ob.uni = bd.obs_uniform[u];
    const double offsettime = t_now - ob.start_time + trajtime;
    const double pt_time = offsettime + t;
    V2 sp, sv, sa;
    if (pt_time < ob.duration) {
      double tt = pt_time;
      const int i = obs_locate(ob, tt);
      sp = piece_pos(ob.cf + 12 * i, tt);
      sv = piece_vel(ob.cf + 12 * i, tt);
      sa = piece_acc(ob.cf + 12 * i, tt);
    } else {
      double tt = ob.duration;
      const int i = obs_locate(ob, tt);
      sa = piece_acc(ob.cf + 12 * i, tt);
      const double exceed = pt_time - ob.duration;
      const V2 vd = piece_vel(ob.cf + 12 * i, tt);
      sv = vd + exceed * sa;
      V2 h = 0.5 * sa;
      h = mk(h.x * exceed, h.y * exceed);
      h = mk(h.x * exceed, h.y * exceed);
      sp = piece_pos(ob.cf + 12 * i, tt) + exceed * vd + h;
    }
    if (norm(sp - sigma) > P.cull_dist) return false;  // traj_optimizer.cpp:1393

    // getR / getRdot at pt_time (quirk Q6: evaluated on the clamped last piece past the duration)
    double tr = pt_time;
    const int ir = obs_locate(ob, tr);
    const V2 rv = piece_vel(ob.cf + 12 * ir, tr);
    const V2 ra = piece_acc(ob.cf + 12 * ir, tr);
    const double rn = norm(rv);
    M2 sur_R = M2{rv.x / rn, -rv.y / rn, rv.y / rn, rv.x / rn};
    // ---- phase B: the eight smooth edge distances and their log-sum-exp (traj_optimizer.cpp:1417-1503).  Nothing
    //      but d_test[8] is kept: most pairs that survive the distance cull are not in collision and leave here.
    const double dnU_inv[4] = {1.0 / norm(vle[1] - vle[0]), 1.0 / norm(vle[2] - vle[1]), 1.0 / norm(vle[3] - vle[2]),
                               1.0 / norm(vle[4] - vle[3])};
    // ---- exp-free rejection.  LSE_{+a}(x) >= max(x) and LSE_{-a}(v) >= min(v) - log(4)/a, hence
    //      costp = d_min - LSE_{+a}(d) <= d_min + log(4)/a - max_k(min_j v_kj + d~_k).  If that bound is negative the
    //      reference takes `continue` at traj_optimizer.cpp:1505 as well; the margin covers rounding.  Most pairs that
    //      pass the 1.5-car-length distance cull are still metres apart and leave here without a single exp/log.
    {
      double best = -1.0e300;
#pragma unroll 1
      for (int e = 0; e < 4; e++) {
        const V2 H_tilde = dnU_inv[e] * (Bh_mul(ego_R) * (vle[e + 1] - vle[e]));
        const V2 hr = rowmul(H_tilde, sur_R);
        double vmin = dot(hr, vle[0]);
#pragma unroll 1
        for (int o = 1; o < 4; o++) vmin = fmin(vmin, dot(hr, vle[o]));
        best = fmax(best, vmin + dot(H_tilde, (sp - sigma) - ego_R * vle[e]));
      }
#pragma unroll 1
      for (int o = 0; o < 4; o++) {
        const V2 H_tilde = dnU_inv[o] * (Bh_mul(sur_R) * (vle[o + 1] - vle[o]));
        const V2 hr = rowmul(H_tilde, ego_R);
        double vmin = dot(hr, vle[0]);
#pragma unroll 1
        for (int e = 1; e < 4; e++) vmin = fmin(vmin, dot(hr, vle[e]));
        best = fmax(best, vmin + dot(H_tilde, (sigma - sp) - sur_R * vle[o]));
      }
      if (d_min + 1.3862943611198906 / alpha - best < -1.0e-9) return false;
    }
    double d_test[8];
#pragma unroll 1
    for (int e = 0; e < 4; e++) {
      const V2 H_tilde = dnU_inv[e] * (Bh_mul(ego_R) * (vle[e + 1] - vle[e]));
      const double d_U_e_tilde = dot(H_tilde, (sp - sigma) - ego_R * vle[e]);
      const V2 hr = rowmul(H_tilde, sur_R);
      double v4[4];
#pragma unroll 1
      for (int o = 0; o < 4; o++) v4[o] = dot(hr, vle[o]);
      double es;
      d_test[e] = lse<4>(-alpha, v4, es) + d_U_e_tilde;
    }
#pragma unroll 1
    for (int o = 0; o < 4; o++) {
      const V2 H_tilde = dnU_inv[o] * (Bh_mul(sur_R) * (vle[o + 1] - vle[o]));
      const double d_E_o_tilde = dot(H_tilde, (sigma - sp) - sur_R * vle[o]);
      const V2 hr = rowmul(H_tilde, ego_R);
      double v4[4];
#pragma unroll 1
      for (int e = 0; e < 4; e++) v4[e] = dot(hr, vle[e]);
      double es;
      d_test[4 + o] = lse<4>(-alpha, v4, es) + d_E_o_tilde;
    }
    double exp_sum_d;
    const double costp = d_min - lse<8>(alpha, d_test, exp_sum_d);  // d_test now holds the exponentials (quirk Q7)
    if (costp <= 0.0) return false;
    double pena, penaD;
    smoothed_l1(P, costp, pena, penaD);

    // ---- phase C: gradients (traj_optimizer.cpp:1510-1646); per-edge quantities are recomputed instead of stored
    V2 pG_sigma = mk(0.0, 0.0), pG_dsigma = mk(0.0, 0.0);
    double pG_that = 0.0;
    // getRdot at pt_time (quirk Q6), traj_optimizer.cpp:1599
    M2 Rud;
    {
      const double p3 = rn * rn * rn;  // pow(|v|,3)
      const double va = dot(rv, ra);
      const M2 ta = M2{ra.x / rn, -ra.y / rn, ra.y / rn, ra.x / rn};
      const M2 tv = M2{rv.x / p3 * va, -rv.y / p3 * va, rv.y / p3 * va, rv.x / p3 * va};
      Rud = ta - tv;
    }
#pragma unroll 1
    for (int e = 0; e < 4; e++) {
      const V2 le = vle[e];
      const V2 delta_le = vle[e + 1] - vle[e];
      const double dnrm = norm(delta_le);
      const V2 H_tilde = (1.0 / dnrm) * (Bh_mul(ego_R) * delta_le);
      const V2 hr = rowmul(H_tilde, sur_R);
      double ex[4], es;
#pragma unroll 1
      for (int o = 0; o < 4; o++) ex[o] = dot(hr, vle[o]);
      (void)lse<4>(-alpha, ex, es);  // ex <- exponentials, es <- their sum
      const double we = d_test[e] / exp_sum_d;
      pG_sigma = pG_sigma - we * (-H_tilde);
      const M2 F_dle = (temp0_reci * ((double)singul_ * transpose(l_Bl(delta_le)))) - (temp3 * outer(dsigma, ego_R * delta_le));
      const M2 F_le = (temp0_reci * ((double)singul_ * transpose(l_Bl(le)))) - (temp3 * outer(dsigma, ego_R * le));
      const V2 a1 = mul_Bh(F_dle) * ((-sp + sigma) + ego_R * le);
      const V2 a2 = (mul_Bh(F_le) * ego_R) * delta_le;
      V2 pdU = (a1 - a2) / dnrm;
      const M2 FdBt = mul_BhT(F_dle);
      double pdUt = dot(H_tilde, sv);
      const V2 nr = rowmul(H_tilde, Rud);
#pragma unroll 1
      for (int o = 0; o < 4; o++) {
        const V2 pd = (FdBt * (sur_R * vle[o])) / dnrm;
        pdU = pdU + (ex[o] / es) * pd;
        pdUt += ex[o] / es * dot(nr, vle[o]);
      }
      pG_dsigma = pG_dsigma - we * pdU;
      pG_that -= we * pdUt;
    }
#pragma unroll 1
    for (int o = 0; o < 4; o++) {
      const V2 lo = vle[o];
      const V2 delta_lo = vle[o + 1] - vle[o];
      const double dnrm = norm(delta_lo);
      const V2 H_tilde = (1.0 / dnrm) * (Bh_mul(sur_R) * delta_lo);
      const V2 hr = rowmul(H_tilde, ego_R);
      double ex[4], es;
#pragma unroll 1
      for (int e = 0; e < 4; e++) ex[e] = dot(hr, vle[e]);
      (void)lse<4>(-alpha, ex, es);
      const double wo = d_test[o + 4] / exp_sum_d;
      pG_sigma = pG_sigma - wo * H_tilde;
      V2 pdE = mk(0.0, 0.0);
      const V2 r1 = (Bh_mul(Rud) * delta_lo) / dnrm;
      const V2 r2 = (Bh_mul(sur_R) * delta_lo) / dnrm;
      double pdEt = dot(r1, (sigma - sp) - sur_R * lo) + dot(r2, -sv - Rud * lo);
#pragma unroll 1
      for (int e = 0; e < 4; e++) {
        const M2 F_le = (temp0_reci * ((double)singul_ * transpose(l_Bl(vle[e])))) - (temp3 * outer(dsigma, ego_R * vle[e]));
        const V2 pd = ((mul_Bh(F_le) * sur_R) * delta_lo) / dnrm;
        pdE = pdE + (ex[e] / es) * pd;
        const V2 row = rowmul(rowmul(ego_R * vle[e], M2{0.0, -1.0, 1.0, 0.0}), Rud);
        pdEt += ex[e] / es * (dot(row, delta_lo) / dnrm);
      }
      pG_dsigma = pG_dsigma - wo * pdE;
      pG_that -= wo * pdEt;
    }
    res.pena = pena, res.penaD = penaD;
    res.tbar = dot(pG_sigma, dsigma) + dot(pG_dsigma, ddsigma);
    res.that = pG_that;
    res.gs = pG_sigma, res.gd = pG_dsigma;
  }
  return true;
}

// the pair's contributions (traj_optimizer.cpp:1508, 1650-1676), added by the lane that owns the sample, in obstacle order
__device__ __forceinline__ void apply_pair(const ParamsDev& P, SampleAcc& acc, const PairOut& r, double omg, double step,
                                           double gama, int pieceid, int trajres, int trajid, int Nseg) {
  const double wei_surround = P.wei_surround;
  acc.pen1 += omg * step * wei_surround * r.pena;
  const double wc = omg * step * wei_surround * r.penaD;
  acc.G0 = acc.G0 + wc * r.gs;
  acc.G1 = acc.G1 + wc * r.gd;
  acc.gdT += omg * wei_surround * (r.pena / trajres + r.penaD * (gama * r.tbar) * step);
  acc.gdT += omg * step * wei_surround * r.that * r.penaD * pieceid;
  acc.gdT += omg * step * wei_surround * gama * r.that * r.penaD;
  for (int idx = 0; idx < trajid; idx++) acc.gdT += omg * step * wei_surround * r.that * r.penaD * Nseg;  // quirk Q5
}

// ------------------------------------------------------------------------------------------------
// constraint-sample sweep of one gear segment (addPVAGradCost2CT, traj_optimizer.cpp:422-779)
// ------------------------------------------------------------------------------------------------
// Regular passes of virtual ranks [r_begin, r_end) (rank-major: all passes of a rank, then the next rank), then - if
// do_end - the end passes (piece-end samples).  Reads the coefficients / segment scalars / obstacle metadata of `ws`
// (the executing warp's own workspace: a helper copies them from its owner first), adds into out.gd and stores the
// scalar sums per virtual rank into out.rscal (both in the OWNER's workspace).
template <bool HAS_OBS>
__device__ __forceinline__ void sweep_segment(const ShapeDev& sh, const ParamsDev& P, const BatchDev& bd, int prob,
                                              const double* __restrict__ planes, WarpWS& ws, const SweepOut& so,
                                              int s, int singul_, int lane, int r_begin, int r_end, bool do_end) {
  const int N = sh.N[s];
  const int K = sh.K, Kd = sh.Kd;
  const int td = (Kd + kTile - 1) / kTile, tk = (K + kTile - 1) / kTile;
  const int ntile = sh.ntile[s];
  const int tlast0 = td + (N - 2) * tk + sh.tpad[s];  // first tile of the last piece
  const int npass_reg = (ntile + 1) >> 1;
  const int npass_end = (N + 31) >> 5;
  const int half = lane >> 4, r = lane & 15;
  const double dT = ws.seg[s].dT;
  const double trajtime = ws.seg[s].trajtime;
  const int sidx = singul_ > 0 ? 0 : 1;
  const double max_vel = P.max_vel[sidx], max_cur = P.max_cur[sidx], max_acc = P.max_acc[sidx];
  const double sg = (double)singul_;
  const double* cseg = ws.c + 2 * sh.row0[s];
  double* gdseg = so.gd + 2 * sh.row0[s];
  double* stage0 = ws.stage;
  double* rsc = so.rscal + (size_t)s * (kVR + 1) * 4;
  // keep these pointers in registers: ptxas otherwise re-derives them from the kernel parameters inside the loop
  asm volatile("" : "+l"(cseg), "+l"(gdseg), "+l"(stage0), "+l"(rsc));
  const int H = sh.H;
  const int nchunk = H / kPlaneChunk;
  // this lane's column of the segment's first pass block
  const double* pl = planes + (size_t)(sh.samp0[s] / 32) * nchunk * 512 + lane;

  double acc_gdT = 0.0, acc_p0 = 0.0, acc_p1 = 0.0, acc_p2 = 0.0;
  const int npass = npass_reg + npass_end;
  const double invK_ = 1.0 / K, invKd_ = 1.0 / Kd;
  double pn[kPlaneChunk][4];
  // ---- pass sequence of this call
  r_end = min(r_end, min(kVR, npass_reg));  // virtual ranks >= npass_reg have no passes
  int vr = r_begin, pass = r_begin;
  bool endpass = false;
  if (vr >= r_end) {
    if (!do_end) return;
    endpass = true;
    vr = kVR;
    pass = npass_reg;
  }
  load_plane_chunk(pn, pl + (size_t)(pass * nchunk) * 512);

  while (true) {
    // successor of this pass: next pass of the rank, first pass of the next rank, first end pass, or nothing
    int pass_next, vr_next = vr;
    bool end_next = endpass, has_next = true;
    if (!endpass) {
      pass_next = pass + kVR;
      if (pass_next >= npass_reg) {
        vr_next = vr + 1;
        pass_next = vr_next;
        if (vr_next >= r_end) {
          end_next = true;
          vr_next = kVR;
          pass_next = npass_reg;
          has_next = do_end;
        }
      }
    } else {
      pass_next = pass + 1;
      has_next = pass_next < npass;
    }
    int piece, j, Kp;
    bool valid;
    if (!endpass) {
      const int t = 2 * pass + half;
      int j0;
      if (t < td) {
        piece = 0;
        j0 = kTile * t;
      } else if (t < td + (N - 2) * tk) {
        const int u = t - td;
        if (tk == 1) {  // K <= 16 (reference config): one tile per inner piece, no division
          piece = 1 + u;
          j0 = 0;
        } else {
          piece = 1 + u / tk;
          j0 = kTile * (u % tk);
        }
      } else {
        const int u = t - tlast0;  // -1: the padding tile (tpad)
        piece = N - 1;
        j0 = kTile * max(u, 0);
        if (u < 0) j0 = Kd;  // invalid
      }
      Kp = (piece == 0 || piece == N - 1) ? Kd : K;
      j = j0 + r;
      valid = (t < ntile) && (j < Kp);
    } else {
      piece = (pass - npass_reg) * 32 + lane;
      valid = piece < N;
      if (!valid) piece = N - 1;
      Kp = (piece == 0 || piece == N - 1) ? Kd : K;
      j = Kp;
    }

    // ---------------- stage A: trajectory derivatives at the sample (traj_optimizer.cpp:501-520)
    const double invKp = (Kp == Kd) ? invKd_ : invK_;  // 1/K: exact for the power-of-two resolutions of the reference config
    const double step = dT * invKp;
    const double s1 = j * step;     // the reference accumulates s1 += step (quirk Q3); differs by O(j) ulp
    const double alpha = invKp * j;
    const double omg = (j == 0 || j == Kp) ? 0.5 : 1.0;
    V2 sigma = mk(0.0, 0.0), dsigma = mk(0.0, 0.0), ddsigma = mk(0.0, 0.0), dddsigma = mk(0.0, 0.0);
    bool active = false;
    if (valid) {
      const double* c = cseg + 12 * piece;
      const double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
      const double b0[6] = {1.0, s1, s2, s3, s4, s5};
      const double b1[6] = {0.0, 1.0, 2.0 * s1, 3.0 * s2, 4.0 * s3, 5.0 * s4};
      const double b2[6] = {0.0, 0.0, 2.0, 6.0 * s1, 12.0 * s2, 20.0 * s3};
      const double b3[6] = {0.0, 0.0, 0.0, 6.0, 24.0 * s1, 60.0 * s2};
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const double cx = c[2 * k], cy = c[2 * k + 1];
        sigma.x += cx * b0[k];
        sigma.y += cy * b0[k];
        dsigma.x += cx * b1[k];
        dsigma.y += cy * b1[k];
        ddsigma.x += cx * b2[k];
        ddsigma.y += cy * b2[k];
        dddsigma.x += cx * b3[k];
        dddsigma.y += cy * b3[k];
      }
      const double v2a = dsigma.x * dsigma.x + dsigma.y * dsigma.y;
      // skip rule of traj_optimizer.cpp:550
      active = !((v2a < P.min_speed * P.min_speed) || (j == 0 && piece == 0) || (piece == N - 1 && j == Kp));
    }

    // ---------------- stage B: kinematic helpers + static corridor
    double z_h0 = 0.0, vel2_reci = 0.0, vel3_2_reci_e = 0.0, vel2_reci_e = 0.0, z_h1 = 0.0, z_h3 = 0.0, v2 = 0.0;
    M2 ego_R = M2{0.0, 0.0, 0.0, 0.0};
    SampleAcc acc;
    acc.G0 = mk(0.0, 0.0);
    acc.G1 = mk(0.0, 0.0);
    acc.G2 = mk(0.0, 0.0);
    acc.gdT = 0.0;
    acc.pen0 = acc.pen1 = acc.pen2 = 0.0;
    if (active) {
      v2 = dsigma.x * dsigma.x + dsigma.y * dsigma.y;
      z_h1 = dot(ddsigma, dsigma);
      z_h3 = ddsigma.y * dsigma.x - ddsigma.x * dsigma.y;
      // |v|, 1/|v|, 1/|v|^2 and (|v|^2+eps)^(-3/2) (traj_optimizer.cpp:527, 556-559).  With eps == 0 (what the plan
      // manager passes, traj_manager.cpp:610) all of them follow from ONE reciprocal square root; a few ulp away
      // from the reference's sqrt + three divisions.
      if (P.help_eps == 0.0) {
        z_h0 = rsqrt(v2);  // 1/|v|
        vel2_reci = z_h0 * z_h0;
        vel2_reci_e = vel2_reci;
        vel3_2_reci_e = vel2_reci * z_h0;
      } else {
        const SpeedHelpers sh_ = speed_helpers_eps(v2, P.help_eps);
        z_h0 = sh_.z_h0;
        vel2_reci = sh_.vel2_reci;
        vel2_reci_e = sh_.vel2_reci_e;
        vel3_2_reci_e = sh_.vel3_2_reci_e;
      }
      const M2 rot = M2{dsigma.x, -dsigma.y, dsigma.y, dsigma.x};
      ego_R = M2{sg * rot.a * z_h0, sg * rot.b * z_h0, sg * rot.c * z_h0, sg * rot.d * z_h0};
      // ---- static corridor (traj_optimizer.cpp:592-622).  Quirk Q1: the vertex list has 5 entries, vertex 0 twice;
      //      the duplicate is folded into a weight of 2 on vertex 0.  Planes are consumed in register chunks of 4;
      //      chunk 0 of this pass was loaded during the previous pass (software prefetch).
      //      Everything after the penalty is LINEAR in the plane normal, so a violated plane only adds
      //      penaD * n and pena to per-vertex sums; the maps to d/dsigma, d/dsigma', d/dt run once per vertex.
      const double wobs = omg * step * P.wei_obs;
      for (int h0 = 0; h0 < H; h0 += kPlaneChunk) {
        if (h0 > 0) load_plane_chunk(pn, pl + (size_t)(pass * nchunk + h0 / kPlaneChunk) * 512);
#pragma unroll 1
        for (int v = 0; v < 4; v++) {  // rolled: measured 161 k solves/s vs 152 k with 2 vertices per trip (19 vs 22 KB loop)
          const V2 le = mk(P.le[v][0], P.le[v][1]);
          const V2 q = ego_R * le;
          const V2 bpt = sigma + q;
          V2 Bn = mk(0.0, 0.0);  // sum penaD * n
          double Sp = 0.0;       // sum pena
          // the four plane tests of a vertex are evaluated together (independent chains) and share one branch: a
          // sample violates a plane only occasionally, and in-order warps pay every branch with exposed latency
          double viol[kPlaneChunk];
#pragma unroll
          for (int hh = 0; hh < kPlaneChunk; hh++)
            viol[hh] = pn[hh][0] * (bpt.x - pn[hh][2]) + pn[hh][1] * (bpt.y - pn[hh][3]);
          if (viol[0] > 0.0 || viol[1] > 0.0 || viol[2] > 0.0 || viol[3] > 0.0) {
#pragma unroll
            for (int hh = 0; hh < kPlaneChunk; hh++) {
              if (viol[hh] > 0.0) {
                double pena, penaD;
                smoothed_l1(P, viol[hh], pena, penaD);
                Bn.x += penaD * pn[hh][0];
                Bn.y += penaD * pn[hh][1];
                Sp += pena;
              }
            }
          }
          if (Sp > 0.0) {
            const double mult = (v == 0) ? 2.0 : 1.0;
            // d(R l)/dt (traj_optimizer.cpp:590), only needed for a violated vertex
            const double f2 = vel2_reci * z_h0 * z_h1;
            const M2 R_dot = M2{sg * (ddsigma.x * z_h0 - rot.a * f2), sg * (-ddsigma.y * z_h0 - rot.b * f2),
                                sg * (ddsigma.y * z_h0 - rot.c * f2), sg * (ddsigma.x * z_h0 - rot.d * f2)};
            const V2 rl = R_dot * le;
            // n^T M,  M = singul*[l,Bl]/|v| - (R l) sigma'^T / |v|^2, applied to Bn
            const double nq = dot(Bn, q);
            const V2 nM = mk(sg * z_h0 * (Bn.x * le.x + Bn.y * le.y) - vel2_reci * nq * dsigma.x,
                             sg * z_h0 * (-Bn.x * le.y + Bn.y * le.x) - vel2_reci * nq * dsigma.y);
            const double gradViolaPt = alpha * dot(Bn, dsigma + rl);
            const double wc = mult * wobs;
            acc.G0 = acc.G0 + wc * Bn;
            acc.G1 = acc.G1 + wc * nM;
            acc.gdT += mult * (omg * P.wei_obs * (gradViolaPt * step + Sp * invKp));
            acc.pen0 += wc * Sp;
          }
        }
      }
    }

    // ---------------- stage C: software prefetch of the next pass's first plane chunk (registers are free now;
    //                  its consumers are most of a pass away)
    if (has_next) load_plane_chunk(pn, pl + (size_t)(pass_next * nchunk) * 512);

    // ---------------- stage D1: moving obstacles (traj_optimizer.cpp:636-638)
    if (HAS_OBS) {
      // (a) light filter, lane <-> sample: obstacle position at the sample time and the distance cull of
      //     traj_optimizer.cpp:1393 (same arithmetic as the reference up to that point) -> bit mask of surviving obstacles;
      // (b) the surviving (sample, obstacle) pairs of the whole pass are compacted into a list (lane-major, obstacle order
      //     inside a lane) and evaluated ONE PAIR PER LANE by dyn_pair, 32 at a time;
      // (c) every lane adds the results of its own pairs to its accumulators in obstacle order - the order, and the
      //     arithmetic, of the reference's loop over the obstacles.
      const int ob0 = bd.obs_ptr[prob], ob1 = bd.obs_ptr[prob + 1];
      const double t_samp = piece * dT + step * j;
      const double t_now = bd.t_now[prob];
#pragma unroll 1
      for (int obase = ob0; obase < ob1; obase += kObsRound) {
        const int nround = min(kObsRound, ob1 - obase);
        unsigned omask = 0;
        if (active) {
#pragma unroll 1
          for (int k = 0; k < nround; k++) {
            const int oi = obase + k;
            double o_start, o_dur, o_uni;
            int oP, op0;
            if (oi - ob0 < kObsCap) {
              const double* m4 = ws.obs + 4 * (oi - ob0);
              o_start = m4[0];
              o_dur = m4[1];
              o_uni = m4[2];
              const long long pk = __double_as_longlong(m4[3]);
              oP = (int)(pk >> 32);
              op0 = (int)(pk & 0xffffffffLL);
            } else {
              const int u = bd.obs_idx[oi];
              o_start = bd.obs_start[u];
              o_dur = bd.obs_dur[u];
              o_uni = bd.obs_uniform[u];
              op0 = bd.obs_piece_ptr[u];
              oP = bd.obs_piece_ptr[u + 1] - op0;
            }
            const double offsettime = t_now - o_start + trajtime;
            const double pt_time = offsettime + t_samp;
            const bool inside = pt_time < o_dur;
            double tt = inside ? pt_time : o_dur;
            // Trajectory::locatePieceIdx (poly_traj_utils.hpp:510-528): sequential subtraction of the piece durations;
            // with bit-identical durations the loop needs no loads and stays exact
            const int idx = locate_piece(o_uni, bd.obs_piece_dur + op0, oP, tt);
            const double* cm = bd.obs_coeffs + 12L * (op0 + idx);
            V2 sp = piece_pos(cm, tt);
            if (!inside) {  // constant-acceleration extrapolation past the prediction (traj_optimizer.cpp:1379-1388)
              const V2 sa = piece_acc(cm, tt);
              const V2 vd = piece_vel(cm, tt);
              const double exceed = pt_time - o_dur;
              V2 h = 0.5 * sa;
              h = mk(h.x * exceed, h.y * exceed);
              h = mk(h.x * exceed, h.y * exceed);
              sp = sp + exceed * vd + h;
            }
            if (!(norm(sp - sigma) > P.cull_dist)) omask |= 1u << k;
          }
        }
        // compaction (all 32 lanes)
        const int cnt = __popc(omask);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += v;
        }
        const int pbase = incl - cnt;
        const int total = __shfl_sync(FULL, incl, 31);
        if (total == 0) continue;
        unsigned short* plist = reinterpret_cast<unsigned short*>(stage0);  // [<= 32 * kObsRound] (lane << 8) | obstacle
        double* pres = stage0 + (32 * kObsRound * 2 + 7) / 8;                // [32][8] results of the current chunk
        {
          unsigned mm = omask;
          int q = pbase;
          while (mm) {
            const int k = __ffs(mm) - 1;
            plist[q++] = (unsigned short)((lane << 8) | k);
            mm &= mm - 1;
          }
        }
        __syncwarp();
#pragma unroll 1
        for (int k0 = 0; k0 < total; k0 += 32) {
          const int pi = k0 + lane;
          const bool have = pi < total;
          const int ent = have ? plist[pi] : 0;
          const int src = ent >> 8;
          // the sample of my pair lives in lane `src`
          const V2 sig_s = mk(__shfl_sync(FULL, sigma.x, src), __shfl_sync(FULL, sigma.y, src));
          const V2 ds_s = mk(__shfl_sync(FULL, dsigma.x, src), __shfl_sync(FULL, dsigma.y, src));
          const V2 dds_s = mk(__shfl_sync(FULL, ddsigma.x, src), __shfl_sync(FULL, ddsigma.y, src));
          const double ra = __shfl_sync(FULL, ego_R.a, src), rc = __shfl_sync(FULL, ego_R.c, src);
          const double ts_s = __shfl_sync(FULL, t_samp, src);
          PairOut r;
          bool hit = false;
          if (have) hit = dyn_pair(P, bd, prob, ts_s, sig_s, ds_s, dds_s, M2{ra, -rc, rc, ra}, singul_, trajtime, obase + (ent & 0xff), r);
          double* mine = pres + 8 * lane;
          mine[0] = hit ? r.pena : -1.0;
          if (hit) mine[1] = r.penaD, mine[2] = r.tbar, mine[3] = r.that, mine[4] = r.gs.x, mine[5] = r.gs.y, mine[6] = r.gd.x, mine[7] = r.gd.y;
          __syncwarp();
          for (int q = max(pbase, k0); q < min(pbase + cnt, k0 + 32); q++) {
            const double* rr = pres + 8 * (q - k0);
            if (rr[0] >= 0.0) {
              PairOut w;
              w.pena = rr[0], w.penaD = rr[1], w.tbar = rr[2], w.that = rr[3];
              w.gs = mk(rr[4], rr[5]), w.gd = mk(rr[6], rr[7]);
              apply_pair(P, acc, w, omg, step, alpha, piece, Kp, s, N);
            }
          }
          __syncwarp();
        }
      }
    }

    // ---------------- stage D2: velocity / acceleration / curvature limits, expansion
    double out[12];
#pragma unroll
    for (int q = 0; q < 12; q++) out[q] = 0.0;
    if (active) {
      const double z_h4 = z_h1 * vel2_reci;
      const double violaVel = v2 - max_vel * max_vel;
      const double violaAcc = z_h1 * z_h1 * vel2_reci - max_acc * max_acc;
      const double cur = z_h3 * vel3_2_reci_e;
      const double violaCurL = cur - max_cur;
      const double violaCurR = -cur - max_cur;
      // ---- velocity (642-653)
      if (violaVel > 0.0) {
        double pena, penaD;
        smoothed_l1(P, violaVel, pena, penaD);
        const double wc = omg * step * P.wei_feas * penaD;
        acc.G1 = acc.G1 + (2.0 * wc) * dsigma;
        acc.gdT += omg * P.wei_feas * (penaD * (2.0 * alpha * z_h1) * step + pena * invKp);
        acc.pen2 += omg * step * P.wei_feas * pena;
      }
      // ---- longitudinal acceleration (655-665)
      if (violaAcc > 0.0) {
        double pena, penaD;
        smoothed_l1(P, violaAcc, pena, penaD);
        const double z_h2 = dot(dddsigma, dsigma);
        const double wc = omg * step * P.wei_feas * penaD;
        const V2 w1 = z_h4 * ddsigma - (z_h4 * z_h4) * dsigma;
        acc.G1 = acc.G1 + (2.0 * wc) * w1;
        acc.G2 = acc.G2 + (2.0 * wc * z_h4) * dsigma;
        const double gradViolaAt = 2.0 * alpha * (z_h4 * (dot(ddsigma, ddsigma) + z_h2) - z_h4 * z_h4 * z_h1);
        acc.gdT += omg * P.wei_feas * (penaD * gradViolaAt * step + pena * invKp);
        acc.pen2 += omg * step * P.wei_feas * pena;
      }
      // ---- curvature, both signs (684-705)
      if (violaCurL > 0.0 || violaCurR > 0.0) {
        const V2 aB = mk(ddsigma.y, -ddsigma.x);
        const V2 Bv = mk(-dsigma.y, dsigma.x);
        const double z_d3 = dddsigma.y * dsigma.x - dddsigma.x * dsigma.y;
        const V2 w1 = vel3_2_reci_e * aB - (3.0 * vel3_2_reci_e * vel2_reci_e * z_h3) * dsigma;
        const double gt = alpha * vel3_2_reci_e * (z_d3 - 3.0 * vel2_reci_e * z_h3 * z_h1);
        const double wf = P.wei_feas * P.cur_weight_scale;
        const bool left = violaCurL > 0.0;  // the two signs are mutually exclusive for max_cur > 0
        double pena, penaD;
        smoothed_l1(P, left ? violaCurL : violaCurR, pena, penaD);
        const double sgn = left ? 1.0 : -1.0;
        const double wc = sgn * (omg * step * wf * penaD);
        acc.G1 = acc.G1 + wc * w1;
        acc.G2 = acc.G2 + (wc * vel3_2_reci_e) * Bv;
        acc.gdT += omg * wf * (penaD * (sgn * gt) * step + pena * invKp);
        acc.pen2 += omg * step * wf * pena;
        if (left && violaCurR > 0.0) {  // only possible for a non-positive curvature limit
          smoothed_l1(P, violaCurR, pena, penaD);
          const double wc2 = omg * step * wf * penaD;
          acc.G1 = acc.G1 - wc2 * w1;
          acc.G2 = acc.G2 - (wc2 * vel3_2_reci_e) * Bv;
          acc.gdT += omg * wf * (penaD * (-gt) * step + pena * invKp);
          acc.pen2 += omg * step * wf * pena;
        }
      }
      acc_gdT += acc.gdT;
      acc_p0 += acc.pen0;
      acc_p1 += acc.pen1;
      acc_p2 += acc.pen2;
      // expansion to the 6x2 coefficient gradient: beta_0 G0^T + beta_1 G1^T + beta_2 G2^T (powers recomputed here
      // instead of being kept live through the whole sample)
      {
        const double s2 = s1 * s1, s3 = s2 * s1, s4 = s2 * s2, s5 = s4 * s1;
        const double b0[6] = {1.0, s1, s2, s3, s4, s5};
        const double b1[6] = {0.0, 1.0, 2.0 * s1, 3.0 * s2, 4.0 * s3, 5.0 * s4};
        const double b2[6] = {0.0, 0.0, 2.0, 6.0 * s1, 12.0 * s2, 20.0 * s3};
#pragma unroll
        for (int k = 0; k < 6; k++) {
          out[2 * k] = b0[k] * acc.G0.x + b1[k] * acc.G1.x + b2[k] * acc.G2.x;
          out[2 * k + 1] = b0[k] * acc.G0.y + b1[k] * acc.G1.y + b2[k] * acc.G2.y;
        }
      }
    }

    // ---------------- stage E: reduction into gdC
    if (endpass) {
      // every lane owns a distinct piece: direct accumulation
      if (valid) {
#pragma unroll
        for (int q = 0; q < 12; q++) gdseg[12 * piece + q] += out[q];
      }
      __syncwarp();
    } else {
      // fixed-order tile reduction: stage[half][q][r], then 12 lanes per half sum their row over r = 0..15
      double* st = stage0 + half * (12 * kStageLd);
#pragma unroll
      for (int q = 0; q < 12; q++) st[q * kStageLd + r] = out[q];
      __syncwarp();
      double sum = 0.0;
      if (r < 12) {  // fixed order: (s0 + ... + s7) + (s8 + ... + s15), two independent chains
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < kTile / 2; i++) {
          s0 += st[r * kStageLd + i];
          s1 += st[r * kStageLd + kTile / 2 + i];
        }
        sum = s0 + s1;
      }
      // the two halves may hit the same piece (Kd > 16): add one after the other
      const int tt = 2 * pass + half;
      const bool tvalid = tt < ntile && !(tt >= td + (N - 2) * tk && tt < tlast0);
      if (half == 0 && r < 12 && tvalid) gdseg[12 * piece + r] += sum;
      __syncwarp();
      if (half == 1 && r < 12 && tvalid) gdseg[12 * piece + r] += sum;
      __syncwarp();
    }
    // ---------------- stage F: leaving a virtual rank (or the end passes): its scalar sums, one interleaved butterfly
    if (!has_next || vr_next != vr) {
      // (inline on purpose: a call inside the sweep loop costs more than the 4 KB of code; measured 172 k vs 188 k solves/s)
      double v4[4] = {acc_gdT, acc_p0, acc_p1, acc_p2};
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (HAS_OBS || q != 2) v4[q] += __shfl_xor_sync(FULL, v4[q], o);
      }
      if (lane < 4) rsc[4 * vr + lane] = lane == 0 ? v4[0] : (lane == 1 ? v4[1] : (lane == 2 ? v4[2] : v4[3]));
      acc_gdT = acc_p0 = acc_p1 = acc_p2 = 0.0;
      __syncwarp();
    }
    if (!has_next) break;
    pass = pass_next;
    vr = vr_next;
    endpass = end_next;
  }
}

// The sweep of all segments: virtual ranks [r_begin, r_end) of the regular passes and, if do_end, the end passes.
template <bool HAS_OBS>
__device__ __forceinline__ void sweep_ranks(const ShapeDev& sh, const ParamsDev& P, const BatchDev& bd, int prob,
                                            const double* __restrict__ planes, WarpWS& ws, const SweepOut& so, int lane,
                                            int r_begin, int r_end, bool do_end) {
  for (int s = 0; s < sh.S; s++)
    sweep_segment<HAS_OBS>(sh, P, bd, prob, planes, ws, so, s, (int)ws.seg[s].singul, lane, r_begin, r_end, do_end);
}

// ------------------------------------------------------------------------------------------------
// one cost + gradient evaluation (costFunctionCallback, traj_optimizer.cpp:206-350)
// x, g: shared-memory vectors of the warp.  Leaves c / bt in the workspace (quirk Q8 relies on that).
// ------------------------------------------------------------------------------------------------
// diagnostic: when non-null, lane 0 of every warp adds the cycles spent in each phase of an evaluation
__device__ unsigned long long* g_phase_clock = nullptr;
// finer split of the L-BFGS part (slots 8..15), only in diagnostic builds (-DDFTPAV_PHASE_FINE=1, tools/lone_phases.py)
#ifndef DFTPAV_PHASE_FINE
#define DFTPAV_PHASE_FINE 0
#endif
#if DFTPAV_PHASE_FINE
#define FINE_T0() long long t_f_ = clock64()
#define FINE_ADD(slot)                                                                                      \
  do {                                                                                                      \
    const long long n_ = clock64();                                                                         \
    if (g_phase_clock && lane == 0) atomicAdd(g_phase_clock + (slot), (unsigned long long)(n_ - t_f_));     \
    t_f_ = n_;                                                                                              \
  } while (0)
#else
#define FINE_T0()
#define FINE_ADD(slot)
#endif
__device__ __noinline__ long long phase_mark(unsigned long long* pc, int idx, long long t_prev, int lane) {
  __syncwarp();
  const long long now = clock64();
  if (lane == 0) atomicAdd(pc + idx, (unsigned long long)(now - t_prev));
  return now;
}
#define PHASE_MARK(idx) \
  if (pc) t_prev = phase_mark(pc, (idx), t_prev, lane);

// ------------------------------------------------------------------------------------------------
// dense MINCO path (ShapeDev::dense_minco): both solves as products with the constant matrix M = A^-1[:, R], R = the
// N+5 rows where the right-hand side is non-zero and the adjoint solution is read (see dense_minco_table() in
// dftpav_api.cu).  Mt[j][row], leading dimension 6N+1.  All 32 lanes work (the banded substitutions keep 2 busy).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int dense_row_of(int j, int N) { return j < 3 ? j : (j < N + 2 ? 6 * (j - 3) + 5 : 5 * N - 5 + j); }

// b~[row] = sum_j Mt[j][row] * bR[j]   (x and y columns together); lanes own rows, 5 row blocks per chunk
__device__ __forceinline__ void dense_forward(const double* __restrict__ Mt, int ld, int ncol, const double* bR,
                                              double* out, int n, int lane) {
  constexpr int RB = 5;
  for (int r0 = 0; r0 < n; r0 += 32 * RB) {
    double ax[RB], ay[RB];
    const double* mp[RB];
#pragma unroll
    for (int k = 0; k < RB; k++) {
      ax[k] = ay[k] = 0.0;
      mp[k] = Mt + min(r0 + lane + 32 * k, n - 1);
    }
#pragma unroll 2
    for (int j = 0; j < ncol; j++) {
      const double2 b = *reinterpret_cast<const double2*>(bR + 2 * j);
#pragma unroll
      for (int k = 0; k < RB; k++) {
        const double m = mp[k][(size_t)j * ld];
        ax[k] = fma(m, b.x, ax[k]);
        ay[k] = fma(m, b.y, ay[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < RB; k++) {
      const int row = r0 + lane + 32 * k;
      if (row < n) *reinterpret_cast<double2*>(out + 2 * row) = make_double2(ax[k], ay[k]);
    }
  }
}

// lambda_R[j] = sum_row Mt[j][row] * gd[row], written over gd at row R[j]; lanes own columns j (ncol <= 64)
__device__ __forceinline__ void dense_adjoint(const double* __restrict__ Mt, int ld, int ncol, double* gd, int n, int N,
                                              int lane) {
  const double* pA = Mt + (size_t)min(lane, ncol - 1) * ld;
  double ax0 = 0.0, ay0 = 0.0, ax1 = 0.0, ay1 = 0.0, bx0 = 0.0, by0 = 0.0, bx1 = 0.0, by1 = 0.0;
  if (ncol <= 32) {
#pragma unroll 2
    for (int r = 0; r < n; r += 2) {  // n = 6N is even
      const double2 g0 = *reinterpret_cast<const double2*>(gd + 2 * r);
      const double2 g1 = *reinterpret_cast<const double2*>(gd + 2 * r + 2);
      const double m0 = pA[r], m1 = pA[r + 1];
      ax0 = fma(m0, g0.x, ax0), ay0 = fma(m0, g0.y, ay0);
      ax1 = fma(m1, g1.x, ax1), ay1 = fma(m1, g1.y, ay1);
    }
  } else {
    const double* pB = Mt + (size_t)min(lane + 32, ncol - 1) * ld;
#pragma unroll 2
    for (int r = 0; r < n; r += 2) {
      const double2 g0 = *reinterpret_cast<const double2*>(gd + 2 * r);
      const double2 g1 = *reinterpret_cast<const double2*>(gd + 2 * r + 2);
      const double m0 = pA[r], m1 = pA[r + 1], q0 = pB[r], q1 = pB[r + 1];
      ax0 = fma(m0, g0.x, ax0), ay0 = fma(m0, g0.y, ay0);
      ax1 = fma(m1, g1.x, ax1), ay1 = fma(m1, g1.y, ay1);
      bx0 = fma(q0, g0.x, bx0), by0 = fma(q0, g0.y, by0);
      bx1 = fma(q1, g1.x, bx1), by1 = fma(q1, g1.y, by1);
    }
  }
  __syncwarp();  // every lane has read gd before the solution overwrites rows of it
  if (lane < ncol) *reinterpret_cast<double2*>(gd + 2 * dense_row_of(lane, N)) = make_double2(ax0 + ax1, ay0 + ay1);
  if (lane + 32 < ncol) *reinterpret_cast<double2*>(gd + 2 * dense_row_of(lane + 32, N)) = make_double2(bx0 + bx1, by0 + by1);
}

// Part 1 (rank 0): decision vector -> boundary states, MINCO coefficients, jerk energy, gdC initialisation.
__device__ __forceinline__ void eval_pre(const ShapeDev& sh, const ParamsDev& P, const double* __restrict__ lu,
                                         WarpWS& ws, const double* x, int lane, long long& t_prev) {
  unsigned long long* pc = g_phase_clock;
  t_prev = pc ? clock64() : 0;
  const int S = sh.S;
  const int tv0 = sh.var0[S - 1] + 2 * (sh.N[S - 1] - 1);  // offset of tau
  const int gb = tv0 + S;                                  // gear positions
  const int ab = gb + 2 * (S - 1);                         // angles

  // ---- per-segment scalars and boundary states (traj_optimizer.cpp:224-234, 269-282)
  if (lane < S) {
    const int s = lane;
    const double VT = x[tv0 + s];
    const double T = VT > 0.0 ? ((0.5 * VT + 1.0) * VT + 1.0) + P.mini_T : 1.0 / ((0.5 * VT - 1.0) * VT + 1.0) + P.mini_T;
    SegScratch& sc = ws.seg[s];
    sc.T = T;
    sc.dT = T / sh.N[s];
#pragma unroll
    for (int k = 0; k < 6; k++) {
      sc.head[k] = sc.ini[k];
      sc.tail[k] = sc.fin[k];
    }
    if (s > 0) {
      const double theta = x[ab + s - 1];
      double sn, cs;
      sincos(theta, &sn, &cs);
      sc.head[0] = x[gb + 2 * (s - 1)];
      sc.head[1] = x[gb + 2 * (s - 1) + 1];
      sc.head[2] = -P.non_sinv * cs;
      sc.head[3] = -P.non_sinv * sn;
    }
    if (s < S - 1) {
      const double theta = x[ab + s];
      double sn, cs;
      sincos(theta, &sn, &cs);
      sc.tail[0] = x[gb + 2 * s];
      sc.tail[1] = x[gb + 2 * s + 1];
      sc.tail[2] = P.non_sinv * cs;
      sc.tail[3] = P.non_sinv * sn;
    }
  }
  const bool dense = sh.dense_minco != 0;
  if (!dense)
    for (int i = lane; i < 2 * sh.rows; i += 32) ws.bt[i] = 0.0;
  __syncwarp();
  if (lane < S) {
    // trajtimes (quirk Q5): T of the previous segment, not cumulative
    ws.seg[lane].trajtime = lane == 0 ? 0.0 : ws.seg[lane - 1].T;
  }
  if (dense) {
    // ---- non-zero rows of the right-hand side, staged in the (still unused) gdC array, then b~ = M b_R
    for (int s = 0; s < S; s++) {
      const int N = sh.N[s];
      const SegScratch& sc = ws.seg[s];
      const double t1 = sc.dT, t2 = t1 * t1;
      double* bR = ws.gd + 2 * sh.row0[s];
      const double* Pin = x + sh.var0[s];
      for (int j = lane; j < N + 5; j += 32) {
        double bx, by;
        if (j >= 3 && j < N + 2) {
          bx = Pin[2 * (j - 3)], by = Pin[2 * (j - 3) + 1];
        } else {
          const double* st = j < 3 ? sc.head : sc.tail;
          const int k = j < 3 ? j : j - (N + 2);
          const double scl = k == 0 ? 1.0 : (k == 1 ? t1 : t2);
          bx = st[2 * k] * scl, by = st[2 * k + 1] * scl;
        }
        *reinterpret_cast<double2*>(bR + 2 * j) = make_double2(bx, by);
      }
    }
    __syncwarp();
    PHASE_MARK(0)
    for (int s = 0; s < S; s++)
      dense_forward(lu + sh.lu_off[s], 6 * sh.N[s] + 1, sh.N[s] + 5, ws.gd + 2 * sh.row0[s], ws.bt + 2 * sh.row0[s],
                    6 * sh.N[s], lane);
    __syncwarp();
    PHASE_MARK(1)
  } else {
  // ---- right-hand side (poly_traj_utils.hpp:968-977)
  for (int s = 0; s < S; s++) {
    const int N = sh.N[s];
    const SegScratch& sc = ws.seg[s];
    const double t1 = sc.dT, t2 = t1 * t1;
    double* b = ws.bt + 2 * sh.row0[s];
    if (lane < 2) {
      b[0 + lane] = sc.head[0 + lane];
      b[2 + lane] = sc.head[2 + lane] * t1;
      b[4 + lane] = sc.head[4 + lane] * t2;
      b[2 * (6 * N - 3) + lane] = sc.tail[0 + lane];
      b[2 * (6 * N - 2) + lane] = sc.tail[2 + lane] * t1;
      b[2 * (6 * N - 1) + lane] = sc.tail[4 + lane] * t2;
    }
    const double* Pin = x + sh.var0[s];
    for (int i = lane; i < 2 * (N - 1); i += 32) b[2 * (6 * (i >> 1) + 5) + (i & 1)] = Pin[i];
  }
  __syncwarp();
  PHASE_MARK(0)
  // ---- banded solve, two lanes per segment
  if (lane < 2 * S) {
    const int s = lane >> 1;
    banded_solve(lu + sh.lu_off[s], ws.bt + 2 * sh.row0[s], 6 * sh.N[s], lane & 1, sh.band_masks_ok != 0);
  }
  __syncwarp();
  PHASE_MARK(1)
  }  // !dense
  // ---- c = b~ / dT^k, jerk energy, gdC / gdT initialisation (poly_traj_utils.hpp:981-984, 998-1035)
  for (int s = 0; s < S; s++) {
    const int N = sh.N[s];
    const double t1 = ws.seg[s].dT, t2 = t1 * t1, t3 = t2 * t1, t4 = t2 * t2, t5 = t4 * t1;
    // 1/dT^k from ONE division (the reference takes t.cwiseInverse(); <= 2 ulp apart)
    const double ti1 = 1.0 / t1, ti2 = ti1 * ti1, ti3 = ti2 * ti1, ti4 = ti2 * ti2, ti5 = ti4 * ti1;
    const double* b = ws.bt + 2 * sh.row0[s];
    double* c = ws.c + 2 * sh.row0[s];
    double* gd = ws.gd + 2 * sh.row0[s];
    double energy = 0.0, gdT = 0.0;
    for (int i = lane; i < N; i += 32) {
      const double* bi = b + 12 * i;
      double* ci = c + 12 * i;
      double* gi = gd + 12 * i;
      const double c0x = bi[0], c0y = bi[1];
      const double c1x = bi[2] * ti1, c1y = bi[3] * ti1;
      const double c2x = bi[4] * ti2, c2y = bi[5] * ti2;
      const double c3x = bi[6] * ti3, c3y = bi[7] * ti3;
      const double c4x = bi[8] * ti4, c4y = bi[9] * ti4;
      const double c5x = bi[10] * ti5, c5y = bi[11] * ti5;
      ci[0] = c0x, ci[1] = c0y, ci[2] = c1x, ci[3] = c1y, ci[4] = c2x, ci[5] = c2y;
      ci[6] = c3x, ci[7] = c3y, ci[8] = c4x, ci[9] = c4y, ci[10] = c5x, ci[11] = c5y;
      const double d33 = c3x * c3x + c3y * c3y, d43 = c4x * c3x + c4y * c3y, d44 = c4x * c4x + c4y * c4y;
      const double d53 = c5x * c3x + c5y * c3y, d54 = c5x * c4x + c5y * c4y, d55 = c5x * c5x + c5y * c5y;
      energy += 36.0 * d33 * t1 + 144.0 * d43 * t2 + 192.0 * d44 * t3 + 240.0 * d53 * t3 + 720.0 * d54 * t4 +
                720.0 * d55 * t5;
      gdT += 36.0 * d33 + 288.0 * d43 * t1 + 576.0 * d44 * t2 + 720.0 * d53 * t2 + 2880.0 * d54 * t3 +
             3600.0 * d55 * t4;
      gi[0] = gi[1] = gi[2] = gi[3] = gi[4] = gi[5] = 0.0;
      gi[6] = 72.0 * c3x * t1 + 144.0 * c4x * t2 + 240.0 * c5x * t3;
      gi[7] = 72.0 * c3y * t1 + 144.0 * c4y * t2 + 240.0 * c5y * t3;
      gi[8] = 144.0 * c3x * t2 + 384.0 * c4x * t3 + 720.0 * c5x * t4;
      gi[9] = 144.0 * c3y * t2 + 384.0 * c4y * t3 + 720.0 * c5y * t4;
      gi[10] = 240.0 * c3x * t3 + 720.0 * c4x * t4 + 1440.0 * c5x * t5;
      gi[11] = 240.0 * c3y * t3 + 720.0 * c4y * t4 + 1440.0 * c5y * t5;
    }
    energy = warp_sum(energy);
    gdT = warp_sum(gdT);
    if (lane == 0) {
      ws.seg[s].jerk = energy;
      ws.seg[s].gdT = gdT;
    }
  }
  __syncwarp();
  PHASE_MARK(2)
}

// Part 2 (rank 0, after the team's sweep): merge of the sweep partials, adjoint MINCO product, gradient assembly, cost.
__device__ __forceinline__ double eval_post(const ShapeDev& sh, const ParamsDev& P, const double* __restrict__ lu,
                                            WarpWS& ws, const double* x, double* g, int lane, long long t_prev) {
  unsigned long long* pc = g_phase_clock;
  const int S = sh.S;
  const int tv0 = sh.var0[S - 1] + 2 * (sh.N[S - 1] - 1);  // offset of tau
  const int gb = tv0 + S;                                  // gear positions
  const int ab = gb + 2 * (S - 1);                         // angles
  const bool dense = sh.dense_minco != 0;
  // fixed-order merge of the sweep's scalar sums: virtual ranks 0 .. kVR-1 in order, then the end passes (whoever
  // computed them: the owner or its drain-phase helpers)
  if (lane < S) {
    const double* rs = ws.rscal + (size_t)lane * (kVR + 1) * 4;
    const int nr = min(kVR, (sh.ntile[lane] + 1) >> 1);
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < nr; r++) {
#pragma unroll
      for (int q = 0; q < 4; q++) a[q] += rs[4 * r + q];
    }
#pragma unroll
    for (int q = 0; q < 4; q++) a[q] += rs[4 * kVR + q];
    ws.seg[lane].gdT += a[0];
    ws.seg[lane].pen[0] = a[1];
    ws.seg[lane].pen[1] = a[2];
    ws.seg[lane].pen[2] = a[3];
  }
  __syncwarp();
  PHASE_MARK(3)

  // ---- MinJerkOpt::calGrads_PT (poly_traj_utils.hpp:1037-1066)
  for (int s = 0; s < S; s++) {
    const int N = sh.N[s];
    const double t1 = ws.seg[s].dT, t2 = t1 * t1, t3 = t2 * t1, t4 = t2 * t2, t5 = t4 * t1;
    const double ti1_ = 1.0 / t1, ti2_ = ti1_ * ti1_;
    double tInv[6] = {1.0, ti1_, ti2_, ti2_ * ti1_, ti2_ * ti2_, ti2_ * ti2_ * ti1_};
    double gdtInv[6] = {0.0, -1.0 * tInv[2], -2.0 * tInv[3], -3.0 * tInv[4], -4.0 * tInv[5], -5.0 * tInv[5] * tInv[1]};
    const double* b = ws.bt + 2 * sh.row0[s];
    double* gd = ws.gd + 2 * sh.row0[s];
    double part = 0.0;
    for (int i = lane; i < N; i += 32) {
      double sacc = 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const int rr = 2 * (6 * i + k);
        sacc += gdtInv[k] * (gd[rr] * b[rr] + gd[rr + 1] * b[rr + 1]);
        gd[rr] *= tInv[k];      // adjScaledGrad, in place
        gd[rr + 1] *= tInv[k];
      }
      part += sacc;
    }
    part = warp_sum(part);
    if (lane == 0) ws.seg[s].pad[0] = part;  // added to gdT after the boundary terms, like the reference
  }
  __syncwarp();
  PHASE_MARK(4)
  if (dense) {
    for (int s = 0; s < S; s++)
      dense_adjoint(lu + sh.lu_off[s], 6 * sh.N[s] + 1, sh.N[s] + 5, ws.gd + 2 * sh.row0[s], 6 * sh.N[s], sh.N[s], lane);
  } else if (lane < 2 * S) {
    const int s = lane >> 1;
    banded_solve_adj(lu + sh.lu_off[s], ws.gd + 2 * sh.row0[s], 6 * sh.N[s], lane & 1, sh.band_masks_ok != 0);
  }
  __syncwarp();
  PHASE_MARK(5)
  // gdP -> g
  for (int s = 0; s < S; s++) {
    const int N = sh.N[s];
    const double* adj = ws.gd + 2 * sh.row0[s];
    double* gP = g + sh.var0[s];
    for (int i = lane; i < 2 * (N - 1); i += 32) gP[i] = adj[2 * (6 * (i >> 1) + 5) + (i & 1)];
  }
  // gdT, time gradient (traj_optimizer.cpp:325, 405-419)
  if (lane < S) {
    const int s = lane;
    const int N = sh.N[s];
    SegScratch& sc = ws.seg[s];
    const double* adj = ws.gd + 2 * sh.row0[s];
    const double t1 = sc.dT;
    double gdT = sc.gdT;
    gdT += sc.head[2] * adj[2] + sc.head[3] * adj[3];
    gdT += (sc.head[4] * adj[4] + sc.head[5] * adj[5]) * 2.0 * t1;
    gdT += sc.tail[2] * adj[2 * (6 * N - 2)] + sc.tail[3] * adj[2 * (6 * N - 2) + 1];
    gdT += (sc.tail[4] * adj[2 * (6 * N - 1)] + sc.tail[5] * adj[2 * (6 * N - 1) + 1]) * 2.0 * t1;
    gdT += sc.pad[0];
    sc.gdT = gdT;
    const double VT = x[tv0 + s];
    double gdVT2Rt;
    if (VT > 0.0) {
      gdVT2Rt = VT + 1.0;
    } else {
      const double den = (0.5 * VT - 1.0) * VT + 1.0;
      gdVT2Rt = (1.0 - VT) / (den * den);
    }
    g[tv0 + s] = (gdT / N + P.wei_time) * gdVT2Rt;
  }
  // gear position / angle gradients (traj_optimizer.cpp:308-319): gear j joins tail of segment j and head of segment j+1
  if (lane < S - 1) {
    const int j = lane;
    double gx = 0.0, gy = 0.0, ga = 0.0;
    if (P.gear_opt) {
      const double theta = x[ab + j];
      double sn, cs;
      sincos(theta, &sn, &cs);
      {  // segment j: gradFin
        const int N = sh.N[j];
        const double* adj = ws.gd + 2 * sh.row0[j];
        const double t1 = ws.seg[j].dT;
        const double f0x = adj[2 * (6 * N - 3)], f0y = adj[2 * (6 * N - 3) + 1];
        const double f1x = adj[2 * (6 * N - 2)] * t1, f1y = adj[2 * (6 * N - 2) + 1] * t1;
        gx += f0x;
        gy += f0y;
        ga += f1x * (-P.non_sinv * sn) + f1y * (P.non_sinv * cs);
      }
      {  // segment j+1: gradIni
        const double* adj = ws.gd + 2 * sh.row0[j + 1];
        const double t1 = ws.seg[j + 1].dT;
        const double h0x = adj[0], h0y = adj[1];
        const double h1x = adj[2] * t1, h1y = adj[3] * t1;
        gx += h0x;
        gy += h0y;
        ga += h1x * (P.non_sinv * sn) + h1y * (-P.non_sinv * cs);
      }
    }
    g[gb + 2 * j] = gx;
    g[gb + 2 * j + 1] = gy;
    g[ab + j] = ga;
  }
  __syncwarp();
  PHASE_MARK(6)
  // ---- total (traj_optimizer.cpp:293-294, 329, 344)
  double sm = 0.0, tc = 0.0, pen = 0.0;
  for (int s = 0; s < S; s++) {
    sm += ws.seg[s].jerk;
    pen += (ws.seg[s].pen[0] + ws.seg[s].pen[1]) + ws.seg[s].pen[2];
    tc += ws.seg[s].T * P.wei_time;
  }
  return sm + tc + pen;
}

// ------------------------------------------------------------------------------------------------
// TMA bulk copy of the LU block into shared memory (one elected thread, mbarrier completion)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_block(double* smem_dst, const double* gmem_src, unsigned bytes,
                                               unsigned long long* mbar) {
  const unsigned mbar_a = (unsigned)__cvta_generic_to_shared(mbar);
  const unsigned dst_a = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
        "l"(gmem_src), "r"(bytes), "r"(mbar_a)
        : "memory");
  }
  // all threads wait for phase 0
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mbar_a)
        : "memory");
  }
}

#if DFTPAV_STRICT
}  // namespace dftpav
#include "dftpav_strict.cuh"
namespace dftpav {
#endif

// ------------------------------------------------------------------------------------------------
// standalone evaluation kernel: f, g at caller-provided x (the parity hook dftpav_eval)
// ------------------------------------------------------------------------------------------------
template <bool HAS_OBS>
__global__ void __launch_bounds__(32 * DFTPAV_MAX_WARPS, 1)
    eval_kernel(const __grid_constant__ ShapeDev sh, const __grid_constant__ ParamsDev P, const BatchDev bd,
                const GroupDev grp, const double* __restrict__ xin, double* __restrict__ fout,
                double* __restrict__ gout, double* __restrict__ coeffs_out, double* __restrict__ dT_out) {
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) unsigned long long mbar;
  tma_load_block(smem, grp.lu, (unsigned)(sh.lu_doubles * sizeof(double)), &mbar);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  WarpWS ws = carve_ws(smem + sh.lu_doubles + warp * sh.ws_doubles, sh);
  const SweepOut so{ws.gd, ws.rscal};
  for (int local = blockIdx.x * nwarp + warp; local < grp.count; local += gridDim.x * nwarp) {
    const int prob = grp.prob_id[local];
    double* planes = grp.planes + (size_t)local * sh.H * 4 * sh.Mpad;
    if (sh.H > 0) layout_problem(sh, bd, prob, planes, lane);  // H == 0: the plane-free ordering pass (planes may still be in flight)
    long long t_prev = 0;
    for (int i = lane; i < sh.n; i += 32) ws.x[i] = xin[(size_t)prob * bd.n_stride + i];
    __syncwarp();
    stage_problem(sh, bd, prob, ws, lane);
    double f;
#if DFTPAV_STRICT
    f = strict_eval(sh, P, planes, smem, ws, ws.adj, ws.x, ws.g, lane);
#else
    if (HAS_OBS) stage_obstacles(bd, prob, ws.obs, lane);
    eval_pre(sh, P, smem, ws, ws.x, lane, t_prev);
    // the whole team's work by this warp: every virtual rank in order, then the end passes (same bits as a helped solve)
    sweep_ranks<HAS_OBS>(sh, P, bd, prob, planes, ws, so, lane, 0, kVR, true);
    f = eval_post(sh, P, smem, ws, ws.x, ws.g, lane, t_prev);
#endif
    __syncwarp();
    if (lane == 0) fout[prob] = f;
    for (int i = lane; i < bd.n_stride; i += 32) gout[(size_t)prob * bd.n_stride + i] = i < sh.n ? ws.g[i] : 0.0;
    if (coeffs_out)
      for (int i = lane; i < 2 * sh.rows; i += 32) coeffs_out[(size_t)prob * bd.row_stride * 2 + i] = ws.c[i];
    if (dT_out && lane < sh.S) dT_out[bd.seg_base[prob] + lane] = ws.seg[lane].dT;
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// L-BFGS two-loop recursion (lbfgs.hpp:716-739)
// ------------------------------------------------------------------------------------------------
// The history keeps 1/ys_j (one division when the pair is stored) and the recursion multiplies by it, instead of the
// reference's division by ys_j in both loops (lbfgs.hpp:722, 733): <= 1.5 ulp per alpha_j / beta_j, and the ~70-cycle
// division leaves the dependent chain of every history step.  The strict build stores ys_j and divides.
#if DFTPAV_STRICT
#define DFTPAV_OVER_YS(num, ysv) ((num) / (ysv))
#define DFTPAV_STORE_YS(ysv) (ysv)
#else
#define DFTPAV_OVER_YS(num, ysv) ((num) * (ysv))
#define DFTPAV_STORE_YS(ysv) (1.0 / (ysv))
#endif

// Register variant: the direction in registers (NPL elements per lane, n <= 32*NPL), the (s_j, y_j, 1/ys_j) stream
// software-prefetched one entry ahead.  Measured on B200 (tools/microbench/two_loop_bench.cu, profiles/README.md): 300-313
// cycles per history step for a lone warp while the history fits the L1, THE SAME with the history in shared memory - the
// step is bound by its dependent chain (2 FMA, a 5-level shuffle butterfly, multiply, FMA).  Inside the solver the L1 is
// ~40 KB (185 KB of shared memory) and a 120-entry history is 90 KB: the step costs 446 cycles (tools/lone_phases.py with
// a -DDFTPAV_PHASE_FINE=1 build: the two-loop is 44 % of a long problem's critical path), i.e. ~130 cycles of the L2 round
// trip are exposed.  Every attempt to hide them lost more than it won (round 2, profiles/README.md):
//   * prefetch distance 2..6 in registers: the noinline callee is held to ~60-80 registers by the caller's live state
//     (whatever the caller re-materialises) and spills the ring - 185 / 161 / 142 / 136 / 114 k solves/s at distance
//     1 / 2 / 3 / 4 / 6 against 198 k; inlined, the kernel's own spills grow from 0.5 to 2.4 KB;
//   * staging ring in shared memory filled by cp.async: 632 cycles per step (~40 instructions on the in-order path), 160 k;
//   * the same ring filled by one TMA bulk copy per entry (history rows interleaved): 691 cycles per step, 172 k;
//   * prefetch.global.L1 hints (CCTL.E.PF1) 2-4 entries ahead: the hint itself stalls the in-order warp - 398 cycles per
//     step with the history IN the L1, 603-612 with the solver's carve-out (441 without the hints); 186-190 k vs 193 k.
// Measured and rejected before: 16 lanes x 3 elements with a 4-level butterfly (297 lone), four entries per step with a
// cached Gram band (219 lone but 144 k vs 188 k solves/s in the solver: the extra loads are exposed), and a full
// Gram-matrix form with the dot products lane <-> entry and the triangular recurrences on shuffle broadcasts (650-980:
// every recurrence step exposes an L2 round trip for its Gram row; 62 k solves/s).  No integer modulo (wrap by compare);
// the trip count is read through a shuffle so that the compiler treats the loop as warp-uniform and emits bare SHFL.
template <int NPL>
__device__ __noinline__ void two_loop(double* d_sm, double* alpha_sm, const double* __restrict__ lm_s,
                                      const double* __restrict__ lm_y, const double* __restrict__ lm_ys, int ld, int m,
                                      int end, int bound_in, int n, int lane, double ys, double yy) {
  const int bound = __shfl_sync(FULL, bound_in, 0);
  double d[NPL], sc_[NPL], yc_[NPL], sn_[NPL], yn_[NPL];
  double ysc, ysn = 0.0;
  bool ok[NPL];
#pragma unroll
  for (int e = 0; e < NPL; e++) {
    const int i = lane + 32 * e;
    ok[e] = i < n;
    d[e] = ok[e] ? d_sm[i] : 0.0;
    sn_[e] = yn_[e] = 0.0;
  }
  // ---- first loop: newest to oldest
  int j = end - 1;
  if (j < 0) j += m;
  {
    const double* sj = lm_s + (size_t)j * ld + lane;
    const double* yj = lm_y + (size_t)j * ld + lane;
#pragma unroll
    for (int e = 0; e < NPL; e++) {
      sc_[e] = ok[e] ? sj[32 * e] : 0.0;
      yc_[e] = ok[e] ? yj[32 * e] : 0.0;
    }
    ysc = lm_ys[j];
  }
#pragma unroll 1
  for (int i = 0; i < bound; ++i) {
    int jn = j - 1;
    if (jn < 0) jn += m;
    if (i + 1 < bound) {
      const double* sj = lm_s + (size_t)jn * ld + lane;
      const double* yj = lm_y + (size_t)jn * ld + lane;
#pragma unroll
      for (int e = 0; e < NPL; e++) {
        sn_[e] = ok[e] ? sj[32 * e] : 0.0;
        yn_[e] = ok[e] ? yj[32 * e] : 0.0;
      }
      ysn = lm_ys[jn];
    }
    double part = 0.0;
#pragma unroll
    for (int e = 0; e < NPL; e++) part += sc_[e] * d[e];
    const double a = DFTPAV_OVER_YS(warp_sum(part), ysc);
    if (lane == 0) alpha_sm[j] = a;
    const double na = -a;
#pragma unroll
    for (int e = 0; e < NPL; e++) {
      d[e] += na * yc_[e];
      sc_[e] = sn_[e];
      yc_[e] = yn_[e];
    }
    ysc = ysn;
    if (i + 1 < bound) j = jn;
  }
  const double scl = ys / yy;
#pragma unroll
  for (int e = 0; e < NPL; e++) d[e] *= scl;
  __syncwarp();
  // ---- second loop: oldest to newest, starting at the entry where the first loop stopped (j)
  {
    const double* sj = lm_s + (size_t)j * ld + lane;
    const double* yj = lm_y + (size_t)j * ld + lane;
#pragma unroll
    for (int e = 0; e < NPL; e++) {
      sc_[e] = ok[e] ? sj[32 * e] : 0.0;
      yc_[e] = ok[e] ? yj[32 * e] : 0.0;
    }
    ysc = lm_ys[j];
  }
#pragma unroll 1
  for (int i = 0; i < bound; ++i) {
    int jn = j + 1;
    if (jn >= m) jn -= m;
    if (i + 1 < bound) {
      const double* sj = lm_s + (size_t)jn * ld + lane;
      const double* yj = lm_y + (size_t)jn * ld + lane;
#pragma unroll
      for (int e = 0; e < NPL; e++) {
        sn_[e] = ok[e] ? sj[32 * e] : 0.0;
        yn_[e] = ok[e] ? yj[32 * e] : 0.0;
      }
      ysn = lm_ys[jn];
    }
    double part = 0.0;
#pragma unroll
    for (int e = 0; e < NPL; e++) part += yc_[e] * d[e];
    const double beta = DFTPAV_OVER_YS(warp_sum(part), ysc);
    const double a = alpha_sm[j] - beta;
#pragma unroll
    for (int e = 0; e < NPL; e++) {
      d[e] += a * sc_[e];
      sc_[e] = sn_[e];
      yc_[e] = yn_[e];
    }
    ysc = ysn;
    j = jn;
  }
#pragma unroll
  for (int e = 0; e < NPL; e++) {
    const int i = lane + 32 * e;
    if (ok[e]) d_sm[i] = d[e];
  }
  __syncwarp();
}

// Generic variant for any n (up to the 517 variables of 4 segments x 64 pieces): the direction stays in shared memory.
// Also the form of the strict build: warp_dot is then the sequential sum of the reference (lbfgs.hpp:719-736).
__device__ __noinline__ void two_loop_generic(double* d_sm, double* alpha_sm, const double* __restrict__ lm_s,
                                              const double* __restrict__ lm_y, const double* __restrict__ lm_ys, int ld,
                                              int m, int end, int bound_in, int n, int lane, double ys, double yy) {
  const int bound = __shfl_sync(FULL, bound_in, 0);
  int j = end;
#pragma unroll 1
  for (int i = 0; i < bound; ++i) {
    j = j == 0 ? m - 1 : j - 1;
    const double* sj = lm_s + (size_t)j * ld;
    const double* yj = lm_y + (size_t)j * ld;
    const double a = DFTPAV_OVER_YS(warp_dot(sj, d_sm, n, lane), lm_ys[j]);
    if (lane == 0) alpha_sm[j] = a;
    const double na = -a;
#pragma unroll 1
    for (int q = lane; q < n; q += 32) d_sm[q] += na * yj[q];
    __syncwarp();
  }
  const double scl = ys / yy;
#pragma unroll 1
  for (int q = lane; q < n; q += 32) d_sm[q] *= scl;
  __syncwarp();
#pragma unroll 1
  for (int i = 0; i < bound; ++i) {
    const double* sj = lm_s + (size_t)j * ld;
    const double* yj = lm_y + (size_t)j * ld;
    const double beta = DFTPAV_OVER_YS(warp_dot(yj, d_sm, n, lane), lm_ys[j]);
    const double a = alpha_sm[j] - beta;
#pragma unroll 1
    for (int q = lane; q < n; q += 32) d_sm[q] += a * sj[q];
    __syncwarp();
    j = j + 1 == m ? 0 : j + 1;
  }
}

__device__ __forceinline__ int wrap_index(int x, int m) {
  x %= m;
  return x < 0 ? x + m : x;
}

// ------------------------------------------------------------------------------------------------
// the solver: persistent warps, each runs lbfgs_optimize (lbfgs.hpp:440-751) for one problem at a time.
// Written as a state machine around ONE call site of eval_cost_grad (keeps the kernel inside the instruction cache):
//   FETCH -> FIRST (evaluation at x0, lbfgs.hpp:524) -> LS (line-search trial, lbfgs.hpp:312-389) -> ... -> FETCH
// All control flow is warp-uniform: every reduction leaves bit-identical values in all lanes.
// ------------------------------------------------------------------------------------------------
enum { PH_FETCH = 0, PH_FIRST = 1, PH_LS = 2 };

// shared-memory flag helpers of the owner / helper protocol (block scope)
__device__ __forceinline__ int ld_flag(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_flag(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }
__device__ __forceinline__ int go_word(int seq, int h) { return (seq << 4) | (h + 1); }  // h = -1: retired

// Work fetch of the arrival-aware longest-first queue (GroupDev::clist): lane c looks at upload chunk c - is it on the
// device, what is the key of its best remaining problem - the warp takes the largest key (NaN first, ties to the lower
// chunk) with one atomicAdd on that chunk's head; a lost race (the head moved past the end) retries.  Returns the local
// index, or -1 when every chunk is exhausted.  While only chunks that are still in flight have problems left, the warp
// sleeps and looks again.  Scheduling only: no effect on any result.
__device__ __noinline__ int fetch_by_chunk(const int* __restrict__ clist, const int* __restrict__ cbeg, int* chead,
                                           const double* __restrict__ key, const int* __restrict__ prob_id, int n_chunks,
                                           const int* ready, int epoch, int lane) {
  while (true) {
    double k = 0.0;
    int state = 0;  // 0 exhausted, 1 available, 2 not yet arrived
    if (lane < n_chunks) {
      const int beg = cbeg[lane], endc = cbeg[lane + 1];
      const int hd = *reinterpret_cast<const volatile int*>(chead + lane);
      if (beg + hd < endc) {
        if (*reinterpret_cast<const volatile int*>(ready + lane) == epoch) {
          state = 1;
          k = key[prob_id[clist[beg + hd]]];
          if (k != k) k = __longlong_as_double(0x7ff0000000000000LL);  // NaN first
        } else {
          state = 2;
        }
      }
    }
    const unsigned avail = __ballot_sync(FULL, state == 1);
    if (avail == 0u) {
      if (__ballot_sync(FULL, state == 2) == 0u) return -1;
      __nanosleep(500);
      continue;
    }
    int best = state == 1 ? lane : 64;
    if (state != 1) k = -__longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ko = __shfl_xor_sync(FULL, k, o);
      const int bo = __shfl_xor_sync(FULL, best, o);
      if (ko > k || (ko == k && bo < best)) k = ko, best = bo;
    }
    int got = -1;
    if (lane == 0) {
      const int beg = cbeg[best], endc = cbeg[best + 1];
      const int old = atomicAdd(chead + best, 1);
      if (beg + old < endc) got = clist[beg + old];
    }
    got = __shfl_sync(FULL, got, 0);
    if (got >= 0) return got;
  }
}

template <bool HAS_OBS>
__global__ void __launch_bounds__(32 * DFTPAV_MAX_WARPS, 1)
    solve_kernel(const __grid_constant__ ShapeDev sh, const __grid_constant__ ParamsDev P, const BatchDev bd,
                 const GroupDev grp, const OutDev out, const HistDev hist, int* __restrict__ work_counter) {
  extern __shared__ __align__(128) double smem[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ TeamCtl ctl[DFTPAV_MAX_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (threadIdx.x < DFTPAV_MAX_WARPS) {
    TeamCtl& c0 = ctl[threadIdx.x];
    c0.go = go_word(0, 0), c0.h = 0, c0.nhelp = 0, c0.done = 0, c0.alive = threadIdx.x < nwarp ? 1 : 0, c0.prob = -1, c0.planes = 0ull;
  }
  tma_load_block(smem, grp.lu, (unsigned)(sh.lu_doubles * sizeof(double)), &mbar);  // (contains a __syncthreads)
  WarpWS ws = carve_ws(smem + sh.lu_doubles + warp * sh.ws_doubles, sh);
  const int slot = blockIdx.x * nwarp + warp;
  const int n = sh.n;
  const int m = P.mem_size;
  double* lm_s = hist.s + (size_t)slot * hist.mem * hist.ld;
  double* lm_y = hist.y + (size_t)slot * hist.mem * hist.ld;
  double* lm_ys = hist.ys + (size_t)slot * hist.mem;
  double* alpha_sm = ws.stage;  // free outside the sweep; the host guarantees mem_size <= 2*12*kStageLd
  const int ld = hist.ld;
  const int npl = (n + 31) >> 5;
  const bool can_help = !kStrict && kMaxHelp > 0 && sh.helpable != 0 && nwarp > 1;

  int phase = PH_FETCH;
  int prob = -1, k = 0, end = 0, bound = 0, evals = 0, count = 0, ret = 0;
  bool brackt = false, touched = false;
  double step = 0.0, fx = 0.0, finit = 0.0, dgtest = 0.0, dstest = 0.0, mu = 0.0, nu = 0.0;
  const double* planes = nullptr;
  long long t_logic = 0;
  // team state: as an owner (seq, done_base), as a helper (own = my owner's warp, hidx = my attach index, seen = last go word)
  bool helper = false;
  int seq = 0, done_base = 0, own = -1, hidx = -1, seen = -1, hprob = -1;

  while (true) {
    long long t_prev = 0;
    int r0 = 0, r1 = kVR, h = 0, prob_t = prob;
    const double* planes_t = planes;
    SweepOut so{ws.gd, ws.rscal};
    if (!helper) {
      // ---------------- owner: take a problem if needed, first part of the evaluation, publish it to the team
      if (phase == PH_FETCH) {
        int local = 0;
        if (grp.clist) {
          local = fetch_by_chunk(grp.clist, grp.cbeg, grp.chead, grp.key, grp.prob_id, grp.n_chunks, bd.ready, bd.epoch, lane);
          if (local < 0) local = grp.count;
        } else {
          if (lane == 0) local = atomicAdd(work_counter, 1);
          local = __shfl_sync(FULL, local, 0);
          if (local < grp.count && grp.order) local = grp.order[local];
        }
        if (local >= grp.count) {
          // the queue is empty: retire as an owner (releases my helpers), help the others or leave
          if (lane == 0) {
            st_flag(&ctl[warp].alive, 0);
            __threadfence_block();
            st_flag(&ctl[warp].go, go_word(++seq, -1));
          }
          __syncwarp();
          if (!can_help) break;
          helper = true;
          continue;
        }
        prob = grp.prob_id[local];
        planes = grp.planes + (size_t)local * sh.H * 4 * sh.Mpad;
        layout_problem(sh, bd, prob, grp.planes + (size_t)local * sh.H * 4 * sh.Mpad, lane);
#pragma unroll 1
        for (int i = lane; i < n; i += 32) ws.x[i] = bd.x0[(size_t)prob * bd.n_stride + i];
        __syncwarp();
        stage_problem(sh, bd, prob, ws, lane);
        if (HAS_OBS) stage_obstacles(bd, prob, ws.obs, lane);
        evals = 0;
        phase = PH_FIRST;
        prob_t = prob;
        planes_t = planes;
      }
      if (g_phase_clock && t_logic) {
        if (lane == 0) atomicAdd(g_phase_clock + 7, (unsigned long long)(clock64() - t_logic));
      }
#if !DFTPAV_STRICT
      eval_pre(sh, P, smem, ws, ws.x, lane, t_prev);
      if (can_help) {
        ++seq;
        int hh = 0;
        if (lane == 0) {
          hh = min(ld_flag(&ctl[warp].nhelp), kMaxHelp);
          ctl[warp].prob = prob;
          ctl[warp].planes = reinterpret_cast<unsigned long long>(planes);
          __threadfence_block();  // coefficients, segment scalars, problem: visible before the go word
          st_flag(&ctl[warp].go, go_word(seq, hh));
        }
        h = __shfl_sync(FULL, hh, 0);
      }
      rank_chunk(0, h + 1, r0, r1);
#endif
    } else {
      // ---------------- helper: find an owner of this CTA that is still solving, serve its evaluations
      if (own < 0) {
        int alive_l = 0, nh_l = 0;
        if (lane < nwarp && lane != warp) {
          alive_l = ld_flag(&ctl[lane].alive);
          nh_l = ld_flag(&ctl[lane].nhelp);
        }
        if (__ballot_sync(FULL, alive_l != 0) == 0u) break;  // nobody left to help
        // fewest helpers first (preferring the owner furthest into its solve instead was measured: no difference)
        int key = (alive_l && nh_l < kMaxHelp) ? nh_l * 64 + lane : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(FULL, key, o));
        if (key == 0x7fffffff) {
          __nanosleep(1000);  // every team is full: look again later
          continue;
        }
        const int o = key & 63;
        int idx = -1;
        if (lane == 0) {
          int old = ld_flag(&ctl[o].nhelp);
          while (old < kMaxHelp) {
            const int prev = atomicCAS(&ctl[o].nhelp, old, old + 1);
            if (prev == old) {
              idx = old;
              break;
            }
            old = prev;
          }
        }
        idx = __shfl_sync(FULL, idx, 0);
        if (idx < 0) continue;
        own = o, hidx = idx, seen = -1, hprob = -1;
      }
      int word = 0;
      if (lane == 0) {
        while ((word = ld_flag(&ctl[own].go)) == seen) __nanosleep(200);
        __threadfence_block();
      }
      word = __shfl_sync(FULL, word, 0);
      __syncwarp();
      seen = word;
      const int hw = (word & 15) - 1;
      if (hw < 0) {  // my owner has retired
        own = -1;
        continue;
      }
      if (hw <= hidx) continue;  // this evaluation was published before I attached
      h = hw;
      rank_chunk(hidx + 1, h + 1, r0, r1);
      // what the sweep reads: the owner's coefficients and segment scalars, copied into my own (idle) workspace
      WarpWS ows = carve_ws(smem + sh.lu_doubles + own * sh.ws_doubles, sh);
      prob_t = ctl[own].prob;
      planes_t = reinterpret_cast<const double*>(ctl[own].planes);
#pragma unroll 1
      for (int i = lane; i < 2 * sh.rows; i += 32) ws.c[i] = ows.c[i];
      if (lane < sh.S) {
        ws.seg[lane].dT = ows.seg[lane].dT;
        ws.seg[lane].trajtime = ows.seg[lane].trajtime;
        ws.seg[lane].singul = ows.seg[lane].singul;
      }
      if (HAS_OBS && hprob != prob_t) {
        for (int i = lane; i < 4 * kObsCap; i += 32) ws.obs[i] = ows.obs[i];
      }
      hprob = prob_t;
      __syncwarp();
      so.gd = ows.gd;
      so.rscal = ows.rscal;
    }
    // ---------------- the single sweep site.  Owner without helpers: all virtual ranks and the end passes in one call.
    //                  Owner with helpers: its chunk of ranks, then - once the helpers have reported - the end passes.
    //                  Helper: its chunk of ranks.
#if !DFTPAV_STRICT
    {
      const int rounds = (!helper && h > 0) ? 2 : 1;
#pragma unroll 1
      for (int round = 0; round < rounds; round++) {
        bool do_end = !helper && h == 0;
        if (round == 1) {
          if (lane == 0) {
            while (ld_flag(&ctl[warp].done) != done_base + h) __nanosleep(100);
            __threadfence_block();
          }
          __syncwarp();
          done_base += h;
          r0 = r1 = kVR;
          do_end = true;
        }
        sweep_ranks<HAS_OBS>(sh, P, bd, prob_t, planes_t, ws, so, lane, r0, r1, do_end);
      }
    }
    if (helper) {
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        atomicAdd(&ctl[own].done, 1);
      }
      continue;
    }
    fx = eval_post(sh, P, smem, ws, ws.x, ws.g, lane, t_prev);
#else
    fx = strict_eval(sh, P, planes, smem, ws, ws.adj, ws.x, ws.g, lane);
#endif
    evals++;
    if (g_phase_clock) t_logic = clock64();
    FINE_T0();

    bool start_ls = false;   // begin a new line search from the current (x, g, d, step)
    bool finish = false;     // problem done: write results
    bool revert = false;     // finish after restoring (xp, gp) (line-search failure, lbfgs.hpp:604-611)

    if (phase == PH_FIRST) {
      // lbfgs.hpp:524-555
      if (lane == 0) ws.pf[0] = fx;
#pragma unroll 1
      for (int i = lane; i < n; i += 32) ws.d[i] = -ws.g[i];
      __syncwarp();
      const double gnorm_inf = warp_maxabs(ws.g, n, lane);
      const double xnorm_inf = warp_maxabs(ws.x, n, lane);
      k = 0;
      if (gnorm_inf / fmax(1.0, xnorm_inf) < P.g_epsilon) {
        ret = DFTPAV_LBFGS_CONVERGENCE;
        finish = true;
      } else {
        step = 1.0 / sqrt(warp_dot(ws.d, ws.d, n, lane));
        k = 1;
        end = 0;
        bound = 0;
        start_ls = true;
      }
    } else {
      // ---------------- one trial of line_search_lewisoverton (lbfgs.hpp:317-388)
      ++count;
      int ls = 0;          // > 0: accepted with `count` evaluations, < 0: error, 0: continue with a new step
      if (isinf(fx) || isnan(fx)) {
        ls = DFTPAV_LBFGSERR_INVALID_FUNCVAL;
      } else if (P.past > 0 && fabs(finit - fx) / (fabs(finit) + 1.0) < P.delta / P.past) {
        ls = count;  // lbfgs.hpp:326-329
      } else {
        bool accepted = false;
        if (fx > finit + step * dgtest) {
          nu = step;
          brackt = true;
        } else if (warp_dot(ws.g, ws.d, n, lane) < dstest) {
          mu = step;
        } else {
          accepted = true;
        }
        if (accepted) {
          ls = count;
        } else if (P.max_linesearch <= count) {
          ls = DFTPAV_LBFGSERR_MAXIMUMLINESEARCH;
        } else if (brackt && (nu - mu) < P.machine_prec * nu) {
          ls = DFTPAV_LBFGSERR_WIDTHTOOSMALL;
        } else {
          if (brackt) {
            step = 0.5 * (mu + nu);
          } else {
            step *= 2.0;
          }
          if (step < P.min_step) {
            ls = DFTPAV_LBFGSERR_MINIMUMSTEP;
          } else if (step > P.max_step) {
            if (touched) {
              ls = DFTPAV_LBFGSERR_MAXIMUMSTEP;
            } else {
              touched = true;
              step = P.max_step;
            }
          }
        }
      }
      if (ls != 0 && out.trace && k - 1 < out.trace_cap) {
        double* tr = out.trace + ((size_t)prob * out.trace_cap + (k - 1)) * 3;
        if (lane == 0) {
          tr[0] = fx;
          tr[1] = step;
          tr[2] = (double)ls;
        }
      }
      FINE_ADD(8);  // line-search trial logic
      if (ls == 0) {
        // next trial point
#pragma unroll 1
        for (int i = lane; i < n; i += 32) ws.x[i] = ws.xp[i] + step * ws.d[i];
        __syncwarp();
        FINE_ADD(9);  // next trial point
        continue;
      }
      if (ls < 0) {
        ret = ls;
        revert = true;
        finish = true;
      } else {
        // ---------------- end of iteration k (lbfgs.hpp:628-743)
        const double gnorm_inf = warp_maxabs(ws.g, n, lane);
        const double xnorm_inf = warp_maxabs(ws.x, n, lane);
        if (gnorm_inf / fmax(1.0, xnorm_inf) < P.g_epsilon) {
          ret = DFTPAV_LBFGS_CONVERGENCE;
          finish = true;
        }
        if (!finish && 0 < P.past) {
          if (P.past <= k) {
            const double rate = fabs(ws.pf[k % P.past] - fx) / fmax(1.0, fabs(fx));
            if (rate < P.delta) {
              ret = DFTPAV_LBFGS_STOP;
              finish = true;
            }
          }
          if (!finish) {
            __syncwarp();
            if (lane == 0) ws.pf[k % P.past] = fx;
            __syncwarp();
          }
        }
        if (!finish && P.max_iterations != 0 && P.max_iterations <= k) {
          ret = DFTPAV_LBFGSERR_MAXIMUMITERATION;
          finish = true;
        }
        FINE_ADD(10);  // end-of-iteration tests
        if (!finish) {
          ++k;
          // history update (lbfgs.hpp:676-687)
          double* se = lm_s + (size_t)end * ld;
          double* ye = lm_y + (size_t)end * ld;
          double r4[4] = {0.0, 0.0, 0.0, 0.0};  // ys, yy, ss, gg
#pragma unroll 1
          for (int i = lane; i < n; i += 32) {
            const double sv = ws.x[i] - ws.xp[i];
            const double yv = ws.g[i] - ws.gp[i];
            se[i] = sv;
            ye[i] = yv;
            r4[0] += yv * sv;
            r4[1] += yv * yv;
            r4[2] += sv * sv;
            r4[3] += ws.gp[i] * ws.gp[i];
            ws.d[i] = -ws.g[i];
          }
          double ys, yy, ss, gg;
          if (kStrict) {  // sequential sums of the reference (lbfgs.hpp:681-704)
            __syncwarp();
            ys = warp_dot(ye, se, n, lane);
            yy = warp_dot(ye, ye, n, lane);
            ss = warp_dot(se, se, n, lane);
            gg = warp_dot(ws.gp, ws.gp, n, lane);
          } else {
            warp_sum4(r4);
            ys = r4[0], yy = r4[1], ss = r4[2], gg = r4[3];
          }
          if (lane == 0) lm_ys[end] = DFTPAV_STORE_YS(ys);
          __syncwarp();
          const double cau = ss * sqrt(gg) * P.cautious_factor;
          FINE_ADD(11);  // history update
          if (ys > cau) {
            ++bound;
            bound = m < bound ? m : bound;
            end = (end + 1) % m;
#if DFTPAV_STRICT
            two_loop_generic(ws.d, alpha_sm, lm_s, lm_y, lm_ys, ld, m, end, bound, n, lane, ys, yy);
#else
            if (npl <= 3) {
              // the direction in registers (n <= 96)
              if (npl == 1) two_loop<1>(ws.d, alpha_sm, lm_s, lm_y, lm_ys, ld, m, end, bound, n, lane, ys, yy);
              else if (npl == 2) two_loop<2>(ws.d, alpha_sm, lm_s, lm_y, lm_ys, ld, m, end, bound, n, lane, ys, yy);
              else two_loop<3>(ws.d, alpha_sm, lm_s, lm_y, lm_ys, ld, m, end, bound, n, lane, ys, yy);
            } else {
              // larger problems (up to 4 x 64 pieces, n = 517): the direction stays in shared memory
              two_loop_generic(ws.d, alpha_sm, lm_s, lm_y, lm_ys, ld, m, end, bound, n, lane, ys, yy);
            }
#endif
          }
          FINE_ADD(12);  // two-loop recursion
          step = 1.0;
          start_ls = true;
        }
      }
    }

    if (start_ls) {
      // head of line_search_lewisoverton (lbfgs.hpp:288-310) + xp/gp snapshot (lbfgs.hpp:560-561)
#pragma unroll 1
      for (int i = lane; i < n; i += 32) {
        ws.xp[i] = ws.x[i];
        ws.gp[i] = ws.g[i];
      }
      __syncwarp();
      if (!(step > 0.0)) {
        ret = DFTPAV_LBFGSERR_INVALIDPARAMETERS;
        finish = true;
      } else {
        const double dginit = warp_dot(ws.gp, ws.d, n, lane);
        if (0.0 < dginit) {
          ret = DFTPAV_LBFGSERR_INCREASEGRADIENT;
          finish = true;
        } else {
          finit = fx;
          dgtest = P.f_dec_coeff * dginit;
          dstest = P.s_curv_coeff * dginit;
          count = 0;
          brackt = false;
          touched = false;
          mu = 0.0;
          nu = P.max_step;
#pragma unroll 1
          for (int i = lane; i < n; i += 32) ws.x[i] = ws.xp[i] + step * ws.d[i];
          __syncwarp();
          phase = PH_LS;
        }
      }
      FINE_ADD(13);  // line-search head
    }

    if (finish) {
      if (revert) {
#pragma unroll 1
        for (int i = lane; i < n; i += 32) {
          ws.x[i] = ws.xp[i];
          ws.g[i] = ws.gp[i];
        }
        __syncwarp();
      }
      // results (traj_optimizer.cpp:176-201); fx keeps the last tried value on failure (quirk Q8)
      bool success = (ret == DFTPAV_LBFGS_CONVERGENCE || ret == DFTPAV_LBFGS_CANCELED || ret == DFTPAV_LBFGS_STOP ||
                      ret == DFTPAV_LBFGSERR_MAXIMUMITERATION || ret == DFTPAV_LBFGSERR_MAXIMUMLINESEARCH);
      if (fx >= P.cost_limit) success = false;
#pragma unroll 1
      for (int i = lane; i < bd.n_stride; i += 32) out.x[(size_t)prob * bd.n_stride + i] = i < n ? ws.x[i] : 0.0;
#pragma unroll 1
      for (int i = lane; i < 2 * bd.row_stride; i += 32)
        out.coeffs[(size_t)prob * bd.row_stride * 2 + i] = i < 2 * sh.rows ? ws.c[i] : 0.0;
      if (lane < sh.S) out.piece_dT[bd.seg_base[prob] + lane] = ws.seg[lane].dT;
      if (lane == 0) {
        out.cost[prob] = fx;
        out.status[prob] = ret;
        out.success[prob] = success ? 1 : 0;
        out.iters[prob] = k;
        out.evals[prob] = evals;
      }
      __syncwarp();
      phase = PH_FETCH;
    }
  }
}

// boundary-state clamps (traj_optimizer.cpp:55-76); x0 is packed on the host (pack_x0 in dftpav_api.cu: glibc atan2)
__global__ void prepare_states_kernel(const ShapeDev sh, ParamsDev P, const int* __restrict__ prob_id, int count,
                                      const int* __restrict__ seg_base, const int* __restrict__ singul,
                                      const double* __restrict__ ini_raw, const double* __restrict__ fin_raw,
                                      double* __restrict__ ini, double* __restrict__ fin) {
  const int local = blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= count) return;
  const int prob = prob_id[local];
  const int g0 = seg_base[prob];
  for (int s = 0; s < sh.S; s++) {
    const int g = g0 + s;
    const int sidx = singul[g] > 0 ? 0 : 1;
    const double max_vel = P.max_vel[sidx], max_acc = P.max_acc[sidx];
    double st[2][6];
    for (int k = 0; k < 6; k++) {
      st[0][k] = ini_raw[6 * g + k];
      st[1][k] = fin_raw[6 * g + k];
    }
    for (int w = 0; w < 2; w++) {
      for (int c = 2; c <= 4; c += 2) {
        const double lim = (c == 2) ? max_vel : max_acc;
        const double vx = st[w][c], vy = st[w][c + 1];
        const double z = vx * vx + vy * vy;
        if (sqrt(z) >= lim) {
          double nx = vx, ny = vy;
          if (z > 0.0) {
            const double sz = sqrt(z);
            nx = vx / sz;
            ny = vy / sz;
          }
          st[w][c] = nx * (lim - 1.0e-2);
          st[w][c + 1] = ny * (lim - 1.0e-2);
        }
      }
    }
    for (int k = 0; k < 6; k++) {
      ini[6 * g + k] = st[0][k];
      fin[6 * g + k] = st[1][k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Read-back side of the path: batched evaluation of the solved trajectories straight from
// the device-resident coefficients.  Replaces the per-sample calls consumers make on plan_utils::Trajectory -
// getPos / getdSigma / getddSigma (poly_traj_utils.hpp:77-87, 179-211, 554-600), getAngle / getCurv / getVel / getAcc /
// getLatAcc / getSteer (233-296, 604-645), GetState (378-406) - e.g. the 0.05 s collision re-check of the executing
// trajectory (traj_server_ros.cpp:385-397).  Polynomial values are accumulated in the reference's order with explicitly
// rounded multiplies / adds (no FMA contraction), so positions, derivatives, speed and accelerations are BIT-IDENTICAL to
// the reference; angle, curvature and steer go through atan2 / pow / atan of the CUDA math library (<= 2 ulp).
// ------------------------------------------------------------------------------------------------
struct SegView {
  const double* c;  // [6N][2] coefficients of the segment, row 6i+k = coefficient of t^k of piece i
  double dT;
  int N, singul;
};
__device__ __forceinline__ SegView seg_view(const SampleDev& sd, int g) {
  SegView v;
  v.N = sd.seg_N[g];
  v.dT = sd.piece_dT[g];
  v.singul = sd.singul[g];
  v.c = sd.coeffs + ((size_t)sd.seg_prob[g] * sd.row_stride + sd.seg_row0[g]) * 2;
  return v;
}
// Trajectory::getTotalDuration (poly_traj_utils.hpp:425-434): the sum, not N * dT
__device__ __forceinline__ double seg_duration(const SegView& v) {
  double tot = 0.0;
  for (int i = 0; i < v.N; i++) tot = __dadd_rn(tot, v.dT);
  return tot;
}
// Trajectory::locatePieceIdx (poly_traj_utils.hpp:510-528)
__device__ __forceinline__ int seg_locate(const SegView& v, double& t) {
  int idx;
  for (idx = 0; idx < v.N && t > v.dT; idx++) t = __dadd_rn(t, -v.dT);
  if (idx == v.N) {
    idx--;
    t = __dadd_rn(t, v.dT);
  }
  return idx;
}
// Piece::getPos / getdSigma / getddSigma: `acc += (factor * tn) * coefficient`, highest-power-first columns
__device__ __forceinline__ void piece_eval(const double* c, double t, V2& pos, V2& ds, V2& dds) {
  pos = ds = dds = mk(0.0, 0.0);
  double tn = 1.0;
#pragma unroll
  for (int k = 0; k <= 5; k++) {  // column i = 5 - k of the reference holds the coefficient of t^k
    pos.x = __dadd_rn(pos.x, __dmul_rn(tn, c[2 * k]));
    pos.y = __dadd_rn(pos.y, __dmul_rn(tn, c[2 * k + 1]));
    tn = __dmul_rn(tn, t);
  }
  tn = 1.0;
#pragma unroll
  for (int k = 1; k <= 5; k++) {
    const double f = __dmul_rn((double)k, tn);
    ds.x = __dadd_rn(ds.x, __dmul_rn(f, c[2 * k]));
    ds.y = __dadd_rn(ds.y, __dmul_rn(f, c[2 * k + 1]));
    tn = __dmul_rn(tn, t);
  }
  tn = 1.0;
#pragma unroll
  for (int k = 2; k <= 5; k++) {
    const double f = __dmul_rn((double)((k - 1) * k), tn);
    dds.x = __dadd_rn(dds.x, __dmul_rn(f, c[2 * k]));
    dds.y = __dadd_rn(dds.y, __dmul_rn(f, c[2 * k + 1]));
    tn = __dmul_rn(tn, t);
  }
}
__device__ __forceinline__ double norm_rn(V2 a) { return sqrt(__dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y))); }

__global__ void sample_kernel(const SampleDev sd, long long nq, const int* __restrict__ seg, const double* __restrict__ tq,
                              int flags, double wheel_base, double* __restrict__ out) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    double* o = out + q * 12;
    const int g = seg[q];
    if (g < 0 || g >= sd.G || sd.seg_N[g] < 1) {
      for (int k = 0; k < 12; k++) o[k] = 0.0;
      continue;
    }
    const SegView v = seg_view(sd, g);
    double t = tq[q];
    if (flags & (DFTPAV_SAMPLE_CLAMP | DFTPAV_SAMPLE_GETSTATE)) {
      const double tot = seg_duration(v);
      if (t > tot) t = tot;
    }
    const int idx = seg_locate(v, t);
    V2 p, d1, d2;
    piece_eval(v.c + 12 * idx, t, p, d1, d2);
    const double sg = (double)v.singul;
    const double nv = norm_rn(d1);
    const double cross = __dadd_rn(__dmul_rn(d1.x, d2.y), -__dmul_rn(d1.y, d2.x));
    const double dotp = __dadd_rn(__dmul_rn(d1.x, d2.x), __dmul_rn(d1.y, d2.y));
    double curv, acc, lat;
    const double vel = __dmul_rn(sg, nv);
    // Piece::getCurv / getAcc / getLatAcc test norm() < 1e-6; getStateExpPos tests abs(vel) with the integer abs (quirk Q11)
    const bool still = (flags & DFTPAV_SAMPLE_GETSTATE) ? (fabs(vel) < 1.0) : (nv < 1e-6);
    if (still) {
      curv = acc = lat = 0.0;
    } else {
      curv = __dmul_rn(sg, cross) / pow(nv, 3.0);
      acc = __dmul_rn(sg, dotp) / nv;
      lat = __dmul_rn(sg, cross) / nv;
    }
    o[0] = p.x, o[1] = p.y, o[2] = d1.x, o[3] = d1.y, o[4] = d2.x, o[5] = d2.y;
    o[6] = atan2(__dmul_rn(sg, d1.y), __dmul_rn(sg, d1.x));
    o[7] = curv;
    o[8] = vel;
    o[9] = acc;
    o[10] = atan(__dmul_rn(wheel_base, curv));
    o[11] = lat;
  }
}

// The sampling loop of the collision re-check (traj_server_ros.cpp:385-397): every segment at t = 0, dt, 2 dt, ... while
// t < duration, t accumulated like the reference's `t += 0.05`; out[g][k] = (x, y, yaw).  One thread per segment.
__global__ void sample_uniform_kernel(const SampleDev sd, double dt, int max_per_seg, double* __restrict__ out,
                                      int* __restrict__ count) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= sd.G) return;
  int k = 0;
  if (sd.seg_N[g] >= 1) {
    const SegView v = seg_view(sd, g);
    const double dur = seg_duration(v);
    double* o = out + (size_t)g * max_per_seg * 3;
    for (double t = 0.0; t < dur && k < max_per_seg; t = __dadd_rn(t, dt), k++) {
      double tl = t;
      const int idx = seg_locate(v, tl);
      V2 p, d1, d2;
      piece_eval(v.c + 12 * idx, tl, p, d1, d2);
      const double sg = (double)v.singul;
      o[3 * k] = p.x;
      o[3 * k + 1] = p.y;
      o[3 * k + 2] = atan2(__dmul_rn(sg, d1.y), __dmul_rn(sg, d1.x));
    }
  }
  count[g] = k;
}

// plan_utils::Piece::getCoeffMat order for the wire format of msg/PolyTraj.msg (order, coef_x, coef_y, duration; float32):
// per piece the six coefficients HIGHEST power first (poly_traj_utils.hpp:19, 987-997), pieces of all segments concatenated
__global__ void pack_polytraj_kernel(const SampleDev sd, float* __restrict__ coef_x, float* __restrict__ coef_y,
                                     float* __restrict__ duration) {
  const int g = blockIdx.x;
  if (g >= sd.G) return;
  const SegView v = seg_view(sd, g);
  const long long base = sd.seg_piece0[g];
  for (int i = threadIdx.x; i < v.N; i += blockDim.x) {
    duration[base + i] = (float)v.dT;
    for (int k = 0; k < 6; k++) {
      coef_x[(base + i) * 6 + k] = (float)v.c[2 * (6 * i + 5 - k)];
      coef_y[(base + i) * 6 + k] = (float)v.c[2 * (6 * i + 5 - k) + 1];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Corridor generation: TrajPlanner::getRectangleConst (traj_manager.cpp:1211-1469) on the
// occupancy grid, one WARP per sampled state.  The expansion state machine (round-robin over +dy, +dx, -dy, -dx, one cell
// per step, three ray-marched segments per step) is warp-uniform; the ray march of a segment - the reference's loop
// `for (dl = 0; dl < |p2-p1|; dl += res/2)` of CheckIfCollisionUsingLine (map_adapter.cpp:117-129) - is spread over the
// lanes: the result of that loop is a boolean OR over its sample points (it returns at the first occupied one), so the
// order does not matter.  Every coordinate is computed with explicitly rounded operations in the reference's
// association; dl comes from a table of the sequentially accumulated sums (built on the host), so the sampled cells are
// exactly the reference's.  Output: four half-planes (n_x, n_y, p_x, p_y) per state = the optimizer's raw plane layout.
// HBM/L2-bound byte lookups; no tensor cores, no shared memory.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool grid_occupied(const CorridorDev& cd, double px, double py) {
  // GridMapND::GetCoordUsingGlobalPosition / CheckCoordInRange / CheckIfEqualUsingGlobalPosition (semantics.cc:169-288)
  const double fx = round(__ddiv_rn(__dadd_rn(px, -cd.origin_x), cd.resolution));
  const double fy = round(__ddiv_rn(__dadd_rn(py, -cd.origin_y), cd.resolution));
  if (!(fx >= 0.0 && fx < (double)cd.size_x && fy >= 0.0 && fy < (double)cd.size_y)) return false;
  return cd.data[(size_t)(int)fx + (size_t)(int)fy * cd.size_x] == cd.occupied;
}
// CheckIfCollisionUsingLine by one warp.  Four chunks of 32 sample points are in flight at a time: a chunk costs two fp64
// divisions and one dependent byte load from L2 per lane, and the strip sides are up to 25 m (250 points) long - issuing the
// chunks one after the other left the warp waiting ~700 cycles per chunk (4.5 ms for one wave of states).
__device__ __forceinline__ bool line_blocked(const CorridorDev& cd, double p1x, double p1y, double p2x, double p2y, int lane) {
  const double dx = __dadd_rn(p2x, -p1x), dy = __dadd_rn(p2y, -p1y);
  const double nrm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  constexpr int kFly = 4;
  bool hit = false;
  for (int k0 = 0; k0 < cd.n_dl && !hit; k0 += 32 * kFly) {
    bool occ = false, all_on = true;
    double dl[kFly];
#pragma unroll
    for (int u = 0; u < kFly; u++) {
      const int k = k0 + 32 * u + lane;
      dl[u] = k < cd.n_dl ? cd.dl[k] : 1.0e300;
    }
#pragma unroll
    for (int u = 0; u < kFly; u++) {
      const bool on = dl[u] < nrm;
      all_on = all_on && on;
      if (on) {
        const double px = __dadd_rn(__ddiv_rn(__dmul_rn(dx, dl[u]), nrm), p1x);
        const double py = __dadd_rn(__ddiv_rn(__dmul_rn(dy, dl[u]), nrm), p1y);
        occ = occ || grid_occupied(cd, px, py);
      }
    }
    if (__any_sync(FULL, occ)) hit = true;
    if (!__all_sync(FULL, all_on)) break;  // dl is increasing: the segment ends inside these chunks
  }
  if (!hit) hit = grid_occupied(cd, p2x, p2y);  // the end point itself (map_adapter.cpp:127)
  return hit;
}

__global__ void corridor_kernel(const CorridorDev cd, long long n, const double* __restrict__ states,
                                const double* __restrict__ cos_sin, double* __restrict__ hpolys) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const double step = cd.resolution, checkl = __ddiv_rn(cd.resolution, 2.0);
  (void)checkl;  // the ray-march spacing lives in cd.dl (accumulated sums of resolution / 2)
  for (long long q = warp0; q < n; q += nwarps) {
    const double x = states[3 * q], y = states[3 * q + 1], yaw = states[3 * q + 2];
    double c, s;
    if (cos_sin) {
      c = cos_sin[2 * q], s = cos_sin[2 * q + 1];
    } else {
      c = cos(yaw), s = sin(yaw);
    }
    const double ms = -s;
    auto RX = [&](double a, double b) { return __dadd_rn(__dmul_rn(c, a), __dmul_rn(ms, b)); };
    auto RY = [&](double a, double b) { return __dadd_rn(__dmul_rn(s, a), __dmul_rn(c, b)); };
    double spx = x, spy = y;
    double width = cd.veh_width, length = cd.veh_length;
    const double dcr = cd.veh_d_cr;
    int notfin = 0xf;
    double ex0 = 0.0, ex1 = 0.0, ex2 = 0.0, ex3 = 0.0;
    while (notfin) {
      for (int i = 0; i < 4; i++) {
        if (!((notfin >> i) & 1)) continue;
        const double hl = __ddiv_rn(length, 2.0), hw = __ddiv_rn(width, 2.0);
        double a1, b1, a2, b2, na1, nb1, na2, nb2;
        if (i == 0) {  // +dy
          a1 = __dadd_rn(hl, dcr), b1 = hw, a2 = __dadd_rn(-hl, dcr), b2 = hw;
          na1 = a1, nb1 = __dadd_rn(hw, step), na2 = a2, nb2 = nb1;
        } else if (i == 1) {  // +dx
          a1 = __dadd_rn(hl, dcr), b1 = -hw, a2 = a1, b2 = hw;
          na1 = __dadd_rn(__dadd_rn(step, hl), dcr), nb1 = -hw, na2 = na1, nb2 = hw;
        } else if (i == 2) {  // -dy
          a1 = __dadd_rn(-hl, dcr), b1 = -hw, a2 = __dadd_rn(hl, dcr), b2 = -hw;
          na1 = a1, nb1 = __dadd_rn(-hw, -step), na2 = a2, nb2 = nb1;
        } else {  // -dx
          a1 = __dadd_rn(-hl, dcr), b1 = hw, a2 = a1, b2 = -hw;
          na1 = __dadd_rn(a1, -step), nb1 = hw, na2 = na1, nb2 = -hw;
        }
        const double p1x = __dadd_rn(spx, RX(a1, b1)), p1y = __dadd_rn(spy, RY(a1, b1));
        const double p2x = __dadd_rn(spx, RX(a2, b2)), p2y = __dadd_rn(spy, RY(a2, b2));
        const double n1x = __dadd_rn(spx, RX(na1, nb1)), n1y = __dadd_rn(spy, RY(na1, nb1));
        const double n2x = __dadd_rn(spx, RX(na2, nb2)), n2y = __dadd_rn(spy, RY(na2, nb2));
        if (line_blocked(cd, p1x, p1y, n1x, n1y, lane) || line_blocked(cd, n1x, n1y, n2x, n2y, lane) ||
            line_blocked(cd, n2x, n2y, p2x, p2y, lane)) {
          notfin &= ~(1 << i);
          continue;
        }
        double e = i == 0 ? ex0 : (i == 1 ? ex1 : (i == 2 ? ex2 : ex3));
        e = __dadd_rn(e, step);
        if (i == 0) ex0 = e; else if (i == 1) ex1 = e; else if (i == 2) ex2 = e; else ex3 = e;
        if (e >= cd.limit_bound) {
          notfin &= ~(1 << i);
          continue;
        }
        const double h2 = __ddiv_rn(step, 2.0);
        double ta, tb;
        if (i == 0) ta = 0.0, tb = h2, width = __dadd_rn(width, step);
        else if (i == 1) ta = h2, tb = 0.0, length = __dadd_rn(length, step);
        else if (i == 2) ta = 0.0, tb = -h2, width = __dadd_rn(width, step);
        else ta = -h2, tb = 0.0, length = __dadd_rn(length, step);
        spx = __dadd_rn(spx, RX(ta, tb));
        spy = __dadd_rn(spy, RY(ta, tb));
      }
    }
    if (lane == 0) {
      const double hl = __ddiv_rn(cd.veh_length, 2.0), hw = __ddiv_rn(cd.veh_width, 2.0);
      const double fa = __dadd_rn(__dadd_rn(hl, dcr), ex1), ra = __dadd_rn(__dadd_rn(-hl, dcr), -ex3);
      const double lb = __dadd_rn(hw, ex0), rb = __dadd_rn(-hw, -ex2);
      double* h = hpolys + 16 * q;
      h[0] = ms, h[1] = c, h[2] = __dadd_rn(x, RX(fa, lb)), h[3] = __dadd_rn(y, RY(fa, lb));
      h[4] = c, h[5] = s, h[6] = __dadd_rn(x, RX(fa, rb)), h[7] = __dadd_rn(y, RY(fa, rb));
      h[8] = s, h[9] = -c, h[10] = __dadd_rn(x, RX(ra, rb)), h[11] = __dadd_rn(y, RY(ra, rb));
      h[12] = -c, h[13] = ms, h[14] = __dadd_rn(x, RX(ra, lb)), h[15] = __dadd_rn(y, RY(ra, lb));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Queue order (scheduling only, no effect on results): the persistent warps take the problems in order of decreasing
// predicted solve length.  The batch time is set by the longest solves, and the initial cost is the one cheap predictor of
// the solve length (Spearman 0.38-0.46 with the evaluation count at config 3; the part of f(x0) that needs NO corridor
// plane - jerk, time, feasibility penalties, moving obstacles - ranks as well as the whole: 0.376 vs 0.378, 768 problems
// through the oracle), so the ordering pass runs while the planes are still being uploaded.  Measured at batch 4096: 20.3 ms
// against 21.9 ms in input order (perfect longest-first: 19.8 ms; tools/order_experiment.py).
// Rank by counting, per upload chunk: clist[#{j: chunk_j < chunk_i} + #{j in my chunk: key_j > key_i}] = i, ties by
// index, NaN first - O(count^2) comparisons, 16 M at batch 4096, on the device; cbeg[c] = first list entry of chunk c.
// ------------------------------------------------------------------------------------------------
constexpr int kOrderTile = 2048;  // keys per shared-memory tile
__global__ void __launch_bounds__(256) chunk_order_kernel(const double* __restrict__ key, const int* __restrict__ prob_id,
                                                          int count, int chunk, int n_chunks, int* __restrict__ clist,
                                                          int* __restrict__ cbeg) {
  // one WARP per problem i: the lanes split the comparisons (a thread per problem looped over all 4096 keys: 0.35 ms)
  __shared__ double tkey[kOrderTile];
  __shared__ int tchunk[kOrderTile];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 8 + warp;
  const double fi = i < count ? key[prob_id[i]] : 0.0;
  const int ci = i < count ? prob_id[i] / chunk : 0;
  const bool ni = fi != fi;
  int pos = 0;  // problems served before i: those of lower chunks, and those of my chunk with a larger key
  for (int j0 = 0; j0 < count; j0 += kOrderTile) {
    const int nj_ = min(kOrderTile, count - j0);
    for (int t = threadIdx.x; t < nj_; t += 256) {
      const int pj = prob_id[j0 + t];
      tkey[t] = key[pj];
      tchunk[t] = pj / chunk;
    }
    __syncthreads();
    for (int t = lane; t < nj_; t += 32) {
      const double fj = tkey[t];
      const int cj = tchunk[t];
      const int j = j0 + t;
      const bool nj = fj != fj;
      const bool before = (nj && !ni) || (nj == ni && (ni ? j < i : (fj > fi || (fj == fi && j < i))));
      pos += (cj < ci || (cj == ci && before)) ? 1 : 0;
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(FULL, pos, o);
  if (i < count && lane == 0) clist[pos] = i;
  if (blockIdx.x == 0 && (int)threadIdx.x <= n_chunks) {  // first list entry of every chunk
    int below = 0;
    for (int j = 0; j < count; j++) below += prob_id[j] / chunk < (int)threadIdx.x ? 1 : 0;
    cbeg[threadIdx.x] = below;
  }
}

// ------------------------------------------------------------------------------------------------
// Batched MINCO fits: MinJerkOpt::reset + generate + getTraj (poly_traj_utils.hpp:880-997) for
// many short trajectories at once - the fits of the predicted obstacle trajectories (TrajPlanner::ConverSurroundTrajFromPoints,
// traj_manager.cpp:743-789) and of the initial guess (traj_manager.cpp:579-581).  One THREAD per trajectory runs the reference's
// banded substitutions on the constant LU factor with explicitly rounded operations (no FMA contraction), so the coefficients
// are BIT-IDENTICAL to the reference's; the work is tiny (a few thousand flops per fit), the batch supplies the parallelism.
// Output: Piece::coeffMat per piece, 2x6 column-major, HIGHEST power first - the obstacle-pool format of dftpav_batch.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double fnma_rn(double a, double b, double c) { return __dadd_rn(c, -__dmul_rn(a, b)); }  // c - a*b, two roundings
__global__ void minco_fit_kernel(int n_traj, const int* __restrict__ piece_ptr, const double* __restrict__ dT,
                                 const double* __restrict__ head, const double* __restrict__ tail,
                                 const double* __restrict__ inner_pts, const long long* __restrict__ inner_ptr,
                                 const double* __restrict__ lu_tables, const int* __restrict__ lu_off,
                                 double* __restrict__ scratch, double* __restrict__ coeffs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_traj) return;
  const int p0 = piece_ptr[k], N = piece_ptr[k + 1] - p0, n = 6 * N;
  const double* C = lu_tables + lu_off[k];  // packed LU columns of this N: A(i,j) = C[j*kLuLd + (i - j + 6)]
  double* b = scratch + 12 * (size_t)p0;    // [6N][2]
  const double t1 = dT[k], t2 = __dmul_rn(t1, t1), t3 = __dmul_rn(t2, t1), t4 = __dmul_rn(t2, t2), t5 = __dmul_rn(t4, t1);
  const double tInv[6] = {1.0, __ddiv_rn(1.0, t1), __ddiv_rn(1.0, t2), __ddiv_rn(1.0, t3), __ddiv_rn(1.0, t4), __ddiv_rn(1.0, t5)};
  // right-hand side (poly_traj_utils.hpp:968-977)
  for (int i = 0; i < 2 * n; i++) b[i] = 0.0;
  const double* P = inner_pts + inner_ptr[k];
  for (int col = 0; col < 2; col++) {
    b[0 + col] = head[6 * k + 0 + col];
    b[2 + col] = __dmul_rn(head[6 * k + 2 + col], t1);
    b[4 + col] = __dmul_rn(head[6 * k + 4 + col], t2);
    for (int i = 0; i < N - 1; i++) b[2 * (6 * i + 5) + col] = P[2 * i + col];
    b[2 * (n - 3) + col] = tail[6 * k + 0 + col];
    b[2 * (n - 2) + col] = __dmul_rn(tail[6 * k + 2 + col], t1);
    b[2 * (n - 1) + col] = __dmul_rn(tail[6 * k + 4 + col], t2);
  }
  // BandedSystem::solve (poly_traj_utils.hpp:805-826)
  for (int j = 0; j <= n - 1; j++) {
    const int iM = min(j + 6, n - 1);
    for (int i = j + 1; i <= iM; i++) {
      const double a = C[(size_t)j * kLuLd + (i - j + 6)];
      if (a != 0.0) {
        b[2 * i] = fnma_rn(a, b[2 * j], b[2 * i]);
        b[2 * i + 1] = fnma_rn(a, b[2 * j + 1], b[2 * i + 1]);
      }
    }
  }
  for (int j = n - 1; j >= 0; j--) {
    const double dj = C[(size_t)j * kLuLd + 6];
    b[2 * j] = __ddiv_rn(b[2 * j], dj);
    b[2 * j + 1] = __ddiv_rn(b[2 * j + 1], dj);
    const int iM = max(0, j - 6);
    for (int i = iM; i <= j - 1; i++) {
      const double a = C[(size_t)j * kLuLd + (i - j + 6)];
      if (a != 0.0) {
        b[2 * i] = fnma_rn(a, b[2 * j], b[2 * i]);
        b[2 * i + 1] = fnma_rn(a, b[2 * j + 1], b[2 * i + 1]);
      }
    }
  }
  // c = b / dT^k (981-984), then getTraj's column reversal (991): highest power first
  for (int i = 0; i < N; i++)
    for (int kk = 0; kk < 6; kk++)
      for (int col = 0; col < 2; col++)
        coeffs[12 * (size_t)(p0 + i) + 2 * (5 - kk) + col] = __dmul_rn(b[2 * (6 * i + kk) + col], tInv[kk]);
}

// ------------------------------------------------------------------------------------------------
// host-callable launchers (used by dftpav_api.cu)
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// Collision re-check of the solved trajectories against the occupancy grid (the consumer loop of
// traj_server_ros.cpp:385-397): sample k of a gear segment is t_k = 0 + dt + ... + dt (the reference's `t += 0.05`, taken
// from a host-built table of the sequentially accumulated sums) while t_k < duration; the vehicle footprint at
// (getPos(t), getAngle(t)) is the vertex set of SemanticMapManager::CheckCollisionUsingPosAndYaw (semantic_map_manager.cc:
// 639-661): box centre = pos + d_cr (cos yaw, sin yaw), corners and edge points every `spacing` metres from
// ShapeUtils::GetDenseVerticesOfOrientedBoundingBox (shapes.cc:110-149), each vertex one grid lookup.  The reference stops
// at the first colliding sample; the result per segment is that sample's index (or -1).  One WARP per segment, lane <->
// sample, every coordinate with explicitly rounded operations in the reference's association (yaw, cos and sin come from
// the CUDA libm: they can differ from glibc by an ulp, which moves a vertex by ~1e-15 m).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool edge_hits(const CorridorDev& cd, double ax, double ay, double bx, double by, double spacing) {
  // for (dl = res; dl < |b - a|; dl += res) p = dl / |b - a| * (b - a) + a      (shapes.cc:128-143)
  const double dx = __dadd_rn(bx, -ax), dy = __dadd_rn(by, -ay);
  const double nrm = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
  bool hit = false;
  for (double dl = spacing; dl < nrm; dl = __dadd_rn(dl, spacing)) {
    const double f = __ddiv_rn(dl, nrm);
    hit = hit || grid_occupied(cd, __dadd_rn(__dmul_rn(f, dx), ax), __dadd_rn(__dmul_rn(f, dy), ay));
  }
  return hit;
}

__global__ void collision_kernel(const SampleDev sd, const CorridorDev cd, const double* __restrict__ t_tab, int n_tab,
                                 double dt, double spacing, const int* __restrict__ seg_prob, int* __restrict__ first_hit,
                                 int* __restrict__ collide) {
  const int lane = threadIdx.x & 31;
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= sd.G) return;
  int first = 0x7fffffff;
  bool truncated = false;
  if (sd.seg_N[g] >= 1) {
    const SegView v = seg_view(sd, g);
    const double dur = seg_duration(v);
    truncated = __dadd_rn(t_tab[n_tab - 1], dt) < dur;  // the caller's sample cap is too small for this segment
    const double sg = (double)v.singul;
    const double hl = __dmul_rn(0.5, cd.veh_length), hw = __dmul_rn(0.5, cd.veh_width);
    for (int k0 = 0; k0 < n_tab; k0 += 32) {
      const int k = k0 + lane;
      const double t = k < n_tab ? t_tab[k] : dur;
      const bool on = t < dur;
      bool hit = false;
      if (on) {
        double tl = t;
        const int idx = seg_locate(v, tl);
        V2 p, d1, d2;
        piece_eval(v.c + 12 * idx, tl, p, d1, d2);
        const double yaw = atan2(__dmul_rn(sg, d1.y), __dmul_rn(sg, d1.x));
        const double c = cos(yaw), s = sin(yaw);
        const double ox = __dadd_rn(p.x, __dmul_rn(cd.veh_d_cr, c)), oy = __dadd_rn(p.y, __dmul_rn(cd.veh_d_cr, s));
        const double lc = __dmul_rn(hl, c), ls = __dmul_rn(hl, s), wc = __dmul_rn(hw, c), ws = __dmul_rn(hw, s);
        const double c1x = __dadd_rn(__dadd_rn(ox, lc), ws), c1y = __dadd_rn(__dadd_rn(oy, ls), -wc);
        const double c2x = __dadd_rn(__dadd_rn(ox, lc), -ws), c2y = __dadd_rn(__dadd_rn(oy, ls), wc);
        const double c3x = __dadd_rn(__dadd_rn(ox, -lc), -ws), c3y = __dadd_rn(__dadd_rn(oy, -ls), wc);
        const double c4x = __dadd_rn(__dadd_rn(ox, -lc), ws), c4y = __dadd_rn(__dadd_rn(oy, -ls), -wc);
        hit = edge_hits(cd, c1x, c1y, c2x, c2y, spacing) || edge_hits(cd, c2x, c2y, c3x, c3y, spacing) ||
              edge_hits(cd, c3x, c3y, c4x, c4y, spacing) || edge_hits(cd, c4x, c4y, c1x, c1y, spacing) ||
              grid_occupied(cd, c1x, c1y) || grid_occupied(cd, c2x, c2y) || grid_occupied(cd, c3x, c3y) ||
              grid_occupied(cd, c4x, c4y);
      }
      const unsigned hits = __ballot_sync(FULL, hit);
      if (hits) {
        first = k0 + __ffs(hits) - 1;
        break;
      }
      if (!__all_sync(FULL, on)) break;  // the table is increasing: the segment ends inside this chunk
    }
  }
  if (lane == 0) {
    first_hit[g] = first != 0x7fffffff ? first : (truncated ? -2 : -1);
    if (first != 0x7fffffff) atomicOr(collide + seg_prob[g], 1);
  }
}

// ------------------------------------------------------------------------------------------------
// Front-end resampling: KinoAstar::evaluatePos / evaluateLength (kino_astar.cpp:468-521,
// 763-795) and the sampling loop of TrajPlanner::RunMINCOParking (traj_manager.cpp:531-567) that turns a searched path into
// the optimizer's inputs (corridor states, inner waypoints, piece counts and durations).  One THREAD per query / state: the
// work per pose is a short scan over the path's gear segments and samples; the batch supplies the parallelism.  Every
// operation explicitly rounded in the reference's association (+ - * / sqrt only; pow(x, 2) is x * x): bit-identical to the
// CPU restatement.
// ------------------------------------------------------------------------------------------------
__device__ double fe_length(double curt, double locallength, double max_vel, double max_acc, double startV, double endV) {
  const double startv2 = __dmul_rn(startV, startV), endv2 = __dmul_rn(endV, endV), maxv2 = __dmul_rn(max_vel, max_vel);
  const double two_a = __dmul_rn(2.0, max_acc), half_a = __dmul_rn(0.5, max_acc);
  const double critical_len = __dadd_rn(__ddiv_rn(__dadd_rn(maxv2, -startv2), two_a), __ddiv_rn(__dadd_rn(maxv2, -endv2), two_a));
  auto ramp = [&](double tq) {  // startV * t + 0.5 * max_acc * t^2
    return __dadd_rn(__dmul_rn(startV, tq), __dmul_rn(half_a, __dmul_rn(tq, tq)));
  };
  if (locallength >= critical_len) {
    const double t1 = __ddiv_rn(__dadd_rn(max_vel, -startV), max_acc);
    const double t2 = __dadd_rn(t1, __ddiv_rn(__dadd_rn(locallength, -critical_len), max_vel));
    if (curt <= t1) return ramp(curt);
    if (curt <= t2) return __dadd_rn(ramp(t1), __dmul_rn(__dadd_rn(curt, -t1), max_vel));
    const double u = __dadd_rn(curt, -t2);
    return __dadd_rn(__dadd_rn(__dadd_rn(ramp(t1), __dmul_rn(__dadd_rn(t2, -t1), max_vel)), __dmul_rn(max_vel, u)),
                     -__dmul_rn(half_a, __dmul_rn(u, u)));
  }
  const double tmpv = sqrt(__dmul_rn(0.5, __dadd_rn(__dadd_rn(startv2, endv2), __dmul_rn(two_a, locallength))));
  const double tmpt = __ddiv_rn(__dadd_rn(tmpv, -startV), max_acc);
  if (curt <= tmpt) return ramp(curt);
  const double u = __dadd_rn(curt, -tmpt);
  return __dadd_rn(__dadd_rn(ramp(tmpt), __dmul_rn(tmpv, u)), -__dmul_rn(half_a, __dmul_rn(u, u)));
}

__device__ void fe_evaluate_pos(const FrontendDev& fe, int p, double t, double* out) {
  const int s0 = fe.shot_ptr[p], ns = fe.shot_ptr[p + 1] - s0;
  const double* st = fe.shot_time + s0;
  const int* si = fe.shot_index + s0 + p;
  const double* pts = fe.pts + 3 * (size_t)fe.pt_ptr[p];
  double total = 0.0;
  for (int i = 0; i < ns; i++) total = __dadd_rn(total, st[i]);
  t = 0.0 < t ? t : 0.0;      // std::max<double>(0, t)
  t = total < t ? total : t;  // std::min<double>(., totalTrajTime)
  int index = -1;
  double tmpT = 0.0, cut = 0.0;
  for (int i = 0; i < ns; i++) {
    tmpT = __dadd_rn(tmpT, st[i]);
    if (tmpT >= t) {
      index = i;
      cut = __dadd_rn(__dadd_rn(t, -tmpT), st[i]);
      break;
    }
  }
  if (index < 0) index = ns - 1;  // (NaN times only; the reference would index with -1)
  const double initv = index == 0 ? fe.start_vel[p] : fe.non_siguav;
  const double finv = index == ns - 1 ? fe.end_vel[p] : fe.non_siguav;
  const double locallength = fe.shot_length[s0 + index];
  const int front = si[index], back = si[index + 1];
  const double* lt = pts + 3 * (size_t)front;
  const int nl = back - front + 1;
  const bool fwd = fe.shot_s[s0 + index] > 0;
  const double arclength = fe_length(cut, locallength, fwd ? fe.max_forward_vel : fe.max_backward_vel,
                                     fwd ? fe.max_forward_acc : fe.max_backward_acc, initv, finv);
  double tmparc = 0.0;
  for (int i = 0; i < nl - 1; i++) {
    const double ax = lt[3 * i], ay = lt[3 * i + 1], ayaw = lt[3 * i + 2];
    const double bx = lt[3 * i + 3], by = lt[3 * i + 4], byaw = lt[3 * i + 5];
    const double dx = __dadd_rn(bx, -ax), dy = __dadd_rn(by, -ay);
    const double l = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    tmparc = __dadd_rn(tmparc, l);
    if (tmparc >= arclength) {
      const double l1 = __dadd_rn(tmparc, -arclength), l2 = __dadd_rn(l, -l1);
      const double w1 = __ddiv_rn(l1, l), w2 = __ddiv_rn(l2, l);
      out[0] = __dadd_rn(__dmul_rn(w1, ax), __dmul_rn(w2, bx));
      out[1] = __dadd_rn(__dmul_rn(w1, ay), __dmul_rn(w2, by));
      double yaw = __dadd_rn(__dmul_rn(w1, ayaw), __dmul_rn(w2, byaw));
      const double kPi = 3.14159265358979323846, k2Pi = 6.28318530717958647692;
      if (fabs(__dadd_rn(byaw, -ayaw)) >= kPi) {
        if (byaw <= 0.0) yaw = __dadd_rn(__dmul_rn(w1, ayaw), __dmul_rn(w2, __dadd_rn(byaw, k2Pi)));
        else if (ayaw <= 0.0) yaw = __dadd_rn(__dmul_rn(w1, __dadd_rn(ayaw, k2Pi)), __dmul_rn(w2, byaw));
      }
      out[2] = yaw;
      return;
    }
  }
  out[0] = lt[3 * (nl - 1)], out[1] = lt[3 * (nl - 1) + 1], out[2] = lt[3 * (nl - 1) + 2];
}

__global__ void frontend_pos_kernel(const FrontendDev fe, long long nq, const int* __restrict__ path,
                                    const double* __restrict__ t, double* __restrict__ out) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= nq) return;
  double o[3];
  fe_evaluate_pos(fe, path[q], t[q], o);
  out[3 * q] = o[0], out[3 * q + 1] = o[1], out[3 * q + 2] = o[2];
}

// one block per gear segment; seg_path / seg_pn / seg_tpp / seg_base / seg_st0 / seg_in0 = path, piece count, piece
// duration, base time (the durations of the path's earlier segments, summed in order), first state, first inner point
__global__ void frontend_resample_kernel(const FrontendDev fe, int n_seg, const int* __restrict__ seg_path,
                                         const int* __restrict__ seg_pn, const double* __restrict__ seg_tpp,
                                         const double* __restrict__ seg_base, const long long* __restrict__ seg_st0,
                                         const long long* __restrict__ seg_in0, int K, int Kd, double* __restrict__ states,
                                         double* __restrict__ inner) {
  const int g = blockIdx.x;
  if (g >= n_seg) return;
  const int pn = seg_pn[g], p = seg_path[g];
  const double tpp = seg_tpp[g], base = seg_base[g];
  const int n_states = 2 * (Kd + 1) + (pn - 2) * (K + 1);
  for (int u = threadIdx.x; u < n_states; u += blockDim.x) {
    int i, k;
    if (u < Kd + 1) {
      i = 0, k = u;
    } else if (u - (Kd + 1) < (pn - 2) * (K + 1)) {
      i = 1 + (u - (Kd + 1)) / (K + 1), k = (u - (Kd + 1)) % (K + 1);
    } else {
      i = pn - 1, k = u - (Kd + 1) - (pn - 2) * (K + 1);
    }
    const int resolution = (i == 0 || i == pn - 1) ? Kd : K;
    double res_time = 0.0;  // res_time += ego_piece_dur_vec[i] after every piece
    for (int q = 0; q < i; q++) res_time = __dadd_rn(res_time, tpp);
    const double t = __dadd_rn(__dadd_rn(base, res_time), __dmul_rn(__ddiv_rn((double)k, (double)resolution), tpp));
    double o[3];
    fe_evaluate_pos(fe, p, t, o);
    double* so = states + 3 * (seg_st0[g] + u);
    so[0] = o[0], so[1] = o[1], so[2] = o[2];
    if (k == resolution && i != pn - 1) {
      double* io = inner + 2 * (seg_in0[g] + i);
      io[0] = o[0], io[1] = o[1];
    }
  }
}

cudaError_t launch_solve(bool has_obs, int grid, int block, size_t smem, cudaStream_t st, const ShapeDev& sh,
                         const ParamsDev& P, const BatchDev& bd, const GroupDev& grp, const OutDev& out,
                         const HistDev& hist, int* counter) {
  if (has_obs) {
    cudaFuncSetAttribute(solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    solve_kernel<true><<<grid, block, smem, st>>>(sh, P, bd, grp, out, hist, counter);
  } else {
    cudaFuncSetAttribute(solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    solve_kernel<false><<<grid, block, smem, st>>>(sh, P, bd, grp, out, hist, counter);
  }
  return cudaGetLastError();
}

cudaError_t launch_eval(bool has_obs, int grid, int block, size_t smem, cudaStream_t st, const ShapeDev& sh,
                        const ParamsDev& P, const BatchDev& bd, const GroupDev& grp, const double* x, double* f,
                        double* g, double* coeffs, double* dT) {
  if (has_obs) {
    cudaFuncSetAttribute(eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    eval_kernel<true><<<grid, block, smem, st>>>(sh, P, bd, grp, x, f, g, coeffs, dT);
  } else {
    cudaFuncSetAttribute(eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    eval_kernel<false><<<grid, block, smem, st>>>(sh, P, bd, grp, x, f, g, coeffs, dT);
  }
  return cudaGetLastError();
}

cudaError_t launch_prepare(int grid, int block, cudaStream_t st, const ShapeDev& sh, const ParamsDev& P,
                           const int* prob_id, int count, const int* seg_base, const int* singul, const double* ini_raw,
                           const double* fin_raw, double* ini, double* fin) {
  prepare_states_kernel<<<grid, block, 0, st>>>(sh, P, prob_id, count, seg_base, singul, ini_raw, fin_raw, ini, fin);
  return cudaGetLastError();
}

cudaError_t launch_sample(cudaStream_t st, const SampleDev& sd, long long nq, const int* seg, const double* t, int flags,
                          double wheel_base, double* out) {
  const int grid = (int)std::min<long long>((nq + 255) / 256, 148LL * 16);
  sample_kernel<<<grid > 0 ? grid : 1, 256, 0, st>>>(sd, nq, seg, t, flags, wheel_base, out);
  return cudaGetLastError();
}
cudaError_t launch_sample_uniform(cudaStream_t st, const SampleDev& sd, double dt, int max_per_seg, double* out, int* count) {
  sample_uniform_kernel<<<(sd.G + 127) / 128, 128, 0, st>>>(sd, dt, max_per_seg, out, count);
  return cudaGetLastError();
}
cudaError_t launch_pack_polytraj(cudaStream_t st, const SampleDev& sd, float* cx, float* cy, float* dur) {
  pack_polytraj_kernel<<<sd.G > 0 ? sd.G : 1, 64, 0, st>>>(sd, cx, cy, dur);
  return cudaGetLastError();
}

cudaError_t launch_minco_fit(cudaStream_t st, int n_traj, const int* piece_ptr, const double* dT, const double* head,
                             const double* tail, const double* inner_pts, const long long* inner_ptr, const double* lu_tables,
                             const int* lu_off, double* scratch, double* coeffs) {
  minco_fit_kernel<<<(n_traj + 63) / 64, 64, 0, st>>>(n_traj, piece_ptr, dT, head, tail, inner_pts, inner_ptr, lu_tables, lu_off,
                                                      scratch, coeffs);
  return cudaGetLastError();
}

cudaError_t launch_chunk_order(cudaStream_t st, const double* key, const int* prob_id, int count, int chunk, int n_chunks,
                               int* clist, int* cbeg) {
  chunk_order_kernel<<<(count + 7) / 8, 256, 0, st>>>(key, prob_id, count, chunk, n_chunks, clist, cbeg);
  return cudaGetLastError();
}

cudaError_t launch_collision(cudaStream_t st, const SampleDev& sd, const CorridorDev& cd, const double* t_tab, int n_tab,
                             double dt, double spacing, const int* seg_prob, int* first_hit, int* collide) {
  const int warps = sd.G, blocks = (warps + 7) / 8;
  collision_kernel<<<blocks > 0 ? blocks : 1, 256, 0, st>>>(sd, cd, t_tab, n_tab, dt, spacing, seg_prob, first_hit, collide);
  return cudaGetLastError();
}

cudaError_t launch_frontend_pos(cudaStream_t st, const FrontendDev& fe, long long nq, const int* path, const double* t,
                                double* out) {
  frontend_pos_kernel<<<(int)((nq + 127) / 128), 128, 0, st>>>(fe, nq, path, t, out);
  return cudaGetLastError();
}
cudaError_t launch_frontend_resample(cudaStream_t st, const FrontendDev& fe, int n_seg, const int* seg_path, const int* seg_pn,
                                     const double* seg_tpp, const double* seg_base, const long long* seg_st0,
                                     const long long* seg_in0, int K, int Kd, double* states, double* inner) {
  frontend_resample_kernel<<<n_seg, 128, 0, st>>>(fe, n_seg, seg_path, seg_pn, seg_tpp, seg_base, seg_st0, seg_in0, K, Kd,
                                                  states, inner);
  return cudaGetLastError();
}

cudaError_t launch_corridor(cudaStream_t st, const CorridorDev& cd, long long n, const double* states, const double* cos_sin,
                            double* hpolys) {
  const long long blocks = std::min<long long>((n + 7) / 8, 148LL * 8);  // 8 warps per block, one warp per state
  corridor_kernel<<<(int)std::max<long long>(1, blocks), 256, 0, st>>>(cd, n, states, cos_sin, hpolys);
  return cudaGetLastError();
}

int ws_doubles_host(int rows, int npad, int S, int has_obs) { return ws_doubles_for(rows, npad, S, has_obs); }
int max_warps_host() { return DFTPAV_MAX_WARPS; }
int virtual_ranks_host() { return kVR; }
int strict_build_host() { return kStrict ? 1 : 0; }

cudaError_t set_phase_clock(unsigned long long* dev_ptr) {
  return cudaMemcpyToSymbol(g_phase_clock, &dev_ptr, sizeof(dev_ptr));
}

}  // namespace dftpav
