// dftpav_api.cu — host side of the C-ABI declared in include/dftpav_b200.h.
//
// Host responsibilities (everything else runs in the kernels of dftpav_kernels.cu):
//   * validation of the reference's preconditions (traj_optimizer.cpp:26-48),
//   * grouping the batch by problem shape (S, N_0..N_{S-1}); one persistent-kernel launch per shape,
//   * MinJerkOpt::reset (poly_traj_utils.hpp:880-951): the time-normalised MINCO matrix depends only on the piece
//     count N, so it is filled and LU-factorised (BandedSystem::factorizeLU, poly_traj_utils.hpp:776-800) ONCE per
//     distinct N on the host and shipped as a packed table that each CTA stages into shared memory with TMA,
//   * device memory, H2D/D2H copies, CUDA-event timings.
// There is no CPU implementation of the solve behind this ABI: without an sm_100 device every call fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/dftpav_b200.h"
#include "dftpav_device.cuh"

namespace dftpav {
cudaError_t launch_solve(bool has_obs, int grid, int block, size_t smem, cudaStream_t st, const ShapeDev& sh,
                         const ParamsDev& P, const BatchDev& bd, const GroupDev& grp, const OutDev& out,
                         const HistDev& hist, int* counter);
cudaError_t launch_eval(bool has_obs, int grid, int block, size_t smem, cudaStream_t st, const ShapeDev& sh,
                        const ParamsDev& P, const BatchDev& bd, const GroupDev& grp, const double* x, double* f,
                        double* g, double* coeffs, double* dT);
cudaError_t launch_layout(int grid, int block, cudaStream_t st, const ShapeDev& sh, const int* prob_id, int count,
                          const long long* samp_base, const long long* plane_ptr, int uniform_H, const double* raw,
                          double* soa);
cudaError_t launch_prepare(int grid, int block, cudaStream_t st, const ShapeDev& sh, const ParamsDev& P,
                           const int* prob_id, int count, const int* seg_base, const int* singul, const double* ini_raw,
                           const double* fin_raw, double* ini, double* fin);
cudaError_t launch_sample(cudaStream_t st, const SampleDev& sd, long long nq, const int* seg, const double* t, int flags,
                          double wheel_base, double* out);
cudaError_t launch_sample_uniform(cudaStream_t st, const SampleDev& sd, double dt, int max_per_seg, double* out, int* count);
cudaError_t launch_pack_polytraj(cudaStream_t st, const SampleDev& sd, float* cx, float* cy, float* dur);
cudaError_t launch_minco_fit(cudaStream_t st, int n_traj, const int* piece_ptr, const double* dT, const double* head,
                             const double* tail, const double* inner_pts, const long long* inner_ptr, const double* lu_tables,
                             const int* lu_off, double* scratch, double* coeffs);
cudaError_t launch_chunk_order(cudaStream_t st, const double* key, const int* prob_id, int count, int chunk, int n_chunks,
                               int* clist, int* cbeg);
cudaError_t launch_frontend_pos(cudaStream_t st, const FrontendDev& fe, long long nq, const int* path, const double* t,
                                double* out);
cudaError_t launch_frontend_resample(cudaStream_t st, const FrontendDev& fe, int n_seg, const int* seg_path, const int* seg_pn,
                                     const double* seg_tpp, const double* seg_base, const long long* seg_st0,
                                     const long long* seg_in0, int K, int Kd, double* states, double* inner);
cudaError_t launch_collision(cudaStream_t st, const SampleDev& sd, const CorridorDev& cd, const double* t_tab, int n_tab,
                             double dt, double spacing, const int* seg_prob, int* first_hit, int* collide);
cudaError_t launch_corridor(cudaStream_t st, const CorridorDev& cd, long long n, const double* states, const double* cos_sin,
                            double* hpolys);
int ws_doubles_host(int rows, int npad, int S, int has_obs);
int max_warps_host();
int virtual_ranks_host();
int strict_build_host();
cudaError_t set_phase_clock(unsigned long long* dev_ptr);
}  // namespace dftpav

using namespace dftpav;

namespace {

// ---------------------------------------------------------------------------------------------
// MinJerkOpt::reset on the host: banded matrix fill + LU (poly_traj_utils.hpp:776-800, 897-948)
// ---------------------------------------------------------------------------------------------
// Time-normalised minimum-jerk collocation matrix of N pieces (poly_traj_utils.hpp:913-960 with every T_i = 1; the reference
// rebuilds and refactorises it per call with the actual durations, here the durations are moved into the right-hand side).
template <class Setter>
static void fill_minco_matrix(int N, Setter&& A) {
  const int n = 6 * N;
  A(0, 0) = 1.0;
  A(1, 1) = 1.0;
  A(2, 2) = 2.0;
  for (int i = 0; i < N - 1; i++) {
    const int r = 6 * i;
    A(r + 3, r + 3) = 6.0, A(r + 3, r + 4) = 24.0, A(r + 3, r + 5) = 60.0, A(r + 3, r + 9) = -6.0;
    A(r + 4, r + 4) = 24.0, A(r + 4, r + 5) = 120.0, A(r + 4, r + 10) = -24.0;
    for (int k = 0; k < 6; k++) A(r + 5, r + k) = 1.0;
    for (int k = 0; k < 6; k++) A(r + 6, r + k) = 1.0;
    A(r + 6, r + 6) = -1.0;
    A(r + 7, r + 1) = 1.0, A(r + 7, r + 2) = 2.0, A(r + 7, r + 3) = 3.0, A(r + 7, r + 4) = 4.0, A(r + 7, r + 5) = 5.0;
    A(r + 7, r + 7) = -1.0;
    A(r + 8, r + 2) = 2.0, A(r + 8, r + 3) = 6.0, A(r + 8, r + 4) = 12.0, A(r + 8, r + 5) = 20.0;
    A(r + 8, r + 8) = -2.0;
  }
  const int m = n;
  for (int k = 0; k < 6; k++) A(m - 3, m - 6 + k) = 1.0;
  A(m - 2, m - 5) = 1.0, A(m - 2, m - 4) = 2.0, A(m - 2, m - 3) = 3.0, A(m - 2, m - 2) = 4.0, A(m - 2, m - 1) = 5.0;
  A(m - 1, m - 4) = 2.0, A(m - 1, m - 3) = 6.0, A(m - 1, m - 2) = 12.0, A(m - 1, m - 1) = 20.0;
}

// Dense path of the MINCO solves.  The right-hand side of the collocation system has only N+5 non-zero rows (three head
// rows, one position row per interior waypoint, three tail rows) and the adjoint solve is only read at the same rows
// (gdP, gradIni, gradFin: poly_traj_utils.hpp:1037-1066), so both solves reduce to products with the constant
// 6N x (N+5) matrix M = A^-1[:, R]:  b~ = M b_R  and  lambda_R = M^T gdC.  M is computed once per N in extended
// precision (Gauss-Jordan with partial pivoting in long double) and rounded to double: its products are closer to the
// exact solution than the reference's unpivoted banded LU (measured 1.5e-16 vs 2.3e-15 relative at N = 24).
// Layout: Mt[j][row] with leading dimension 6N+1 (odd: conflict-free both along rows and along columns).
static int dense_row_of(int j, int N) { return j < 3 ? j : (j < N + 2 ? 6 * (j - 3) + 5 : 5 * N - 5 + j); }
static int dense_ld(int N) { return 6 * N + 1; }
static size_t dense_doubles(int N) { return ((size_t)(N + 5) * dense_ld(N) + 1) / 2 * 2; }
static const std::vector<double>& dense_minco_table(int N) {
  static std::mutex mu;
  static std::map<int, std::vector<double>> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(N);
  if (it != cache.end()) return it->second;
  const int n = 6 * N, nc = N + 5, w = n + nc;
  std::vector<long double> a((size_t)n * w, 0.0L);
  fill_minco_matrix(N, [&](int i, int j) -> long double& { return a[(size_t)i * w + j]; });
  for (int j = 0; j < nc; j++) a[(size_t)dense_row_of(j, N) * w + n + j] = 1.0L;
  for (int k = 0; k < n; k++) {
    int p = k;
    for (int i = k + 1; i < n; i++)
      if (fabsl(a[(size_t)i * w + k]) > fabsl(a[(size_t)p * w + k])) p = i;
    if (p != k)
      for (int j = 0; j < w; j++) std::swap(a[(size_t)k * w + j], a[(size_t)p * w + j]);
    const long double inv = 1.0L / a[(size_t)k * w + k];
    for (int j = k; j < w; j++) a[(size_t)k * w + j] *= inv;
    for (int i = 0; i < n; i++) {
      const long double f = a[(size_t)i * w + k];
      if (i == k || f == 0.0L) continue;
      for (int j = k; j < w; j++) a[(size_t)i * w + j] -= f * a[(size_t)k * w + j];
    }
  }
  std::vector<double> t(dense_doubles(N), 0.0);
  const int ld = dense_ld(N);
  for (int j = 0; j < nc; j++)
    for (int i = 0; i < n; i++) t[(size_t)j * ld + i] = (double)a[(size_t)i * w + n + j];
  return cache.emplace(N, std::move(t)).first->second;
}

// LU factor of the banded collocation matrix for the fallback path (tables too large for the dense products) and for the
// strict verification build.  Doolittle without pivoting, computed row by row ("ikj" order): row i is reduced by the
// already final rows k = i-6 .. i-1 in ascending k, so every entry receives exactly the updates
// A(i,j) -= L(i,k) U(k,j), k ascending, followed (below the diagonal) by one division by U(k,k) - the same rounded
// operations per entry as BandedSystem::factorizeLU (poly_traj_utils.hpp:776-800), whose right-looking loop nest visits
// them column by column; the factors are bit-identical (tests/test_minco_dense_table.py compares them with the oracle's).
struct BandLU {
  int n;
  std::vector<double> d;  // band storage d[(i-j+6)*n + j], lower = upper = 6
  explicit BandLU(int N) : n(6 * N), d((size_t)13 * 6 * N, 0.0) {
    auto A = [&](int i, int j) -> double& { return d[(size_t)(i - j + 6) * n + j]; };
    fill_minco_matrix(N, A);
    for (int i = 1; i < n; i++) {
      for (int k = std::max(0, i - 6); k < i; k++) {
        double& lik = A(i, k);
        if (lik == 0.0) continue;  // structural zero: no fill-in from row k (the reference skips exact zeros as well)
        lik /= A(k, k);
        const int jmax = std::min(k + 6, n - 1);
        for (int j = k + 1; j <= jmax; j++) {
          const double ukj = A(k, j);
          if (ukj != 0.0) A(i, j) -= lik * ukj;
        }
      }
    }
  }
  // The kernels skip structurally-zero LU entries with compile-time masks of period 6 (kBandMask in
  // dftpav_kernels.cu); this checks that every non-zero of the factor is covered (L passes: all but the last piece).
  bool masks_hold() const {
    static const int LM[6] = {0x60, 0x70, 0x78, 0x3c, 0x1e, 0x0e}, UM[6] = {0x00, 0x00, 0x00, 0x56, 0x6e, 0x06};
    static const int ULT[6] = {0x00, 0x0c, 0x06, 0x46, 0x42, 0x30}, LLT[6] = {0x7e, 0x7c, 0x78, 0x00, 0x00, 0x3e};
    auto A = [&](int i, int j) -> double { return (i < 0 || j < 0 || i >= n || j >= n) ? 0.0 : d[(size_t)(i - j + 6) * n + j]; };
    for (int j = 0; j < n; j++)
      for (int k = 1; k <= 6; k++) {
        if (j < n - 6 && A(j + k, j) != 0.0 && !((LM[j % 6] >> k) & 1)) return false;
        if (j < n - 6 && A(j, j - k) != 0.0 && !((LLT[j % 6] >> k) & 1)) return false;
        if (A(j - k, j) != 0.0 && !((UM[j % 6] >> k) & 1)) return false;
        if (A(j, j + k) != 0.0 && !((ULT[j % 6] >> k) & 1)) return false;
      }
    return true;
  }
  // packed column table: C[j][0..12] = band rows 0..12 of column j, C[j][13] = 1/U(j,j)
  void pack(double* C) const {
    for (int j = 0; j < n; j++) {
      double* Cj = C + (size_t)j * kLuLd;
      for (int r = 0; r < 13; r++) {
        const int i = j + r - 6;  // matrix row of band row r
        Cj[r] = (i >= 0 && i < n) ? d[(size_t)r * n + j] : 0.0;
      }
      Cj[13] = 1.0 / Cj[6];
      Cj[14] = 0.0;
      Cj[15] = 0.0;
    }
  }
};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct Group {
  ShapeDev sh;
  std::vector<int> probs;
  DevBuf d_prob_id, d_planes, d_lu, d_order, d_clist, d_cq;  // d_cq: cbeg[kMaxChunks+1] then chead[kMaxChunks]
  int warps = 0, grid = 0;
  size_t smem = 0;
};

inline int samples_of(int N, int K, int Kd) { return (N - 2) * (K + 1) + 2 * (Kd + 1); }

constexpr int kMaxChunks = 32;  // chunks of the plane upload that dftpav_solve overlaps with the solve

}  // namespace

struct dftpav_handle {
  dftpav_params prm;
  ParamsDev pd;
  int device = 0;
  int num_sms = 0;
  size_t smem_optin = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_copy = nullptr;
  int epoch = 0, n_chunks = 1, chunk = 1;   // plane-upload hand-shake (BatchDev::ready / laid)
  int* h_epoch = nullptr;                    // pinned [kMaxChunks], all = epoch (source of the arrival-flag copies)
  std::vector<long long> h_samp_base, h_plane_ptr;
  std::vector<double> h_x0;
  long long total_planes = 0;
  int uniform_H = 0;                         // -1: ragged corridor (h_plane_ptr)
  int min_n_stride = 0, min_row_stride = 0;   // dftpav_solve_multi: a shard uses the strides of the whole batch
  bool force_banded = false;                 // test hook: banded LU substitutions even where the dense tables fit
  std::string err;
  bool uploaded = false, solved = false, has_obs = false;
  dftpav_dims dims{};
  dftpav_stats stats{};
  std::vector<Group*> groups, group_pool;
  // host copies needed later
  std::vector<int> h_seg_ptr, h_piece_num, h_n, h_evalbytes, h_status0;
  int param_status = 0;   // LBFGSERR_INVALID_* of the parameter block (lbfgs.hpp:456-501), 0 = parameters usable
  int n_rejected = 0;     // problems violating a reference precondition (traj_optimizer.cpp:26-48)
  // device buffers
  DevBuf seg_base, t_now, obs_ptr, obs_idx, singul, ini_raw, fin_raw, ini, fin, samp_base,
      plane_ptr, planes_raw, obs_start, obs_dur, obs_piece_ptr, obs_piece_dur, obs_coeffs, obs_uniform, x0, xin, fout, gout;
  DevBuf o_x, o_cost, o_status, o_success, o_iters, o_evals, o_coeffs, o_dT, counter;
  DevBuf hist_s, hist_y, hist_ys, trace, ready, laid;
  DevBuf seg_prob, seg_row0, seg_N, seg_piece0, q_seg, q_t, q_out, q_cnt;  // trajectory evaluator (dftpav_sample*)
  long long pieces_total = 0;
  // corridor generator: device copy of the occupancy grid, ray-march offsets, and whether planes_raw holds generated polygons
  DevBuf map_data, map_dl, cor_states, cor_cs;
  DevBuf fit_pp, fit_dT, fit_head, fit_tail, fit_inner, fit_iptr, fit_lu, fit_luoff, fit_scratch, fit_coeffs;  // dftpav_minco_fit
  DevBuf fe_arena, fe_out;  // dftpav_frontend_*: all inputs in one upload, outputs
  dftpav_gridmap map_meta{};
  int map_n_dl = 0;
  std::vector<int> dbg_order;  // experiment hook (dftpav_debug_set_order): serve the work queue in this order of problem ids
  bool planes_from_corridor = false;
  int dbg_order_key = 0;  // experiment hook: 0 = the caller's choice, 1 = plane-free key, 2 = the whole f(x0)
  long long corridor_n = -1;  // >= 0: planes_raw holds corridor_n generated rectangles (4 planes each)
  int trace_cap = 0;
  int hist_ld = 0, hist_slots = 0;
  int upload_launches = 0;
  BatchDev bd{};
};

namespace {

int fail(dftpav_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

#define CU_TRY(h, call)                                                                                  \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(h, DFTPAV_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));              \
  } while (0)

// lbfgs_optimize's own parameter checks, in its order (lbfgs.hpp:456-501): the code every solve of the batch would return
int lbfgs_param_status(const dftpav_params& p) {
  if (p.lbfgs_mem_size <= 0) return DFTPAV_LBFGSERR_INVALID_MEMSIZE;
  if (p.lbfgs_g_epsilon < 0.0) return DFTPAV_LBFGSERR_INVALID_GEPSILON;
  if (p.lbfgs_past < 0) return DFTPAV_LBFGSERR_INVALID_TESTPERIOD;
  if (p.lbfgs_delta < 0.0) return DFTPAV_LBFGSERR_INVALID_DELTA;
  if (p.lbfgs_min_step < 0.0) return DFTPAV_LBFGSERR_INVALID_MINSTEP;
  if (p.lbfgs_max_step < p.lbfgs_min_step) return DFTPAV_LBFGSERR_INVALID_MAXSTEP;
  if (!(p.lbfgs_f_dec_coeff > 0.0 && p.lbfgs_f_dec_coeff < 1.0)) return DFTPAV_LBFGSERR_INVALID_FDECCOEFF;
  if (!(p.lbfgs_s_curv_coeff < 1.0 && p.lbfgs_s_curv_coeff > p.lbfgs_f_dec_coeff)) return DFTPAV_LBFGSERR_INVALID_SCURVCOEFF;
  if (!(p.lbfgs_machine_prec > 0.0)) return DFTPAV_LBFGSERR_INVALID_MACHINEPREC;
  if (p.lbfgs_max_linesearch <= 0) return DFTPAV_LBFGSERR_INVALID_MAXLINESEARCH;
  return 0;
}

// Validation shared by dftpav_batch_dims / upload.  Returns 0 or an API error (malformed arrays, compiled-in limits);
// fills dims.  A problem that violates a precondition of OptimizeTrajectory itself (traj_optimizer.cpp:26-48: a segment
// with a single piece, initTs < mini_T) does NOT fail the batch: the reference rejects that ONE call (`return false`), so
// the problem is flagged in rejected[p] and reported with status DFTPAV_PROBLEM_REJECTED / success 0.
int validate(const dftpav_params* prm, const dftpav_batch* b, dftpav_dims* dims, std::string* why,
             std::vector<unsigned char>* rejected) {
  auto bad = [&](int code, const std::string& s) {
    if (why) *why = s;
    return code;
  };
  if (!prm || !b) return bad(DFTPAV_ERR_INVALID_ARG, "null params or batch");
  if (prm->abi_version != DFTPAV_ABI_VERSION) return bad(DFTPAV_ERR_INVALID_ARG, "abi_version mismatch");
  if (b->B < 0) return bad(DFTPAV_ERR_INVALID_ARG, "B < 0");
  if (b->B > 0 && (!b->seg_ptr || !b->piece_num || !b->singul || !b->ini_state || !b->fin_state || !b->init_T || !b->inner_pts))
    return bad(DFTPAV_ERR_INVALID_ARG, "null array in batch");
  // planes == NULL: the corridor comes from dftpav_corridor_rectangles (device-resident, 4 planes per sample)
  if (b->B > 0 && !b->planes && (b->plane_cnt || b->uniform_H != 4))
    return bad(DFTPAV_ERR_INVALID_ARG, "planes == NULL needs uniform_H == 4 and plane_cnt == NULL (generated corridor)");
  if (prm->traj_resolution < 1 || prm->des_traj_resolution < 1) return bad(DFTPAV_ERR_INVALID_ARG, "resolution < 1");
  if (prm->lbfgs_past > 8) return bad(DFTPAV_ERR_UNSUPPORTED, "lbfgs_past > 8");
  if (prm->lbfgs_mem_size > 2 * 12 * kStageLd)  // alpha array lives in the sweep staging buffer
    return bad(DFTPAV_ERR_UNSUPPORTED, "lbfgs_mem_size > 408");
  if (!b->plane_cnt && b->uniform_H < 0) return bad(DFTPAV_ERR_INVALID_ARG, "uniform_H < 0");
  if (b->B > 0 && b->seg_ptr[0] != 0) return bad(DFTPAV_ERR_INVALID_ARG, "seg_ptr[0] != 0");
  if (rejected) rejected->assign((size_t)b->B, 0);
  int n_stride = 0, row_stride = 0;
  long long samples = 0;
  for (int p = 0; p < b->B; p++) {
    const int g0 = b->seg_ptr[p], g1 = b->seg_ptr[p + 1];
    const int S = g1 - g0;
    if (S < 1) return bad(DFTPAV_ERR_INVALID_ARG, "problem " + std::to_string(p) + ": no segments (seg_ptr not increasing)");
    if (S > kMaxSeg) return bad(DFTPAV_ERR_UNSUPPORTED, "problem " + std::to_string(p) + ": more than 4 gear segments");
    int n = 0, rows = 0;
    bool rej = false;
    for (int g = g0; g < g1; g++) {
      const int N = b->piece_num[g];
      if (N < 1) return bad(DFTPAV_ERR_INVALID_ARG, "problem " + std::to_string(p) + ": piece_num < 1");
      if (N > kMaxPieces) return bad(DFTPAV_ERR_UNSUPPORTED, "problem " + std::to_string(p) + ": more than 64 pieces");
      if (N < 2) rej = true;                            // traj_optimizer.cpp:38-41 "There is only a piece?"
      if (!(b->init_T[g] >= prm->mini_T)) rej = true;   // traj_optimizer.cpp:30-33
      n += 2 * (N - 1);
      rows += 6 * N;
      samples += samples_of(N, prm->traj_resolution, prm->des_traj_resolution);
    }
    if (rej && rejected) (*rejected)[p] = 1;
    n += S + 3 * (S - 1);
    n_stride = std::max(n_stride, n);
    row_stride = std::max(row_stride, rows);
  }
  if (b->plane_cnt)
    for (long long q = 0; q < samples; q++)
      if (b->plane_cnt[q] < 0) return bad(DFTPAV_ERR_INVALID_ARG, "negative plane_cnt");
  // obstacle CSR arrays (setSurroundTrajs, traj_optimizer.h:108): checked here so that nothing downstream dereferences them blindly
  if (b->obs_ptr && b->B > 0) {
    if (b->obs_ptr[0] != 0) return bad(DFTPAV_ERR_INVALID_ARG, "obs_ptr[0] != 0");
    for (int p = 0; p < b->B; p++)
      if (b->obs_ptr[p + 1] < b->obs_ptr[p]) return bad(DFTPAV_ERR_INVALID_ARG, "obs_ptr not monotone");
    if (b->obs_ptr[b->B] > 0) {
      const int U = b->n_obs_pool;
      if (U <= 0 || !b->obs_idx || !b->obs_start_time || !b->obs_duration || !b->obs_piece_ptr || !b->obs_piece_dur ||
          !b->obs_coeffs)
        return bad(DFTPAV_ERR_INVALID_ARG, "obstacle arrays missing");
      if (b->obs_piece_ptr[0] != 0) return bad(DFTPAV_ERR_INVALID_ARG, "obs_piece_ptr[0] != 0");
      for (int u = 0; u < U; u++)
        if (b->obs_piece_ptr[u + 1] - b->obs_piece_ptr[u] < 1)
          return bad(DFTPAV_ERR_INVALID_ARG, "obstacle " + std::to_string(u) + " has no pieces");
      for (int k = 0; k < b->obs_ptr[b->B]; k++)
        if (b->obs_idx[k] < 0 || b->obs_idx[k] >= U) return bad(DFTPAV_ERR_INVALID_ARG, "obs_idx out of range");
    }
  }
  if (dims) {
    dims->B = b->B;
    dims->n_stride = n_stride;
    dims->row_stride = row_stride;
    dims->seg_total = b->B ? b->seg_ptr[b->B] : 0;
    dims->sample_total = samples;
  }
  return DFTPAV_OK;
}

void fill_params_dev(const dftpav_params& p, ParamsDev& d) {
  d.wei_obs = p.wei_obs;
  d.wei_surround = p.wei_surround;
  d.wei_feas = p.wei_feas;
  d.wei_time = p.wei_time;
  d.surround_clearance = p.surround_clearance;
  d.d_min = p.surround_clearance + std::log(8.0) / p.lse_alpha;  // traj_optimizer.cpp:1336
  d.max_vel[0] = p.max_forward_vel, d.max_vel[1] = p.max_backward_vel;
  d.max_cur[0] = p.max_forward_cur, d.max_cur[1] = p.max_backward_cur;
  d.max_acc[0] = p.max_forward_acc, d.max_acc[1] = p.max_backward_acc;
  d.non_sinv = p.non_sinv;
  d.mini_T = p.mini_T;
  d.help_eps = p.help_eps;
  d.min_speed = p.min_speed;
  d.smooth_eps = p.smooth_eps;
  d.cur_weight_scale = p.cur_weight_scale;
  d.lse_alpha = p.lse_alpha;
  // vehicle box inflated by 2*half_margin (traj_optimizer.cpp:1749-1768), vertices clockwise from front-left
  const double width = p.veh_width + 2 * p.half_margin, length = p.veh_length + 2 * p.half_margin;
  d.cull_dist = length * p.obs_cull_factor;
  d.cost_limit = p.cost_limit;
  d.le[0][0] = p.veh_d_cr + length / 2.0, d.le[0][1] = width / 2.0;
  d.le[1][0] = p.veh_d_cr + length / 2.0, d.le[1][1] = -width / 2.0;
  d.le[2][0] = p.veh_d_cr - length / 2.0, d.le[2][1] = -width / 2.0;
  d.le[3][0] = p.veh_d_cr - length / 2.0, d.le[3][1] = width / 2.0;
  {  // positiveSmoothedL1 constants, same expressions as traj_optimizer.cpp:785-790
    const double pe = p.smooth_eps;
    d.l1_half = 0.5 * pe;
    d.l1_f3c = 1.0 / (pe * pe);
    d.l1_f4c = -0.5 * d.l1_f3c / pe;
    d.l1_d2c = 3.0 * d.l1_f3c;
    d.l1_d3c = 4.0 * d.l1_f4c;
  }
  d.gear_opt = p.gear_opt;
  d.mem_size = p.lbfgs_mem_size;
  d.past = p.lbfgs_past;
  d.max_iterations = p.lbfgs_max_iterations;
  d.max_linesearch = p.lbfgs_max_linesearch;
  d.delta = p.lbfgs_delta;
  d.g_epsilon = p.lbfgs_g_epsilon;
  d.min_step = p.lbfgs_min_step;
  d.max_step = p.lbfgs_max_step;
  d.f_dec_coeff = p.lbfgs_f_dec_coeff;
  d.s_curv_coeff = p.lbfgs_s_curv_coeff;
  d.cautious_factor = p.lbfgs_cautious_factor;
  d.machine_prec = p.lbfgs_machine_prec;
}

// Groups are recycled across uploads so that their device buffers (the 231 MB plane array at config 3) are allocated
// once: a planner calls dftpav_solve every cycle and cudaMalloc/cudaFree of that size costs milliseconds.
void recycle_groups(dftpav_handle* h) {
  for (Group* g : h->groups) {
    g->probs.clear();
    h->group_pool.push_back(g);
  }
  h->groups.clear();
}
void free_groups(dftpav_handle* h) {
  recycle_groups(h);
  for (Group* g : h->group_pool) {
    g->d_prob_id.release();
    g->d_planes.release();
    g->d_lu.release();
    g->d_order.release();
    g->d_clist.release();
    g->d_cq.release();
    delete g;
  }
  h->group_pool.clear();
}

// host twin of prepare_states_kernel's clamp (traj_optimizer.cpp:55-76)
void clamp_state(double* st, double max_vel, double max_acc) {
  for (int c = 2; c <= 4; c += 2) {
    const double lim = (c == 2) ? max_vel : max_acc;
    const double vx = st[c], vy = st[c + 1];
    const double z = vx * vx + vy * vy;
    if (std::sqrt(z) >= lim) {
      double nx = vx, ny = vy;
      if (z > 0.0) {
        const double sz = std::sqrt(z);
        nx = vx / sz;
        ny = vy / sz;
      }
      st[c] = nx * (lim - 1.0e-2);
      st[c + 1] = ny * (lim - 1.0e-2);
    }
  }
}

// Initial decision vector of problem p exactly as traj_optimizer.cpp:96-115 (inner points, RealT2VirtualT 361-369, gear
// positions and atan2 headings from the clamped final states 55-76).  Host code on purpose: x0 must carry glibc's atan2,
// like the reference (the CUDA math library's atan2 differs from it by an ulp now and then).  Returns the inner-point
// offset of the next problem.
long long pack_x0(const dftpav_params* prm, const dftpav_batch* b, int p, long long inner, int n_stride, double* x) {
  std::fill(x, x + n_stride, 0.0);
  const int g0 = b->seg_ptr[p], g1 = b->seg_ptr[p + 1], S = g1 - g0;
  int off = 0;
  for (int g = g0; g < g1; g++) {
    const int np = 2 * (b->piece_num[g] - 1);
    std::memcpy(x + off, b->inner_pts + inner, sizeof(double) * np);
    off += np;
    inner += np;
  }
  for (int s = 0; s < S; s++) {  // RealT2VirtualT, traj_optimizer.cpp:361-369
    const double RT = b->init_T[g0 + s], mini_T = prm->mini_T;
    x[off + s] = RT > 1.0 + mini_T ? (std::sqrt(2.0 * RT - 1.0 - 2 * mini_T) - 1.0)
                                   : (1.0 - std::sqrt(2.0 / (RT - mini_T) - 1.0));
  }
  off += S;
  for (int s = 0; s < S - 1; s++) {
    double st[6];
    std::memcpy(st, b->fin_state + 6 * (g0 + s), sizeof(st));
    const bool fwd = b->singul[g0 + s] > 0;
    clamp_state(st, fwd ? prm->max_forward_vel : prm->max_backward_vel, fwd ? prm->max_forward_acc : prm->max_backward_acc);
    x[off + 2 * s] = st[0];
    x[off + 2 * s + 1] = st[1];
    x[off + 2 * (S - 1) + s] = std::atan2(st[3], st[2]);
  }
  return inner;
}

}  // namespace

extern "C" {

const char* dftpav_version(void) { return "dftpav_b200 0.1 (sm_100a, abi 1)"; }

void dftpav_default_params(dftpav_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->abi_version = DFTPAV_ABI_VERSION;
  p->traj_resolution = 16;
  p->des_traj_resolution = 32;
  p->gear_opt = 1;
  p->wei_obs = 1000.0;
  p->wei_surround = 5000.0;
  p->wei_feas = 2500.0;
  p->wei_time = 500.0;
  p->surround_clearance = 0.4;
  p->half_margin = 0.15;
  p->veh_width = 1.90;
  p->veh_length = 4.88;
  p->veh_d_cr = 1.015;
  p->max_forward_vel = 5.0;
  p->max_backward_vel = 2.0;
  p->max_forward_cur = 1.0;
  p->max_backward_cur = 1.0;
  p->max_forward_acc = 8.0;
  p->max_backward_acc = 4.0;
  p->non_sinv = 0.24;
  p->mini_T = 0.1;
  p->help_eps = 0.0;
  p->min_speed = 1.0e-4;
  p->smooth_eps = 1.0e-4;
  p->cur_weight_scale = 10.0;
  p->lse_alpha = 100.0;
  p->obs_cull_factor = 1.5;
  p->cost_limit = 50000.0;
  p->lbfgs_mem_size = 256;
  p->lbfgs_past = 3;
  p->lbfgs_max_iterations = 12000;
  p->lbfgs_max_linesearch = 64;
  p->lbfgs_delta = 1.0e-4;
  p->lbfgs_g_epsilon = 1.0e-16;
  p->lbfgs_min_step = 1.0e-32;
  p->lbfgs_max_step = 1.0e20;
  p->lbfgs_f_dec_coeff = 1.0e-4;
  p->lbfgs_s_curv_coeff = 0.9;
  p->lbfgs_cautious_factor = 1.0e-6;
  p->lbfgs_machine_prec = 1.0e-16;
}

int dftpav_batch_dims(const dftpav_params* prm, const dftpav_batch* batch, dftpav_dims* dims) {
  return validate(prm, batch, dims, nullptr, nullptr);
}

int dftpav_batch_check(const dftpav_params* prm, const dftpav_batch* batch, int32_t* rejected) {
  std::vector<unsigned char> rej;
  const int rc = validate(prm, batch, nullptr, nullptr, &rej);
  if (rc == DFTPAV_OK && rejected)
    for (int p = 0; p < batch->B; p++) rejected[p] = rej[p];
  return rc;
}

int dftpav_initial_x(const dftpav_params* prm, const dftpav_batch* b, double* x0) {
  dftpav_dims dims;
  int rc = validate(prm, b, &dims, nullptr, nullptr);
  if (rc) return rc;
  if (!x0) return DFTPAV_ERR_INVALID_ARG;
  long long inner = 0;
  for (int p = 0; p < b->B; p++) inner = pack_x0(prm, b, p, inner, dims.n_stride, x0 + (size_t)p * dims.n_stride);
  return DFTPAV_OK;
}

int dftpav_create(const dftpav_params* prm, int device, dftpav_handle** out) {
  if (!prm || !out) return DFTPAV_ERR_INVALID_ARG;
  *out = nullptr;
  if (prm->abi_version != DFTPAV_ABI_VERSION) return DFTPAV_ERR_INVALID_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return DFTPAV_ERR_NO_DEVICE;
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) return DFTPAV_ERR_NO_DEVICE;
  }
  if (device >= count) return DFTPAV_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DFTPAV_ERR_NO_DEVICE;
  if (prop.major != 10) return DFTPAV_ERR_NO_DEVICE;  // kernels are built for sm_100a only
  dftpav_handle* h = new dftpav_handle();
  h->prm = *prm;
  fill_params_dev(*prm, h->pd);
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  // the solver stream outranks the plane-copy stream: where the hardware arbitrates by priority, the small uploads and the
  // kernels of a call do not queue behind its own 233 MB of corridor planes
  int prio_lo = 0, prio_hi = 0;
  cudaSetDevice(device);
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->own_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
    delete h;
    return DFTPAV_ERR_CUDA;
  }
  h->stream = h->own_stream;
  for (auto& e : h->ev) cudaEventCreate(&e);
  if (cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming) != cudaSuccess ||
      cudaMallocHost(&h->h_epoch, sizeof(int) * kMaxChunks) != cudaSuccess) {
    dftpav_destroy(h);
    return DFTPAV_ERR_CUDA;
  }
  *out = h;
  return DFTPAV_OK;
}

int dftpav_destroy(dftpav_handle* h) {
  if (!h) return DFTPAV_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  free_groups(h);
  DevBuf* bufs[] = {&h->seg_base, &h->t_now, &h->obs_ptr, &h->obs_idx, &h->singul, &h->ini_raw, &h->fin_raw, &h->ini,
                    &h->fin, &h->samp_base, &h->plane_ptr, &h->planes_raw,
                    &h->obs_start, &h->obs_dur, &h->obs_piece_ptr, &h->obs_piece_dur, &h->obs_coeffs, &h->obs_uniform, &h->x0, &h->xin,
                    &h->fout, &h->gout, &h->o_x, &h->o_cost, &h->o_status, &h->o_success, &h->o_iters, &h->o_evals,
                    &h->o_coeffs, &h->o_dT, &h->counter, &h->hist_s, &h->hist_y, &h->hist_ys, &h->trace, &h->ready, &h->laid, &h->seg_prob, &h->seg_row0, &h->seg_N,
                    &h->seg_piece0, &h->q_seg, &h->q_t, &h->q_out, &h->q_cnt, &h->map_data, &h->map_dl, &h->cor_states, &h->cor_cs, &h->fit_pp, &h->fit_dT, &h->fit_head, &h->fit_tail,
                    &h->fit_inner, &h->fit_iptr, &h->fit_lu, &h->fit_luoff, &h->fit_scratch, &h->fit_coeffs, &h->fe_arena, &h->fe_out};
  for (DevBuf* b : bufs) b->release();
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->h_epoch) cudaFreeHost(h->h_epoch);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return DFTPAV_OK;
}

int dftpav_set_stream(dftpav_handle* h, void* cuda_stream) {
  if (!h) return DFTPAV_ERR_INVALID_ARG;
  h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return DFTPAV_OK;
}

int dftpav_get_dims(const dftpav_handle* h, dftpav_dims* dims) {
  if (!h || !dims) return DFTPAV_ERR_INVALID_ARG;
  if (!h->uploaded) return DFTPAV_ERR_STATE;
  *dims = h->dims;
  return DFTPAV_OK;
}

}  // extern "C" (reopened below)

namespace {

// Everything of an upload except the corridor planes themselves (the dominant bytes): validation, prefixes, shape groups,
// MINCO tables, small arrays, boundary-state / x0 kernel, launch configurations, history.  The planes follow through
// issue_plane_copies() - before the solve for dftpav_upload_batch, AFTER the solve kernels have been launched for
// dftpav_solve, so that the copy runs under the solve (the kernels wait per upload chunk, layout_problem()).
int issue_plane_copies(dftpav_handle* h, const dftpav_batch* b, cudaStream_t cs);

// early_copy_stream != nullptr (dftpav_solve): the plane copies are queued on that stream as soon as the plane prefixes are
// known - BEFORE the host-side packing of x0, the MINCO tables and the small uploads (~0.4 ms of host work), which then run
// under the copy
int upload_impl(dftpav_handle* h, const dftpav_batch* b, int n_chunks_wanted, cudaStream_t early_copy_stream = nullptr) {
  std::string why;
  dftpav_dims dims;
  std::vector<unsigned char> rejected;
  int rc = validate(&h->prm, b, &dims, &why, &rejected);
  if (rc) return fail(h, rc, why);
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  h->uploaded = false;
  h->solved = false;
  recycle_groups(h);
  dims.n_stride = std::max(dims.n_stride, h->min_n_stride);
  dims.row_stride = std::max(dims.row_stride, h->min_row_stride);
  h->dims = dims;
  h->param_status = lbfgs_param_status(h->prm);
  std::memset(&h->stats, 0, sizeof(h->stats));
  CU_TRY(h, cudaEventRecord(h->ev[0], st));
  const int B = b->B, G = dims.seg_total;
  const int K = h->prm.traj_resolution, Kd = h->prm.des_traj_resolution;
  const long long Q = dims.sample_total;
  long long h2d = 0;

  // ---- per-problem prefixes and shape groups
  std::vector<int> seg_base(B);
  std::vector<long long>& samp_base = h->h_samp_base;
  samp_base.assign((size_t)B + 1, 0);
  std::map<std::vector<int>, Group*> by_shape;
  long long samp = 0, inner = 0;
  h->h_seg_ptr.assign(b->seg_ptr, b->seg_ptr + B + 1);
  h->h_piece_num.assign(b->piece_num, b->piece_num + G);
  h->h_n.resize(B);
  h->h_status0.assign(B, 0);
  h->n_rejected = 0;
  Group* last_grp = nullptr;  // batches are mostly runs of one shape: the map is consulted only when the shape changes
  int last_g0 = 0, last_S = -1;
  for (int p = 0; p < B; p++) {
    const int g0 = b->seg_ptr[p], g1 = b->seg_ptr[p + 1];
    seg_base[p] = g0;
    samp_base[p] = samp;
    int n = 0;
    for (int g = g0; g < g1; g++) {
      samp += samples_of(b->piece_num[g], K, Kd);
      inner += 2LL * (b->piece_num[g] - 1);
      n += 2 * (b->piece_num[g] - 1);
    }
    n += (g1 - g0) + 3 * (g1 - g0 - 1);
    h->h_n[p] = n;
    if (rejected[p]) {  // the reference returns false for this call alone (traj_optimizer.cpp:26-48)
      h->h_status0[p] = DFTPAV_PROBLEM_REJECTED;
      h->n_rejected++;
      continue;
    }
    if (h->param_status) h->h_status0[p] = h->param_status;
    if (last_grp && g1 - g0 == last_S && std::equal(b->piece_num + g0, b->piece_num + g1, b->piece_num + last_g0)) {
      last_grp->probs.push_back(p);
      continue;
    }
    std::vector<int> key(b->piece_num + g0, b->piece_num + g1);
    Group*& grp = by_shape[key];
    if (!grp) {
      if (!h->group_pool.empty()) {
        grp = h->group_pool.back();
        h->group_pool.pop_back();
      } else {
        grp = new Group();
      }
      h->groups.push_back(grp);
    }
    grp->probs.push_back(p);
    last_grp = grp, last_g0 = g0, last_S = g1 - g0;
  }
  samp_base[B] = samp;
  // plane prefix (only for ragged corridors)
  std::vector<long long>& plane_ptr = h->h_plane_ptr;
  plane_ptr.clear();
  if (b->plane_cnt) {
    plane_ptr.resize(Q + 1);
    plane_ptr[0] = 0;
    for (long long q = 0; q < Q; q++) plane_ptr[q + 1] = plane_ptr[q] + b->plane_cnt[q];
    h->total_planes = plane_ptr[Q];
  } else {
    h->total_planes = Q * b->uniform_H;
  }
  h->uniform_H = b->plane_cnt ? -1 : b->uniform_H;
  // algorithmic bytes per evaluation of each problem
  h->h_evalbytes.resize(B);
  for (int p = 0; p < B; p++) {
    const int g0 = b->seg_ptr[p], g1 = b->seg_ptr[p + 1];
    const long long q0 = samp_base[p], q1 = samp_base[p + 1];
    const long long planes = b->plane_cnt ? plane_ptr[q1] - plane_ptr[q0] : (q1 - q0) * b->uniform_H;
    long long obs = 0;
    if (b->obs_ptr)
      for (int k = b->obs_ptr[p]; k < b->obs_ptr[p + 1]; k++) {
        const int u = b->obs_idx[k];
        obs += b->obs_piece_ptr[u + 1] - b->obs_piece_ptr[u];
      }
    h->h_evalbytes[p] = (int)(16LL * h->h_n[p] + 32LL * planes + 96LL * (g1 - g0) + 104LL * obs);
  }
  h->has_obs = (b->obs_ptr != nullptr) && B > 0 && b->obs_ptr[B] > 0;
  if (h->has_obs && strict_build_host())
    return fail(h, DFTPAV_ERR_UNSUPPORTED, "the strict verification build does not evaluate moving obstacles (exp/log)");

  // ---- the corridor planes first: buffer, arrival flags, and - for the overlapped call - the copies themselves
  h->planes_from_corridor = (b->planes == nullptr) && B > 0;
  if (h->planes_from_corridor) {
    if (h->corridor_n != Q)
      return fail(h, DFTPAV_ERR_STATE, "planes == NULL but dftpav_corridor_rectangles(keep) has not produced one polygon per sample of this batch");
  } else {
    h->corridor_n = -1;  // the buffer is about to be overwritten by uploaded planes
    CU_TRY(h, h->planes_raw.ensure(sizeof(double) * 4 * (size_t)h->total_planes));
  }
  // ---- arrival flags of the plane upload and per-problem layout flags (epoch-stamped: nothing to clear between uploads)
  h->epoch += 1;
  h->n_chunks = std::max(1, std::min(std::min(n_chunks_wanted, kMaxChunks), B / 64));
  h->chunk = std::max(1, (B + h->n_chunks - 1) / h->n_chunks);
  h->n_chunks = B > 0 ? (B + h->chunk - 1) / h->chunk : 1;
  if (h->ready.cap < sizeof(int) * kMaxChunks) {
    CU_TRY(h, h->ready.ensure(sizeof(int) * kMaxChunks));
    CU_TRY(h, cudaMemsetAsync(h->ready.p, 0, sizeof(int) * kMaxChunks, st));
  }
  if (h->laid.cap < sizeof(int) * (size_t)std::max(1, B)) {
    CU_TRY(h, h->laid.ensure(sizeof(int) * (size_t)std::max(1, B)));
    CU_TRY(h, cudaMemsetAsync(h->laid.p, 0, h->laid.cap, st));
  }
  for (int c = 0; c < kMaxChunks; c++) h->h_epoch[c] = h->epoch;
  if (early_copy_stream) {
    // the copy stream starts after everything already queued on the main stream (previous results are read, the old plane
    // buffer is free, the flag arrays are cleared)
    CU_TRY(h, cudaEventRecord(h->ev_copy, st));
    CU_TRY(h, cudaStreamWaitEvent(early_copy_stream, h->ev_copy, 0));
    rc = issue_plane_copies(h, b, early_copy_stream);
    if (rc) return rc;
  }

  // ---- raw uploads (everything but the planes)
  auto up = [&](DevBuf& buf, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = buf.ensure(bytes);
    if (e != cudaSuccess) return e;
    if (bytes) e = cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, st);
    h2d += (long long)bytes;
    return e;
  };
  {  // segment directory of the trajectory evaluator
    std::vector<int> sp(G), sr(G);
    std::vector<long long> pc(G);
    long long pieces = 0;
    for (int p = 0; p < B; p++) {
      int row = 0;
      for (int g = b->seg_ptr[p]; g < b->seg_ptr[p + 1]; g++) {
        sp[g] = p;
        sr[g] = row;
        pc[g] = pieces;
        row += 6 * b->piece_num[g];
        pieces += b->piece_num[g];
      }
    }
    h->pieces_total = pieces;
    CU_TRY(h, up(h->seg_prob, sp.data(), sizeof(int) * G));
    CU_TRY(h, up(h->seg_row0, sr.data(), sizeof(int) * G));
    CU_TRY(h, up(h->seg_piece0, pc.data(), sizeof(long long) * G));
    CU_TRY(h, up(h->seg_N, b->piece_num, sizeof(int) * G));
  }
  CU_TRY(h, up(h->seg_base, seg_base.data(), sizeof(int) * B));
  CU_TRY(h, up(h->samp_base, samp_base.data(), sizeof(long long) * (B + 1)));
  CU_TRY(h, up(h->singul, b->singul, sizeof(int) * G));
  CU_TRY(h, up(h->ini_raw, b->ini_state, sizeof(double) * 6 * G));
  CU_TRY(h, up(h->fin_raw, b->fin_state, sizeof(double) * 6 * G));
  {  // x0 packed on the host (pack_x0), rejected problems all zero
    std::vector<double>& x0 = h->h_x0;
    x0.assign((size_t)B * dims.n_stride, 0.0);
    long long in_off = 0;
    for (int p = 0; p < B; p++) {
      if (rejected[p]) {
        for (int g = b->seg_ptr[p]; g < b->seg_ptr[p + 1]; g++) in_off += 2LL * (b->piece_num[g] - 1);
        continue;
      }
      in_off = pack_x0(&h->prm, b, p, in_off, dims.n_stride, x0.data() + (size_t)p * dims.n_stride);
    }
    CU_TRY(h, up(h->x0, x0.data(), sizeof(double) * x0.size()));
  }
  if (b->plane_cnt) CU_TRY(h, up(h->plane_ptr, plane_ptr.data(), sizeof(long long) * (Q + 1)));
  std::vector<double> tnow(B, 0.0);
  if (b->t_now) std::copy(b->t_now, b->t_now + B, tnow.begin());
  CU_TRY(h, up(h->t_now, tnow.data(), sizeof(double) * B));
  if (h->has_obs) {
    const int U = b->n_obs_pool;
    const int NP = b->obs_piece_ptr[U];
    CU_TRY(h, up(h->obs_ptr, b->obs_ptr, sizeof(int) * (B + 1)));
    CU_TRY(h, up(h->obs_idx, b->obs_idx, sizeof(int) * b->obs_ptr[B]));
    CU_TRY(h, up(h->obs_start, b->obs_start_time, sizeof(double) * U));
    CU_TRY(h, up(h->obs_dur, b->obs_duration, sizeof(double) * U));
    CU_TRY(h, up(h->obs_piece_ptr, b->obs_piece_ptr, sizeof(int) * (U + 1)));
    CU_TRY(h, up(h->obs_piece_dur, b->obs_piece_dur, sizeof(double) * NP));
    CU_TRY(h, up(h->obs_coeffs, b->obs_coeffs, sizeof(double) * 12 * NP));
    // uniform piece duration per obstacle (enables the load-free exact locate loop in the kernel)
    std::vector<double> uni(U);
    for (int u = 0; u < U; u++) {
      const int p0 = b->obs_piece_ptr[u], p1 = b->obs_piece_ptr[u + 1];
      double d0 = b->obs_piece_dur[p0];
      for (int q = p0 + 1; q < p1; q++)
        if (b->obs_piece_dur[q] != d0) d0 = std::nan("");
      uni[u] = d0;
    }
    CU_TRY(h, up(h->obs_uniform, uni.data(), sizeof(double) * U));
  }
  CU_TRY(h, h->ini.ensure(sizeof(double) * 6 * G));
  CU_TRY(h, h->fin.ensure(sizeof(double) * 6 * G));
  // outputs
  CU_TRY(h, h->o_x.ensure(sizeof(double) * (size_t)B * dims.n_stride));
  CU_TRY(h, h->o_cost.ensure(sizeof(double) * B));
  CU_TRY(h, h->o_status.ensure(sizeof(int) * B));
  CU_TRY(h, h->o_success.ensure(sizeof(int) * B));
  CU_TRY(h, h->o_iters.ensure(sizeof(int) * B));
  CU_TRY(h, h->o_evals.ensure(sizeof(int) * B));
  CU_TRY(h, h->o_coeffs.ensure(sizeof(double) * (size_t)B * dims.row_stride * 2));
  CU_TRY(h, h->o_dT.ensure(sizeof(double) * G));
  CU_TRY(h, h->counter.ensure(sizeof(int) * 64));
  if (h->n_rejected > 0 || h->param_status) {
    // problems no kernel touches: status from the host, everything else zero
    CU_TRY(h, cudaMemsetAsync(h->o_x.p, 0, sizeof(double) * (size_t)B * dims.n_stride, st));
    CU_TRY(h, cudaMemsetAsync(h->o_cost.p, 0, sizeof(double) * B, st));
    CU_TRY(h, cudaMemsetAsync(h->o_success.p, 0, sizeof(int) * B, st));
    CU_TRY(h, cudaMemsetAsync(h->o_iters.p, 0, sizeof(int) * B, st));
    CU_TRY(h, cudaMemsetAsync(h->o_evals.p, 0, sizeof(int) * B, st));
    CU_TRY(h, cudaMemsetAsync(h->o_coeffs.p, 0, sizeof(double) * (size_t)B * dims.row_stride * 2, st));
    CU_TRY(h, cudaMemsetAsync(h->o_dT.p, 0, sizeof(double) * G, st));
    CU_TRY(h, cudaMemcpyAsync(h->o_status.p, h->h_status0.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
  }

  // ---- per-group tables, state kernel, launch configuration
  const int VR = virtual_ranks_host();
  const bool strict = strict_build_host() != 0;
  int max_slots = 0, max_npad = 0;
  int launches = 0;
  for (Group* grp : h->groups) {
    ShapeDev& sh = grp->sh;
    std::memset(&sh, 0, sizeof(sh));
    const int p0 = grp->probs[0];
    const int g0 = b->seg_ptr[p0], g1 = b->seg_ptr[p0 + 1];
    sh.S = g1 - g0;
    sh.K = K;
    sh.Kd = Kd;
    sh.has_obs = h->has_obs ? 1 : 0;
    const int td = (Kd + kTile - 1) / kTile, tk = (K + kTile - 1) / kTile;
    // regular sweep passes touch disjoint pieces when every inner piece is one tile and the first / last piece fit one
    // pass (two tiles); the tiles of the last piece are aligned to a pass by an empty tile where needed
    sh.helpable = (tk == 1 && td <= 2) ? 1 : 0;
    std::map<int, int> lu_of_N;
    int rows = 0, var = 0, sampoff = 0, luoff = 0, pieces = 0;
    for (int s = 0; s < sh.S; s++) {
      const int N = b->piece_num[g0 + s];
      sh.N[s] = N;
      sh.row0[s] = rows;
      sh.var0[s] = var;
      sh.samp0[s] = sampoff;
      sh.tpad[s] = (sh.helpable && td == 2 && ((td + (N - 2) * tk) & 1)) ? 1 : 0;
      sh.ntile[s] = 2 * td + (N - 2) * tk + sh.tpad[s];
      rows += 6 * N;
      var += 2 * (N - 1);
      pieces += N;
      // storage slots: 32 per sweep pass (two 16-sample tiles per regular pass, then the piece-end samples)
      sampoff += 32 * ((sh.ntile[s] + 1) / 2 + (N + 31) / 32);
      auto it = lu_of_N.find(N);
      if (it == lu_of_N.end()) {
        // table layout per distinct N: [8 zero columns][6N packed columns][8 zero columns] so that the software-
        // pipelined update-form substitutions may address columns j-7 .. j+7 without bounds checks
        lu_of_N[N] = luoff + 8 * kLuLd;
        sh.lu_off[s] = luoff + 8 * kLuLd;
        luoff += (6 * N + 16) * kLuLd;
      } else {
        sh.lu_off[s] = it->second;
      }
    }
    sh.rows = rows;
    sh.pieces = pieces;
    sh.n = var + sh.S + 3 * (sh.S - 1);
    sh.npad = (sh.n + 3) / 4 * 4;
    sh.Mpad = sampoff;
    sh.lu_doubles = luoff;
    sh.ws_doubles = ws_doubles_host(sh.rows, sh.npad, sh.S, sh.has_obs);
    // launch shape: one CTA per SM (or per problem), W warps; a CTA with fewer problems than warps fills up with
    // helper-only warps (up to VR warps per trajectory), so small batches are swept by teams from the first evaluation
    const int count = (int)grp->probs.size();
    const size_t cap = h->smem_optin > 2048 ? h->smem_optin - 1024 : 0;
    const int grid = std::max(1, std::min(count, h->num_sms));
    const int per_cta = (count + grid - 1) / grid;
    const int wantW = std::max(1, std::min(max_warps_host(), per_cta * ((sh.helpable && !strict) ? VR : 1)));
    // dense MINCO tables instead of the LU tables when they fit beside the workspaces of the wanted warps
    {
      size_t dd = 0;
      bool cols_ok = true;
      for (auto& kv : lu_of_N) dd += dense_doubles(kv.first), cols_ok = cols_ok && kv.first + 5 <= 64;
      const bool want_dense = !strict && !h->force_banded;
      if (want_dense && cols_ok && sizeof(double) * (dd + (size_t)sh.ws_doubles * wantW) <= cap) {
        sh.dense_minco = 1;
        size_t off = 0;
        std::map<int, int> off_of_N;
        for (auto& kv : lu_of_N) off_of_N[kv.first] = (int)off, off += dense_doubles(kv.first);
        for (int s = 0; s < sh.S; s++) sh.lu_off[s] = off_of_N[sh.N[s]];
        sh.lu_doubles = (int)dd;
      }
    }
    // padded plane count of the group
    int Hmax = 0;
    if (b->plane_cnt) {
      for (int p : grp->probs)
        for (long long q = samp_base[p]; q < samp_base[p + 1]; q++) Hmax = std::max(Hmax, (int)b->plane_cnt[q]);
    } else {
      Hmax = b->uniform_H;
    }
    sh.H = (Hmax + 3) / 4 * 4;  // padded to the register chunk of the sweep (null planes never violate)
    // table block
    std::vector<double> lu((size_t)sh.lu_doubles, 0.0);
    sh.band_masks_ok = 1;
    if (sh.dense_minco) {
      size_t off = 0;
      for (auto& kv : lu_of_N) {
        const std::vector<double>& t = dense_minco_table(kv.first);
        std::copy(t.begin(), t.end(), lu.begin() + off);
        off += t.size();
      }
    } else {
      for (auto& kv : lu_of_N) {
        BandLU f(kv.first);
        f.pack(lu.data() + kv.second);
        if (!f.masks_hold()) sh.band_masks_ok = 0;
      }
    }
    CU_TRY(h, up(grp->d_lu, lu.data(), sizeof(double) * lu.size()));
    CU_TRY(h, up(grp->d_prob_id, grp->probs.data(), sizeof(int) * grp->probs.size()));
    const size_t plane_bytes = sizeof(double) * (size_t)count * std::max(1, sh.H) * 4 * sh.Mpad;
    CU_TRY(h, grp->d_planes.ensure(plane_bytes));
    const size_t ws_bytes = sizeof(double) * sh.ws_doubles, lu_bytes = sizeof(double) * sh.lu_doubles;
    if (lu_bytes + ws_bytes > cap) return fail(h, DFTPAV_ERR_UNSUPPORTED, "problem shape does not fit in shared memory");
    const int W = std::max(1, (int)std::min<size_t>((size_t)wantW, (cap - lu_bytes) / ws_bytes));
    grp->grid = grid;
    grp->warps = W;
    grp->smem = lu_bytes + ws_bytes * W;
    max_slots = std::max(max_slots, grp->grid * W);
    max_npad = std::max(max_npad, sh.npad);
    CU_TRY(h, launch_prepare((count + 127) / 128, 128, st, sh, h->pd, grp->d_prob_id.as<int>(), count,
                             h->seg_base.as<int>(), h->singul.as<int>(), h->ini_raw.as<double>(),
                             h->fin_raw.as<double>(), h->ini.as<double>(), h->fin.as<double>()));
    launches += 1;
  }
  // L-BFGS history ring per resident warp
  h->hist_ld = std::max(1, max_npad);
  h->hist_slots = max_slots;
  const size_t mem = (size_t)std::max(1, h->prm.lbfgs_mem_size);
  const size_t hbytes = sizeof(double) * (size_t)max_slots * mem * h->hist_ld;
  CU_TRY(h, h->hist_s.ensure(hbytes));
  CU_TRY(h, h->hist_y.ensure(hbytes));
  CU_TRY(h, h->hist_ys.ensure(sizeof(double) * (size_t)max_slots * mem));


  BatchDev& bd = h->bd;
  bd.seg_base = h->seg_base.as<int>();
  bd.t_now = h->t_now.as<double>();
  bd.obs_ptr = h->has_obs ? h->obs_ptr.as<int>() : nullptr;
  bd.obs_idx = h->obs_idx.as<int>();
  bd.singul = h->singul.as<int>();
  bd.ini = h->ini.as<double>();
  bd.fin = h->fin.as<double>();
  bd.obs_start = h->obs_start.as<double>();
  bd.obs_dur = h->obs_dur.as<double>();
  bd.obs_piece_ptr = h->obs_piece_ptr.as<int>();
  bd.obs_piece_dur = h->obs_piece_dur.as<double>();
  bd.obs_coeffs = h->obs_coeffs.as<double>();
  bd.obs_uniform = h->obs_uniform.as<double>();
  bd.x0 = h->x0.as<double>();
  bd.n_stride = dims.n_stride;
  bd.row_stride = dims.row_stride;
  bd.planes_raw = h->planes_raw.as<double>();
  bd.samp_base = h->samp_base.as<long long>();
  bd.plane_ptr = b->plane_cnt ? h->plane_ptr.as<long long>() : nullptr;
  bd.ready = h->ready.as<int>();
  bd.laid = h->laid.as<int>();
  bd.uniform_H = b->uniform_H;
  bd.chunk = h->chunk;
  bd.epoch = h->epoch;

  CU_TRY(h, cudaEventRecord(h->ev[1], st));
  h->stats.h2d_bytes += h2d;  // (the early plane copies have already added theirs)
  h->upload_launches = launches;
  h->stats.kernel_launches = launches;
  h->stats.n_shapes = (int)h->groups.size();
  h->uploaded = true;
  return DFTPAV_OK;
}

// The corridor planes: n_chunks copies, each followed (in stream order) by its arrival flag.
int issue_plane_copies(dftpav_handle* h, const dftpav_batch* b, cudaStream_t cs) {
  const int B = h->dims.B;
  for (int c = 0; c < h->n_chunks; c++) {
    const int p0 = c * h->chunk, p1 = std::min(B, p0 + h->chunk);
    if (p0 >= p1) break;
    const long long q0 = h->h_samp_base[p0], q1 = h->h_samp_base[p1];
    const long long e0 = h->uniform_H < 0 ? h->h_plane_ptr[q0] : q0 * h->uniform_H;
    const long long e1 = h->uniform_H < 0 ? h->h_plane_ptr[q1] : q1 * h->uniform_H;
    if (e1 > e0 && !h->planes_from_corridor)
      CU_TRY(h, cudaMemcpyAsync(h->planes_raw.as<double>() + 4 * e0, b->planes + 4 * e0, sizeof(double) * 4 * (size_t)(e1 - e0),
                                cudaMemcpyHostToDevice, cs));
    CU_TRY(h, cudaMemcpyAsync(h->ready.as<int>() + c, h->h_epoch + c, sizeof(int), cudaMemcpyHostToDevice, cs));
  }
  if (!h->planes_from_corridor) h->stats.h2d_bytes += (long long)sizeof(double) * 4 * h->total_planes;
  return DFTPAV_OK;
}

// launches the solver kernels of every shape group on the handle's stream (no synchronisation)
// order_key: 0 = input order, 1 = longest-predicted-first keyed by f(x0) WITHOUT the corridor term (the overlapped call: no
// plane needs to have arrived), 2 = keyed by the whole f(x0) (resident batches; measured equal at configs 2, 3, 5 and 1-4 %
// better at config 4, where the corridor term separates the problems more)
int launch_solve_kernels(dftpav_handle* h, int* launches_out, int order_key) {
  cudaStream_t st = h->stream;
  CU_TRY(h, cudaMemsetAsync(h->counter.p, 0, sizeof(int) * 64, st));
  OutDev out;
  out.x = h->o_x.as<double>();
  out.cost = h->o_cost.as<double>();
  out.status = h->o_status.as<int>();
  out.success = h->o_success.as<int>();
  out.iters = h->o_iters.as<int>();
  out.evals = h->o_evals.as<int>();
  out.coeffs = h->o_coeffs.as<double>();
  out.piece_dT = h->o_dT.as<double>();
  out.trace = nullptr;
  out.trace_cap = 0;
  if (h->trace_cap > 0) {
    const size_t tb = sizeof(double) * 3 * (size_t)h->trace_cap * h->dims.B;
    CU_TRY(h, h->trace.ensure(tb));
    CU_TRY(h, cudaMemsetAsync(h->trace.p, 0, tb, st));
    out.trace = h->trace.as<double>();
    out.trace_cap = h->trace_cap;
  }
  HistDev hist;
  hist.s = h->hist_s.as<double>();
  hist.y = h->hist_y.as<double>();
  hist.ys = h->hist_ys.as<double>();
  hist.ld = h->hist_ld;
  hist.mem = h->prm.lbfgs_mem_size;
  int launches = 0, gi = 0;
  // scheduling only: serve the queue longest-predicted-first (chunk_order_kernel, fetch_by_chunk).  The key is f(x0)
  // WITHOUT the corridor term: one extra evaluation per problem that reads no plane, so it also runs in the overlapped
  // dftpav_solve while the planes are still arriving; the queue only hands out problems of chunks that have arrived.
  const bool lpt = order_key != 0 && !strict_build_host() && !h->param_status && h->dbg_order.empty() &&
                   h->dims.B > 4 * h->num_sms;
  const bool full_key = h->dbg_order_key ? h->dbg_order_key == 2 : order_key == 2;
  if (lpt) {
    const int B = h->dims.B;
    CU_TRY(h, h->fout.ensure(sizeof(double) * B));
    CU_TRY(h, h->gout.ensure(sizeof(double) * (size_t)B * h->dims.n_stride));
    for (Group* grp : h->groups) {
      GroupDev gd{};
      gd.prob_id = grp->d_prob_id.as<int>();
      gd.count = (int)grp->probs.size();
      gd.planes = grp->d_planes.as<double>();
      gd.lu = grp->d_lu.as<double>();
      ShapeDev sh0 = grp->sh;
      if (!full_key) sh0.H = 0;  // no corridor term, no plane read, no waiting for the upload
      CU_TRY(h, launch_eval(h->has_obs, grp->grid, 32 * grp->warps, grp->smem, st, sh0, h->pd, h->bd, gd,
                            h->x0.as<double>(), h->fout.as<double>(), h->gout.as<double>(), nullptr, nullptr));
      CU_TRY(h, grp->d_clist.ensure(sizeof(int) * gd.count));
      CU_TRY(h, grp->d_cq.ensure(sizeof(int) * (2 * kMaxChunks + 1)));
      CU_TRY(h, cudaMemsetAsync(grp->d_cq.p, 0, sizeof(int) * (2 * kMaxChunks + 1), st));
      CU_TRY(h, launch_chunk_order(st, h->fout.as<double>(), gd.prob_id, gd.count, h->chunk, h->n_chunks,
                                   grp->d_clist.as<int>(), grp->d_cq.as<int>()));
      launches += 2;
    }
  }
  if (h->param_status) {
    // lbfgs_optimize returns before touching x (lbfgs.hpp:456-501): the decision vector stays the initial packing
    CU_TRY(h, cudaMemcpyAsync(h->o_x.p, h->x0.p, sizeof(double) * (size_t)h->dims.B * h->dims.n_stride,
                              cudaMemcpyDeviceToDevice, st));
  } else {
    for (Group* grp : h->groups) {
      if (gi >= 64) return fail(h, DFTPAV_ERR_UNSUPPORTED, "more than 64 distinct shapes in one batch");
      GroupDev gd{};
      gd.prob_id = grp->d_prob_id.as<int>();
      gd.count = (int)grp->probs.size();
      gd.planes = grp->d_planes.as<double>();
      gd.lu = grp->d_lu.as<double>();
      gd.order = nullptr;
      if (lpt) {
        gd.clist = grp->d_clist.as<int>();
        gd.cbeg = grp->d_cq.as<int>();
        gd.chead = grp->d_cq.as<int>() + kMaxChunks + 1;
        gd.key = h->fout.as<double>();
        gd.n_chunks = h->n_chunks;
      }
      if (!h->dbg_order.empty()) {  // experiment hook: queue order supplied by the caller (no effect on results)
        std::vector<int> local_of(h->dims.B, -1), ord;
        for (int i = 0; i < (int)grp->probs.size(); i++) local_of[grp->probs[i]] = i;
        for (int pid : h->dbg_order)
          if (pid >= 0 && pid < h->dims.B && local_of[pid] >= 0) ord.push_back(local_of[pid]), local_of[pid] = -1;
        for (int i = 0; i < (int)grp->probs.size(); i++)
          if (local_of[grp->probs[i]] >= 0) ord.push_back(i);
        CU_TRY(h, grp->d_order.ensure(sizeof(int) * ord.size()));
        CU_TRY(h, cudaMemcpyAsync(grp->d_order.p, ord.data(), sizeof(int) * ord.size(), cudaMemcpyHostToDevice, st));
        CU_TRY(h, cudaStreamSynchronize(st));
        gd.order = grp->d_order.as<int>();
      }
      CU_TRY(h, launch_solve(h->has_obs, grp->grid, 32 * grp->warps, grp->smem, st, grp->sh, h->pd, h->bd, gd, out, hist,
                             h->counter.as<int>() + gi));
      launches++;
      gi++;
    }
  }
  *launches_out = launches;
  return DFTPAV_OK;
}

}  // namespace

extern "C" {

int dftpav_upload_batch(dftpav_handle* h, const dftpav_batch* b) {
  if (!h) return DFTPAV_ERR_INVALID_ARG;
  int rc = upload_impl(h, b, 1);
  if (rc) return rc;
  rc = issue_plane_copies(h, b, h->stream);
  if (rc) {
    h->uploaded = false;
    return rc;
  }
  CU_TRY(h, cudaEventRecord(h->ev[1], h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->stats.upload_ms = ms;
  return DFTPAV_OK;
}

int dftpav_eval(dftpav_handle* h, const double* x, double* f, double* g) {
  if (!h || !x || !f || !g) return DFTPAV_ERR_INVALID_ARG;
  if (!h->uploaded) return fail(h, DFTPAV_ERR_STATE, "dftpav_eval before dftpav_upload_batch");
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int B = h->dims.B, ns = h->dims.n_stride;
  const size_t xb = sizeof(double) * (size_t)B * ns;
  CU_TRY(h, h->xin.ensure(xb));
  CU_TRY(h, h->gout.ensure(xb));
  CU_TRY(h, h->fout.ensure(sizeof(double) * B));
  CU_TRY(h, cudaMemcpyAsync(h->xin.p, x, xb, cudaMemcpyHostToDevice, st));
  if (h->n_rejected > 0) {  // rejected problems are not evaluated: f = 0, g = 0
    CU_TRY(h, cudaMemsetAsync(h->fout.p, 0, sizeof(double) * B, st));
    CU_TRY(h, cudaMemsetAsync(h->gout.p, 0, xb, st));
  }
  CU_TRY(h, cudaEventRecord(h->ev[0], st));
  int launches = 0;
  for (Group* grp : h->groups) {
    GroupDev gd{};
    gd.prob_id = grp->d_prob_id.as<int>();
    gd.count = (int)grp->probs.size();
    gd.planes = grp->d_planes.as<double>();
    gd.lu = grp->d_lu.as<double>();
    gd.order = nullptr;
    CU_TRY(h, launch_eval(h->has_obs, grp->grid, 32 * grp->warps, grp->smem, st, grp->sh, h->pd, h->bd, gd,
                          h->xin.as<double>(), h->fout.as<double>(), h->gout.as<double>(), nullptr, nullptr));
    launches++;
  }
  CU_TRY(h, cudaEventRecord(h->ev[1], st));
  CU_TRY(h, cudaMemcpyAsync(f, h->fout.p, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(g, h->gout.p, xb, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->stats.solve_ms = ms;
  h->stats.kernel_launches = launches;
  h->stats.total_evals = B - h->n_rejected;
  h->stats.total_iters = 0;
  long long ab = 0;
  for (int p = 0; p < B; p++)
    if (h->h_status0[p] != DFTPAV_PROBLEM_REJECTED) ab += h->h_evalbytes[p];
  h->stats.algorithmic_bytes = ab;
  return DFTPAV_OK;
}

int dftpav_optimize(dftpav_handle* h) {
  if (!h) return DFTPAV_ERR_INVALID_ARG;
  if (!h->uploaded) return fail(h, DFTPAV_ERR_STATE, "dftpav_optimize before dftpav_upload_batch");
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  CU_TRY(h, cudaEventRecord(h->ev[0], st));
  int launches = 0;
  int rc = launch_solve_kernels(h, &launches, 2);
  if (rc) return rc;
  CU_TRY(h, cudaEventRecord(h->ev[1], st));
  CU_TRY(h, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->stats.solve_ms = ms;
  h->stats.kernel_launches = h->upload_launches + launches;
  h->solved = true;
  return DFTPAV_OK;
}

int dftpav_download(dftpav_handle* h, const dftpav_result* out) {
  if (!h || !out) return DFTPAV_ERR_INVALID_ARG;
  if (!h->solved) return fail(h, DFTPAV_ERR_STATE, "dftpav_download before dftpav_optimize");
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const int B = h->dims.B, ns = h->dims.n_stride, rs = h->dims.row_stride, G = h->dims.seg_total;
  long long d2h = 0;
  CU_TRY(h, cudaEventRecord(h->ev[2], st));
  auto down = [&](void* dst, const DevBuf& src, size_t bytes) -> cudaError_t {
    if (!dst || !bytes) return cudaSuccess;
    d2h += (long long)bytes;
    return cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, st);
  };
  std::vector<int> iters(B), evals(B);
  CU_TRY(h, down(out->x, h->o_x, sizeof(double) * (size_t)B * ns));
  CU_TRY(h, down(out->cost, h->o_cost, sizeof(double) * B));
  CU_TRY(h, down(out->status, h->o_status, sizeof(int) * B));
  CU_TRY(h, down(out->success, h->o_success, sizeof(int) * B));
  CU_TRY(h, down(out->coeffs, h->o_coeffs, sizeof(double) * (size_t)B * rs * 2));
  CU_TRY(h, down(out->piece_dT, h->o_dT, sizeof(double) * G));
  CU_TRY(h, down(iters.data(), h->o_iters, sizeof(int) * B));
  CU_TRY(h, down(evals.data(), h->o_evals, sizeof(int) * B));
  CU_TRY(h, cudaEventRecord(h->ev[3], st));
  CU_TRY(h, cudaStreamSynchronize(st));
  if (out->iters) std::copy(iters.begin(), iters.end(), out->iters);
  if (out->evals) std::copy(evals.begin(), evals.end(), out->evals);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]);
  h->stats.download_ms = ms;
  h->stats.d2h_bytes = d2h;
  long long ti = 0, te = 0, ab = 0;
  const long long m = h->prm.lbfgs_mem_size;
  for (int p = 0; p < B; p++) {
    ti += iters[p];
    te += evals[p];
    const long long n = h->h_n[p];
    ab += (long long)evals[p] * h->h_evalbytes[p];
    // B_iter(k) = 32 n min(k-1, m) + 112 n, k = 1..iters  
    const long long k = iters[p];
    const long long kk = std::min(k, m + 1);  // iterations with k-1 <= m
    ab += 32 * n * (kk * (kk - 1) / 2 + (k - kk) * m) + 112 * n * k;
  }
  h->stats.total_iters = ti;
  h->stats.total_evals = te;
  h->stats.algorithmic_bytes = ab;
  return DFTPAV_OK;
}

// One planner call: small arrays and state kernels on the main stream, the corridor planes (95 % of the bytes) in chunks on a
// second stream, and the solver kernels launched right behind them - the solve runs under the copy: a warp waits for the
// arrival flag of its problem's chunk (layout_problem in dftpav_kernels.cu), the first problems start after the first
// chunk.  Replaces the deep copy of hPoly_container at the top of OptimizeTrajectory (traj_optimizer.cpp:15).
int dftpav_solve(dftpav_handle* h, const dftpav_batch* batch, const dftpav_result* out) {
  if (!h) return DFTPAV_ERR_INVALID_ARG;
  // The plane copies are queued on the copy stream from INSIDE upload_impl, as soon as the plane prefixes are known and before
  // the solver kernels are launched: with pinned host buffers they return at once and run under the host-side packing and
  // the kernels; and whatever serialises the device work (a profiler, CUDA_LAUNCH_BLOCKING, pageable source buffers that the
  // driver stages synchronously) only loses the overlap - the kernels can never be left polling for planes nobody has queued.
  int rc = upload_impl(h, batch, kMaxChunks, h->copy_stream);
  if (rc) {
    cudaStreamSynchronize(h->copy_stream);
    h->uploaded = false;
    return rc;
  }
  cudaStream_t st = h->stream;
  CU_TRY(h, cudaEventRecord(h->ev[4], st));
  int launches = 0;
  rc = launch_solve_kernels(h, &launches, 1);
  if (rc) {
    cudaStreamSynchronize(h->copy_stream);
    return rc;
  }
  CU_TRY(h, cudaEventRecord(h->ev[5], st));
  h->stats.kernel_launches = h->upload_launches + launches;
  h->solved = true;
  rc = dftpav_download(h, out);  // queued behind the kernels; synchronises the main stream
  CU_TRY(h, cudaStreamSynchronize(h->copy_stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]);
  h->stats.solve_ms = ms;  // includes waiting for the plane chunks
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->stats.upload_ms = ms;  // small arrays, tables and the state kernel; the plane copy runs under the solve
  return rc;
}

// The local half of the global convergence flag: how many problems of the last solve ended with
// flag_success (traj_optimizer.cpp:176-201).  Reads B ints back; the caller all-reduces (MIN of n_success == B) over ranks.
int dftpav_success_count(dftpav_handle* h, int32_t* n_success) {
  if (!h || !n_success) return DFTPAV_ERR_INVALID_ARG;
  if (!h->solved) return fail(h, DFTPAV_ERR_STATE, "dftpav_success_count before dftpav_optimize");
  CU_TRY(h, cudaSetDevice(h->device));
  std::vector<int> ok(h->dims.B);
  CU_TRY(h, cudaMemcpyAsync(ok.data(), h->o_success.p, sizeof(int) * h->dims.B, cudaMemcpyDeviceToHost, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  int n = 0;
  for (int v : ok) n += v != 0;
  *n_success = n;
  return DFTPAV_OK;
}

void dftpav_state_to_flat_output(const double* state, double* flat) {
  // TrajPlanner::state_to_flat_output (traj_manager.cpp:139-158); state = (x, y, angle, velocity, acceleration, curvature)
  double vel = state[3];
  const double angle = state[2], acc = state[4];
  const double c = std::cos(angle), s = std::sin(angle);
  if (vel == 0) vel = 1e-5;
  flat[0] = state[0], flat[1] = state[1];
  flat[2] = c * vel + (-s) * 0.0, flat[3] = s * vel + c * 0.0;
  const double lat = state[5] * std::pow(vel, 2);
  flat[4] = c * acc + (-s) * lat, flat[5] = s * acc + c * lat;
}

int dftpav_minco_fit(dftpav_handle* h, int32_t n_traj, const int32_t* piece_ptr, const double* dT, const double* head,
                     const double* tail, const double* inner_pts, double* coeffs) {
  if (!h || n_traj < 0 || (n_traj > 0 && (!piece_ptr || !dT || !head || !tail || !inner_pts || !coeffs))) return DFTPAV_ERR_INVALID_ARG;
  if (n_traj == 0) return DFTPAV_OK;
  if (piece_ptr[0] != 0) return fail(h, DFTPAV_ERR_INVALID_ARG, "piece_ptr[0] != 0");
  // LU table per distinct piece count (MinJerkOpt::reset: the factor depends only on N), inner-point prefix
  std::map<int, int> off_of_N;
  std::vector<double> tables;
  std::vector<int> lu_off(n_traj);
  std::vector<long long> iptr(n_traj + 1, 0);
  for (int k = 0; k < n_traj; k++) {
    const int N = piece_ptr[k + 1] - piece_ptr[k];
    if (N < 2) return fail(h, DFTPAV_ERR_INVALID_ARG, "a fit needs at least 2 pieces");
    if (N > 4 * kMaxPieces) return fail(h, DFTPAV_ERR_UNSUPPORTED, "more than 256 pieces in one fit");
    auto it = off_of_N.find(N);
    if (it == off_of_N.end()) {
      const int off = (int)tables.size();
      tables.resize(tables.size() + (size_t)6 * N * kLuLd);
      BandLU f(N);
      f.pack(tables.data() + off);
      it = off_of_N.emplace(N, off).first;
    }
    lu_off[k] = it->second;
    iptr[k + 1] = iptr[k] + 2LL * (N - 1);
  }
  const long long NP = piece_ptr[n_traj];
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  auto up = [&](DevBuf& buf, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = buf.ensure(bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, st);
  };
  CU_TRY(h, up(h->fit_pp, piece_ptr, sizeof(int) * (n_traj + 1)));
  CU_TRY(h, up(h->fit_dT, dT, sizeof(double) * n_traj));
  CU_TRY(h, up(h->fit_head, head, sizeof(double) * 6 * n_traj));
  CU_TRY(h, up(h->fit_tail, tail, sizeof(double) * 6 * n_traj));
  CU_TRY(h, up(h->fit_inner, inner_pts, sizeof(double) * (size_t)iptr[n_traj]));
  CU_TRY(h, up(h->fit_iptr, iptr.data(), sizeof(long long) * (n_traj + 1)));
  CU_TRY(h, up(h->fit_lu, tables.data(), sizeof(double) * tables.size()));
  CU_TRY(h, up(h->fit_luoff, lu_off.data(), sizeof(int) * n_traj));
  CU_TRY(h, h->fit_scratch.ensure(sizeof(double) * 12 * (size_t)NP));
  CU_TRY(h, h->fit_coeffs.ensure(sizeof(double) * 12 * (size_t)NP));
  CU_TRY(h, launch_minco_fit(st, n_traj, h->fit_pp.as<int>(), h->fit_dT.as<double>(), h->fit_head.as<double>(),
                             h->fit_tail.as<double>(), h->fit_inner.as<double>(), h->fit_iptr.as<long long>(),
                             h->fit_lu.as<double>(), h->fit_luoff.as<int>(), h->fit_scratch.as<double>(),
                             h->fit_coeffs.as<double>()));
  CU_TRY(h, cudaMemcpyAsync(coeffs, h->fit_coeffs.p, sizeof(double) * 12 * (size_t)NP, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));  // (also keeps the host staging vectors alive until the copies are done)
  return DFTPAV_OK;
}

int dftpav_corridor_set_map(dftpav_handle* h, const dftpav_gridmap* map) {
  if (!h || !map || !map->data || map->size_x < 1 || map->size_y < 1 || !(map->resolution > 0.0)) return DFTPAV_ERR_INVALID_ARG;
  CU_TRY(h, cudaSetDevice(h->device));
  const size_t bytes = (size_t)map->size_x * map->size_y;
  CU_TRY(h, h->map_data.ensure(bytes));
  CU_TRY(h, cudaMemcpyAsync(h->map_data.p, map->data, bytes, cudaMemcpyHostToDevice, h->stream));
  // ray-march offsets of CheckIfCollisionUsingLine: dl = 0; dl += resolution / 2 (sequential sums, like the reference's loop)
  const double checkl = map->resolution / 2.0;
  const double longest = 64.0 + 4.0 * map->resolution;  // > vehicle length + 2 * limit_bound for any sensible box
  std::vector<double> dl;
  for (double v = 0.0; v < longest && dl.size() < (1u << 16); v += checkl) dl.push_back(v);
  CU_TRY(h, h->map_dl.ensure(sizeof(double) * dl.size()));
  CU_TRY(h, cudaMemcpyAsync(h->map_dl.p, dl.data(), sizeof(double) * dl.size(), cudaMemcpyHostToDevice, h->stream));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  h->map_meta = *map;
  h->map_meta.data = nullptr;
  h->map_n_dl = (int)dl.size();
  return DFTPAV_OK;
}

int dftpav_corridor_rectangles(dftpav_handle* h, const dftpav_corridor_params* cp, int64_t n, const double* states,
                               const double* cos_sin, double* hpolys, int32_t keep) {
  if (!h || !cp || n < 0 || (n > 0 && !states)) return DFTPAV_ERR_INVALID_ARG;
  if (h->map_n_dl == 0) return fail(h, DFTPAV_ERR_STATE, "dftpav_corridor_rectangles before dftpav_corridor_set_map");
  if (!(cp->limit_bound > 0.0) || (cp->veh_length + 2.0 * cp->limit_bound + 4.0 * h->map_meta.resolution) >= 64.0)
    return fail(h, DFTPAV_ERR_UNSUPPORTED, "vehicle length + 2 * limit_bound must stay below 64 m");
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  h->corridor_n = -1;
  if (n == 0) return DFTPAV_OK;
  CU_TRY(h, h->cor_states.ensure(sizeof(double) * 3 * (size_t)n));
  CU_TRY(h, cudaMemcpyAsync(h->cor_states.p, states, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st));
  if (cos_sin) {
    CU_TRY(h, h->cor_cs.ensure(sizeof(double) * 2 * (size_t)n));
    CU_TRY(h, cudaMemcpyAsync(h->cor_cs.p, cos_sin, sizeof(double) * 2 * (size_t)n, cudaMemcpyHostToDevice, st));
  }
  CU_TRY(h, h->planes_raw.ensure(sizeof(double) * 16 * (size_t)n));
  CorridorDev cd;
  cd.data = h->map_data.as<unsigned char>();
  cd.dl = h->map_dl.as<double>();
  cd.size_x = h->map_meta.size_x, cd.size_y = h->map_meta.size_y, cd.n_dl = h->map_n_dl;
  cd.occupied = h->map_meta.occupied_value;
  cd.origin_x = h->map_meta.origin_x, cd.origin_y = h->map_meta.origin_y, cd.resolution = h->map_meta.resolution;
  cd.veh_width = cp->veh_width, cd.veh_length = cp->veh_length, cd.veh_d_cr = cp->veh_d_cr, cd.limit_bound = cp->limit_bound;
  CU_TRY(h, cudaEventRecord(h->ev[0], st));
  CU_TRY(h, launch_corridor(st, cd, n, h->cor_states.as<double>(), cos_sin ? h->cor_cs.as<double>() : nullptr,
                            h->planes_raw.as<double>()));
  CU_TRY(h, cudaEventRecord(h->ev[1], st));
  if (hpolys)
    CU_TRY(h, cudaMemcpyAsync(hpolys, h->planes_raw.p, sizeof(double) * 16 * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->stats.solve_ms = ms;
  h->stats.kernel_launches = 1;
  if (keep) h->corridor_n = n;
  return DFTPAV_OK;
}

static SampleDev sample_view(const dftpav_handle* h) {
  SampleDev sd;
  sd.coeffs = h->o_coeffs.as<double>();
  sd.piece_dT = h->o_dT.as<double>();
  sd.singul = h->singul.as<int>();
  sd.seg_prob = h->seg_prob.as<int>();
  sd.seg_row0 = h->seg_row0.as<int>();
  sd.seg_N = h->seg_N.as<int>();
  sd.seg_piece0 = h->seg_piece0.as<long long>();
  sd.G = h->dims.seg_total;
  sd.row_stride = h->dims.row_stride;
  return sd;
}

int dftpav_sample(dftpav_handle* h, int64_t nq, const int32_t* seg, const double* t, int32_t flags, double wheel_base,
                  double* out) {
  if (!h || nq < 0 || (nq > 0 && (!seg || !t || !out))) return DFTPAV_ERR_INVALID_ARG;
  if (!h->solved) return fail(h, DFTPAV_ERR_STATE, "dftpav_sample before dftpav_optimize");
  if (nq == 0) return DFTPAV_OK;
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  CU_TRY(h, h->q_seg.ensure(sizeof(int) * (size_t)nq));
  CU_TRY(h, h->q_t.ensure(sizeof(double) * (size_t)nq));
  CU_TRY(h, h->q_out.ensure(sizeof(double) * DFTPAV_STATE_DOUBLES * (size_t)nq));
  CU_TRY(h, cudaMemcpyAsync(h->q_seg.p, seg, sizeof(int) * (size_t)nq, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(h->q_t.p, t, sizeof(double) * (size_t)nq, cudaMemcpyHostToDevice, st));
  CU_TRY(h, launch_sample(st, sample_view(h), nq, h->q_seg.as<int>(), h->q_t.as<double>(), flags, wheel_base,
                          h->q_out.as<double>()));
  CU_TRY(h, cudaMemcpyAsync(out, h->q_out.p, sizeof(double) * DFTPAV_STATE_DOUBLES * (size_t)nq, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return DFTPAV_OK;
}

int dftpav_sample_uniform(dftpav_handle* h, double dt, int32_t max_per_seg, double* xyyaw, int32_t* count) {
  if (!h || !(dt > 0.0) || max_per_seg < 1 || !xyyaw || !count) return DFTPAV_ERR_INVALID_ARG;
  if (!h->solved) return fail(h, DFTPAV_ERR_STATE, "dftpav_sample_uniform before dftpav_optimize");
  const int G = h->dims.seg_total;
  if (G == 0) return DFTPAV_OK;
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const size_t ob = sizeof(double) * 3 * (size_t)G * max_per_seg;
  CU_TRY(h, h->q_out.ensure(ob));
  CU_TRY(h, h->q_cnt.ensure(sizeof(int) * G));
  CU_TRY(h, cudaMemsetAsync(h->q_out.p, 0, ob, st));
  CU_TRY(h, launch_sample_uniform(st, sample_view(h), dt, max_per_seg, h->q_out.as<double>(), h->q_cnt.as<int>()));
  CU_TRY(h, cudaMemcpyAsync(xyyaw, h->q_out.p, ob, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(count, h->q_cnt.p, sizeof(int) * G, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return DFTPAV_OK;
}

// ---- front-end resampling (row N3) ----
namespace {

// host-side packer: every array of one call goes into ONE staging vector (8-byte aligned slots) and one H2D copy
struct Arena {
  std::vector<unsigned char> host;
  size_t put(const void* src, size_t bytes) {
    const size_t off = (host.size() + 7) & ~(size_t)7;
    host.resize(off + bytes);
    if (bytes) std::memcpy(host.data() + off, src, bytes);
    return off;
  }
};

int check_frontend(dftpav_handle* h, const dftpav_frontend* fe) {
  if (!fe || fe->n_paths < 0) return fail(h, DFTPAV_ERR_INVALID_ARG, "frontend: null or negative n_paths");
  if (fe->n_paths == 0) return DFTPAV_OK;
  if (!fe->pt_ptr || !fe->pts || !fe->shot_ptr || !fe->shot_index || !fe->shot_time || !fe->shot_length || !fe->shot_s ||
      !fe->start_vel || !fe->end_vel)
    return fail(h, DFTPAV_ERR_INVALID_ARG, "frontend: null array");
  if (fe->pt_ptr[0] != 0 || fe->shot_ptr[0] != 0) return fail(h, DFTPAV_ERR_INVALID_ARG, "frontend: CSR pointers must start at 0");
  for (int p = 0; p < fe->n_paths; p++) {
    const int np = fe->pt_ptr[p + 1] - fe->pt_ptr[p], ns = fe->shot_ptr[p + 1] - fe->shot_ptr[p];
    if (np < 2 || ns < 1) return fail(h, DFTPAV_ERR_INVALID_ARG, "frontend: a path needs two points and one gear segment");
    const int32_t* si = fe->shot_index + fe->shot_ptr[p] + p;
    for (int i = 0; i <= ns; i++)
      if (si[i] < 0 || si[i] >= np || (i > 0 && si[i] < si[i - 1]))
        return fail(h, DFTPAV_ERR_INVALID_ARG, "frontend: shot_index out of range or decreasing");
  }
  return DFTPAV_OK;
}

// packs the front end into the arena; returns the device view (pointers relative to `base`)
FrontendDev pack_frontend(Arena& a, const dftpav_frontend* fe, size_t (&off)[9]) {
  const int n = fe->n_paths, PT = fe->pt_ptr[n], SH = fe->shot_ptr[n];
  off[0] = a.put(fe->pt_ptr, sizeof(int) * (n + 1));
  off[1] = a.put(fe->pts, sizeof(double) * 3 * (size_t)PT);
  off[2] = a.put(fe->shot_ptr, sizeof(int) * (n + 1));
  off[3] = a.put(fe->shot_index, sizeof(int) * ((size_t)SH + n));
  off[4] = a.put(fe->shot_time, sizeof(double) * SH);
  off[5] = a.put(fe->shot_length, sizeof(double) * SH);
  off[6] = a.put(fe->shot_s, sizeof(int) * SH);
  off[7] = a.put(fe->start_vel, sizeof(double) * n);
  off[8] = a.put(fe->end_vel, sizeof(double) * n);
  FrontendDev d{};
  d.n_paths = n;
  d.max_forward_vel = fe->max_forward_vel, d.max_forward_acc = fe->max_forward_acc;
  d.max_backward_vel = fe->max_backward_vel, d.max_backward_acc = fe->max_backward_acc, d.non_siguav = fe->non_siguav;
  return d;
}

void bind_frontend(FrontendDev& d, const unsigned char* base, const size_t (&off)[9]) {
  d.pt_ptr = reinterpret_cast<const int*>(base + off[0]);
  d.pts = reinterpret_cast<const double*>(base + off[1]);
  d.shot_ptr = reinterpret_cast<const int*>(base + off[2]);
  d.shot_index = reinterpret_cast<const int*>(base + off[3]);
  d.shot_time = reinterpret_cast<const double*>(base + off[4]);
  d.shot_length = reinterpret_cast<const double*>(base + off[5]);
  d.shot_s = reinterpret_cast<const int*>(base + off[6]);
  d.start_vel = reinterpret_cast<const double*>(base + off[7]);
  d.end_vel = reinterpret_cast<const double*>(base + off[8]);
}

}  // namespace

double dftpav_trapezoid_duration(double length, double max_vel, double max_acc, double startV, double endV) {
  // KinoAstar::evaluateDuration (kino_astar.cpp:744-762)
  const double startv2 = startV * startV, endv2 = endV * endV, maxv2 = max_vel * max_vel;
  const double critical_len = (maxv2 - startv2) / (2 * max_acc) + (maxv2 - endv2) / (2 * max_acc);
  if (length >= critical_len) return (max_vel - startV) / max_acc + (max_vel - endV) / max_acc + (length - critical_len) / max_vel;
  const double tmpv = std::sqrt(0.5 * (startv2 + endv2 + 2 * max_acc * length));
  return (tmpv - startV) / max_acc + (tmpv - endV) / max_acc;
}

int dftpav_frontend_evaluate_pos(dftpav_handle* h, const dftpav_frontend* fe, int64_t nq, const int32_t* path, const double* t,
                                 double* out) {
  if (!h || nq < 0 || (nq > 0 && (!path || !t || !out))) return DFTPAV_ERR_INVALID_ARG;
  int rc = check_frontend(h, fe);
  if (rc || nq == 0) return rc;
  for (int64_t q = 0; q < nq; q++)
    if (path[q] < 0 || path[q] >= fe->n_paths) return fail(h, DFTPAV_ERR_INVALID_ARG, "frontend: query path out of range");
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  Arena a;
  size_t off[9];
  FrontendDev d = pack_frontend(a, fe, off);
  const size_t o_path = a.put(path, sizeof(int) * (size_t)nq), o_t = a.put(t, sizeof(double) * (size_t)nq);
  CU_TRY(h, h->fe_arena.ensure(a.host.size()));
  CU_TRY(h, h->fe_out.ensure(sizeof(double) * 3 * (size_t)nq));
  CU_TRY(h, cudaMemcpyAsync(h->fe_arena.p, a.host.data(), a.host.size(), cudaMemcpyHostToDevice, st));
  const unsigned char* base = h->fe_arena.as<unsigned char>();
  bind_frontend(d, base, off);
  CU_TRY(h, launch_frontend_pos(st, d, nq, reinterpret_cast<const int*>(base + o_path),
                                reinterpret_cast<const double*>(base + o_t), h->fe_out.as<double>()));
  CU_TRY(h, cudaMemcpyAsync(out, h->fe_out.p, sizeof(double) * 3 * (size_t)nq, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return DFTPAV_OK;
}

int dftpav_frontend_resample(dftpav_handle* h, const dftpav_frontend* fe, const double* seg_duration,
                             double traj_piece_duration, int32_t K, int32_t Kd, int32_t* piece_num, double* piece_dur,
                             int64_t states_cap, double* states, double* inner, int64_t* n_states) {
  if (!h || !seg_duration || !(traj_piece_duration > 0.0) || K < 1 || Kd < 1 || !piece_num || !piece_dur || !n_states)
    return DFTPAV_ERR_INVALID_ARG;
  int rc = check_frontend(h, fe);
  if (rc) return rc;
  const int n = fe->n_paths, SH = n ? fe->shot_ptr[n] : 0;
  // per gear segment (host arithmetic of traj_manager.cpp:531-541, 566): piece count, piece duration, base time, offsets
  std::vector<int> seg_path(SH), seg_pn(SH);
  std::vector<double> seg_tpp(SH), seg_base(SH);
  std::vector<long long> seg_st0(SH), seg_in0(SH);
  long long ns = 0, ni = 0;
  for (int p = 0; p < n; p++) {
    double basetime = 0.0;
    for (int g = fe->shot_ptr[p]; g < fe->shot_ptr[p + 1]; g++) {
      const double D = seg_duration[g];
      if (!(D > 0.0) || !(D / traj_piece_duration < 1.0e6)) return fail(h, DFTPAV_ERR_INVALID_ARG, "frontend: bad segment duration");
      const int pn = std::max(int(D / traj_piece_duration + 0.5), 2);
      seg_path[g] = p, seg_pn[g] = pn, seg_tpp[g] = D / pn, seg_base[g] = basetime;
      seg_st0[g] = ns, seg_in0[g] = ni;
      piece_num[g] = pn, piece_dur[g] = seg_tpp[g];
      ns += 2LL * (Kd + 1) + (long long)(pn - 2) * (K + 1);
      ni += pn - 1;
      basetime += D;
    }
  }
  *n_states = ns;
  if (ns > states_cap) return fail(h, DFTPAV_ERR_INVALID_ARG, "frontend: states_cap is too small (see *n_states)");
  if (SH == 0) return DFTPAV_OK;
  if (!states || !inner) return DFTPAV_ERR_INVALID_ARG;
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  Arena a;
  size_t off[9];
  FrontendDev d = pack_frontend(a, fe, off);
  const size_t o_path = a.put(seg_path.data(), sizeof(int) * SH), o_pn = a.put(seg_pn.data(), sizeof(int) * SH);
  const size_t o_tpp = a.put(seg_tpp.data(), sizeof(double) * SH), o_base = a.put(seg_base.data(), sizeof(double) * SH);
  const size_t o_st0 = a.put(seg_st0.data(), sizeof(long long) * SH), o_in0 = a.put(seg_in0.data(), sizeof(long long) * SH);
  CU_TRY(h, h->fe_arena.ensure(a.host.size()));
  const size_t sb = sizeof(double) * 3 * (size_t)ns, ib = sizeof(double) * 2 * (size_t)ni;
  CU_TRY(h, h->fe_out.ensure(sb + ib));
  CU_TRY(h, cudaMemcpyAsync(h->fe_arena.p, a.host.data(), a.host.size(), cudaMemcpyHostToDevice, st));
  const unsigned char* base = h->fe_arena.as<unsigned char>();
  bind_frontend(d, base, off);
  double* d_states = h->fe_out.as<double>();
  double* d_inner = d_states + 3 * (size_t)ns;
  CU_TRY(h, cudaEventRecord(h->ev[0], st));
  CU_TRY(h, launch_frontend_resample(st, d, SH, reinterpret_cast<const int*>(base + o_path),
                                     reinterpret_cast<const int*>(base + o_pn), reinterpret_cast<const double*>(base + o_tpp),
                                     reinterpret_cast<const double*>(base + o_base),
                                     reinterpret_cast<const long long*>(base + o_st0),
                                     reinterpret_cast<const long long*>(base + o_in0), K, Kd, d_states, d_inner));
  CU_TRY(h, cudaEventRecord(h->ev[1], st));
  CU_TRY(h, cudaMemcpyAsync(states, d_states, sb, cudaMemcpyDeviceToHost, st));
  if (ib) CU_TRY(h, cudaMemcpyAsync(inner, d_inner, ib, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->stats.solve_ms = ms;
  h->stats.kernel_launches = 1;
  return DFTPAV_OK;
}

int dftpav_collision_check(dftpav_handle* h, const dftpav_corridor_params* vp, double dt, double vertex_spacing,
                           int32_t max_samples, int32_t* collide, int32_t* first_hit) {
  if (!h || !vp || !(dt > 0.0) || !(vertex_spacing > 0.0) || max_samples < 1 || !collide) return DFTPAV_ERR_INVALID_ARG;
  if (!h->solved) return fail(h, DFTPAV_ERR_STATE, "dftpav_collision_check before dftpav_optimize");
  if (!h->map_data.p) return fail(h, DFTPAV_ERR_STATE, "dftpav_collision_check before dftpav_corridor_set_map");
  const int G = h->dims.seg_total, B = h->dims.B;
  if (G == 0) return DFTPAV_OK;
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  // sample times of `for (t = 0; t < duration; t += dt)`: sequentially accumulated sums
  std::vector<double> tt((size_t)max_samples);
  double t = 0.0;
  for (int k = 0; k < max_samples; k++, t += dt) tt[k] = t;
  CU_TRY(h, h->q_t.ensure(sizeof(double) * tt.size()));
  CU_TRY(h, cudaMemcpyAsync(h->q_t.p, tt.data(), sizeof(double) * tt.size(), cudaMemcpyHostToDevice, st));
  CU_TRY(h, h->q_cnt.ensure(sizeof(int) * ((size_t)G + B)));
  int* d_first = h->q_cnt.as<int>();
  int* d_col = d_first + G;
  CU_TRY(h, cudaMemsetAsync(d_col, 0, sizeof(int) * B, st));
  CorridorDev cd;
  cd.data = h->map_data.as<unsigned char>();
  cd.dl = h->map_dl.as<double>();
  cd.size_x = h->map_meta.size_x, cd.size_y = h->map_meta.size_y, cd.n_dl = h->map_n_dl;
  cd.occupied = h->map_meta.occupied_value;
  cd.origin_x = h->map_meta.origin_x, cd.origin_y = h->map_meta.origin_y, cd.resolution = h->map_meta.resolution;
  cd.veh_width = vp->veh_width, cd.veh_length = vp->veh_length, cd.veh_d_cr = vp->veh_d_cr, cd.limit_bound = vp->limit_bound;
  CU_TRY(h, cudaEventRecord(h->ev[0], st));
  CU_TRY(h, launch_collision(st, sample_view(h), cd, h->q_t.as<double>(), max_samples, dt, vertex_spacing,
                             h->seg_prob.as<int>(), d_first, d_col));
  CU_TRY(h, cudaEventRecord(h->ev[1], st));
  std::vector<int> fh((size_t)G);
  CU_TRY(h, cudaMemcpyAsync(fh.data(), d_first, sizeof(int) * G, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(collide, d_col, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->stats.solve_ms = ms;
  h->stats.kernel_launches = 1;
  bool truncated = false;
  for (int g = 0; g < G; g++) {
    truncated = truncated || fh[g] == -2;
    if (first_hit) first_hit[g] = fh[g];
  }
  if (truncated) return fail(h, DFTPAV_ERR_INVALID_ARG, "dftpav_collision_check: max_samples * dt is shorter than a segment");
  return DFTPAV_OK;
}

int dftpav_pack_polytraj(dftpav_handle* h, float* coef_x, float* coef_y, float* duration) {
  if (!h || !coef_x || !coef_y || !duration) return DFTPAV_ERR_INVALID_ARG;
  if (!h->solved) return fail(h, DFTPAV_ERR_STATE, "dftpav_pack_polytraj before dftpav_optimize");
  const long long NP = h->pieces_total;
  if (NP == 0) return DFTPAV_OK;
  CU_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  CU_TRY(h, h->q_out.ensure(sizeof(float) * 13 * (size_t)NP));
  float* cx = h->q_out.as<float>();
  float* cy = cx + 6 * NP;
  float* du = cy + 6 * NP;
  CU_TRY(h, launch_pack_polytraj(st, sample_view(h), cx, cy, du));
  CU_TRY(h, cudaMemcpyAsync(coef_x, cx, sizeof(float) * 6 * (size_t)NP, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(coef_y, cy, sizeof(float) * 6 * (size_t)NP, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(duration, du, sizeof(float) * (size_t)NP, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return DFTPAV_OK;
}

int dftpav_last_stats(const dftpav_handle* h, dftpav_stats* stats) {
  if (!h || !stats) return DFTPAV_ERR_INVALID_ARG;
  *stats = h->stats;
  return DFTPAV_OK;
}

// Internal (dftpav_multi.cu): lower bounds for the output strides, so that a shard's rows line up with the whole batch's arrays.
int dftpav_internal_set_min_strides(dftpav_handle* h, int n_stride, int row_stride) {
  if (!h) return DFTPAV_ERR_INVALID_ARG;
  h->min_n_stride = n_stride;
  h->min_row_stride = row_stride;
  return DFTPAV_OK;
}

// Test hook (not part of the drop-in ABI): use the banded LU substitutions even where the dense MINCO tables fit; takes
// effect at the next upload.
int dftpav_debug_force_banded(dftpav_handle* h, int on) {
  if (!h) return DFTPAV_ERR_INVALID_ARG;
  h->force_banded = on != 0;
  return DFTPAV_OK;
}
// Experiment hook (not part of the drop-in ABI): order in which the persistent warps take the problems (ids, longest-expected
// first); NULL / n = 0 restores the input order.  Results do not depend on it.
int dftpav_debug_set_order(dftpav_handle* h, const int32_t* order, int32_t n) {
  if (!h) return DFTPAV_ERR_INVALID_ARG;
  h->dbg_order.assign(order ? order : nullptr, order ? order + n : nullptr);
  return DFTPAV_OK;
}
// Build variant of the loaded library: bit 0 strict-order verification build; bits 8.. virtual ranks
int dftpav_debug_build_flags(void) {
  return (strict_build_host() ? 1 : 0) | (virtual_ranks_host() << 8);
}

// Diagnostic (not part of the drop-in ABI): per-iteration (fx, step, ls) traces of the next dftpav_optimize calls.
int dftpav_debug_set_trace(dftpav_handle* h, int trace_cap) {
  if (!h || trace_cap < 0) return DFTPAV_ERR_INVALID_ARG;
  h->trace_cap = trace_cap;
  return DFTPAV_OK;
}
int dftpav_debug_get_trace(dftpav_handle* h, double* out) {
  if (!h || !out || h->trace_cap <= 0 || !h->solved) return DFTPAV_ERR_STATE;
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaMemcpy(out, h->trace.p, sizeof(double) * 3 * (size_t)h->trace_cap * h->dims.B, cudaMemcpyDeviceToHost));
  return DFTPAV_OK;
}

// Experiment hook (scheduling only): key of the longest-first queue, 0 = default (the whole f(x0) for resident batches, f(x0)
// without the corridor term for the overlapped dftpav_solve), 1 = always plane-free, 2 = always the whole f(x0).
int dftpav_debug_set_order_key(dftpav_handle* h, int mode) {
  if (!h) return DFTPAV_ERR_INVALID_ARG;
  h->dbg_order_key = mode;
  return DFTPAV_OK;
}

// Diagnostic (not part of the drop-in ABI): enables per-phase cycle counters of the evaluation and returns them.
// phases: 0 rhs build, 1 banded solve, 2 coefficients + jerk, 3 sweep, 4 adjoint rhs, 5 adjoint solve, 6 gradient assembly
int dftpav_debug_phase_clock(dftpav_handle* h, int enable, unsigned long long* cycles8) {
  if (!h) return DFTPAV_ERR_INVALID_ARG;
  static unsigned long long* d_clk = nullptr;
  CU_TRY(h, cudaSetDevice(h->device));
  CU_TRY(h, cudaStreamSynchronize(h->stream));
  if (!d_clk) CU_TRY(h, cudaMalloc(&d_clk, 16 * sizeof(unsigned long long)));
  // slots 8..15 (filled by -DDFTPAV_PHASE_FINE=1 builds only) are returned when enable == 2 (cycles8 then holds 16 values)
  if (cycles8)
    CU_TRY(h, cudaMemcpy(cycles8, d_clk, (enable == 2 ? 16 : 8) * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (enable == 2) return DFTPAV_OK;
  CU_TRY(h, cudaMemset(d_clk, 0, 16 * sizeof(unsigned long long)));
  CU_TRY(h, set_phase_clock(enable ? d_clk : nullptr));
  return DFTPAV_OK;
}

// Diagnostic (host only, no GPU needed): the dense MINCO table of N pieces as the kernels read it, Mt[j][row] with
// leading dimension 6N+1, (N+5) columns; rows_of_col[j] = matrix row of column j.  Returns the number of doubles written
// (or needed, when out is null).
long long dftpav_debug_minco_dense(int N, double* out, int* rows_of_col) {
  if (N < 1 || N > kMaxPieces) return -1;
  const std::vector<double>& t = dense_minco_table(N);
  if (out) std::copy(t.begin(), t.end(), out);
  if (rows_of_col)
    for (int j = 0; j < N + 5; j++) rows_of_col[j] = dense_row_of(j, N);
  return (long long)t.size();
}

// Diagnostic (host only): the banded LU factor of N pieces as the fallback / strict kernels read it, packed columns
// C[j][0..12] = band rows of column j (A(i,j) = C[j][i-j+6]), C[j][13] = 1/U(j,j); 6N x 16 doubles.
long long dftpav_debug_minco_band(int N, double* out) {
  if (N < 1 || N > kMaxPieces) return -1;
  if (out) {
    BandLU f(N);
    f.pack(out);
  }
  return (long long)6 * N * kLuLd;
}

const char* dftpav_last_error(const dftpav_handle* h) { return h ? h->err.c_str() : "null handle"; }

}  // extern "C"
