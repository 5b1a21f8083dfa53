// dftpav_device.cuh — shared host/device structures of the batched MINCO / L-BFGS solver (sm_100a).
//
// Mapping of the hot path onto the GPU:
//   * one warp owns one trajectory problem for its whole L-BFGS solve (no inter-warp communication);
//   * MINCO forward solve / adjoint solve: products with the constant matrix A^-1[:, R] (dense path, 32 lanes) or banded
//     substitutions with the constant LU factor (fallback, 2 lanes per gear segment); the table is staged in shared
//     memory once per CTA by a TMA bulk copy (cp.async.bulk);
//   * constraint-sample sweep: lane <-> sample, 16-sample tiles, plane data read coalesced from a tiled SoA
//     layout, fixed-order (deterministic) tile reduction of the 6x2 coefficient gradient in shared memory;
//   * L-BFGS two-loop recursion + Lewis-Overton line search: warp-shuffle dot products, (s,y) history ring in
//     global memory (L2-resident), no host synchronisation until the batch is done.
#pragma once
#include <cstdint>

namespace dftpav {

constexpr int kMaxSeg = 4;        // max gear segments per problem (compiled-in limit)
constexpr int kMaxPieces = 64;    // max pieces per segment
constexpr int kTile = 16;         // samples per tile (half warp)
constexpr int kLuLd = 16;         // doubles per packed LU column (13 band entries + 1/diag + pad)
constexpr int kStageLd = 17;      // leading dimension of the tile-reduction staging rows
constexpr int kObsCap = 16;       // obstacles per problem whose metadata is staged in shared memory (more: global reads)

// Problem "shape": everything that fixes control flow and shared-memory layout.  All problems of one kernel
// launch share one shape (the host groups the batch by shape).
struct ShapeDev {
  int S;                 // gear segments
  int n;                 // decision variables
  int rows;              // sum 6*N_i
  int pieces;            // sum N_i
  int H;                 // padded planes per sample (group-wide max)
  int Mpad;              // padded sample-storage length of a problem (multiple of 16)
  int K, Kd;             // traj_resolution, des_traj_resolution
  int N[kMaxSeg];        // pieces per segment
  int row0[kMaxSeg];     // first coefficient row of the segment
  int var0[kMaxSeg];     // offset of the segment's inner points in x
  int lu_off[kMaxSeg];   // offset (doubles) of the segment's packed LU table in the LU block
  int samp0[kMaxSeg];    // sample-storage offset of the segment
  int ntile[kMaxSeg];    // regular 16-sample tiles of the segment (including the padding tile below)
  int tpad[kMaxSeg];     // 1: an empty tile precedes the tiles of the last piece so that they share one sweep pass
  int lu_doubles;        // size of the LU block (all distinct N of this shape)
  int npad;              // n rounded up to a multiple of 4
  int ws_doubles;        // per-warp workspace size in doubles
  int band_masks_ok;     // 1 if the LU tables satisfy the compile-time sparsity masks of the banded passes
  int dense_minco;       // 1: the table block holds the dense matrices M = A^-1[:, R] (lu_off = their offsets), not LU columns
  int has_obs;           // the batch carries moving obstacles (obstacle metadata array in the warp workspace)
  int helpable;          // 1: regular sweep passes touch disjoint pieces (K <= 16 < Kd <= 32 or Kd <= 16): drain-phase helpers allowed
};

// Scalar parameters reaching the arithmetic (mirror of dftpav_params, device side).
struct ParamsDev {
  double wei_obs, wei_surround, wei_feas, wei_time;
  double surround_clearance, d_min;      // d_min = clearance + log(8)/alpha  (OPT.cpp:1336)
  double max_vel[2], max_cur[2], max_acc[2];  // [0] forward, [1] backward
  double non_sinv, mini_T, help_eps, min_speed, smooth_eps, cur_weight_scale, lse_alpha, cull_dist, cost_limit;
  double le[4][2];                        // ego / obstacle box vertices (OPT.cpp:1765-1768), clockwise
  double l1_half, l1_f3c, l1_f4c, l1_d2c, l1_d3c;  // constants of positiveSmoothedL1 (OPT.cpp:785-790), precomputed
  int gear_opt;
  int mem_size, past, max_iterations, max_linesearch;
  double delta, g_epsilon, min_step, max_step, f_dec_coeff, s_curv_coeff, cautious_factor, machine_prec;
};

// Device views of the uploaded batch.
struct BatchDev {
  // per problem (global problem id)
  const int* seg_base;          // [B] index of the problem's first segment
  const double* t_now;          // [B]
  const int* obs_ptr;           // [B+1] or nullptr
  const int* obs_idx;
  // per segment
  const int* singul;            // [G]
  const double* ini;            // [G][6] clamped
  const double* fin;            // [G][6] clamped
  // obstacle pool
  const double* obs_start;
  const double* obs_dur;
  const int* obs_piece_ptr;
  const double* obs_piece_dur;
  const double* obs_coeffs;     // [pieces][12], 2x6 column-major, highest power first
  const double* obs_uniform;    // [pool] common piece duration if all pieces of the obstacle have bit-identical durations, else NaN
  // decision vectors
  const double* x0;             // [B][n_stride]
  int n_stride, row_stride;
  // raw corridor planes as uploaded ([sample][h][4], normals not yet unit) and the hand-shake with the (possibly still
  // running) upload: the planes of problems [c*chunk, (c+1)*chunk) have arrived when ready[c] == epoch; a problem's tiled
  // SoA copy exists when laid[p] == epoch (written by the warp that first takes the problem)
  const double* planes_raw;
  const long long* samp_base;   // [B] first raw sample of the problem
  const long long* plane_ptr;   // [Q+1] plane prefix of the samples, or nullptr (uniform_H planes per sample)
  const int* ready;             // [chunks]
  int* laid;                    // [B]
  int uniform_H, chunk, epoch;
};

// One launch = one shape group.
struct GroupDev {
  const int* prob_id;           // [count] global problem ids of this group
  int count;
  double* planes;               // tiled SoA: [count][H][4][Mpad], filled by the kernels on first use (layout_problem)
  const double* lu;             // packed LU block of this shape
  const int* order;             // work list: order[w] = local index served w-th, or nullptr (input order)
  // longest-predicted-first queue that respects the arrival of the plane chunks (nullptr: the plain counter above).
  // The problems of upload chunk c (global problem id / chunk size) are listed in clist[cbeg[c] .. cbeg[c+1]) by
  // decreasing key; a fetching warp takes the best head among the chunks that have ARRIVED (BatchDev::ready).
  const int* clist;             // [count] local indices
  const int* cbeg;              // [n_chunks + 1]
  int* chead;                   // [n_chunks] entries already taken per chunk (zeroed before the launch)
  const double* key;            // [B] by global problem id: f(x0) without the corridor term (needs no plane)
  int n_chunks;
};

struct OutDev {
  double* x;        // [B][n_stride]
  double* cost;     // [B]
  int* status;      // [B]
  int* success;     // [B]
  int* iters;       // [B]
  int* evals;       // [B]
  double* coeffs;   // [B][row_stride][2]
  double* piece_dT; // [G]
  double* trace;    // optional [B][trace_cap][3] = (fx, step, ls) per L-BFGS iteration (diagnostic), or nullptr
  int trace_cap;
};

// Device view for the batched trajectory evaluator (dftpav_sample*): results of the last solve + segment directory.
struct SampleDev {
  const double* coeffs;         // [B][row_stride][2]
  const double* piece_dT;       // [G]
  const int* singul;            // [G]
  const int* seg_prob;          // [G] problem of the segment
  const int* seg_row0;          // [G] first coefficient row of the segment inside its problem
  const int* seg_N;             // [G] pieces
  const long long* seg_piece0;  // [G] pieces of all earlier segments (PolyTraj packing)
  int G, row_stride;
};

// Occupancy grid + vehicle box for the corridor generator (dftpav_corridor_*).
struct CorridorDev {
  const unsigned char* data;    // [size_y][size_x]
  const double* dl;             // [n_dl] accumulated ray-march offsets: 0, r/2, r/2 + r/2, ... (sequential sums, built on the host)
  int size_x, size_y, n_dl;
  unsigned char occupied;
  double origin_x, origin_y, resolution;
  double veh_width, veh_length, veh_d_cr, limit_bound;
};

// Device view of dftpav_frontend (the searched paths of KinoAstar after getKinoNode) for the front-end resampling kernels.
struct FrontendDev {
  int n_paths;
  const int* pt_ptr;
  const double* pts;
  const int* shot_ptr;
  const int* shot_index;
  const double* shot_time;
  const double* shot_length;
  const int* shot_s;
  const double* start_vel;
  const double* end_vel;
  double max_forward_vel, max_forward_acc, max_backward_vel, max_backward_acc, non_siguav;
};

struct HistDev {
  double* s;        // [slots][mem][npad_max]
  double* y;
  double* ys;       // [slots][mem]
  int ld;           // npad_max
  int mem;
};

}  // namespace dftpav
