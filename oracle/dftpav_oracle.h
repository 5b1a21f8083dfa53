/*
 * dftpav_oracle.h — CPU oracle for the Dftpav MINCO / L-BFGS hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * there only as the checker / the CPU arm.  The product path (dftpav_b200/csrc, the C-ABI of
 * include/dftpav_b200.h) never calls into this library and has no CPU fallback.
 *
 * PARITY STATUS: PINNED against the reference's own source.  The reference ships no golden vectors, KATs or
 * fixtures for this path and its build needs Eigen3, ROS1 and protobuf (all absent), but
 * its hot-path SOURCES compile here, unmodified and where they lie under /root/reference, against a test-only
 * Eigen-compatible layer (oracle/eigen_shim2) and inert stubs of the ROS / protobuf / RViz dependencies
 * (oracle/ref_stubs); recipe oracle/Makefile, outputs only into oracle/_ref/:
 *   - libref_optimizer.so = src/traj_optimizer.cpp + plan_manage/traj_optimizer.h + poly_traj_utils.hpp +
 *     traj_container.hpp + lbfgs.hpp, driven through the reference's public API (setParam, setSurroundTrajs,
 *     OptimizeTrajectory, getMinJerkOptPtr).  This restatement is BIT-IDENTICAL to it: cost, gradient and
 *     coefficients of costFunctionCallback at initial and perturbed points, and final x / cost / return code /
 *     success flag / iteration and evaluation counts of complete solves, on all five BASELINE configurations
 *     (tests/test_oracle_vs_reference_optimizer.py, live and against tests/golden/optimizer_reference.npz);
 *   - libref_minco.so (MinJerkOpt, BandedSystem, Piece, Trajectory) and libref_lbfgs.so (lbfgs.hpp alone, against
 *     the older minimal shim): bit-identical as well (tests/test_oracle_vs_reference_minco.py, test_oracle_lbfgs.py).
 * The caveat that remains: the Eigen layer evaluates every expression element by element in the written order, real
 * Eigen may vectorise reductions differently (rounding-level only; the algorithm is ulp-chaotic anyway).
 * Independent checks kept as well: central finite differences of every cost term, MINCO identities, banded LU vs
 * dense solve, an independent numpy restatement of the cost (tests/np_restatement.py).
 *
 * Input structs are those of include/dftpav_b200.h (the oracle exports the same batch format so
 * tests and the baseline timer are backend-agnostic).
 */
#ifndef DFTPAV_ORACLE_H
#define DFTPAV_ORACLE_H

#include "../include/dftpav_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One cost+gradient evaluation (costFunctionCallback, OPT.cpp:206-350) of problem `p` at x.
 * x and g have n(p) entries.  If coeffs != NULL it receives the rows*2 MINCO coefficients
 * (row-major pairs), if piece_dT != NULL the S per-piece durations.  terms[5] (optional) receives
 * {jerk, time, corridor, dynamic-obstacle, feasibility} cost parts.  Returns 0 or a DFTPAV_ERR_*. */
int oracle_eval(const dftpav_params* prm, const dftpav_batch* batch, int p, const double* x,
                double* f, double* g, double* coeffs, double* piece_dT, double* terms);

/* Per-problem sizes: n, coefficient rows, samples. */
int oracle_problem_dims(const dftpav_params* prm, const dftpav_batch* batch, int p, int* n,
                        int* rows, int* samples);

/* Full OptimizeTrajectory (OPT.cpp:7-202) for problems [p0, p1) using `threads` OpenMP threads
 * (each problem is solved by one thread, exactly like the single-threaded reference).
 * `out` arrays are indexed like the C-ABI's (absolute problem index, strides from
 * dftpav_batch_dims).  trace (optional): [B][trace_cap][3] = (fx, step, ls) per L-BFGS iteration. */
int oracle_optimize(const dftpav_params* prm, const dftpav_batch* batch, int p0, int p1,
                    int threads, const dftpav_result* out, double* trace, int trace_cap);

/* Generic L-BFGS restatement (LBFGS.hpp:276-390, 440-751) on a C callback; used to pin the
 * restatement against the reference header (oracle/_ref).  trace: [trace_cap][3]. */
typedef double (*oracle_fg_t)(void* user, const double* x, double* g, int n);
int oracle_lbfgs(int n, double* x, double* f, oracle_fg_t fg, void* user, const dftpav_params* prm,
                 int* iters, int* evals, double* trace, int trace_cap);

/* MINCO building blocks exposed for identity tests (MINCO.hpp:855-1095). */
/* Fills the dense 6N x 6N coefficient matrix (row-major) and the banded LU factor
 * (band storage ptr[(i-j+6)*6N + j], 13*6N doubles). */
int oracle_minco_matrix(int N, double* dense, double* lu_band);
/* generate(): inner points P (2 x (N-1) col-major), dT, head/tail 2x3 col-major -> c (6N x 2 row-major). */
int oracle_minco_generate(int N, const double* P, double dT, const double* head, const double* tail,
                          double* c);

/* The MinJerkOpt call sequence of one cost evaluation (generate, getTrajJerkCost, initSmGradCost, [gdC += extra_gdC,
 * gdT += extra_gdT], calGrads_PT): same signature as ref_minco_run in oracle/ref_minco_driver.cpp, which runs the
 * reference's own header.  Matrices 6N x 2 row-major, P 2x(N-1) / head / tail 2x3 column-major. */
int oracle_minco_run(int N, const double* P, double dT, const double* head, const double* tail,
                     const double* extra_gdC, double extra_gdT, double* c, double* energy, double* gdC0,
                     double* gdT0, double* gdP, double* gdHead, double* gdTail, double* gdT);
/* The trajectory of a generate() call sampled at t with the oracle's obstacle-trajectory evaluators (same output layout
 * as ref_traj_eval): pos(2) dsigma(2) ddsigma(2) R(4) Rdot(4) angle piece_index local_t. */
int oracle_traj_eval(int N, const double* P, double dT, const double* head, const double* tail, int singul, double t,
                     double* out);

/* Trajectory / Piece evaluators on coefficients c (6N x 2 row-major, row 6i+k = coefficient of t^k of piece i, every piece
 * of duration dT): out[12] = x y dsigma(2) ddsigma(2) angle curvature velocity acceleration steer lateral-acceleration at
 * time t (clamped to the total duration if clamp != 0); gs[6] (optional) = Trajectory::GetState's angle curvature velocity
 * acceleration steer, position x.  Same signature as ref_traj_state (oracle/ref_minco_driver.cpp, the reference's header). */
int oracle_traj_state(int N, const double* c, double dT, int singul, double t, int clamp, double wheel_base, double* out,
                      double* gs);

/* Synthetic-test analytic functions for oracle_lbfgs / the reference header: id 0 = Rosenbrock
 * chain, 1 = ill-conditioned quadratic, 2 = nonsmooth |x| mix.  Same code is compiled into
 * oracle/_ref so both sides evaluate bit-identical f,g. */
double oracle_testfunc(int id, const double* x, double* g, int n);

double oracle_testfunc_fg(void* user /* int* id */, const double* x, double* g, int n);

/* Problem handles: expose one problem's cost function as an oracle_fg_t so that an external optimizer
 * loop (the reference lbfgs.hpp in oracle/_ref) can drive it. */
void* oracle_problem_create(const dftpav_params* prm, const dftpav_batch* batch, int p);
void oracle_problem_free(void* h);
int oracle_problem_n(void* h);
void oracle_problem_initial_x(void* h, double* x);
double oracle_problem_fg(void* h, const double* x, double* g, int n);

int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
