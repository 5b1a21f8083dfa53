// ref_lbfgs_driver.cpp — C wrapper around the UNMODIFIED reference L-BFGS header
// (/root/reference/src/Plan/traj_planner/include/geo_utils2d/lbfgs.hpp), compiled in place against the
// test-only Eigen shim (oracle/eigen_shim).  Built by oracle/Makefile into oracle/_ref/libref_lbfgs.so.
// TEST INFRASTRUCTURE ONLY: used to pin the oracle's L-BFGS restatement and, through a C callback, to drive
// the oracle's cost function with the reference's own optimizer loop.  No reference source is copied.
#include <geo_utils2d/lbfgs.hpp>

#include <vector>

extern "C" {

typedef double (*ref_fg_t)(void* user, const double* x, double* g, int n);

struct RefCtx {
  ref_fg_t fg;
  void* user;
  long evals;
  std::vector<double>* trace;
  int trace_cap;
};

static double ref_eval(void* instance, const Eigen::VectorXd& x, Eigen::VectorXd& g) {
  RefCtx* c = reinterpret_cast<RefCtx*>(instance);
  c->evals++;
  return c->fg(c->user, x.data(), g.data(), x.size());
}

static int ref_progress(void* instance, const Eigen::VectorXd&, const Eigen::VectorXd&, const double fx,
                        const double step, const int k, const int ls) {
  RefCtx* c = reinterpret_cast<RefCtx*>(instance);
  if (c->trace && k - 1 < c->trace_cap) {
    (*c->trace)[3 * (k - 1)] = fx;
    (*c->trace)[3 * (k - 1) + 1] = step;
    (*c->trace)[3 * (k - 1) + 2] = (double)ls;
  }
  return 0;
}

// Same parameter block as OPT.cpp:127-134 builds; explicit so the test controls every field.
int ref_lbfgs(int n, double* x, double* f, ref_fg_t fg, void* user, int mem_size, int past, double g_epsilon,
              double min_step, double delta, int max_iterations, int with_progress, int* evals, double* trace,
              int trace_cap) {
  lbfgs::lbfgs_parameter_t prm;
  prm.mem_size = mem_size;
  prm.past = past;
  prm.g_epsilon = g_epsilon;
  prm.min_step = min_step;
  prm.delta = delta;
  prm.max_iterations = max_iterations;
  Eigen::VectorXd xv(n);
  for (int i = 0; i < n; i++) xv(i) = x[i];
  std::vector<double> tr((size_t)3 * (trace_cap > 0 ? trace_cap : 1), 0.0);
  RefCtx ctx{fg, user, 0, trace ? &tr : nullptr, trace_cap};
  double fx = 0.0;
  // the progress callback only observes (returns 0); OPT.cpp:163-164 passes NULL for both callbacks
  int ret = lbfgs::lbfgs_optimize(xv, fx, ref_eval, nullptr, with_progress ? ref_progress : nullptr, &ctx, prm);
  for (int i = 0; i < n; i++) x[i] = xv(i);
  *f = fx;
  if (evals) *evals = (int)ctx.evals;
  if (trace)
    for (int i = 0; i < 3 * trace_cap; i++) trace[i] = tr[i];
  return ret;
}

const char* ref_lbfgs_strerror(int err) { return lbfgs::lbfgs_strerror(err); }

}  // extern "C"
