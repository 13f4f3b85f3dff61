// TEST INFRASTRUCTURE ONLY.  The Newton / step-control restatement (xyce_b200/csrc/tran_driver.h) around a SCRIPTED
// backend: the norms every residual evaluation "measures" and the linear-solver return codes are given by the test,
// so that the control flow -- return codes of DampedNewton::converged_, step-size and order selection of
// OneStep::completeStep / rejectStep -- can be pinned on tables worked by hand from the reference source
// (N_NLS_DampedNewton.C:1191-1397, N_TIA_OneStep.C completeStep / rejectStep) instead of being compared with itself.
#include <cstring>
#include <vector>

#include "../../xyce_b200/csrc/tran_driver.h"

using namespace xb::sim;

namespace {
struct ScriptedBackend {
  // script of residual evaluations: index 0 = initial residual, k = after the k-th Newton update
  std::vector<double> n2, ninf, dxw; std::vector<int> devconv, lin;
  int evals = 0, solves = 0;
  double wrms = 0.0, inf = 1.0;
  int n() const { return 1; }
  void copy(int, int) {} void fill(int, double) {} void scale(int, double) {}
  void axpby(int, double, int, double, int) {} void axpy(int, double, int) {}
  double norm2(int) { return 0.0; } double norm_inf(int) { return inf; }
  double wmax_norm(int, int) { return 0.0; } double wrms_norm(int, int) { return wrms; }
  void sol_weights(int, double, double, int, int) {} void abs_weights(int, double, double, int) {}
  bool load_rhs(const Flags &, double) { return true; }
  void load_jacobian(double, double) {}
  int solve() { const int k = solves++; return k < (int)lin.size() ? lin[k] : 0; }
  void residual_and_norms(const ResidualForm &, NewtonNorms &out) {
    const int k = evals < (int)n2.size() ? evals : (int)n2.size() - 1;
    ++evals;
    out.rhs_norm2 = n2[k]; out.rhs_norm_inf = ninf[k]; out.dx_wmax = dxw[k]; out.devices_converged = devconv[k] != 0;
  }
  bool limiter_active() const { return false; }
  void accept_state() {} void record(double) {}
  void breakpoints(double, std::vector<double> &) {}
  double max_source_step(double) { return 1e99; }
};
}  // namespace

namespace xb { namespace sim {
// access to the driver's private step-control state (friend of TransientDriver)
struct DriverProbe {
  template <class D> static int newton(D &d, bool dc, int *iters) { const int st = d.newton_solve(dc); *iters = d.nIterations; return st; }
  template <class D> static void set_step_state(D &d, double t, double h, double last_h, int order, int nsteps, double est, double stop,
                                                double hmin, double hmax, int nef, int newton_status, int step_number) {
    d.currentTime = t; d.nextTime = t + h; d.currentTimeStep = h; d.lastTimeStep = last_h; d.currentOrder = order; d.numberOfSteps = nsteps;
    d.estOverTol = est; d.stopTime = stop; d.finalTime = stop; d.minTimeStep = hmin; d.maxTimeStep = hmax; d.nef = nef;
    d.newtonConvergenceStatus = newton_status; d.stepNumber = step_number; d.psi[0] = h; d.psi[1] = last_h; d.psi[2] = last_h;
  }
  template <class D> static void complete(D &d) { d.complete_step(); }
  template <class D> static bool reject(D &d) { return d.reject_step(); }
  template <class D> static void get(D &d, double *out6) {
    out6[0] = d.currentTimeStep; out6[1] = d.currentOrder; out6[2] = d.nextTime; out6[3] = d.currentTime; out6[4] = d.savedTimeStep; out6[5] = d.nef;
  }
};
}}

extern "C" {

// One DampedNewton::solve on the scripted norms.  Returns the status code; *iters = Newton steps taken.
int xbh_newton_script(int dc, int n_evals, const double *n2, const double *ninf, const double *dxw, const int *devconv,
                      int n_lin, const int *lin_rc, int max_newton, int *iters) {
  ScriptedBackend B;
  B.n2.assign(n2, n2 + n_evals); B.ninf.assign(ninf, ninf + n_evals); B.dxw.assign(dxw, dxw + n_evals); B.devconv.assign(devconv, devconv + n_evals);
  B.lin.assign(lin_rc, lin_rc + n_lin);
  TranParams P; P.tstop = 1.0; P.tstep = 1e-3;
  if (max_newton > 0) { P.maxNewtonStep = max_newton; P.dcMaxNewtonStep = max_newton; }
  TransientDriver<ScriptedBackend> d(B, P);
  DriverProbe::set_step_state(d, 0.0, 1e-3, 1e-3, 1, 0, 0.0, 1.0, 1e-15, 0.1, 0, 0, 1);
  return DriverProbe::newton(d, dc != 0, iters);
}

// OneStep::completeStep (accept = 1) or rejectStep (accept = 0) from a given step-control state.
// out6 = {new step, new order, nextTime, currentTime, savedTimeStep, nef}; returns rejectStep's "can continue".
int xbh_step_control(int accept, double t, double h, double last_h, int order, int nsteps, double est, double stop, double hmin,
                     double hmax, int nef, int newton_status, int max_order, double *out6) {
  ScriptedBackend B;
  B.n2 = {0}; B.ninf = {0}; B.dxw = {0}; B.devconv = {1};
  TranParams P; P.tstop = stop; P.tstep = h; P.maxOrder = max_order;
  TransientDriver<ScriptedBackend> d(B, P);
  DriverProbe::set_step_state(d, t, h, last_h, order, nsteps, est, stop, hmin, hmax, nef, newton_status, 5);
  int ok = 1;
  if (accept) DriverProbe::complete(d); else ok = DriverProbe::reject(d) ? 1 : 0;
  DriverProbe::get(d, out6);
  return ok;
}

}  // extern "C"
