// xyce_b200 -- host-side caller of the hot path: Newton solve and variable-step transient loop.
//
// This is the thin layer that sits directly on top of the loader in the reference (SURVEY.md rows
// a18, a20 and "next" row f2).  It is a from-scratch restatement of
//   Nonlinear::DampedNewton::solve / converged_ / updateWeights_   (src/NonlinearSolverPKG/N_NLS_DampedNewton.C:362-515, :1191-1397, :297-360)
//   TimeIntg::OneStep (trapezoid, variable order 1-2): updateCoeffs, obtainPredictor, obtainResidual,
//     obtainJacobian, initialize, completeStep, rejectStep, updateHistory
//                                                                  (src/TimeIntegrationPKG/N_TIA_OneStep.C:135-281, :469-515, :1390-1488, :1584-1669, :1684-1850, :1946-2110, :2124-2310)
//   TimeIntg::StepErrorControl::evaluateStepError, DataStore::setErrorWtVector / WRMS_errorNorm
//                                                                  (N_TIA_StepErrorControl.C:470-560, N_TIA_DataStore.C:1300-1520)
//   Analysis::Transient::doLoopProcess / takeAnIntegrationStep_     (src/AnalysisPKG/N_ANP_Transient.C:1184-1330, :3262-3308)
//   TimeIntg::Gear12 (BDF, variable order 1-2; `method = 8`): updateCoeffs, obtainPredictor, obtainResidual, obtainJacobian,
//     initialize, updateHistory, rejectStep, completeStep       (N_TIA_Gear12.C:131-190, :208-262, :477-500, :822-880, :1050-1100, :1183-1300, :1439-1560, :1639-1800)
//   TimeIntg::NoTimeIntegration::obtainResidual / obtainJacobian (DC operating point, `dcop = 1`; N_TIA_NoTimeIntegration.C:161-173, :291-298)
//     with DampedNewton in DC_OP mode (defaults N_NLS_NLParams.h:462-663; weights re-evaluated every iteration, N_NLS_DampedNewton.C:473-474)
//     and the initJct / initFix flags of SolverState (Core/N_DEV_SolverState.C:374-420)
// with the transient-mode defaults of N_NLS_NLParams.C:101-112 and N_TIA_TIAParams.C:79-111.
// The driver is a template over a Backend that owns the vectors and performs the loads, BLAS-1
// kernels and the linear solve: the product instantiates it with the CUDA backend (sim_gpu.cu); the
// oracle instantiates the same control flow around the *reference's* device code and Kundert Sparse
// (oracle/sim_ref.cpp) so that Newton iteration counts and waveforms can be compared.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace xb {
namespace sim {

// vector slots owned by the backend
enum Vec {
  vNextSol, vCurrSol, vF, vQ, vB, vFlim, vQlim, vRHS, vDX, vSolWt, vErrWt, vQErrWt, vXn0, vQn0,
  vXh0, vXh1, vXh2, vQh0, vQh1, vQh2, vNewtCorr, vQNewtCorr, vTmp, kNumVec
};

// ---- independent-source waveforms (src/DeviceModelPKG/Core/N_DEV_SourceData.C) ----
// type 0 DC          p = {v}
//      1 PULSE       p = {v1, v2, td, tr, tf, pw, per}            PulseData::updateSource :1168-1248, getBreakPoints :1442-1500
//      2 SIN         p = {v0, va, freq, td, theta, phase_deg}     SinData::updateSource :475-496
//      3 EXP         p = {v1, v2, td1, tau1, td2, tau2}           ExpData::updateSource :811-835
//      4 SFFM        p = {v0, va, fc, mdi, fs}                    SFFMData::updateSource :2908-2922
//      5 PWL         p = {td, offset, count, repeat, repeattime}; the (time, value) pairs sit at pwl[2 offset ..)
//                                                                 PWLinData::updateSource :1770-1886, getBreakPoints :2044-2110
// bpTol = the break-point tolerance of the step control (2 minTimeStep, N_TIA_StepErrorControl.C:756): the PULSE corner
// tests are tolerant to it exactly as the reference's are.
inline double pulse_value(const double *p, double t, double bpTol = 0.0) {
  const double V1 = p[0], V2 = p[1], TD = p[2], TR = p[3], TF = p[4], PW = p[5], PER = p[6];
  double time = t - TD;
  if (time > PER && PER != 0.0) time -= PER * std::floor(time / PER);
  if (time <= 0 || (time > (TR + PW + TF) && (std::fabs(time - (TR + PW + TF)) > bpTol))) return V1;
  if ((time > TR && std::fabs(time - TR) > bpTol) && (time < (TR + PW) || std::fabs(time - (TR + PW)) < bpTol)) return V2;
  if (time > 0 && (time < TR || std::fabs(time - TR) < bpTol)) return TR != 0.0 ? V1 + (V2 - V1) * time / TR : V1;
  return TF != 0.0 ? V2 + (V1 - V2) * (time - (TR + PW)) / TF : V2;
}

// SPICE SIN(v0 va freq td theta phase_deg) (SinData, N_DEV_SourceData.C)
inline double sin_value(const double *p, double t) {
  const double v0 = p[0], va = p[1], f = p[2], td = p[3], theta = p[4], ph = p[5];
  const double two_pi = 6.283185307179586;
  if (t < td) return v0 + va * std::sin(two_pi * ph / 360.0);
  return v0 + va * std::sin(two_pi * (f * (t - td) + ph / 360.0)) * std::exp(-(t - td) * theta);
}
inline double exp_value(const double *p, double t) {
  const double V1 = p[0], V2 = p[1], TD1 = p[2], TAU1 = p[3], TD2 = p[4], TAU2 = p[5];
  if (t <= TD1) return V1;
  if (t <= TD2) return V1 + (V2 - V1) * (1 - std::exp(-(t - TD1) / TAU1));
  return V1 + (V2 - V1) * (1 - std::exp(-(t - TD1) / TAU1)) + (V1 - V2) * (1 - std::exp(-(t - TD2) / TAU2));
}
inline double sffm_value(const double *p, double t) {
  const double mpi = 3.14159265358979323846;
  return p[0] + p[1] * std::sin((2 * mpi * p[2] * t) + p[3] * std::sin(2 * mpi * p[4] * t));
}
inline double pwl_value(const double *p, const double *pwl, double t) {
  const double TD = p[0];
  const int NUM = (int)p[2];
  const bool REPEAT = p[3] != 0.0;
  const double REPEATTIME = p[4];
  const double *tv = pwl + 2 * (size_t)p[1];
  if (!(t >= TD) || NUM <= 0) return 0.0;
  double time = t - TD, time1, time2, voltage1, voltage2;
  int loc = 0;
  if (time <= tv[2 * (NUM - 1)]) {
    for (int i = 0; i < NUM; ++i) if (time < tv[2 * i]) { loc = i; break; }
    if (loc == 0) { time1 = 0.0; voltage1 = 0.0; } else { time1 = tv[2 * (loc - 1)]; voltage1 = tv[2 * (loc - 1) + 1]; }
    time2 = tv[2 * loc]; voltage2 = tv[2 * loc + 1];
  } else if (!REPEAT) {
    time1 = 0.0; time2 = 1.0; voltage1 = voltage2 = tv[2 * (NUM - 1) + 1];
  } else {
    const double looptime = tv[2 * (NUM - 1)] - REPEATTIME;
    time -= tv[2 * (NUM - 1)];
    time -= looptime * std::floor(time / looptime);
    time += REPEATTIME;
    for (int i = 0; i < NUM; ++i) if (time < tv[2 * i]) { loc = i; break; }
    if (time == REPEATTIME) { time1 = 0.0; time2 = 1.0; voltage1 = voltage2 = tv[2 * (NUM - 1) + 1]; }
    else {
      if (loc == 0) { time1 = REPEATTIME; voltage1 = tv[2 * (NUM - 1) + 1]; } else { time1 = tv[2 * (loc - 1)]; voltage1 = tv[2 * (loc - 1) + 1]; }
      time2 = tv[2 * loc]; voltage2 = tv[2 * loc + 1];
    }
  }
  if (time1 == time2) return voltage2;
  const double length = time2 - time1;
  double v = (time2 - time) * voltage1 / length;
  v += (-time1 + time) * voltage2 / length;
  return v;
}
inline double source_value(int type, const double *p, double t, const double *pwl = nullptr, double bpTol = 0.0) {
  switch (type) {
    case 1: return pulse_value(p, t, bpTol);
    case 2: return sin_value(p, t);
    case 3: return exp_value(p, t);
    case 4: return sffm_value(p, t);
    case 5: return pwl ? pwl_value(p, pwl, t) : 0.0;
    default: return p[0];
  }
}
// Break points a source announces at circuit time t (Loader::getBreakPoints -> SourceData::getBreakPoints): the corners
// of the current and the next PULSE period, the PWL points (the current and the next repetition when it repeats).
inline void source_breakpoints(int type, const double *p, const double *pwl, double t, std::vector<double> &out) {
  if (type == 1) {
    const double TD = p[2], TR = p[3], TF = p[4], PW = p[5], PER = p[6];
    double time = t - TD, basetime = 0.0;
    if (time >= PER && PER != 0.0) basetime = PER * (double)(int)std::floor(time / PER);
    out.push_back(basetime + TD); out.push_back(basetime + TD + TR); out.push_back(basetime + TD + TR + PW); out.push_back(basetime + TD + TR + PW + TF);
    if (PER != 0.0) {
      out.push_back(basetime + TD + PER); out.push_back(basetime + TD + PER + TR); out.push_back(basetime + TD + PER + TR + PW);
      out.push_back(basetime + TD + PER + TR + PW + TF); out.push_back(basetime + TD + PER + PER);
    }
  } else if (type == 5 && pwl) {
    const double TD = p[0], REPEATTIME = p[4];
    const int NUM = (int)p[2];
    const double *tv = pwl + 2 * (size_t)p[1];
    const double time = t - TD;
    if (NUM <= 0) return;
    if (p[3] != 0.0 && time >= tv[2 * (NUM - 1)]) {
      const double looptime = tv[2 * (NUM - 1)] - REPEATTIME;
      const double loopBaseTime = looptime * (1.0 + std::floor((time - tv[2 * (NUM - 1)]) / looptime));
      for (int i = 0; i < NUM; ++i) if (tv[2 * i] >= REPEATTIME) out.push_back(tv[2 * i] + loopBaseTime + TD);
    } else {
      for (int i = 0; i < NUM; ++i) out.push_back(tv[2 * i] + TD);
    }
  }
}
// SourceData::getMaxTimeStepSize (:242, PulseData :1518-1537): a PULSE source caps the step at a tenth of its period
// (of its delay while still in the delay); everything else imposes nothing.  Values <= 0 are ignored by the caller.
inline double source_max_step(int type, const double *p, double t) {
  if (type == 1) return t < p[2] ? 0.1 * p[2] : 0.1 * p[6];
  return 1.0e99;
}

struct TranParams {
  double tstop = 0, tstep = 0, delmax = 0;
  // Newton (transient mode defaults)
  int maxNewtonStep = 20;
  double deltaXTol = 0.33, absTol = 1e-6, relTol = 1e-2, RHSTol = 1e-2, smallUpdateTol = 1e-6;
  int enforceDeviceConv = 1;
  // time integration
  double relErrorTol = 1e-3, absErrorTol = 1e-6, errTolAcceptance = 1.0;
  int maxOrder = 2, minOrder = 1;
  int maxSteps = 1000000;
  int method = 7;      // .OPTIONS TIMEINT METHOD: 7 = trapezoid (OneStep, the reference default), 8 = Gear (Gear12)
  int dcop = 0;        // 0 = start from x0 as given (.TRAN ... NOOP / UIC), 1 = DC operating point from x0 first
  // Newton, DC_OP mode defaults (NLParams constructor)
  int dcMaxNewtonStep = 200;
  double dcDeltaXTol = 1.0, dcAbsTol = 1e-12, dcRelTol = 1e-3, dcRHSTol = 1e-6;
  // Verification mode (used by the oracle side of the at-size tests): follow a given sequence of accepted steps
  // -- step size and integration order of each -- instead of the driver's own step-size / order selection and
  // LTE accept test.  Everything else (coefficients, predictor, Newton solve, history updates) runs unchanged, so
  // a sub-circuit can be integrated on exactly the time points a larger run chose.  Empty = normal operation.
  std::vector<double> replay_h;
  std::vector<int> replay_order;
};

struct StepRecord { double t, h; int newton_iters, order, status; };

struct TranStats {
  int accepted = 0, rejected = 0, newton_total = 0, jacobian_loads = 0, residual_loads = 0, linear_solves = 0;
  int failed = 0;
  int dcop_newton = 0, dcop_status = 0;
  int breakpoints = 0;      // accepted steps that landed on a source break point
};

// Backend concept:
//   int n(); void copy(Vec dst, Vec src); void fill(Vec, double); void scale(Vec, double);
//   void axpby(Vec dst, double a, Vec x, double b, Vec y);          dst = a x + b y
//   void axpy(Vec dst, double a, Vec x);                              dst += a x
//   double norm2(Vec); double norm_inf(Vec); double wmax_norm(Vec x, Vec w); double wrms_norm(Vec x, Vec w);
//   void sol_weights(Vec dst, double rel, double abs, Vec a, Vec b);  dst_i = rel*max(|a_i|,|b_i|) + abs
//   void abs_weights(Vec dst, double rel, double abs, Vec a);         dst_i = rel*|a_i| + abs
//   bool load_rhs(const Flags&, double time);    evaluates devices at vNextSol: fills vF, vQ, vB, vFlim, vQlim
//   void load_jacobian(double qscalar, double fscalar);               J = qscalar dQdx + fscalar dFdx
//   int  solve();                                                      J vDX = vRHS   (0 ok)
//   void residual_and_norms(const ResidualForm &f, NewtonNorms &out);
//        form 0 (OneStep):  vRHS = -[(vQ - vQh0) inv_h + fs (vF - vB) (+ 1/2 vQh2)] (+ qlim_coef vQlim + fs vFlim)
//        form 1 (Gear12):   vRHS = -[(a0 vQ + a1 vQh0 (+ a2 vQh1)) inv_h + vF - vB] (+ qlim_coef vQlim + vFlim)
//        form 2 (DC):       vRHS = -(vF - vB) (+ vFlim)
//        element by element in the operation order of the vector-update sequence it replaces; then ||RHS||_2,
//        ||RHS||_inf, max |vDX / vSolWt| and the AND of the devices' convergence flags
//   bool limiter_active();
//   void breakpoints(double t, std::vector<double> &out);   break points the sources announce at time t
//   double max_source_step(double t);                        smallest step cap of the sources (<= 0: none)
//   void accept_state();                                               curr state/store <- next
//   void record(double t);                                             sample probes
// everything DampedNewton::converged_ looks at after one residual evaluation, fetched in one go
struct NewtonNorms { double rhs_norm2 = 0, rhs_norm_inf = 0, dx_wmax = 0; bool devices_converged = true; };

struct ResidualForm {
  int form = 0;                     // 0 OneStep, 1 Gear12, 2 DC
  double inv_h = 0, fs = 1, qlim_coef = 0, a0 = 1, a1 = -1, a2 = 0;
  bool order2 = false, limiter = false;
};

struct Flags {
  int dcop = 0, tranop = 0, transient = 1, initTran = 0, newtonIter = 0, initJct = 0, initFix = 0;
  double currTimeStep = 0, lastTimeStep = 0;
  int beginIntegration = 0;   // SolverState::beginIntegrationFlag_ (first step out of a break point, t = 0 included)
  double bpTol = 0;       // break-point tolerance, for the sources' corner tests
};

struct DriverProbe;      // test access to the step-control state (tests/host_mirror/driver_host.cpp)

template <class Backend>
class TransientDriver {
  friend struct DriverProbe;
 public:
  TransientDriver(Backend &b, const TranParams &p) : B(b), P(p) {}
  std::vector<StepRecord> steps;
  TranStats stats;

  // x(0) must already be in vNextSol and vCurrSol ("NOOP / UIC" start, as .TRAN ... NOOP does).
  int run() {
    const double machEps = 2.220446049250313e-16;
    // --- StepErrorControl state (N_TIA_StepErrorControl.C:100-170 defaults) ---
    initialTime = 0.0; currentTime = 0.0; finalTime = P.tstop; stopTime = finalTime;
    startingTimeStep = P.tstep;
    minTimeStep = (finalTime - initialTime) * 4.0 * machEps;
    bpTol = 2.0 * minTimeStep;                      // StepErrorControl::updateBreakPoints (:756)
    update_max_time_step();
    update_stop_time();
    beginningIntegration = true; stepAttemptStatus = true; stepNumber = 0; iNumCalls = 0;
    nef = 0;
    if (P.dcop) {      // Transient::doInit: DC operating point, then the solution becomes the current one
      const int st = newton_solve(true);
      stats.dcop_status = st; stats.dcop_newton = nIterations;
      if (st <= 0) { stats.failed = 1; return 4; }
      B.copy(vCurrSol, vNextSol);
      B.accept_state();
      iNumCalls = 0;
    }
    // initial load: Q, F, B at x(0)
    Flags fl; fl.initTran = 1; fl.newtonIter = 0; fl.currTimeStep = startingTimeStep; fl.bpTol = bpTol;
    B.load_rhs(fl, 0.0); ++stats.residual_loads;
    set_error_weights();
    B.fill(vQh1, 0.0);
    initialize_integrator();          // Transient::doInit
    B.record(0.0);
    while (!(currentTime >= finalTime - minTimeStep)) {
      if ((int)steps.size() >= P.maxSteps) return 3;
      if (replay() && replayIdx_ >= P.replay_h.size()) break;
      if (beginningIntegration && stepAttemptStatus) { update_max_time_step(); initialize_integrator(); }
      if (replay() && !beginningIntegration) {      // the recorded step instead of the one completeStep selected
        currentTimeStep = P.replay_h[replayIdx_];
        currentOrder = P.replay_order[replayIdx_];
        nextTime = currentTime + currentTimeStep;
        currentTimeStepRatio = currentTimeStep / lastTimeStep;
        currentTimeStepSum = currentTimeStep + lastTimeStep;
      }
      update_coeffs();
      // predictor (OneStep::obtainPredictor)
      if (gear()) {       // Gear12::obtainPredictor: sum over i = 0..order of beta_i * history_i
        B.fill(vXn0, 0.0); B.fill(vQn0, 0.0);
        for (int i = 0; i <= currentOrder; ++i) {
          B.axpy(vXn0, beta[i], i == 0 ? vXh0 : (i == 1 ? vXh1 : vXh2));
          B.axpy(vQn0, beta[i], i == 0 ? vQh0 : vQh1);     // qHistory[2] is never filled by Gear12::updateHistory (stays 0)
        }
      } else {
        B.copy(vXn0, vXh0); B.copy(vQn0, vQh0);
        for (int i = 1; i <= currentOrder; ++i) B.axpy(vXn0, beta[i], i == 1 ? vXh1 : vXh2);
      }
      B.copy(vNextSol, vXn0);
      const int status = newton_solve(false);
      // stepLinearCombo + evaluateStepError
      B.axpby(vNewtCorr, 1.0, vNextSol, -1.0, vXn0);
      B.axpby(vQNewtCorr, 1.0, vQ, -1.0, vQn0);
      bool ok = status > 0;
      newtonConvergenceStatus = status;
      const bool testError = ok && stepNumber >= 1 && !beginningIntegration;
      if (testError) {
        estOverTol = ck * B.wrms_norm(vNewtCorr, vErrWt);
        ok = estOverTol <= P.errTolAcceptance || replay();
      }
      stepAttemptStatus = ok;
      StepRecord rec{nextTime, currentTimeStep, nIterations, currentOrder, status};
      if (ok) {
        ++replayIdx_;
        complete_step();
        B.copy(vCurrSol, vNextSol);
        B.accept_state();
        ++stepNumber; ++stats.accepted;
        // a step that lands on a break point (not the final time) restarts the integration there: order 1, fresh
        // initial step, no LTE test on the first step (Transient::doHandlePredictor / processSuccessfulStep,
        // N_ANP_Transient.C:1900-1948)
        if (std::fabs(currentTime - stopTime) <= bpTol && std::fabs(currentTime - finalTime) > bpTol) {
          beginningIntegration = true; ++stats.breakpoints;
          update_stop_time();
        } else {
          beginningIntegration = false;
        }
        set_error_weights();
        B.record(currentTime);
      } else {
        ++stats.rejected;
        rec.status = status > 0 ? -100 : status;    // -100: rejected by the LTE test
        if (!reject_step()) { steps.push_back(rec); stats.failed = 1; return 2; }
      }
      steps.push_back(rec);
    }
    return 0;
  }

 private:
  Backend &B;
  TranParams P;
  // StepErrorControl / OneStep members (same names as the reference where possible)
  double initialTime, currentTime, nextTime, stopTime, lastTime;
  double currentTimeStep, lastTimeStep, currentTimeStepRatio, currentTimeStepSum, savedTimeStep = 0;
  double startingTimeStep, minTimeStep, maxTimeStep;
  double finalTime = 0.0, bpTol = 0.0;
  std::vector<double> bps_;

  // StepErrorControl::updateMaxTimeStep (:879-940): DELMAX or a tenth of the run, capped by the sources' own limits
  void update_max_time_step() {
    maxTimeStep = (P.delmax > 0.0) ? P.delmax : 0.1 * (finalTime - initialTime);
    const double maxDevStep = B.max_source_step(currentTime);
    if (maxDevStep > 0.0) maxTimeStep = std::min(maxTimeStep, maxDevStep);
  }
  // StepErrorControl::updateBreakPoints + updateStopTime (:742-850, :370-420): the stop time is the first break point
  // after the current time, or the final time
  void update_stop_time() {
    bps_.clear();
    B.breakpoints(currentTime, bps_);
    std::sort(bps_.begin(), bps_.end());
    stopTime = finalTime;
    for (double b : bps_) if (b > currentTime + bpTol && b < finalTime) { stopTime = b; break; }
  }

  double psi[3] = {0, 0, 0}, beta[3] = {1, 0, 0}, alpha[3] = {1, -1, 0}, alphas = -1.0, ck = 1.0, estOverTol = 0.0;
  bool gear() const { return P.method == 8; }
  bool replay() const { return !P.replay_h.empty(); }
  size_t replayIdx_ = 0;
  int currentOrder = 1, usedOrder = 1, numberOfSteps = 0, nef = 0, stepNumber = 0, nIterations = 0;
  int newtonConvergenceStatus = 0, iNumCalls = 0;
  bool beginningIntegration = true, stepAttemptStatus = true;
  const double tolAimFac = 0.5, r_min = 0.25, r_max = 0.9, r_hincr_test = 2.0, r_hincr = 2.0;
  const double h0_safety = 2.0, h0_max_factor = 0.005;
  const int maxNumfail = 15;

  void set_error_weights() {      // DataStore::setErrorWtVector, newLte = 1 ("point global")
    const double m = B.norm_inf(vCurrSol);
    B.fill(vErrWt, P.relErrorTol * m + P.absErrorTol);
    B.abs_weights(vQErrWt, P.relErrorTol, P.absErrorTol, vQ);
  }

  void initialize_integrator() {  // OneStep::initialize
    const double time_to_stop = stopTime - currentTime;
    // the reference measures qHistory[1] as it stands on entry (zero before the very first call, the
    // previous call's -h (F - B) afterwards: Transient::doInit calls initialize once before the loop)
    const double dnorm_q = B.wrms_norm(vQh1, vQErrWt);
    double h;
    if (dnorm_q > 0.0) {
      if (currentTime == initialTime) h = std::min(h0_max_factor * std::fabs(time_to_stop), std::sqrt(2.0) / (h0_safety * dnorm_q));
      else h = 0.1 * std::min(savedTimeStep, std::fabs(time_to_stop));
    } else {
      if (currentTime == initialTime) h = h0_max_factor * std::fabs(time_to_stop);
      else h = 0.1 * std::min(savedTimeStep, std::fabs(time_to_stop));
    }
    if (startingTimeStep > 0.0 && currentTime == initialTime) h = std::min(startingTimeStep, h);
    if (replay() && replayIdx_ < P.replay_h.size()) h = P.replay_h[replayIdx_];
    if (currentTime != initialTime) currentTimeStep = std::min(currentTimeStep, h);
    else currentTimeStep = h;
    currentTimeStep = std::max(currentTimeStep, minTimeStep);
    currentTimeStep = std::min(currentTimeStep, maxTimeStep);
    currentTimeStepRatio = 1.0;
    currentTimeStepSum = 2.0 * currentTimeStep;
    lastTimeStep = currentTimeStep;
    nextTime = currentTime + currentTimeStep;
    B.copy(vXh0, vCurrSol);
    B.copy(vQh0, vQ);
    if (gear()) {         // Gear12::initialize (:1262-1266)
      B.copy(vXh1, vCurrSol);
      B.copy(vQh1, vQ);
    } else {
      B.fill(vXh1, 0.0);
      B.axpby(vQh1, 1.0, vF, -1.0, vB);
      B.scale(vQh1, -currentTimeStep);
    }
    numberOfSteps = 0; currentOrder = 1; usedOrder = 1;
    psi[0] = currentTimeStep;
    nef = 0;
  }

  void update_coeffs() {          // OneStep::updateCoeffs
    const double t1 = currentTimeStep;
    if (currentOrder == 2) psi[2] = psi[1];
    psi[1] = psi[0];
    psi[0] = t1;
    beta[0] = 1.0; alphas = -1.0;
    if (gear()) {         // Gear12::updateCoeffs (:1072-1100)
      if (currentOrder == 2) {
        beta[2] = t1 / psi[2] * (t1 + psi[1]) / (psi[1] + psi[2]);
        beta[1] = -t1 / psi[1] - beta[2] * (psi[1] + psi[2]) / psi[1];
        beta[0] = 1.0 - beta[2] - beta[1];
        alpha[2] = -t1 / psi[1] * t1 / (2 * t1 + psi[1]);
        alpha[1] = 1 - alpha[2];
        alpha[0] = -alpha[1] - alpha[2] * (1 + psi[1] / t1);
        alpha[2] = alpha[2] / alpha[0];
        alpha[1] = alpha[1] / alpha[0];
        alpha[0] = -1 / alpha[0];
        ck = currentTimeStep / (t1 + psi[1] + psi[2]);
      } else {
        beta[0] = 1.0 + t1 / psi[1];
        beta[1] = -t1 / psi[1];
        alpha[0] = 1.0; alpha[1] = -1.0;
        ck = currentTimeStep / (t1 + psi[1]);
      }
      return;
    }
    if (currentOrder == 2) {
      const double t2 = psi[1];
      beta[1] = t1 / t2 + (t1 / t2) * (t1 / t2) / 2;
      beta[2] = -1.0 * t1 * t1 / t2 / psi[2] / 2;
      ck = (currentTimeStep / currentTimeStepSum) / 3.0;
    } else {
      beta[1] = t1 / psi[1];
      ck = currentTimeStep / currentTimeStepSum;
    }
  }

  // residual of OneStep::obtainResidual: RHS = -[(Q - q0)/h + f (F - B) (+ 1/2 qHistory[2])] + limiter terms
  // Gear12::obtainResidual: RHS = -[(a0 Q + a1 qHistory[0] (+ a2 qHistory[1]))/h + F - B] + (a0/h) dQdxdVp + dFdxdVp
  // NoTimeIntegration::obtainResidual: RHS = -(F - B) + dFdxdVp
  double residual(const Flags &fl, bool dc) {
    B.load_rhs(fl, dc ? 0.0 : nextTime); ++stats.residual_loads;
    ResidualForm f;
    f.limiter = B.limiter_active();
    if (dc) {
      f.form = 2;
    } else if (gear()) {
      f.form = 1; f.inv_h = 1.0 / currentTimeStep; f.order2 = currentOrder == 2;
      f.a0 = alpha[0]; f.a1 = alpha[1]; f.a2 = alpha[2]; f.qlim_coef = alpha[0] / currentTimeStep;
    } else {
      f.form = 0; f.inv_h = 1.0 / currentTimeStep; f.fs = (currentOrder == 2) ? 0.5 : 1.0; f.order2 = currentOrder == 2;
      f.qlim_coef = -alphas / currentTimeStep;
    }
    B.residual_and_norms(f, nn);
    return nn.rhs_norm2;
  }
  NewtonNorms nn;
  int stagCount_ = 0;
  double tmpConvRate_ = 0.0;
  const bool trace_ = std::getenv("XB_TRAN_TRACE") != nullptr;     // diagnostics: one line per Newton iteration on stderr

  // DampedNewton::solve with FULL search (step length 1); dc = DC_OP mode on NoTimeIntegration
  int newton_solve(bool dc) {
    Flags fl;
    if (dc) { fl.dcop = 1; fl.tranop = 1; fl.initJct = 1; fl.currTimeStep = 0.0; }
    else { fl.initTran = (stepNumber == 0); fl.currTimeStep = currentTimeStep; fl.lastTimeStep = lastTimeStep; }
    fl.beginIntegration = beginningIntegration;
    fl.bpTol = bpTol;
    const int maxNewtonStep = dc ? P.dcMaxNewtonStep : P.maxNewtonStep;
    const double deltaXTol = dc ? P.dcDeltaXTol : P.deltaXTol, RHSTol = dc ? P.dcRHSTol : P.RHSTol;
    const double relTol = dc ? P.dcRelTol : P.relTol, absTol = dc ? P.dcAbsTol : P.absTol;
    int nlStep = 0;
    fl.newtonIter = 0;
    double normRHS = residual(fl, dc);
    double normRHS_old = normRHS, normRHS_init = normRHS;
    if (!dc) B.sol_weights(vSolWt, relTol, absTol, vNextSol, vCurrSol);     // updateWeights_ (transient: once per solve)
    int status = 0;
    int &count = stagCount_;            // DampedNewton::count / tmpConvRate are class members (N_NLS_DampedNewton.h:199-203):
    double &tmpConvRate = tmpConvRate_; // they persist from one solve() to the next
    const double fs = (currentOrder == 2) ? 0.5 : 1.0;
    while (status == 0) {
      ++nlStep;
      if (dc) B.load_jacobian(1.0e-20, 1.0);
      else if (gear()) B.load_jacobian(alpha[0] / currentTimeStep, 1.0);
      else B.load_jacobian(-alphas / currentTimeStep, fs);
      ++stats.jacobian_loads;
      const int lin = B.solve(); ++stats.linear_solves;
      B.axpy(vNextSol, 1.0, vDX);
      if (dc) {           // updateWeights_ every iteration outside TRANSIENT mode (:473-474, :297-313)
        if (iNumCalls == 0 && B.norm_inf(vNextSol) <= 2.2250738585072014e-308) B.fill(vSolWt, relTol + absTol);
        else B.sol_weights(vSolWt, relTol, absTol, vNextSol, vCurrSol);
      }
      fl.newtonIter = nlStep;
      if (dc) { fl.initJct = 0; fl.initFix = nn.devices_converged ? 0 : 1; }     // SolverState.C:387-420
      normRHS = residual(fl, dc);
      // ---- converged_ ----
      if (lin != 0) { status = -9; break; }
      if (P.enforceDeviceConv) {
        const bool conv = nn.devices_converged;
        if (trace_ && !conv) std::fprintf(stderr, "NEWTON t=%.9e it=%d devices not converged ||rhs||2=%.17g\n", nextTime, nlStep, normRHS);
        if (!conv && nlStep < maxNewtonStep) continue;
        if (!conv && nlStep >= maxNewtonStep) { status = -1; break; }
      }
      if (!(normRHS == normRHS)) { status = -6; break; }
      const double maxNormRHS = nn.rhs_norm_inf;
      const double wtNormDX = nn.dx_wmax;
      if (normRHS < 2.220446049250313e-16) { status = 1; break; }     // normTooSmall
      const double normRHS_rel = normRHS / normRHS_init;
      const double resConvRate = normRHS / normRHS_old;
      normRHS_old = normRHS;
      const double updateSize = wtNormDX;
      if (trace_) std::fprintf(stderr, "NEWTON t=%.9e it=%d ||rhs||2=%.17g ||rhs||inf=%.17g ||dx||w=%.17g devconv=%d\n", nextTime,
                               nlStep, normRHS, maxNormRHS, updateSize, (int)nn.devices_converged);
      if (maxNormRHS <= RHSTol && updateSize <= deltaXTol) { status = 2; break; }
      // "near converged" and the stagnation test exist in TRANSIENT mode only (N_NLS_DampedNewton.C:1316-1323, :1341-1358)
      if (!dc && nlStep >= maxNewtonStep && normRHS_rel <= 0.9 && resConvRate <= 1.0) { status = 3; break; }
      if (updateSize <= P.smallUpdateTol) { status = 4; break; }
      if (nlStep >= maxNewtonStep) { status = -1; break; }
      if (resConvRate > 0.5 * 1.7976931348623157e308) { status = -2; break; }
      if (!dc && std::fabs(resConvRate - 1.0) <= 1.0e-3) {
        if (count == 0 || resConvRate < tmpConvRate) tmpConvRate = resConvRate;
        ++count;
      } else {
        count = 0;
      }
      if (!dc && count == 5) {
        count = 0;
        status = (normRHS_rel < 0.9 && tmpConvRate <= 1.0) ? 3 : -3;
        break;
      }
    }
    nIterations = nlStep;
    stats.newton_total += nlStep;
    ++iNumCalls;
    return status;
  }

  void update_history() {         // OneStep::updateHistory / Gear12::updateHistory
    if (gear()) {
      if (currentOrder == 2) B.copy(vXh2, vXh1);
      B.copy(vQh1, vQh0);
      B.copy(vXh1, vXh0);
      B.copy(vXh0, vNextSol);
      B.copy(vQh0, vQ);
      return;
    }
    if (currentOrder == 2) {
      B.copy(vXh2, vXh1);
      B.axpby(vQh2, 1.0, vF, -1.0, vB);
    }
    B.axpby(vXh1, 1.0, vNextSol, -1.0, vXh0);
    B.axpby(vQh1, 1.0, vQ, -1.0, vQh0);
    B.copy(vXh0, vNextSol);
    B.copy(vQh0, vQ);
  }

  void complete_step() {          // OneStep::completeStep (LOCAL_TRUNCATED_ESTIMATES)
    ++numberOfSteps;
    nef = 0;
    lastTime = currentTime;
    currentTime = nextTime;
    double newTimeStep = currentTimeStep;
    lastTimeStep = currentTimeStep;
    usedOrder = currentOrder;
    double rr = tolAimFac / (estOverTol + 0.0001);
    rr = std::pow(rr, 1.0 / (currentOrder + 1.0));
    if (numberOfSteps >= 2 && P.maxOrder == 2) {
      if (currentOrder == 1) {
        currentOrder = 2;
        rr = tolAimFac / (estOverTol + 0.0001);
        rr = std::pow(rr, 1.0 / (currentOrder + 1.0));
        if (rr <= 1.05) currentOrder = P.minOrder;
      }
    }
    if (rr >= r_hincr_test) { rr = r_hincr; newTimeStep = rr * currentTimeStep; }
    else if (rr <= 1) { rr = std::max(r_min, std::min(r_max, rr)); newTimeStep = rr * currentTimeStep; }
    if (replay() && replayIdx_ < P.replay_order.size()) currentOrder = P.replay_order[replayIdx_];
    update_history();     // with the order selected for the NEXT step, as the reference does (:2228)
    newTimeStep = std::max(newTimeStep, minTimeStep);
    newTimeStep = std::min(newTimeStep, maxTimeStep);
    if ((stopTime - currentTime) >= minTimeStep) {
      double nextTimePt = currentTime + newTimeStep;
      if (nextTimePt > stopTime) {
        savedTimeStep = newTimeStep;
        nextTimePt = stopTime;
        newTimeStep = stopTime - currentTime;
      }
      nextTime = nextTimePt;
      currentTimeStepRatio = newTimeStep / lastTimeStep;
      currentTimeStepSum = newTimeStep + lastTimeStep;
      currentTimeStep = newTimeStep;
    }
  }

  bool reject_step() {            // OneStep::rejectStep
    double newTimeStep = currentTimeStep;
    ++nef;
    for (int i = 1; i <= currentOrder; ++i) psi[i - 1] = psi[i];     // restoreHistory
    if (nef >= maxNumfail) return false;
    if (newtonConvergenceStatus <= 0) {
      newTimeStep = currentTimeStep / 8;
      currentOrder = P.minOrder;
    } else if (nef == 1) {
      if (!gear()) currentOrder = P.minOrder;     // Gear12::rejectStep keeps the order on the first error-test failure (:1484-1491)
      double rr = tolAimFac / (estOverTol + 0.0001);
      rr = std::pow(rr, 1.0 / (currentOrder + 1.0));
      rr = std::max(r_min, std::min(r_max, rr));
      newTimeStep = rr * currentTimeStep;
    } else {
      newTimeStep = r_min * currentTimeStep;
      currentOrder = P.minOrder;
    }
    newTimeStep = std::max(newTimeStep, minTimeStep);
    newTimeStep = std::min(newTimeStep, maxTimeStep);
    double nextTimePt = currentTime + newTimeStep;
    if (nextTimePt > stopTime) { nextTimePt = stopTime; newTimeStep = stopTime - currentTime; }
    nextTime = nextTimePt;
    currentTimeStepRatio = newTimeStep / lastTimeStep;
    currentTimeStepSum = newTimeStep + lastTimeStep;
    currentTimeStep = newTimeStep;
    if (currentTimeStep <= minTimeStep) return false;
    return true;
  }
};

}  // namespace sim
}  // namespace xb
