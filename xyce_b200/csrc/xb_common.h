// xyce_b200 -- common definitions shared by every compact-model evaluator.
//
// Single-source: this header is compiled by nvcc for sm_100a (the product) and,
// for CPU-side unit tests only, by g++ (tests/host_mirror).  It holds no state.
//
// Constants follow the reference's values so that results are comparable at
// 1e-12: src/DeviceModelPKG/Core/N_DEV_Const.h:42-113 and the BSIM4-local
// overrides in src/DeviceModelPKG/OpenModels/N_DEV_MOSFET_B4p82.C:58-105.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define XB_HD __host__ __device__ __forceinline__
#define XB_D __device__ __forceinline__
#else
#define XB_HD inline
#define XB_D inline
#endif

// Scalar type of all model arithmetic.  The product uses plain double; tests/host_mirror can
// substitute an operation-counting wrapper (XB_REAL) to measure executed flops per evaluation.
#ifndef XB_REAL
#define XB_REAL double
#endif
// Helpers that are called from several places of an evaluator (junction diode for source and drain, limiters,
// GIDL / GISL, ...): real functions in the fast device build -- the second call finds the code in the instruction
// caches -- and inlined everywhere else (strict builds, host mirror).
// They take and return values only (no references into the caller's frame).
#if defined(__CUDACC__) && defined(XB_ARITH) && XB_ARITH == 2 && !defined(XB_HELPERS_INLINE)
#define XB_HELPER static __host__ __device__ __noinline__
#else
#define XB_HELPER XB_HD
#endif

// Lock-step builds (XB_LOCKSTEP=1, device only) put a block-wide barrier between the sections of the
// large evaluators so that all warps of a block execute the same ~1-2k instruction window and share
// instruction-cache lines.  Every thread of the block must reach each XB_SYNC_POINT().
#ifndef XB_LOCKSTEP
#define XB_LOCKSTEP 0
#endif
// XB_SYNC_POINT(level): level 1 = coarse grid (about every 1k instructions), 2 = fine grid; active when
// level <= XB_LOCKSTEP.  XB_SYNC_POINT_U marks points inside branches on model-card / bin / solver-flag
// values: legal only when those are uniform per block (the uniform-record kernel).
#if XB_LOCKSTEP && defined(__CUDA_ARCH__)
#define XB_SYNC_POINT(level) do { if ((level) <= XB_LOCKSTEP) __syncthreads(); } while (0)
#else
#define XB_SYNC_POINT(level) ((void)0)
#endif
#define XB_SYNC_POINT_U(level) XB_SYNC_POINT(level)

namespace xb {

using real = XB_REAL;
struct Real2 { real a, b; };
struct Real4 { real a, b, c, d; };
XB_HD double to_double(double v) { return v; }

// ---- physical constants (N_DEV_Const.h) -----------------------------------
constexpr double kQ        = 1.6021918e-19;
constexpr double kCtoK     = 273.15;
constexpr double kBoltz    = 1.3806226e-23;
constexpr double kKoverQ   = kBoltz / kQ;
constexpr double kVt0      = kBoltz * (27.0 + kCtoK) / kQ;   // CONSTvt0
constexpr double kEpsOx    = 3.453133e-11;
constexpr double kEpsSi    = 1.03594e-10;
constexpr double kEps0B4   = 8.85418e-12;                    // B4p82.C:58
constexpr double kChargeQ  = 1.60219e-19;                    // B4p82.C:60
constexpr double kMaxExp   = 5.834617425e14;
constexpr double kMinExp   = 1.713908431e-15;
constexpr double kExpThr   = 34.0;
constexpr double kMaxExpL  = 2.688117142e+43;
constexpr double kMinExpL  = 3.720075976e-44;
constexpr double kExpLThr  = 100.0;
constexpr double kDelta1   = 0.02;
constexpr double kDelta3   = 0.02;
constexpr double kDelta4   = 0.02;
constexpr double kMachEps  = 2.220446049250313e-16;          // MachineDependentParams::MachinePrecision()
constexpr int    kMM       = 3;                              // B4 smoothing coefficient

// ---- per-launch solver/device flag block ----------------------------------
// Mirrors the subset of Device::SolverState (Core/N_DEV_SolverState.h:114-217)
// and Device::DeviceOptions (Core/N_DEV_DeviceOptions.C:79-150) that is read
// inside model evaluation.
struct SolverFlags {
  int dcopFlag;
  int tranopFlag;
  int acopFlag;
  int transientFlag;
  int dcsweepFlag;
  int initJctFlag;
  int initFixFlag;
  int initTranFlag;
  int newtonIter;
  int locaEnabledFlag;
  int artParameterFlag;
  int voltageLimiterFlag;
  double gmin;
  double gainScale;
  double nltermScale;
  double vgstConst;
  double vdsScaleMin;
  double sizeScale;      // MOS1 homotopy (unused by BSIM4)
  double currTimeStep;   // unused by BSIM4 eval
  double lastTimeStep;   // BJT excess phase (Weil's approximation, N_DEV_BJT.C:2742-2799)
  int beginIntegrationFlag;   // first step out of a break point, t = 0 included (N_DEV_SolverState.h:169)
};

// ---- smoothed exponentials (B4p82.C:82-105) --------------------------------
XB_HD void dexp(real a, real &b, real &c) {
#if defined(XB_DEXP_SELECT) && defined(__CUDA_ARCH__)
  // branch-free: the exponential is evaluated unconditionally (its argument saturates inside fm::exp) and the
  // linear / floor continuations are selected afterwards
  const real ex = exp(a);
  const bool hi = a > kExpThr, lo = a < -kExpThr;
  b = hi ? kMaxExp * (1.0 + a - kExpThr) : (lo ? real(kMinExp) : ex);
  c = hi ? real(kMaxExp) : (lo ? real(0.0) : ex);
  return;
#endif
  if (a > kExpThr) { b = kMaxExp * (1.0 + a - kExpThr); c = kMaxExp; }
  else if (a < -kExpThr) { b = kMinExp; c = 0.0; }
  else { b = exp(a); c = b; }
}
XB_HD real dexp2(real a) {
#if defined(XB_DEXP_SELECT) && defined(__CUDA_ARCH__)
  const real ex = exp(a);
  return a > kExpThr ? kMaxExp * (1.0 + a - kExpThr) : (a < -kExpThr ? real(kMinExp) : ex);
#endif
  if (a > kExpThr) return kMaxExp * (1.0 + a - kExpThr);
  if (a < -kExpThr) return kMinExp;
  return exp(a);
}

XB_HD double rpow(double a, double b) { return pow(a, b); }      // overloaded for the wrapper scalar types
XB_HD double rtan(double a) { return tan(a); }
XB_HD real dmax(real a, real b) { return a < b ? b : a; }   // std::max semantics
XB_HD real dmin(real a, real b) { return b < a ? b : a; }   // std::min semantics

// ---- SPICE3 Newton limiters (Core/N_DEV_DeviceSupport.C:161-393) ------------
XB_HD real limvds(real vnew, real vold) {
  if (vold >= 3.5) {
    if (vnew > vold) vnew = dmin(vnew, 3.0 * vold + 2.0);
    else if (vnew < 3.5) vnew = dmax(vnew, 2.0);
  } else {
    if (vnew > vold) vnew = dmin(vnew, 4.0);
    else vnew = dmax(vnew, -0.5);
  }
  return vnew;
}

XB_HD real pnjlim(real vnew, real vold, real vt, real vcrit, int &icheck) {
  if ((vnew > vcrit) && (fabs(vnew - vold) > (vt + vt))) {
    if (vold > 0) {
      real arg = 1 + (vnew - vold) / vt;
      vnew = (arg > 0) ? vold + vt * log(arg) : vcrit;
    } else {
      vnew = vt * log(vnew / vt);
    }
    icheck = 1;
  } else {
    icheck = 0;
  }
  return vnew;
}

XB_HD real fetlim(real vnew, real vold, real vto) {
  const real vtsthi = fabs(2 * (vold - vto)) + 2;
  const real vtstlo = vtsthi / 2 + 2;
  const real vtox = vto + 3.5;
  const real delv = vnew - vold;
  if (vold >= vto) {
    if (vold >= vtox) {
      if (delv <= 0) {
        if (vnew >= vtox) { if (-delv > vtstlo) vnew = vold - vtstlo; }
        else vnew = dmax(vnew, vto + 2.0);
      } else {
        if (delv >= vtsthi) vnew = vold + vtsthi;
      }
    } else {
      vnew = (delv <= 0) ? dmax(vnew, vto - 0.5) : dmin(vnew, vto + 4.0);
    }
  } else {
    if (delv <= 0) {
      if (-delv > vtsthi) vnew = vold - vtsthi;
    } else {
      const real vtemp = vto + 0.5;
      if (vnew <= vtemp) { if (delv > vtstlo) vnew = vold + vtstlo; }
      else vnew = vtemp;
    }
  }
  return vnew;
}

}  // namespace xb
