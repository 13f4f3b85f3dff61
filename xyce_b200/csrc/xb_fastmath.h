// xyce_b200 -- device-side fp64 elementary functions of the "fast" arithmetic variant (XB_ARITH == 2).
//
// Why they exist: the BSIM4 evaluation kernel is bound by instruction delivery, and
// the CUDA math library's exp / log expand to 60-75 instructions each, a third of them moves that
// materialise polynomial coefficients as immediates.  The versions below keep their coefficients in
// constant memory (used directly as instruction operands), have no slow-path branches on the common
// path, and are accurate to <= 2 ulp -- four orders of magnitude inside the 1e-12 parity budget, which
// tests/test_gpu_bsim4_parity.py checks against the reference for every arithmetic variant.
//
// Algorithms (textbook):
//   exp(x)  = 2^n * exp(r),  n = rint(x / ln 2), r = x - n ln2 (two-term Cody-Waite),
//             exp(r) by its degree-13 Taylor polynomial on |r| <= ln2 / 2 (truncation < 4e-18)
//   log(x)  = e ln2 + 2 atanh(f), x = 2^e m, m in [sqrt(1/2), sqrt(2)), f = (m - 1) / (m + 1),
//             atanh by its odd Taylor series up to f^21 on |f| <= 0.1716 (truncation < 2e-18)
//   a / b   = a * (1/b); 1/b = reciprocal seed (MUFU.RCP64H) + two Newton steps.  The reciprocal depends on b alone, so
//             the compiler shares it between all divisions by the same value (the evaluators divide by the same
//             denominators again and again) and the quotient is one multiply after `a` is known
//             (XB_DIV_MODE 0 = the round-1 form: one Newton step, quotient, one residual correction).
//             Measured on B200 (profiles/r02_b4_eval_ilp_experiments.md): C2 kernel 62.5 -> 59.4 us
//   sqrt(x) = reciprocal-root seed (MUFU.RSQ64H), one coupled Goldschmidt step, one residual correction, inlined and
//             branch-free (a call costs ~8 instructions and ends the scheduling block): 59.4 -> 56.3 us
#pragma once
#if defined(__CUDACC__)

namespace xb {
namespace fm {

// 1/k!, k = 2 .. 13
static __constant__ double kExpC[12] = {
    1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0, 1.0 / 362880.0,
    1.0 / 3628800.0, 1.0 / 39916800.0, 1.0 / 479001600.0, 1.0 / 6227020800.0};
// log2(e), ln2 split into a 32-bit-mantissa head and the remainder, clamp bounds
static __constant__ double kExpK[5] = {1.4426950408889634074, 6.93147180369123816490e-01, 1.90821492927058770002e-10,
                                       -745.0, 709.78};
// 2/(2k+1), k = 1 .. 10
static __constant__ double kLogC[10] = {2.0 / 3.0, 2.0 / 5.0, 2.0 / 7.0, 2.0 / 9.0, 2.0 / 11.0, 2.0 / 13.0, 2.0 / 15.0,
                                        2.0 / 17.0, 2.0 / 19.0, 2.0 / 21.0};
static __constant__ double kLogK[2] = {6.93147180369123816490e-01, 1.90821492927058770002e-10};

__device__ __forceinline__ double rcp_seed(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  return r;
}

// Reciprocal to <= 1 ulp: seed + two Newton steps.  Depends on b only, so the compiler shares one reciprocal between
// every division by the same value (XB_DIV_MODE 1).
__device__ __forceinline__ double rcp(double b) {
  const double r = rcp_seed(b);
#if defined(XB_RCP_NEWTON2)
  const double r1 = fma(fma(-b, r, 1.0), r, r);
  return fma(fma(-b, r1, 1.0), r1, r1);
#else
  // one third-order step: e = 1 - b r (|e| < 2^-20 from the seed), 1/b = r (1 + e + e^2 + ...): r + r (e + e^2) leaves e^3
  const double e = fma(-b, r, 1.0);
  return fma(r, fma(e, e, e), r);
#endif
}

// <= 1 ulp for normal operands (mode 1: <= 2 ulp); b = 0, inf, denormal give NaN/inf like the seed does (the model code
// guards its denominators; the strict variants keep IEEE division)
#ifndef XB_DIV_MODE
#define XB_DIV_MODE 1
#endif
__device__ __forceinline__ double div(double a, double b) {
#if XB_DIV_MODE == 1
  return a * rcp(b);
#else
  double r = rcp_seed(b);
  r = fma(fma(-b, r, 1.0), r, r);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
#endif
}

__device__ __forceinline__ double rsqrt_seed(double b) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  return r;
}
// Branch-free square root, <= 1 ulp for normal x: reciprocal-root seed, one coupled (Goldschmidt) step for sqrt and
// 1/(2 sqrt), one residual correction.  0 is passed through by a select; x < 0, NaN and +inf give NaN;
// denormal x (never an argument of the model code) is not supported.
__device__ __forceinline__ double sqrt_inline(double x) {
  const double y = rsqrt_seed(x);
  double g = x * y, h = 0.5 * y;
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  g = fma(fma(-g, g, x), h, g);
#if defined(XB_SQRT_INF)
  return (x == 0.0 || x == __longlong_as_double(0x7ff0000000000000LL)) ? x : g;
#else
  return x == 0.0 ? x : g;      // +inf gives NaN (0 * inf in the first product): never an argument of the model code
#endif
}

// exp / log / sqrt are real functions (not inlined): the evaluation kernel is bound by instruction delivery, and
// one 40-instruction body that stays in the instruction caches beats 50 inlined copies (measured: C2 0.087 -> 0.063 ms)
#ifndef XB_FM_INLINE
#define XB_FM_INLINE __noinline__
#endif
#if defined(XB_SQRT_CALL)
static __device__ XB_FM_INLINE double sqrt(double x) { return ::sqrt(x); }
#else
static __device__ __forceinline__ double sqrt(double x) { return sqrt_inline(x); }
#endif
#ifndef XB_EXP_INLINE
#define XB_EXP_INLINE XB_FM_INLINE
#endif
#ifndef XB_LOG_INLINE
#define XB_LOG_INLINE XB_FM_INLINE
#endif
static __device__ XB_EXP_INLINE double exp_general(double x);
// exp on the range the model code uses (|x| < 700: results are normal numbers), inlined: no clamps, 2^n by an integer
// add on the exponent field.  Everything else (NaN, +-inf, |x| >= 700) takes the saturating general version below, a
// call.  ~30 instructions instead of ~52 (clamps with their constant loads, two-factor scaling, callee spills);
// line-level attribution: profiles/r02_b4_eval_kernel_v4_ncu_summary.md.
__device__ __forceinline__ double exp_lean(double x) {
  if (!(fabs(x) < 700.0)) return exp_general(x);
  const double shifter = 6755399441055744.0;              // 1.5 * 2^52
  const double t = fma(x, kExpK[0], shifter);
  const double fn = t - shifter;
  double r = fma(-fn, kExpK[1], x);
  r = fma(-fn, kExpK[2], r);
  double p = kExpC[11];
#pragma unroll
  for (int k = 10; k >= 0; --k) p = fma(p, r, kExpC[k]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return __hiloint2double(__double2hiint(p) + (__double2loint(t) << 20), __double2loint(p));
}
// Measured on B200 (C2 kernel / 1M instances, profiles/r02_b4_eval_ilp_experiments.md): general version through a call
// 54.3 / 375.8 us, lean version inlined at its 36 call sites 54.3 / 373.8 us (+1 250 static instructions), lean version
// as ONE real function 54.3 / 372.7 us -> the default.
#if defined(XB_EXP_CALL)
static __device__ __forceinline__ double exp(double x) { return exp_general(x); }
#elif defined(XB_EXP_LEAN_INLINE)
static __device__ __forceinline__ double exp(double x) { return exp_lean(x); }
#else
static __device__ XB_EXP_INLINE double exp(double x) { return exp_lean(x); }
#endif
static __device__ XB_EXP_INLINE double exp_general(double x) {
  x = x < kExpK[3] ? kExpK[3] : x;          // NaN stays NaN (comparisons false)
  x = x > kExpK[4] ? kExpK[4] : x;
  const double shifter = 6755399441055744.0;              // 1.5 * 2^52
  const double t = fma(x, kExpK[0], shifter);
  const int n = __double2loint(t);
  const double fn = t - shifter;
  double r = fma(-fn, kExpK[1], x);
  r = fma(-fn, kExpK[2], r);
  double p = kExpC[11];
#pragma unroll
  for (int k = 10; k >= 0; --k) p = fma(p, r, kExpC[k]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  // 2^n in two factors so that denormal results and n = 1024 stay exact
  const int n1 = n >> 1, n2 = n - n1;
  const double s1 = __hiloint2double((n1 + 1023) << 20, 0), s2 = __hiloint2double((n2 + 1023) << 20, 0);
  return p * s1 * s2;
}

static __device__ XB_LOG_INLINE double log(double x) {
  if (!(x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308)) return ::log(x);   // 0, < 0, denormal, inf, NaN
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;                      // m in [1, 2)
  if (hi >= 0x3ff6a09f) { hi -= 0x00100000; ++e; }          // m >= sqrt(2) (to 20 bits): halve
  const double m = __hiloint2double(hi, lo);
  const double f = div(m - 1.0, m + 1.0);
  const double s = f * f;
  double p = kLogC[9];
#pragma unroll
  for (int k = 8; k >= 0; --k) p = fma(p, s, kLogC[k]);
  const double fe = (double)e;
  // log m = 2f + f s p ;  result = e ln2_hi + (2f + (f s p + e ln2_lo))
  const double tail = fma(fe, kLogK[1], f * s * p);
  return fma(fe, kLogK[0], fma(2.0, f, tail));
}

}  // namespace fm
}  // namespace xb
#endif
