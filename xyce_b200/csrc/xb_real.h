// xyce_b200 -- alternative scalar types for the single-source model evaluators.
//   xb::CountReal   host only: counts executed fp64 operations (each + - * / sqrt exp log = 1),
//                   used once to establish the executed-flop figure of the roofline unit U1
//                   (BASELINE.md section 4).
//   xb::FastReal    device: identical to double except that a/b is computed as a branch-free
//                   reciprocal-Newton sequence (1 MUFU + 7 DFMA/DMUL, <= 2 ulp) instead of the
//                   ~20-instruction IEEE division with its slow-path branch.
// Include before xb_common.h and define XB_REAL to select one of them.
#pragma once
#include <math.h>
#include "xb_fastmath.h"
#if defined(__CUDACC__)
#define XBR_HD __host__ __device__ __forceinline__
#else
#define XBR_HD inline
#endif

namespace xb {

struct OpCounts { unsigned long long add, mul, div, sqrt_, exp_, log_, cmp; };
#if !defined(__CUDA_ARCH__)
inline OpCounts &op_counts() { static thread_local OpCounts c{}; return c; }
#endif

struct CountPolicy {
#if !defined(__CUDA_ARCH__)
  static double add(double a, double b) { ++op_counts().add; return a + b; }
  static double sub(double a, double b) { ++op_counts().add; return a - b; }
  static double mul(double a, double b) { ++op_counts().mul; return a * b; }
  static double div(double a, double b) { ++op_counts().div; return a / b; }
  static double sqrt_(double a) { ++op_counts().sqrt_; return ::sqrt(a); }
  static double exp_(double a) { ++op_counts().exp_; return ::exp(a); }
  static double log_(double a) { ++op_counts().log_; return ::log(a); }
  static void cmp() { ++op_counts().cmp; }
#endif
};

struct FastDivPolicy {
  static XBR_HD double add(double a, double b) { return a + b; }
  static XBR_HD double sub(double a, double b) { return a - b; }
  static XBR_HD double mul(double a, double b) { return a * b; }
  static XBR_HD double div(double a, double b) {
#if defined(__CUDA_ARCH__)
    return fm::div(a, b);
#else
    return a / b;
#endif
  }
#if defined(__CUDA_ARCH__)
  static XBR_HD double sqrt_(double a) { return fm::sqrt(a); }
  static XBR_HD double exp_(double a) { return fm::exp(a); }
  static XBR_HD double log_(double a) { return fm::log(a); }
#else
  static XBR_HD double sqrt_(double a) { return ::sqrt(a); }
  static XBR_HD double exp_(double a) { return ::exp(a); }
  static XBR_HD double log_(double a) { return ::log(a); }
#endif
  static XBR_HD void cmp() {}
};

template <class P>
struct RealT {
  double v;
  XBR_HD RealT() {}
  XBR_HD RealT(double x) : v(x) {}
  XBR_HD RealT &operator+=(const RealT &o) { v = P::add(v, o.v); return *this; }
  XBR_HD RealT &operator-=(const RealT &o) { v = P::sub(v, o.v); return *this; }
  XBR_HD RealT &operator*=(const RealT &o) { v = P::mul(v, o.v); return *this; }
  XBR_HD RealT &operator/=(const RealT &o) { v = P::div(v, o.v); return *this; }
  friend XBR_HD RealT operator-(const RealT &a) { return RealT(-a.v); }
  friend XBR_HD RealT operator+(const RealT &a) { return a; }
  friend XBR_HD RealT operator+(const RealT &a, const RealT &b) { return RealT(P::add(a.v, b.v)); }
  friend XBR_HD RealT operator-(const RealT &a, const RealT &b) { return RealT(P::sub(a.v, b.v)); }
  friend XBR_HD RealT operator*(const RealT &a, const RealT &b) { return RealT(P::mul(a.v, b.v)); }
  friend XBR_HD RealT operator/(const RealT &a, const RealT &b) { return RealT(P::div(a.v, b.v)); }
  friend XBR_HD bool operator<(const RealT &a, const RealT &b) { P::cmp(); return a.v < b.v; }
  friend XBR_HD bool operator>(const RealT &a, const RealT &b) { P::cmp(); return a.v > b.v; }
  friend XBR_HD bool operator<=(const RealT &a, const RealT &b) { P::cmp(); return a.v <= b.v; }
  friend XBR_HD bool operator>=(const RealT &a, const RealT &b) { P::cmp(); return a.v >= b.v; }
  friend XBR_HD bool operator==(const RealT &a, const RealT &b) { P::cmp(); return a.v == b.v; }
  friend XBR_HD bool operator!=(const RealT &a, const RealT &b) { P::cmp(); return a.v != b.v; }
  friend XBR_HD RealT sqrt(const RealT &a) { return RealT(P::sqrt_(a.v)); }
  friend XBR_HD RealT exp(const RealT &a) { return RealT(P::exp_(a.v)); }
  friend XBR_HD RealT log(const RealT &a) { return RealT(P::log_(a.v)); }
  friend XBR_HD RealT rpow(const RealT &a, const RealT &b) { return RealT(::pow(a.v, b.v)); }
  friend XBR_HD RealT rtan(const RealT &a) { return RealT(::tan(a.v)); }
  friend XBR_HD RealT fabs(const RealT &a) { return RealT(::fabs(a.v)); }
  // the rest of the Verilog-A math library as the translated ADMS models call it (plain library versions)
  friend XBR_HD RealT sin(const RealT &a) { return RealT(::sin(a.v)); }
  friend XBR_HD RealT cos(const RealT &a) { return RealT(::cos(a.v)); }
  friend XBR_HD RealT tan(const RealT &a) { return RealT(::tan(a.v)); }
  friend XBR_HD RealT atan(const RealT &a) { return RealT(::atan(a.v)); }
  friend XBR_HD RealT asin(const RealT &a) { return RealT(::asin(a.v)); }
  friend XBR_HD RealT acos(const RealT &a) { return RealT(::acos(a.v)); }
  friend XBR_HD RealT sinh(const RealT &a) { return RealT(::sinh(a.v)); }
  friend XBR_HD RealT cosh(const RealT &a) { return RealT(::cosh(a.v)); }
  friend XBR_HD RealT tanh(const RealT &a) { return RealT(::tanh(a.v)); }
  friend XBR_HD RealT asinh(const RealT &a) { return RealT(::asinh(a.v)); }
  friend XBR_HD RealT acosh(const RealT &a) { return RealT(::acosh(a.v)); }
  friend XBR_HD RealT atanh(const RealT &a) { return RealT(::atanh(a.v)); }
  friend XBR_HD RealT hypot(const RealT &a, const RealT &b) { return RealT(::hypot(a.v, b.v)); }
  friend XBR_HD RealT log10(const RealT &a) { return RealT(::log10(a.v)); }
  friend XBR_HD RealT atan2(const RealT &a, const RealT &b) { return RealT(::atan2(a.v, b.v)); }
  friend XBR_HD RealT floor(const RealT &a) { return RealT(::floor(a.v)); }
  friend XBR_HD RealT ceil(const RealT &a) { return RealT(::ceil(a.v)); }
  friend XBR_HD double to_double(const RealT &a) { return a.v; }
};

// xb::TaintReal   host only (analysis tool): every value carries a flag "depends on the bias point" (node voltages,
//                   previous limiting voltages); operations are counted separately for bias-dependent and
//                   bias-independent results, and divisions whose DIVISOR is bias-independent are counted on their own:
//                   the work that could move into the per-bin / per-instance records (scripts/count_hoistable.py).
#if !defined(__CUDA_ARCH__)
struct TaintCounts { unsigned long long ops[2][6]; unsigned long long div_const_divisor; };
#if defined(XB_TAINT_TRACE)      // scripts/hoistable_lines.py: where the bias-independent operations are (call-stack sampling)
extern "C" void xb_taint_event(int kind);
#define XB_TAINT_EVENT(k) xb_taint_event(k)
#else
#define XB_TAINT_EVENT(k) ((void)0)
#endif
inline TaintCounts &taint_counts() { static thread_local TaintCounts c{}; return c; }
struct TaintReal {
  double v; bool b;
  TaintReal() : v(0.0), b(false) {}
  TaintReal(double x) : v(x), b(false) {}
  TaintReal(double x, bool bias) : v(x), b(bias) {}
  static TaintReal mk(double x, bool bias, int kind) { ++taint_counts().ops[bias ? 1 : 0][kind]; if (!bias) XB_TAINT_EVENT(kind); return TaintReal(x, bias); }
  TaintReal &operator+=(const TaintReal &o) { *this = mk(v + o.v, b || o.b, 0); return *this; }
  TaintReal &operator-=(const TaintReal &o) { *this = mk(v - o.v, b || o.b, 0); return *this; }
  TaintReal &operator*=(const TaintReal &o) { *this = mk(v * o.v, b || o.b, 1); return *this; }
  TaintReal &operator/=(const TaintReal &o) { if (b && !o.b) { ++taint_counts().div_const_divisor; XB_TAINT_EVENT(6); } *this = mk(v / o.v, b || o.b, 2); return *this; }
  friend TaintReal operator-(const TaintReal &a) { return TaintReal(-a.v, a.b); }
  friend TaintReal operator+(const TaintReal &a) { return a; }
  friend TaintReal operator+(TaintReal a, const TaintReal &o) { a += o; return a; }
  friend TaintReal operator-(TaintReal a, const TaintReal &o) { a -= o; return a; }
  friend TaintReal operator*(TaintReal a, const TaintReal &o) { a *= o; return a; }
  friend TaintReal operator/(TaintReal a, const TaintReal &o) { a /= o; return a; }
  friend bool operator<(const TaintReal &a, const TaintReal &o) { return a.v < o.v; }
  friend bool operator>(const TaintReal &a, const TaintReal &o) { return a.v > o.v; }
  friend bool operator<=(const TaintReal &a, const TaintReal &o) { return a.v <= o.v; }
  friend bool operator>=(const TaintReal &a, const TaintReal &o) { return a.v >= o.v; }
  friend bool operator==(const TaintReal &a, const TaintReal &o) { return a.v == o.v; }
  friend bool operator!=(const TaintReal &a, const TaintReal &o) { return a.v != o.v; }
  friend TaintReal sqrt(const TaintReal &a) { return mk(::sqrt(a.v), a.b, 3); }
  friend TaintReal exp(const TaintReal &a) { return mk(::exp(a.v), a.b, 4); }
  friend TaintReal log(const TaintReal &a) { return mk(::log(a.v), a.b, 5); }
  friend TaintReal rpow(const TaintReal &a, const TaintReal &o) { return mk(::pow(a.v, o.v), a.b || o.b, 4); }
  friend TaintReal rtan(const TaintReal &a) { return TaintReal(::tan(a.v), a.b); }
  friend TaintReal fabs(const TaintReal &a) { return TaintReal(::fabs(a.v), a.b); }
  friend double to_double(const TaintReal &a) { return a.v; }
};
#endif

using CountReal = RealT<CountPolicy>;
using FastReal = RealT<FastDivPolicy>;

}  // namespace xb
