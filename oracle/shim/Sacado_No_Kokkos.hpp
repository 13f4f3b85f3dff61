// oracle build only: a small forward-mode dual number standing in for Sacado::Fad::SFad so that the
// reference's N_DEV_Diode.C / N_DEV_BJT.C (which use it ONLY in their parameter-sensitivity code,
// never on the Newton path) compile without Trilinos.  Written from scratch; not Sacado source.
#ifndef XB_ORACLE_SACADO_SHIM
#define XB_ORACLE_SACADO_SHIM
#include <cmath>
namespace Sacado {
namespace Fad {
template <class T, int N>
class SFad {
 public:
  typedef T value_type;
  SFad() : v_(T()) { for (int i = 0; i < N; ++i) d_[i] = T(); }
  SFad(const T &v) : v_(v) { for (int i = 0; i < N; ++i) d_[i] = T(); }
  SFad(int v) : v_(T(v)) { for (int i = 0; i < N; ++i) d_[i] = T(); }
  SFad(int, int i, const T &v) : v_(v) { for (int k = 0; k < N; ++k) d_[k] = T(); d_[i] = T(1); }
  void diff(int i, int) { for (int k = 0; k < N; ++k) d_[k] = T(); d_[i] = T(1); }
  const T &val() const { return v_; }
  T &val() { return v_; }
  const T &dx(int i) const { return d_[i]; }
  T &fastAccessDx(int i) { return d_[i]; }
  const T &fastAccessDx(int i) const { return d_[i]; }
  int size() const { return N; }
  SFad &operator=(const T &v) { v_ = v; for (int i = 0; i < N; ++i) d_[i] = T(); return *this; }
  SFad &operator+=(const SFad &o) { v_ += o.v_; for (int i = 0; i < N; ++i) d_[i] += o.d_[i]; return *this; }
  SFad &operator-=(const SFad &o) { v_ -= o.v_; for (int i = 0; i < N; ++i) d_[i] -= o.d_[i]; return *this; }
  SFad &operator*=(const SFad &o) { for (int i = 0; i < N; ++i) d_[i] = d_[i] * o.v_ + v_ * o.d_[i]; v_ *= o.v_; return *this; }
  SFad &operator/=(const SFad &o) { for (int i = 0; i < N; ++i) d_[i] = (d_[i] * o.v_ - v_ * o.d_[i]) / (o.v_ * o.v_); v_ /= o.v_; return *this; }
  friend SFad operator-(const SFad &a) { SFad r; r.v_ = -a.v_; for (int i = 0; i < N; ++i) r.d_[i] = -a.d_[i]; return r; }
  friend SFad operator+(const SFad &a) { return a; }
  friend SFad operator+(const SFad &a, const SFad &b) { SFad r(a); r += b; return r; }
  friend SFad operator-(const SFad &a, const SFad &b) { SFad r(a); r -= b; return r; }
  friend SFad operator*(const SFad &a, const SFad &b) { SFad r(a); r *= b; return r; }
  friend SFad operator/(const SFad &a, const SFad &b) { SFad r(a); r /= b; return r; }
  friend bool operator<(const SFad &a, const SFad &b) { return a.v_ < b.v_; }
  friend bool operator>(const SFad &a, const SFad &b) { return a.v_ > b.v_; }
  friend bool operator<=(const SFad &a, const SFad &b) { return a.v_ <= b.v_; }
  friend bool operator>=(const SFad &a, const SFad &b) { return a.v_ >= b.v_; }
  friend bool operator==(const SFad &a, const SFad &b) { return a.v_ == b.v_; }
  friend bool operator!=(const SFad &a, const SFad &b) { return a.v_ != b.v_; }
  friend bool operator!(const SFad &a) { return a.v_ == T(); }
  friend SFad max(const SFad &a, const SFad &b) { return (a.v_ < b.v_) ? b : a; }
  friend SFad min(const SFad &a, const SFad &b) { return (b.v_ < a.v_) ? b : a; }
  friend SFad chain(const SFad &a, const T &f, const T &df) { SFad r; r.v_ = f; for (int i = 0; i < N; ++i) r.d_[i] = df * a.d_[i]; return r; }
  friend SFad exp(const SFad &a) { const T e = std::exp(a.v_); return chain(a, e, e); }
  friend SFad log(const SFad &a) { return chain(a, std::log(a.v_), T(1) / a.v_); }
  friend SFad sqrt(const SFad &a) { const T s = std::sqrt(a.v_); return chain(a, s, T(0.5) / s); }
  friend SFad fabs(const SFad &a) { return chain(a, std::fabs(a.v_), a.v_ < T(0) ? T(-1) : T(1)); }
  friend SFad abs(const SFad &a) { return fabs(a); }
  friend SFad tan(const SFad &a) { const T t = std::tan(a.v_); return chain(a, t, T(1) + t * t); }
  friend SFad sin(const SFad &a) { return chain(a, std::sin(a.v_), std::cos(a.v_)); }
  friend SFad cos(const SFad &a) { return chain(a, std::cos(a.v_), -std::sin(a.v_)); }
  friend SFad atan(const SFad &a) { return chain(a, std::atan(a.v_), T(1) / (T(1) + a.v_ * a.v_)); }
  friend SFad tanh(const SFad &a) { const T t = std::tanh(a.v_); return chain(a, t, T(1) - t * t); }
  friend SFad pow(const SFad &a, const SFad &b) {
    const T p = std::pow(a.v_, b.v_);
    SFad r; r.v_ = p;
    for (int i = 0; i < N; ++i) {
      T d = T();
      if (a.d_[i] != T()) d += b.v_ * std::pow(a.v_, b.v_ - T(1)) * a.d_[i];
      if (b.d_[i] != T()) d += p * std::log(a.v_) * b.d_[i];
      r.d_[i] = d;
    }
    return r;
  }
  friend SFad pow(const SFad &a, const T &b) { return pow(a, SFad(b)); }
  friend SFad pow(const T &a, const SFad &b) { return pow(SFad(a), b); }
 private:
  T v_;
  T d_[N];
};
}  // namespace Fad
}  // namespace Sacado
// mixed scalar / dual min and max, as the reference's templated device code calls them through std::
namespace std {
template <class T, int N> Sacado::Fad::SFad<T, N> min(const T &a, const Sacado::Fad::SFad<T, N> &b) { return (b.val() < a) ? b : Sacado::Fad::SFad<T, N>(a); }
template <class T, int N> Sacado::Fad::SFad<T, N> min(const Sacado::Fad::SFad<T, N> &a, const T &b) { return (b < a.val()) ? Sacado::Fad::SFad<T, N>(b) : a; }
template <class T, int N> Sacado::Fad::SFad<T, N> max(const T &a, const Sacado::Fad::SFad<T, N> &b) { return (a < b.val()) ? b : Sacado::Fad::SFad<T, N>(a); }
template <class T, int N> Sacado::Fad::SFad<T, N> max(const Sacado::Fad::SFad<T, N> &a, const T &b) { return (a.val() < b) ? Sacado::Fad::SFad<T, N>(b) : a; }
template <class T, int N> Sacado::Fad::SFad<T, N> min(const Sacado::Fad::SFad<T, N> &a, const Sacado::Fad::SFad<T, N> &b) { return (b.val() < a.val()) ? b : a; }
template <class T, int N> Sacado::Fad::SFad<T, N> max(const Sacado::Fad::SFad<T, N> &a, const Sacado::Fad::SFad<T, N> &b) { return (a.val() < b.val()) ? b : a; }
}  // namespace std
#endif
