// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
//
// Forward-mode dual numbers ("Jets").  The reference evaluates both residual
// families through ceres::AutoDiffCostFunction (vicalibrator.h:412-453, :620-621);
// Ceres itself is not vendored in /root/reference, so this restates the published
// Jet algebra: a value plus an N-vector of partial derivatives, every elementary
// operation carrying the chain rule.
#ifndef VICALIB_ORACLE_DUAL_H_
#define VICALIB_ORACLE_DUAL_H_

#include <cmath>

namespace vo {

template <int N>
struct Dual {
  double a;
  double v[N];

  Dual() : a(0.0) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  Dual(double s) : a(s) {  // NOLINT (implicit on purpose, like ceres::Jet)
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
  Dual(double s, int k) : a(s) {
    for (int i = 0; i < N; ++i) v[i] = 0.0;
    v[k] = 1.0;
  }
  Dual& operator+=(const Dual& o) {
    a += o.a;
    for (int i = 0; i < N; ++i) v[i] += o.v[i];
    return *this;
  }
  Dual& operator-=(const Dual& o) {
    a -= o.a;
    for (int i = 0; i < N; ++i) v[i] -= o.v[i];
    return *this;
  }
  Dual& operator*=(const Dual& o) { return *this = *this * o; }
  Dual& operator/=(const Dual& o) { return *this = *this / o; }
};

template <int N>
inline Dual<N> operator+(const Dual<N>& x, const Dual<N>& y) {
  Dual<N> r;
  r.a = x.a + y.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i];
  return r;
}
template <int N>
inline Dual<N> operator-(const Dual<N>& x, const Dual<N>& y) {
  Dual<N> r;
  r.a = x.a - y.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i];
  return r;
}
template <int N>
inline Dual<N> operator-(const Dual<N>& x) {
  Dual<N> r;
  r.a = -x.a;
  for (int i = 0; i < N; ++i) r.v[i] = -x.v[i];
  return r;
}
template <int N>
inline Dual<N> operator*(const Dual<N>& x, const Dual<N>& y) {
  Dual<N> r;
  r.a = x.a * y.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a;
  return r;
}
template <int N>
inline Dual<N> operator/(const Dual<N>& x, const Dual<N>& y) {
  Dual<N> r;
  const double inv = 1.0 / y.a;
  r.a = x.a * inv;
  for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv;
  return r;
}
#define VO_MIXED(op)                                                        \
  template <int N>                                                          \
  inline Dual<N> operator op(const Dual<N>& x, double s) {                  \
    return x op Dual<N>(s);                                                 \
  }                                                                         \
  template <int N>                                                          \
  inline Dual<N> operator op(double s, const Dual<N>& x) {                  \
    return Dual<N>(s) op x;                                                 \
  }
VO_MIXED(+)
VO_MIXED(-)
VO_MIXED(*)
VO_MIXED(/)
#undef VO_MIXED

#define VO_CMP(op)                                                          \
  template <int N>                                                          \
  inline bool operator op(const Dual<N>& x, const Dual<N>& y) {             \
    return x.a op y.a;                                                      \
  }                                                                         \
  template <int N>                                                          \
  inline bool operator op(const Dual<N>& x, double y) {                     \
    return x.a op y;                                                        \
  }                                                                         \
  template <int N>                                                          \
  inline bool operator op(double x, const Dual<N>& y) {                     \
    return x op y.a;                                                        \
  }
VO_CMP(<)
VO_CMP(>)
VO_CMP(<=)
VO_CMP(>=)
VO_CMP(==)
VO_CMP(!=)
#undef VO_CMP

template <int N>
inline Dual<N> chain(const Dual<N>& x, double f, double df) {
  Dual<N> r;
  r.a = f;
  for (int i = 0; i < N; ++i) r.v[i] = df * x.v[i];
  return r;
}
template <int N>
inline Dual<N> sqrt(const Dual<N>& x) {
  const double s = std::sqrt(x.a);
  return chain(x, s, 0.5 / s);
}
template <int N>
inline Dual<N> sin(const Dual<N>& x) {
  return chain(x, std::sin(x.a), std::cos(x.a));
}
template <int N>
inline Dual<N> cos(const Dual<N>& x) {
  return chain(x, std::cos(x.a), -std::sin(x.a));
}
template <int N>
inline Dual<N> tan(const Dual<N>& x) {
  const double t = std::tan(x.a);
  return chain(x, t, 1.0 + t * t);
}
template <int N>
inline Dual<N> atan(const Dual<N>& x) {
  return chain(x, std::atan(x.a), 1.0 / (1.0 + x.a * x.a));
}
template <int N>
inline Dual<N> asin(const Dual<N>& x) {
  return chain(x, std::asin(x.a), 1.0 / std::sqrt(1.0 - x.a * x.a));
}
template <int N>
inline Dual<N> log(const Dual<N>& x) {
  return chain(x, std::log(x.a), 1.0 / x.a);
}
template <int N>
inline Dual<N> atan2(const Dual<N>& y, const Dual<N>& x) {
  Dual<N> r;
  const double d = 1.0 / (x.a * x.a + y.a * y.a);
  r.a = std::atan2(y.a, x.a);
  for (int i = 0; i < N; ++i) r.v[i] = d * (x.a * y.v[i] - y.a * x.v[i]);
  return r;
}
template <int N>
inline Dual<N> abs(const Dual<N>& x) {
  return x.a < 0.0 ? -x : x;
}

// scalar overloads so templated code can call vo::sqrt etc. on plain doubles
inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
inline double tan(double x) { return std::tan(x); }
inline double atan(double x) { return std::atan(x); }
inline double asin(double x) { return std::asin(x); }
inline double log(double x) { return std::log(x); }
inline double atan2(double y, double x) { return std::atan2(y, x); }
inline double abs(double x) { return std::fabs(x); }

// interpolation-buffer.h:26-34 — index logic uses the scalar part only
template <int N>
inline double scalar_of(const Dual<N>& x) {
  return x.a;
}
inline double scalar_of(double x) { return x; }

}  // namespace vo
#endif
