// ORACLE — TEST INFRASTRUCTURE ONLY.
// The two pieces of Ceres the reference's cost-functor and local-parameterisation headers need to compile:
// ceres::Jet<T, N> (forward-mode dual numbers, as published in ceres/jet.h) and the LocalParameterization interface.
#pragma once
#include <Eigen/Core>
#include <cmath>

namespace ceres {

template <class T, int N>
struct Jet {
  T a;
  Eigen::Matrix<T, N, 1> v;
  Jet() : a() {}
  Jet(const T& value) : a(value) {}  // NOLINT
  Jet(const T& value, int k) : a(value) { v[k] = T(1.0); }
  Jet(const T& value, const Eigen::Matrix<T, N, 1>& d) : a(value), v(d) {}
  Jet& operator+=(const Jet& y) { *this = *this + y; return *this; }
  Jet& operator-=(const Jet& y) { *this = *this - y; return *this; }
  Jet& operator*=(const Jet& y) { *this = *this * y; return *this; }
  Jet& operator/=(const Jet& y) { *this = *this / y; return *this; }
};
#define VS_JET template <class T, int N> inline
VS_JET Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
VS_JET Jet<T, N> operator-(const Jet<T, N>& f) { return Jet<T, N>(-f.a, -f.v); }
VS_JET Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) { return Jet<T, N>(f.a + g.a, f.v + g.v); }
VS_JET Jet<T, N> operator+(const Jet<T, N>& f, T s) { return Jet<T, N>(f.a + s, f.v); }
VS_JET Jet<T, N> operator+(T s, const Jet<T, N>& f) { return Jet<T, N>(f.a + s, f.v); }
VS_JET Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) { return Jet<T, N>(f.a - g.a, f.v - g.v); }
VS_JET Jet<T, N> operator-(const Jet<T, N>& f, T s) { return Jet<T, N>(f.a - s, f.v); }
VS_JET Jet<T, N> operator-(T s, const Jet<T, N>& f) { return Jet<T, N>(s - f.a, -f.v); }
VS_JET Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) { return Jet<T, N>(f.a * g.a, f.a * g.v + f.v * g.a); }
VS_JET Jet<T, N> operator*(const Jet<T, N>& f, T s) { return Jet<T, N>(f.a * s, f.v * s); }
VS_JET Jet<T, N> operator*(T s, const Jet<T, N>& f) { return Jet<T, N>(f.a * s, f.v * s); }
VS_JET Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  const T g_a_inverse = T(1.0) / g.a;
  const T f_a_by_g_a = f.a * g_a_inverse;
  return Jet<T, N>(f_a_by_g_a, (f.v - f_a_by_g_a * g.v) * g_a_inverse);
}
VS_JET Jet<T, N> operator/(T s, const Jet<T, N>& g) {
  const T minus_s_g_a_inverse2 = -s / (g.a * g.a);
  return Jet<T, N>(s / g.a, g.v * minus_s_g_a_inverse2);
}
VS_JET Jet<T, N> operator/(const Jet<T, N>& f, T s) {
  const T s_inverse = T(1.0) / s;
  return Jet<T, N>(f.a * s_inverse, f.v * s_inverse);
}
#define VS_JET_CMP(op)                                                                    \
  VS_JET bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; } \
  VS_JET bool operator op(const T& s, const Jet<T, N>& g) { return s op g.a; }           \
  VS_JET bool operator op(const Jet<T, N>& f, const T& s) { return f.a op s; }
VS_JET_CMP(<)
VS_JET_CMP(<=)
VS_JET_CMP(>)
VS_JET_CMP(>=)
VS_JET_CMP(==)
VS_JET_CMP(!=)
#undef VS_JET_CMP
VS_JET Jet<T, N> abs(const Jet<T, N>& f) { return f.a < T(0.0) ? -f : f; }
VS_JET Jet<T, N> sqrt(const Jet<T, N>& f) {
  const T tmp = std::sqrt(f.a);
  return Jet<T, N>(tmp, f.v * (T(1.0) / (T(2.0) * tmp)));
}
VS_JET Jet<T, N> sin(const Jet<T, N>& f) { return Jet<T, N>(std::sin(f.a), f.v * std::cos(f.a)); }
VS_JET Jet<T, N> cos(const Jet<T, N>& f) { return Jet<T, N>(std::cos(f.a), f.v * (-std::sin(f.a))); }
VS_JET Jet<T, N> tan(const Jet<T, N>& f) {
  const T t = std::tan(f.a);
  return Jet<T, N>(t, f.v * (T(1.0) + t * t));
}
VS_JET Jet<T, N> atan(const Jet<T, N>& f) { return Jet<T, N>(std::atan(f.a), f.v * (T(1.0) / (T(1.0) + f.a * f.a))); }
VS_JET Jet<T, N> asin(const Jet<T, N>& f) { return Jet<T, N>(std::asin(f.a), f.v * (T(1.0) / std::sqrt(T(1.0) - f.a * f.a))); }
VS_JET Jet<T, N> acos(const Jet<T, N>& f) { return Jet<T, N>(std::acos(f.a), f.v * (T(-1.0) / std::sqrt(T(1.0) - f.a * f.a))); }
VS_JET Jet<T, N> atan2(const Jet<T, N>& g, const Jet<T, N>& f) {
  const T tmp = T(1.0) / (f.a * f.a + g.a * g.a);
  return Jet<T, N>(std::atan2(g.a, f.a), (g.v * f.a - f.v * g.a) * tmp);
}
VS_JET Jet<T, N> exp(const Jet<T, N>& f) {
  const T tmp = std::exp(f.a);
  return Jet<T, N>(tmp, f.v * tmp);
}
VS_JET Jet<T, N> log(const Jet<T, N>& f) { return Jet<T, N>(std::log(f.a), f.v * (T(1.0) / f.a)); }
VS_JET Jet<T, N> pow(const Jet<T, N>& f, double g) {
  const T tmp = g * std::pow(f.a, g - T(1.0));
  return Jet<T, N>(std::pow(f.a, g), f.v * tmp);
}
#undef VS_JET

class LocalParameterization {
 public:
  virtual ~LocalParameterization() {}
  virtual bool Plus(const double* x, const double* delta, double* x_plus_delta) const = 0;
  virtual bool ComputeJacobian(const double* x, double* jacobian) const = 0;
  virtual int GlobalSize() const = 0;
  virtual int LocalSize() const = 0;
};

}  // namespace ceres
