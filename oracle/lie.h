// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
//
// SO3 / SE3 arithmetic with the semantics of the (un-vendored, un-pinned) Sophus
// `SO3Group<T>` / `SE3Group<T>` API the reference is written against
// (vicalibrator.h:48, local-param-se3.h:20-24, ceres-cost-functions.h:42-51,361-367,468).
// Storage follows Sophus/Eigen: quaternion coefficients (x,y,z,w); SE3 = [q(4) | t(3)];
// tangent order (upsilon, omega)  (local-param-se3.h:34-37).
#ifndef VICALIB_ORACLE_LIE_H_
#define VICALIB_ORACLE_LIE_H_

#include "dual.h"

namespace vo {

constexpr double kSophusEps = 1e-10;  // SophusConstants<double>::epsilon()

template <class T>
struct Vec3 {
  T x, y, z;
  T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class T>
inline Vec3<T> operator+(const Vec3<T>& a, const Vec3<T>& b) {
  return {a.x + b.x, a.y + b.y, a.z + b.z};
}
template <class T>
inline Vec3<T> operator-(const Vec3<T>& a, const Vec3<T>& b) {
  return {a.x - b.x, a.y - b.y, a.z - b.z};
}
template <class T>
inline Vec3<T> operator*(const Vec3<T>& a, const T& s) {
  return {a.x * s, a.y * s, a.z * s};
}
template <class T>
inline Vec3<T> cross(const Vec3<T>& a, const Vec3<T>& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
template <class T>
inline T dot(const Vec3<T>& a, const Vec3<T>& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z;
}

template <class T>
struct Quat {
  T x, y, z, w;
};

// Eigen quaternion product a*b (Hamilton).
template <class T>
inline Quat<T> qmul(const Quat<T>& a, const Quat<T>& b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
          a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
template <class T>
inline Quat<T> qconj(const Quat<T>& q) {
  return {-q.x, -q.y, -q.z, q.w};
}
// Eigen QuaternionBase::_transformVector: v + w*uv + q.vec x uv with uv = 2 (q.vec x v)
template <class T>
inline Vec3<T> qrot(const Quat<T>& q, const Vec3<T>& v) {
  const Vec3<T> qv{q.x, q.y, q.z};
  Vec3<T> uv = cross(qv, v);
  uv = uv + uv;
  return v + uv * q.w + cross(qv, uv);
}
// Eigen toRotationMatrix (row-major 3x3 in m[9]); SO3::Adj() == matrix().
template <class T>
inline void qmat(const Quat<T>& q, T m[9]) {
  const T tx = T(2.0) * q.x, ty = T(2.0) * q.y, tz = T(2.0) * q.z;
  const T twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const T txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const T tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  m[0] = T(1.0) - (tyy + tzz);
  m[1] = txy - twz;
  m[2] = txz + twy;
  m[3] = txy + twz;
  m[4] = T(1.0) - (txx + tzz);
  m[5] = tyz - twx;
  m[6] = txz - twy;
  m[7] = tyz + twx;
  m[8] = T(1.0) - (txx + tyy);
}
template <class T>
inline Vec3<T> mat_mul(const T m[9], const Vec3<T>& v) {
  return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z,
          m[6] * v.x + m[7] * v.y + m[8] * v.z};
}
template <class T>
inline Quat<T> qnormalized(const Quat<T>& q) {
  const T n = vo::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return {q.x / n, q.y / n, q.z / n, q.w / n};
}

// Sophus SO3::expAndTheta (theta<eps Taylor branch)
template <class T>
inline Quat<T> so3_exp(const Vec3<T>& om, T* theta_out = nullptr) {
  const T theta_sq = dot(om, om);
  T imag, real, theta;
  if (scalar_of(theta_sq) < kSophusEps * kSophusEps) {
    // sqrt(0) would poison the dual part; the Taylor branch only needs theta^2.
    const T theta_po4 = theta_sq * theta_sq;
    imag = T(0.5) - T(1.0 / 48.0) * theta_sq + T(1.0 / 3840.0) * theta_po4;
    real = T(1.0) - T(1.0 / 8.0) * theta_sq + T(1.0 / 384.0) * theta_po4;
    theta = scalar_of(theta_sq) > 0.0 ? vo::sqrt(theta_sq) : T(0.0);
  } else {
    theta = vo::sqrt(theta_sq);
    const T half = T(0.5) * theta;
    imag = vo::sin(half) / theta;
    real = vo::cos(half);
  }
  if (theta_out) *theta_out = theta;
  return {imag * om.x, imag * om.y, imag * om.z, real};
}

// Sophus SO3::logAndTheta
template <class T>
inline Vec3<T> so3_log(const Quat<T>& q, T* theta_out = nullptr) {
  const T sq_n = q.x * q.x + q.y * q.y + q.z * q.z;
  const T w = q.w;
  T f;
  T n = T(0.0);
  if (scalar_of(sq_n) < kSophusEps * kSophusEps) {
    const T sq_w = w * w;
    f = T(2.0) / w - T(2.0) * sq_n / (w * sq_w);
    if (scalar_of(sq_n) > 0.0) n = vo::sqrt(sq_n);
  } else {
    n = vo::sqrt(sq_n);
    if (std::fabs(scalar_of(w)) < kSophusEps) {
      f = (scalar_of(w) > 0.0 ? T(M_PI) : T(-M_PI)) / n;
    } else {
      f = T(2.0) * vo::atan(n / w) / n;
    }
  }
  if (theta_out) *theta_out = f * n;
  return {f * q.x, f * q.y, f * q.z};
}

template <class T>
struct SE3 {
  Quat<T> q;
  Vec3<T> t;
};
template <class T>
inline SE3<T> se3_from(const T* p) {  // [qx qy qz qw tx ty tz]
  return {{p[0], p[1], p[2], p[3]}, {p[4], p[5], p[6]}};
}
template <class T>
inline void se3_to(const SE3<T>& s, T* p) {
  p[0] = s.q.x; p[1] = s.q.y; p[2] = s.q.z; p[3] = s.q.w;
  p[4] = s.t.x; p[5] = s.t.y; p[6] = s.t.z;
}
template <class T>
inline SE3<T> se3_mul(const SE3<T>& a, const SE3<T>& b) {
  return {qmul(a.q, b.q), a.t + qrot(a.q, b.t)};
}
template <class T>
inline SE3<T> se3_inv(const SE3<T>& a) {
  const Quat<T> qi = qconj(a.q);
  const Vec3<T> ti = qrot(qi, a.t);
  return {qi, {-ti.x, -ti.y, -ti.z}};
}
template <class T>
inline Vec3<T> se3_act(const SE3<T>& a, const Vec3<T>& p) {
  return qrot(a.q, p) + a.t;
}

template <class T>
inline void hat_sq(const Vec3<T>& w, T O[9], T O2[9]) {
  O[0] = T(0.0); O[1] = -w.z; O[2] = w.y;
  O[3] = w.z; O[4] = T(0.0); O[5] = -w.x;
  O[6] = -w.y; O[7] = w.x; O[8] = T(0.0);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      T s = T(0.0);
      for (int k = 0; k < 3; ++k) s = s + O[i * 3 + k] * O[k * 3 + j];
      O2[i * 3 + j] = s;
    }
}

// Sophus SE3::exp: t = V(omega) * upsilon
template <class T>
inline SE3<T> se3_exp(const T* d) {  // d = (upsilon, omega)
  const Vec3<T> ups{d[0], d[1], d[2]}, om{d[3], d[4], d[5]};
  T theta;
  const Quat<T> q = so3_exp(om, &theta);
  T O[9], O2[9], V[9];
  hat_sq(om, O, O2);
  if (scalar_of(theta) < kSophusEps) {
    qmat(q, V);  // "V = so3.matrix()" in Sophus' small-angle branch
  } else {
    const T th2 = theta * theta;
    const T c1 = (T(1.0) - vo::cos(theta)) / th2;
    const T c2 = (theta - vo::sin(theta)) / (th2 * theta);
    for (int i = 0; i < 9; ++i) V[i] = c1 * O[i] + c2 * O2[i];
    V[0] = V[0] + T(1.0); V[4] = V[4] + T(1.0); V[8] = V[8] + T(1.0);
  }
  return {q, mat_mul(V, ups)};
}

// Sophus SE3::log -> (upsilon, omega)
template <class T>
inline void se3_log(const SE3<T>& s, T out[6]) {
  T theta;
  const Vec3<T> om = so3_log(s.q, &theta);
  T O[9], O2[9], Vi[9];
  hat_sq(om, O, O2);
  T c;
  if (std::fabs(scalar_of(theta)) < kSophusEps) {
    c = T(1.0 / 12.0);
  } else {
    c = (T(1.0) - theta / (T(2.0) * vo::tan(theta / T(2.0)))) / (theta * theta);
  }
  for (int i = 0; i < 9; ++i) Vi[i] = T(-0.5) * O[i] + c * O2[i];
  Vi[0] = Vi[0] + T(1.0); Vi[4] = Vi[4] + T(1.0); Vi[8] = Vi[8] + T(1.0);
  const Vec3<T> u = mat_mul(Vi, s.t);
  out[0] = u.x; out[1] = u.y; out[2] = u.z;
  out[3] = om.x; out[4] = om.y; out[5] = om.z;
}

}  // namespace vo
#endif
