// Device-side math for the calibration solve: quaternions / SE3 with Sophus semantics, the
// camera models' Project with analytic Jacobians, loss functions.
//
// Reference semantics being reproduced (the reference gets Jacobians from ceres autodiff; here
// they are derived by hand and checked against the dual-number oracle in tests/):
//   ImuReprojectionCostFunctor      ceres-cost-functions.h:342-377
//   Calibu Project (un-vendored)    call site ceres-cost-functions.h:369, SURVEY App. A.2
//   LocalParamSe3/So3 Plus+Jacobian local-param-se3.h:14-91, 107-157
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace vc {

enum { kLinear = 0, kFov = 1, kPoly2 = 2, kPoly3 = 3, kKb4 = 4 };
constexpr double kSophusEps = 1e-10;

__host__ __device__ inline int num_intr(int model) {
  return model == kLinear ? 4 : model == kFov ? 5 : model == kPoly2 ? 6 : model == kPoly3 ? 7 : 8;
}

struct Q4 {
  double x, y, z, w;
};
struct V3 {
  double x, y, z;
};

__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
__device__ __forceinline__ Q4 qconj(Q4 q) { return {-q.x, -q.y, -q.z, q.w}; }
__device__ __forceinline__ Q4 qnormalized(Q4 q) {
  const double n = 1.0 / sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return {q.x * n, q.y * n, q.z * n, q.w * n};
}
// rotation matrix (row-major) of a unit quaternion
__device__ __forceinline__ void qmat(Q4 q, double R[9]) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}
__device__ __forceinline__ V3 mat_mul(const double R[9], V3 v) {
  return {R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z,
          R[6] * v.x + R[7] * v.y + R[8] * v.z};
}
__device__ __forceinline__ V3 mat_tmul(const double R[9], V3 v) {
  return {R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z,
          R[2] * v.x + R[5] * v.y + R[8] * v.z};
}
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
  const V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = uv + uv;
  return v + uv * q.w + cross(qv, uv);
}

// Sophus SO3::exp (theta < eps Taylor branch)
__device__ inline Q4 so3_exp(V3 om, double* theta_out = nullptr) {
  const double th2 = dot(om, om);
  const double th = sqrt(th2);
  double imag, real;
  if (th < kSophusEps) {
    const double th4 = th2 * th2;
    imag = 0.5 - th2 * (1.0 / 48.0) + th4 * (1.0 / 3840.0);
    real = 1.0 - th2 * (1.0 / 8.0) + th4 * (1.0 / 384.0);
  } else {
    double s, c;
    sincos(0.5 * th, &s, &c);
    imag = s / th;
    real = c;
  }
  if (theta_out) *theta_out = th;
  return {imag * om.x, imag * om.y, imag * om.z, real};
}

// LocalParamSe3::Plus: T * exp(upsilon, omega), quaternion renormalised like Sophus' group product
__device__ inline void se3_plus(const double* x, const double* d, double* out) {
  const Q4 q{x[0], x[1], x[2], x[3]};
  const V3 ups{d[0], d[1], d[2]}, om{d[3], d[4], d[5]};
  double th;
  const Q4 dq = so3_exp(om, &th);
  // t' = t + R (V upsilon)
  V3 vu;
  if (th < kSophusEps) {
    double Rd[9];
    qmat(dq, Rd);
    vu = mat_mul(Rd, ups);
  } else {
    const double th2 = th * th;
    double s, c;
    sincos(th, &s, &c);
    const double c1 = (1.0 - c) / th2, c2 = (th - s) / (th2 * th);
    const V3 ou = cross(om, ups);
    const V3 oou = cross(om, ou);
    vu = ups + ou * c1 + oou * c2;
  }
  const V3 t = V3{x[4], x[5], x[6]} + qrot(q, vu);
  const Q4 qo = qnormalized(qmul(q, dq));
  out[0] = qo.x; out[1] = qo.y; out[2] = qo.z; out[3] = qo.w;
  out[4] = t.x; out[5] = t.y; out[6] = t.z;
}
__device__ inline void so3_plus(const double* x, const double* d, double* out) {
  const Q4 qo = qnormalized(qmul(Q4{x[0], x[1], x[2], x[3]}, so3_exp(V3{d[0], d[1], d[2]})));
  out[0] = qo.x; out[1] = qo.y; out[2] = qo.z; out[3] = qo.w;
}

// ------------------------------------------------------------------ camera models
// project<MODEL>(pc, intr, z, dz_dp[2x3], dz_di[2xK]); dz_* may be null (residual only).
// Every model is written as  z = (fu * s * X' + u0, fv * s * Y' + v0)  and the derivatives
// follow by the chain rule; branch conditions mirror the reference model's own branches so the
// autodiff oracle and this code take the same side.
template <int MODEL>
struct Cam;

template <>
struct Cam<kLinear> {
  static constexpr int K = 4;
  __device__ static inline void project(V3 p, const double* in, double z[2], double* dzp, double* dzi) {
    const double iz = 1.0 / p.z, u = p.x * iz, v = p.y * iz;
    z[0] = in[0] * u + in[2];
    z[1] = in[1] * v + in[3];
    if (!dzp) return;
    dzp[0] = in[0] * iz; dzp[1] = 0.0; dzp[2] = -in[0] * u * iz;
    dzp[3] = 0.0; dzp[4] = in[1] * iz; dzp[5] = -in[1] * v * iz;
    dzi[0] = u; dzi[1] = 0.0; dzi[2] = 1.0; dzi[3] = 0.0;
    dzi[K + 0] = 0.0; dzi[K + 1] = v; dzi[K + 2] = 0.0; dzi[K + 3] = 1.0;
  }
};

// shared tail for the "radial factor on normalised coordinates" models:
// f = factor, q = (1/rad) df/drad
template <int K>
__device__ __forceinline__ void radial_tail(const double* in, double u, double v, double iz, double r2,
                                            double f, double q, double z[2], double* dzp, double* dzi) {
  z[0] = in[0] * f * u + in[2];
  z[1] = in[1] * f * v + in[3];
  if (!dzp) return;
  const double fz = f + r2 * q;
  dzp[0] = in[0] * (f + u * u * q) * iz;
  dzp[1] = in[0] * (u * v * q) * iz;
  dzp[2] = -in[0] * iz * u * fz;
  dzp[3] = in[1] * (u * v * q) * iz;
  dzp[4] = in[1] * (f + v * v * q) * iz;
  dzp[5] = -in[1] * iz * v * fz;
  dzi[0] = f * u; dzi[1] = 0.0; dzi[2] = 1.0; dzi[3] = 0.0;
  dzi[K + 0] = 0.0; dzi[K + 1] = f * v; dzi[K + 2] = 0.0; dzi[K + 3] = 1.0;
}

template <>
struct Cam<kFov> {
  static constexpr int K = 5;
  __device__ static inline void project(V3 p, const double* in, double z[2], double* dzp, double* dzi) {
    const double iz = 1.0 / p.z, u = p.x * iz, v = p.y * iz, r2 = u * u + v * v;
    const double w = in[4];
    double f = 1.0, q = 0.0, fw = 0.0;
    if (w * w > 1e-5) {
      const double t = tan(0.5 * w), m = 2.0 * t, dm = 1.0 + t * t;
      if (r2 < 1e-5) {
        f = m / w;
        fw = (dm * w - m) / (w * w);
      } else {
        const double rad = sqrt(r2), a = atan(rad * m), den = 1.0 + r2 * m * m;
        f = a / (rad * w);
        q = (m * rad / den - a) / (r2 * rad * w);
        fw = dm / (den * w) - f / w;
      }
    }
    radial_tail<K>(in, u, v, iz, r2, f, q, z, dzp, dzi);
    if (!dzp) return;
    dzi[4] = in[0] * u * fw;
    dzi[K + 4] = in[1] * v * fw;
  }
};

template <int ORDER>
struct PolyCam {
  static constexpr int K = 4 + ORDER;
  __device__ static inline void project(V3 p, const double* in, double z[2], double* dzp, double* dzi) {
    const double iz = 1.0 / p.z, u = p.x * iz, v = p.y * iz, r2 = u * u + v * v;
    const double k1 = in[4], k2 = in[5], k3 = (ORDER == 3) ? in[6] : 0.0;
    const double r4 = r2 * r2;
    const double f = 1.0 + k1 * r2 + k2 * r4 + k3 * r4 * r2;
    const double q = 2.0 * (k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4);
    radial_tail<K>(in, u, v, iz, r2, f, q, z, dzp, dzi);
    if (!dzp) return;
    const double au = in[0] * u, av = in[1] * v;
    dzi[4] = au * r2; dzi[K + 4] = av * r2;
    dzi[5] = au * r4; dzi[K + 5] = av * r4;
    if (ORDER == 3) { dzi[6] = au * r4 * r2; dzi[K + 6] = av * r4 * r2; }
  }
};
template <>
struct Cam<kPoly2> : PolyCam<2> {};
template <>
struct Cam<kPoly3> : PolyCam<3> {};

template <>
struct Cam<kKb4> {
  static constexpr int K = 8;
  __device__ static inline void project(V3 p, const double* in, double z[2], double* dzp, double* dzi) {
    const double rho2 = p.x * p.x + p.y * p.y;
    const double rho = sqrt(rho2);
    const double th = atan2(rho, p.z), th2 = th * th;
    const double d = th * (1.0 + th2 * (in[4] + th2 * (in[5] + th2 * (in[6] + th2 * in[7]))));
    // rho -> 0 limit: d/rho -> 1/Z (pinhole); the autodiff reference is singular there.
    const bool axis = rho2 < 1e-30;
    const double irho = axis ? 0.0 : 1.0 / rho;
    const double c = axis ? 1.0 : p.x * irho, s = axis ? 0.0 : p.y * irho;
    z[0] = in[0] * d * c + in[2];
    z[1] = in[1] * d * s + in[3];
    if (!dzp) return;
    const double dp = 1.0 + th2 * (3.0 * in[4] + th2 * (5.0 * in[5] + th2 * (7.0 * in[6] + th2 * 9.0 * in[7])));
    const double in2 = 1.0 / (rho2 + p.z * p.z);
    const double tx = p.z * in2 * c, ty = p.z * in2 * s, tz = -rho * in2;  // dtheta/d(X,Y,Z)
    const double dor = axis ? 1.0 / p.z : d * irho;
    dzp[0] = in[0] * (dp * tx * c + dor * s * s);
    dzp[1] = in[0] * (dp * ty * c - dor * c * s);
    dzp[2] = in[0] * dp * tz * c;
    dzp[3] = in[1] * (dp * tx * s - dor * c * s);
    dzp[4] = in[1] * (dp * ty * s + dor * c * c);
    dzp[5] = in[1] * dp * tz * s;
    dzi[0] = d * c; dzi[1] = 0.0; dzi[2] = 1.0; dzi[3] = 0.0;
    dzi[K + 0] = 0.0; dzi[K + 1] = d * s; dzi[K + 2] = 0.0; dzi[K + 3] = 1.0;
    const double th3 = th2 * th, ac = in[0] * c, as = in[1] * s;
    double tp = th3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dzi[4 + i] = ac * tp;
      dzi[K + 4 + i] = as * tp;
      tp *= th2;
    }
  }
};

// Ceres SoftLOneLoss(0.5) / CauchyLoss(100) (vicalibrator.h:127,133): rho and rho'
__device__ __forceinline__ void soft_l_one(double s, double* rho0, double* rho1) {
  const double b = 0.25, sum = 1.0 + s * (1.0 / b), tmp = sqrt(sum);
  *rho0 = 2.0 * b * (tmp - 1.0);
  *rho1 = 1.0 / tmp;
}
__device__ __forceinline__ void cauchy(double s, double* rho0, double* rho1) {
  const double b = 1e4, sum = 1.0 + s * (1.0 / b);
  *rho0 = b * log(sum);
  *rho1 = 1.0 / sum;
}

}  // namespace vc
