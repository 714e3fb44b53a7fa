// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
//
// Restatement of the per-iteration IMU weight update:
//   ViCalibrator::UpdateImuWeights                       vicalibrator.h:723-799
//   ImuResidualT::IntegratePose / GetPoseDerivative /
//     IntegrateImu / IntegrateResidual (double + Jacobians + covariance)
//                                                        types.h:330-378, 380-425, 427-595, 611-687
//   dLog_dq, dqExp_dw, dq1q2_dq2, dq1q2_dq1, dqx_dq, dt1t2_dt1, dLog_dSE3
//                                                        vicalibrator-utils.h:106-154,187-202,214-230,234-274,307-434
// Formulas are restated AS WRITTEN, approximations included (dk_dx ignores the scale factors,
// types.h:413-423; dqExp_dw is a truncated series; dLog_dSE3's small-angle oddity at
// vicalibrator-utils.h:372).  Only the values that reach an output are computed:
// per IMU step C <- A C A^T + G R G^T with A = dy_dy0, G = dy_db the full-step RK4 Jacobians.
#ifndef VICALIB_ORACLE_IMU_WEIGHTS_H_
#define VICALIB_ORACLE_IMU_WEIGHTS_H_

#include <cmath>
#include <cstring>

#include "cost_functors.h"

namespace vo {

template <int R, int C>
struct Mat {
  double m[R * C];
  Mat() { for (int i = 0; i < R * C; ++i) m[i] = 0.0; }
  double& operator()(int i, int j) { return m[i * C + j]; }
  double operator()(int i, int j) const { return m[i * C + j]; }
  static Mat Identity() {
    Mat r;
    for (int i = 0; i < (R < C ? R : C); ++i) r(i, i) = 1.0;
    return r;
  }
};
template <int R, int K, int C>
inline Mat<R, C> operator*(const Mat<R, K>& a, const Mat<K, C>& b) {
  Mat<R, C> r;
  for (int i = 0; i < R; ++i)
    for (int k = 0; k < K; ++k) {
      const double x = a(i, k);
      if (x == 0.0) continue;
      for (int j = 0; j < C; ++j) r(i, j) += x * b(k, j);
    }
  return r;
}
template <int R, int C>
inline Mat<R, C> operator+(const Mat<R, C>& a, const Mat<R, C>& b) {
  Mat<R, C> r;
  for (int i = 0; i < R * C; ++i) r.m[i] = a.m[i] + b.m[i];
  return r;
}
template <int R, int C>
inline Mat<R, C> operator*(const Mat<R, C>& a, double s) {
  Mat<R, C> r;
  for (int i = 0; i < R * C; ++i) r.m[i] = a.m[i] * s;
  return r;
}
template <int R, int C>
inline Mat<C, R> Tr(const Mat<R, C>& a) {
  Mat<C, R> r;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) r(j, i) = a(i, j);
  return r;
}
template <int R, int C, int BR, int BC>
inline void SetBlock(Mat<R, C>* dst, int r0, int c0, const Mat<BR, BC>& src) {
  for (int i = 0; i < BR; ++i)
    for (int j = 0; j < BC; ++j) (*dst)(r0 + i, c0 + j) = src(i, j);
}

inline double powi(double x, int y) {  // vicalibrator-utils.h:69-82
  if (y == 0) return 1.0;
  if (y < 0) return 1.0 / powi(x, -y);
  double ret = x;
  for (int ii = 1; ii < y; ++ii) ret *= x;
  return ret;
}

// vicalibrator-utils.h:106-154
inline Mat<3, 4> dLog_dq(const Quat<double>& q) {
  const double x = q.x, y = q.y, z = q.z, w = q.w;
  const double vec_squarednorm = powi(x, 2) + powi(y, 2) + powi(z, 2);
  const double vec_norm = std::sqrt(vec_squarednorm);
  Mat<3, 4> r;
  if (vec_norm < 1e-9) {
    const double s1 = 2 * vec_squarednorm;
    const double s2 = 1.0 / powi(w, 3);
    const double s3 = (3 * s1) / powi(w, 4) - 2 / powi(w, 2);
    const double s4 = 2 / w;
    const double v[12] = {-4 * s2 * powi(x, 2) + s4 - s1 * s2, -4 * x * y * s2, -4 * x * z * s2, x * s3,
                          -4 * x * y * s2, -4 * s2 * powi(y, 2) + s4 - s1 * s2, -4 * y * z * s2, y * s3,
                          -4 * x * z * s2, -4 * y * z * s2, -4 * s2 * powi(z, 2) + s4 - s1 * s2, z * s3};
    std::memcpy(r.m, v, sizeof v);
  } else {
    const double s1 = vec_squarednorm;
    const double s2 = 1 / (s1 / powi(w, 2) + 1);
    const double s3 = std::atan(std::sqrt(s1) / w);
    const double s4 = 1 / std::pow(s1, (3.0 / 2.0));
    const double s5 = 1 / s1;
    const double s6 = 1 / w;
    const double s7 = (2 * s3) / std::sqrt(s1);
    const double s8 = 2 * y * z * s2 * s5 * s6 - 2 * y * z * s3 * s4;
    const double s9 = 2 * x * z * s2 * s5 * s6 - 2 * x * z * s3 * s4;
    const double s10 = 2 * x * y * s2 * s5 * s6 - 2 * x * y * s3 * s4;
    const double v[12] = {s7 - 2 * powi(x, 2) * s3 * s4 + 2 * powi(x, 2) * s2 * s5 * s6, s10, s9,
                          -(2 * x * s2) / powi(w, 2),
                          s10, s7 - 2 * powi(y, 2) * s3 * s4 + 2 * powi(y, 2) * s2 * s5 * s6, s8,
                          -(2 * y * s2) / powi(w, 2),
                          s9, s8, s7 - 2 * powi(z, 2) * s3 * s4 + 2 * powi(z, 2) * s2 * s5 * s6,
                          -(2 * z * s2) / powi(w, 2)};
    std::memcpy(r.m, v, sizeof v);
  }
  return r;
}

// vicalibrator-utils.h:187-202
inline Mat<4, 3> dqExp_dw(const Vec3<double>& w) {
  const double t = std::sqrt(dot(w, w));
  const double s1 = t / 20 - 1;
  const double s2 = powi(t, 2) / 48 - 0.5;
  const double s3 = (s1 * w[1] * w[2]) / 24;
  const double s4 = (s1 * w[0] * w[2]) / 24;
  const double s5 = (s1 * w[0] * w[1]) / 24;
  const double s6 = powi(t, 2);
  Mat<4, 3> r;
  const double v[12] = {(s1 * powi(w[0], 2)) / 24 - s6 / 48 + 0.5, s5, s4,
                        s5, (s1 * powi(w[1], 2)) / 24 - s6 / 48 + 0.5, s3,
                        s4, s3, (s1 * powi(w[2], 2)) / 24 - s6 / 48 + 0.5,
                        (s2 * w[0]) / 2, (s2 * w[1]) / 2, (s2 * w[2]) / 2};
  std::memcpy(r.m, v, sizeof v);
  return r;
}
// vicalibrator-utils.h:214-220
inline Mat<4, 4> dq1q2_dq2(const Quat<double>& q1) {
  Mat<4, 4> r;
  const double v[16] = {q1.w, -q1.z, q1.y, q1.x, q1.z, q1.w, -q1.x, q1.y,
                        -q1.y, q1.x, q1.w, q1.z, -q1.x, -q1.y, -q1.z, q1.w};
  std::memcpy(r.m, v, sizeof v);
  return r;
}
// vicalibrator-utils.h:224-230
inline Mat<4, 4> dq1q2_dq1(const Quat<double>& q2) {
  Mat<4, 4> r;
  const double v[16] = {q2.w, q2.z, -q2.y, q2.x, -q2.z, q2.w, q2.x, q2.y,
                        q2.y, -q2.x, q2.w, q2.z, -q2.x, -q2.y, -q2.z, q2.w};
  std::memcpy(r.m, v, sizeof v);
  return r;
}
// vicalibrator-utils.h:234-254
inline Mat<3, 4> dqx_dq(const Quat<double>& q, const Vec3<double>& vec) {
  const double x = vec[0], y = vec[1], z = vec[2];
  const double s1 = 2 * q.x * y;
  const double s2 = 2 * q.y * y;
  const double s3 = 2 * q.x * x;
  const double s4 = 2 * q.z * x;
  const double s5 = 2 * q.y * z;
  const double s6 = 2 * q.z * z;
  Mat<3, 4> r;
  const double v[12] = {s2 + s6, s1 - 4 * q.y * x + 2 * q.w * z, 2 * q.x * z - 2 * q.w * y - 4 * q.z * x,
                        s5 - 2 * q.z * y,
                        2 * q.y * x - 4 * q.x * y - 2 * q.w * z, s3 + s6, s5 + 2 * q.w * x - 4 * q.z * y,
                        s4 - 2 * q.x * z,
                        s4 + 2 * q.w * y - 4 * q.x * z, 2 * q.z * y - 2 * q.w * x - 4 * q.y * z, s2 + s3,
                        s1 - 2 * q.y * x};
  std::memcpy(r.m, v, sizeof v);
  return r;
}
// vicalibrator-utils.h:260-274 (7-vector = translation(3), quaternion(4))
inline Mat<7, 7> dt1t2_dt1(const SE3<double>& t1, const SE3<double>& t2) {
  Mat<7, 7> r;
  SetBlock(&r, 0, 0, Mat<3, 3>::Identity());
  SetBlock(&r, 0, 3, dqx_dq(t1.q, t2.t));
  SetBlock(&r, 3, 3, dq1q2_dq1(t2.q));
  return r;
}

// vicalibrator-utils.h:307-434
inline Mat<6, 7> dLog_dSE3(const SE3<double>& t) {
  const Mat<3, 4> dw_dq = dLog_dq(t.q);
  const double x = t.t.x, y = t.t.y, z = t.t.z;
  double theta;
  const Vec3<double> w = so3_log(t.q, &theta);
  const double wx = w.x, wy = w.y, wz = w.z;
  double O[9], O2[9];
  hat_sq(w, O, O2);
  const bool close_to_zero = std::fabs(theta) < kSophusEps;
  Mat<3, 3> v_inv;
  {
    const double c = close_to_zero ? 1. / 12.
                                   : (1.0 - theta / (2.0 * std::tan(theta / 2.0))) / (theta * theta);
    for (int i = 0; i < 9; ++i) v_inv.m[i] = -0.5 * O[i] + c * O2[i];
    v_inv(0, 0) += 1.0; v_inv(1, 1) += 1.0; v_inv(2, 2) += 1.0;
  }
  Mat<6, 7> dlog;
  SetBlock(&dlog, 0, 0, v_inv);
  SetBlock(&dlog, 3, 3, dw_dq);
  Mat<3, 3> dlog_dw;
  if (close_to_zero) {
    const double div_12 = 1. / 12, div_6 = 1. / 6.;
    const double wx_x = wx * x, wy_x = wy * x, wz_x = wz * x;
    const double wx_y = wx * y, wy_y = wy * y, wz_y = wz * y;
    const double wx_z = wx * z, wy_z = wy * z, wz_z = wz * z;
    const double v[9] = {div_12 * (wy_y + wz_z), div_12 * wx_y - div_6 * wy_x - 0.5 * z,
                         0.5 * y - div_6 * wz_x + div_12 * wx_z,
                         0.5 * z + div_12 * wy_x - div_6 * wx_y, div_12 * (wx_x + wz_z),
                         div_12 * wy_z - div_6 * wz_y - 0.5 * x,
                         div_12 * wz_x - div_6 * wx_z - 0.5 * y, 0.5 * x + div_12 * wz_y - div_6 * wy_z,
                         div_12 * (wx_x * wy_y)};
    std::memcpy(dlog_dw.m, v, sizeof v);
  } else {
    const double s1 = powi(wx, 2) + powi(wy, 2) + powi(wz, 2);
    const double s2 = std::tan(std::sqrt(s1) / 2);
    const double s3 = std::sqrt(s1) / (2 * s2) - 1;
    const double s4 = wz / (2 * std::sqrt(s1) * s2) - (wz * (powi(s2, 2) + 1)) / (4 * powi(s2, 2));
    const double s5 = wy / (2 * std::sqrt(s1) * s2) - (wy * (powi(s2, 2) + 1)) / (4 * powi(s2, 2));
    const double s6 = wx / (2 * std::sqrt(s1) * s2) - (wx * (powi(s2, 2) + 1)) / (4 * powi(s2, 2));
    const double s7 = 1 / s1;
    const double s8 = 1 / powi(s1, 2);
    const double s9 = powi(wx, 2) + powi(wy, 2);
    const double s10 = powi(wx, 2) + powi(wz, 2);
    const double s11 = powi(wy, 2) + powi(wz, 2);
    const double s12 = 2 * s3 * s8 * wx * wy * wz;
    const double s13 = -2 * s3 * s8 * wy * powi(wz, 2) + s4 * s7 * wy * wz + s3 * s7 * wy;
    const double s14 = -2 * s3 * s8 * wx * powi(wz, 2) + s4 * s7 * wx * wz + s3 * s7 * wx;
    const double s15 = -2 * s3 * s8 * wz * powi(wy, 2) + s5 * s7 * wz * wy + s3 * s7 * wz;
    const double s16 = -2 * s3 * s8 * wz * powi(wx, 2) + s6 * s7 * wz * wx + s3 * s7 * wz;
    const double s17 = -2 * s3 * s8 * wx * powi(wy, 2) + s5 * s7 * wx * wy + s3 * s7 * wx;
    const double s18 = -2 * s3 * s8 * wy * powi(wx, 2) + s6 * s7 * wy * wx + s3 * s7 * wy;
    const double s19 = 2 * s3 * s7 * wy;
    const double s20 = 2 * s3 * s7 * wx;
    const double v[9] = {
        x * (s6 * s7 * s11 - 2 * s3 * s8 * s11 * wx) - s18 * y - s16 * z,
        x * (s19 + s5 * s7 * s11 - 2 * s3 * s8 * s11 * wy) - s17 * y -
            z * (s5 * s7 * wx * wz - 2 * s3 * s8 * wx * wy * wz + 0.5),
        x * (s4 * s7 * s11 + 2 * s3 * s7 * wz - 2 * s3 * s8 * s11 * wz) - s14 * z +
            y * (s12 - s4 * s7 * wx * wy + 0.5),
        y * (s20 + s6 * s7 * s10 - 2 * s3 * s8 * s10 * wx) - s18 * x + z * (s12 - s6 * s7 * wy * wz + 0.5),
        y * (s5 * s7 * s10 - 2 * s3 * s8 * s10 * wy) - s17 * x - s15 * z,
        y * (s4 * s7 * s10 + 2 * s3 * s7 * wz - 2 * s3 * s8 * s10 * wz) - s13 * z -
            x * (s4 * s7 * wx * wy - s12 + 0.5),
        z * (s20 + s6 * s7 * s9 - 2 * s3 * s8 * s9 * wx) - s16 * x - y * (s6 * s7 * wy * wz - s12 + 0.5),
        z * (s19 + s5 * s7 * s9 - 2 * s3 * s8 * s9 * wy) - s15 * y + x * (s12 - s5 * s7 * wx * wz + 0.5),
        z * (s4 * s7 * s9 - 2 * s3 * s8 * s9 * wz) - s14 * x - s13 * y};
    std::memcpy(dlog_dw.m, v, sizeof v);
  }
  SetBlock(&dlog, 0, 3, dlog_dw * dw_dq);
  return dlog;
}

// ---------------------------------------------------------------- double integrator
struct DPose {  // ImuPoseT<double> restricted to what the weight update reads
  SE3<double> t_wp;
  Vec3<double> v_w;
};

// types.h:330-378
inline DPose IntegratePoseD(const DPose& pose, const double k[9], double dt, Mat<10, 9>* dy_dk,
                            Mat<10, 10>* dy_dy) {
  const Vec3<double> wdt{k[3] * dt, k[4] * dt, k[5] * dt};
  const Quat<double> r = so3_exp(wdt);
  DPose y = pose;
  y.t_wp.t = pose.t_wp.t + Vec3<double>{k[0] * dt, k[1] * dt, k[2] * dt};
  y.t_wp.q = qmul(r, pose.t_wp.q);
  y.v_w = pose.v_w + Vec3<double>{k[6] * dt, k[7] * dt, k[8] * dt};
  if (dy_dk) {
    *dy_dk = Mat<10, 9>();
    SetBlock(dy_dk, 0, 0, Mat<3, 3>::Identity() * dt);
    SetBlock(dy_dk, 3, 3, (dq1q2_dq1(pose.t_wp.q) * dqExp_dw(wdt)) * dt);
    SetBlock(dy_dk, 7, 6, Mat<3, 3>::Identity() * dt);
  }
  if (dy_dy) {
    *dy_dy = Mat<10, 10>();
    SetBlock(dy_dy, 0, 0, Mat<3, 3>::Identity());
    SetBlock(dy_dy, 3, 3, dq1q2_dq2(r));
    SetBlock(dy_dy, 7, 7, Mat<3, 3>::Identity());
  }
  return y;
}

// types.h:380-425
inline void GetPoseDerivativeD(const DPose& pose, const Vec3<double>& g_w, const ImuMeas<double>& z0,
                               const ImuMeas<double>& z1, const Vec3<double>& bg, const Vec3<double>& ba,
                               const double sf[6], double dt, double deriv[9], Mat<9, 6>* dk_db,
                               Mat<9, 10>* dk_dx) {
  const double alpha = (z1.time - (z0.time + dt)) / (z1.time - z0.time);
  const Vec3<double> zg = z0.w * alpha + z1.w * (1.0 - alpha);
  const Vec3<double> za = z0.a * alpha + z1.a * (1.0 - alpha);
  double R[9];
  qmat(pose.t_wp.q, R);
  deriv[0] = pose.v_w.x; deriv[1] = pose.v_w.y; deriv[2] = pose.v_w.z;
  const Vec3<double> w = mat_mul(R, Vec3<double>{zg.x * sf[0] + bg.x, zg.y * sf[1] + bg.y, zg.z * sf[2] + bg.z});
  deriv[3] = w.x; deriv[4] = w.y; deriv[5] = w.z;
  const Vec3<double> a =
      qrot(pose.t_wp.q, Vec3<double>{za.x * sf[3] + ba.x, za.y * sf[4] + ba.y, za.z * sf[5] + ba.z}) - g_w;
  deriv[6] = a.x; deriv[7] = a.y; deriv[8] = a.z;
  if (dk_db) {
    *dk_db = Mat<9, 6>();
    Mat<3, 3> Rm;
    std::memcpy(Rm.m, R, sizeof R);
    SetBlock(dk_db, 3, 0, Rm);
    SetBlock(dk_db, 6, 3, Rm);
  }
  if (dk_dx) {
    *dk_dx = Mat<9, 10>();
    SetBlock(dk_dx, 0, 7, Mat<3, 3>::Identity());
    SetBlock(dk_dx, 3, 3, dqx_dq(pose.t_wp.q, zg) + dqx_dq(pose.t_wp.q, bg));
    SetBlock(dk_dx, 6, 3, dqx_dq(pose.t_wp.q, za) + dqx_dq(pose.t_wp.q, ba));
  }
}

// types.h:427-595 (Jacobian + covariance branch; only *c_prior is an output)
inline DPose IntegrateImuD(const DPose& pose, const ImuMeas<double>& z0, const ImuMeas<double>& z1,
                           const Vec3<double>& bg, const Vec3<double>& ba, const double sf[6],
                           const Vec3<double>& g, Mat<10, 6>* dy_db_out, Mat<10, 10>* dy_dy0_out,
                           Mat<10, 10>* c_prior, const Mat<6, 6>& cov_meas) {
  const double dt = z1.time - z0.time;
  if (dt == 0) {
    // The reference returns with its Jacobian outputs untouched (uninitialised); treat the
    // degenerate zero-length step as the identity map.
    *dy_db_out = Mat<10, 6>();
    *dy_dy0_out = Mat<10, 10>::Identity();
    return pose;
  }
  Mat<10, 6> dy_db;
  Mat<10, 10> dy_dy0 = Mat<10, 10>::Identity();
  Mat<9, 6> dk_db;
  Mat<9, 10> dk_dy;
  Mat<10, 9> dy_dk;
  Mat<10, 10> dy_dy;
  double k1[9], k2[9], k3[9], k4[9], k[9];

  GetPoseDerivativeD(pose, g, z0, z1, bg, ba, sf, 0, k1, &dk_db, &dk_dy);
  const Mat<9, 6> dk1_db = dk_db + dk_dy * dy_db;
  const Mat<9, 10> dk1_dy = dk_dy * dy_dy0;
  const DPose y1 = IntegratePoseD(pose, k1, dt * 0.5, &dy_dk, &dy_dy);
  dy_db = dy_dk * dk1_db;
  dy_dy0 = dy_dy + dy_dk * dk1_dy;

  GetPoseDerivativeD(y1, g, z0, z1, bg, ba, sf, dt / 2, k2, &dk_db, &dk_dy);
  const Mat<9, 6> dk2_db = dk_db + dk_dy * dy_db;
  const Mat<9, 10> dk2_dy = dk_dy * dy_dy0;
  const DPose y2 = IntegratePoseD(pose, k2, dt * 0.5, &dy_dk, &dy_dy);
  dy_db = dy_dk * dk2_db;
  dy_dy0 = dy_dy + dy_dk * dk2_dy;

  GetPoseDerivativeD(y2, g, z0, z1, bg, ba, sf, dt / 2, k3, &dk_db, &dk_dy);
  const Mat<9, 6> dk3_db = dk_db + dk_dy * dy_db;
  const Mat<9, 10> dk3_dy = dk_dy * dy_dy0;
  const DPose y3 = IntegratePoseD(pose, k3, dt, &dy_dk, &dy_dy);
  dy_db = dy_dk * dk3_db;
  dy_dy0 = dy_dy + dy_dk * dk3_dy;

  GetPoseDerivativeD(y3, g, z0, z1, bg, ba, sf, dt, k4, &dk_db, &dk_dy);
  const Mat<9, 6> dk4_db = dk_db + dk_dy * dy_db;
  const Mat<9, 10> dk4_dy = dk_dy * dy_dy0;

  for (int i = 0; i < 9; ++i) k[i] = k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i];
  const Mat<9, 6> dk_total_db = dk1_db + dk2_db * 2.0 + dk3_db * 2.0 + dk4_db;
  const Mat<9, 10> dk_total_dy = dk1_dy + dk2_dy * 2.0 + dk3_dy * 2.0 + dk4_dy;

  const DPose res = IntegratePoseD(pose, k, dt / 6.0, &dy_dk, &dy_dy);
  dy_db = dy_dk * dk_total_db;
  dy_dy0 = dy_dy + dy_dk * dk_total_dy;
  if (c_prior) {
    const Mat<10, 10> c_prop = dy_dy0 * (*c_prior) * Tr(dy_dy0);
    *c_prior = c_prop + dy_db * cov_meas * Tr(dy_db);
  }
  *dy_db_out = dy_db;
  *dy_dy0_out = dy_dy0;
  return res;
}

// general inverse by LU with partial pivoting (Eigen's .inverse() for 9x9)
template <int N>
inline bool Inverse(const Mat<N, N>& a, Mat<N, N>* out) {
  double M[N][2 * N];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) { M[i][j] = a(i, j); M[i][N + j] = (i == j) ? 1.0 : 0.0; }
  for (int c = 0; c < N; ++c) {
    int p = c;
    for (int r = c + 1; r < N; ++r) if (std::fabs(M[r][c]) > std::fabs(M[p][c])) p = r;
    if (M[p][c] == 0.0) return false;
    if (p != c) for (int j = 0; j < 2 * N; ++j) std::swap(M[p][j], M[c][j]);
    const double inv = 1.0 / M[c][c];
    for (int j = 0; j < 2 * N; ++j) M[c][j] *= inv;
    for (int r = 0; r < N; ++r) {
      if (r == c) continue;
      const double f = M[r][c];
      if (f == 0.0) continue;
      for (int j = 0; j < 2 * N; ++j) M[r][j] -= f * M[c][j];
    }
  }
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) (*out)(i, j) = M[i][N + j];
  return true;
}

// principal square root of a symmetric PSD matrix by Jacobi eigen-decomposition
// (Eigen MatrixFunctions `.sqrt()` at vicalibrator.h:796; equal to it for SPD input).
// Rotation order: round-robin tournament — round r holds the disjoint pairs {i, j}, i + j = r (mod N), i < j —
// with all rotations of a round computed from the same matrix and applied together (columns, then rows).
// Rotations on disjoint index pairs commute, so this is the classical cyclic method in another sweep order; it is
// the order the CUDA kernel runs four pairs at a time (vc_imu_weights.cuh).
template <int N>
inline Mat<N, N> SqrtSym(const Mat<N, N>& a_in) {
  static_assert(N % 2 == 1, "round-robin schedule written for odd N");
  Mat<N, N> a;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) a(i, j) = 0.5 * (a_in(i, j) + a_in(j, i));
  Mat<N, N> V = Mat<N, N>::Identity();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < N; ++i) {
      diag += a(i, i) * a(i, i);
      for (int j = i + 1; j < N; ++j) off += a(i, j) * a(i, j);
    }
    if (off <= 1e-30 * diag) break;
    for (int r = 0; r < N; ++r) {
      int P[N / 2], Q[N / 2], np = 0;
      double C[N / 2], S[N / 2];
      for (int i = 0; i < N; ++i) {
        const int j = ((r - i) % N + N) % N;
        if (i >= j) continue;
        const double apq = a(i, j);
        double c = 1.0, sn = 0.0;
        if (apq != 0.0) {
          const double tau = (a(j, j) - a(i, i)) / (2.0 * apq);
          const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
          c = 1.0 / std::sqrt(1.0 + t * t);
          sn = t * c;
        }
        P[np] = i; Q[np] = j; C[np] = c; S[np] = sn; ++np;
      }
      for (int t = 0; t < np; ++t)
        for (int k = 0; k < N; ++k) {
          const double akp = a(k, P[t]), akq = a(k, Q[t]);
          a(k, P[t]) = C[t] * akp - S[t] * akq;
          a(k, Q[t]) = S[t] * akp + C[t] * akq;
          const double vkp = V(k, P[t]), vkq = V(k, Q[t]);
          V(k, P[t]) = C[t] * vkp - S[t] * vkq;
          V(k, Q[t]) = S[t] * vkp + C[t] * vkq;
        }
      for (int t = 0; t < np; ++t)
        for (int k = 0; k < N; ++k) {
          const double apk = a(P[t], k), aqk = a(Q[t], k);
          a(P[t], k) = C[t] * apk - S[t] * aqk;
          a(Q[t], k) = S[t] * apk + C[t] * aqk;
        }
    }
  }
  Mat<N, N> r;
  for (int k = 0; k < N; ++k) {
    const double l = a(k, k) > 0 ? std::sqrt(a(k, k)) : 0.0;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) r(i, j) += V(i, k) * l * V(j, k);
  }
  return r;
}

// One iteration of the loop body at vicalibrator.h:726-797.  w_sqrt (9x9 row-major) is left
// untouched when the interval has no measurements (:731-733).
inline void UpdateOneImuWeight(const InterpolationBuffer& buf, double t_start, double t_end, double ts,
                               const double* T_w1, const double* v1, const double* T_w2, const double* v2,
                               const double b[6], const double sf[6], const double g2[2], double sigma_g,
                               double sigma_a, double* w_sqrt, double* mahalanobis = nullptr) {
  (void)v2;
  std::vector<ImuMeas<double>> meas;
  buf.GetRange(t_start, t_end, ts, &meas);
  if (meas.empty()) return;
  const SE3<double> t_2w = se3_inv(se3_from(T_w2));
  DPose pose{se3_from(T_w1), {v1[0], v1[1], v1[2]}};
  Mat<10, 10> c;
  Mat<6, 6> r;
  for (int i = 0; i < 3; ++i) { r(i, i) = powi(sigma_g, 2); r(3 + i, 3 + i) = powi(sigma_a, 2); }
  const Vec3<double> bg{b[0], b[1], b[2]}, ba{b[3], b[4], b[5]};
  const Vec3<double> gv = GetGravityVector<double>(g2, gravity());
  // types.h:611-687 (IntegrateResidual): chain of IntegrateImu steps
  const ImuMeas<double>* prev = nullptr;
  for (const ImuMeas<double>& m : meas) {
    if (prev) {
      Mat<10, 6> dy_db;
      Mat<10, 10> dy_dy;
      pose = IntegrateImuD(pose, *prev, m, bg, ba, sf, gv, &dy_db, &dy_dy, &c, r);
    }
    prev = &m;
  }
  const SE3<double> t12 = se3_mul(pose.t_wp, t_2w);
  const Mat<6, 7> dlog = dLog_dSE3(t12);
  const Mat<7, 7> dmul = dt1t2_dt1(pose.t_wp, t_2w);
  const Mat<6, 7> dse3 = dlog * dmul;
  Mat<9, 10> Jt;
  SetBlock(&Jt, 0, 0, dse3);
  SetBlock(&Jt, 6, 7, Mat<3, 3>::Identity());
  const Mat<9, 9> P = Jt * c * Tr(Jt);
  Mat<9, 9> info;
  if (!Inverse(P, &info)) return;
  if (mahalanobis) {
    double res[9];
    se3_log(t12, res);
    res[6] = pose.v_w.x - v2[0]; res[7] = pose.v_w.y - v2[1]; res[8] = pose.v_w.z - v2[2];
    double d = 0;
    for (int i = 0; i < 9; ++i)
      for (int j = 0; j < 9; ++j) d += res[i] * info(i, j) * res[j];
    *mahalanobis = d;
  }
  const Mat<9, 9> W = SqrtSym(info);
  std::memcpy(w_sqrt, W.m, sizeof W.m);
}

}  // namespace vo
#endif
