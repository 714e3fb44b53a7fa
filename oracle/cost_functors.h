// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
//
// CPU restatement of the reference's two residual families and their support code:
//   * ImuReprojectionCostFunctor          ceres-cost-functions.h:342-377
//   * SwitchedFullImuCostFunction         ceres-cost-functions.h:379-490
//   * IntegratePoseJet/GetPoseDerivativeJet/IntegrateImuJet/IntegrateResidualJet
//                                         ceres-cost-functions.h:38-56, 79-105, 138-177, 199-227
//   * InterpolationBufferT                interpolation-buffer.h:50-227
//   * LocalParamSe3 / LocalParamSo3       local-param-se3.h:28-91, 121-157
//   * GetGravityVector / gravity()        types.h:40-42, 93-104
// All templated on the scalar so they run on double and on Dual<N> ("Jet").
#ifndef VICALIB_ORACLE_COST_FUNCTORS_H_
#define VICALIB_ORACLE_COST_FUNCTORS_H_

#include <algorithm>
#include <cstddef>
#include <limits>
#include <vector>

#include "camera_models.h"
#include "lie.h"

namespace vo {

inline double gravity() { return 9.8007; }  // types.h:40-42

// types.h:93-104
template <class T>
inline Vec3<T> GetGravityVector(const T* dir, const T& g) {
  const T sp = vo::sin(dir[0]), cp = vo::cos(dir[0]);
  const T sq = vo::sin(dir[1]), cq = vo::cos(dir[1]);
  const T mg = -g;
  return {cp * sq * mg, -sp * mg, cp * cq * mg};
}

// ---------------------------------------------------------------- reprojection
// ceres-cost-functions.h:350-373
template <class Cam, class T>
inline void ReprojectionResidual(const T* t_wk, const T* r_ck, const T* t_ck, const T* cam_params,
                                 const double* p_w, const double* p_c, T* residuals) {
  const SE3<T> T_wk = se3_from(t_wk);
  const SE3<T> T_kw = se3_inv(T_wk);
  const SE3<T> T_ck{{r_ck[0], r_ck[1], r_ck[2], r_ck[3]}, {t_ck[0], t_ck[1], t_ck[2]}};
  const Vec3<T> pw{T(p_w[0]), T(p_w[1]), T(p_w[2])};
  const Vec3<T> p_cv = se3_act(T_ck, se3_act(T_kw, pw));
  const T ray[3] = {p_cv.x, p_cv.y, p_cv.z};
  T z[2];
  Cam::Project(ray, cam_params, z);
  residuals[0] = z[0] - T(p_c[0]);
  residuals[1] = z[1] - T(p_c[1]);
}

// ---------------------------------------------------------------- local params
// local-param-se3.h:28-91: 7x6 row-major, global = [q(4) t(3)], local = [dt(3) w(3)]
inline void LocalParamSe3Jacobian(const double* x, double* J /*7x6*/) {
  for (int i = 0; i < 42; ++i) J[i] = 0.0;
  const double q1 = x[0], q2 = x[1], q3 = x[2], q0 = x[3];
  const double h0 = 0.5 * q0, h1 = 0.5 * q1, h2 = 0.5 * q2, h3 = 0.5 * q3;
  J[3] = h0;  J[4] = -h3; J[5] = h2;
  J[9] = h3;  J[10] = h0; J[11] = -h1;
  J[15] = -h2; J[16] = h1; J[17] = h0;
  J[21] = -h1; J[22] = -h2; J[23] = -h3;
  J[24] = 1.0 - 2.0 * (q2 * q2 + q3 * q3);
  J[25] = 2.0 * (q1 * q2 - q0 * q3);
  J[26] = 2.0 * (q1 * q3 + q0 * q2);
  J[30] = 2.0 * (q1 * q2 + q0 * q3);
  J[31] = 1.0 - 2.0 * (q1 * q1 + q3 * q3);
  J[32] = 2.0 * (q2 * q3 - q0 * q1);
  J[36] = 2.0 * (q1 * q3 - q0 * q2);
  J[37] = 2.0 * (q2 * q3 + q0 * q1);
  J[38] = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
}
// local-param-se3.h:121-157: 4x3 row-major
inline void LocalParamSo3Jacobian(const double* x, double* J /*4x3*/) {
  const double q1 = x[0], q2 = x[1], q3 = x[2], q0 = x[3];
  const double h0 = 0.5 * q0, h1 = 0.5 * q1, h2 = 0.5 * q2, h3 = 0.5 * q3;
  J[0] = h0;  J[1] = -h3; J[2] = h2;
  J[3] = h3;  J[4] = h0;  J[5] = -h1;
  J[6] = -h2; J[7] = h1;  J[8] = h0;
  J[9] = -h1; J[10] = -h2; J[11] = -h3;
}
// local-param-se3.h:14-26 / 107-119: Plus = T * exp(delta)
inline void LocalParamSe3Plus(const double* x, const double* delta, double* out) {
  const SE3<double> T = se3_from(x);
  se3_to(se3_mul(T, se3_exp(delta)), out);
  // Sophus group multiplication renormalises the quaternion
  const Quat<double> qn = qnormalized(Quat<double>{out[0], out[1], out[2], out[3]});
  out[0] = qn.x; out[1] = qn.y; out[2] = qn.z; out[3] = qn.w;
}
inline void LocalParamSo3Plus(const double* x, const double* delta, double* out) {
  const Quat<double> q{x[0], x[1], x[2], x[3]};
  const Vec3<double> w{delta[0], delta[1], delta[2]};
  const Quat<double> r = qnormalized(qmul(q, so3_exp(w)));
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// ---------------------------------------------------------------- IMU types
template <class T>
struct ImuMeas {  // types.h:210-253
  Vec3<T> w, a;
  T time;
};
template <class T>
struct ImuPose {  // types.h:175-207 (t_wp_, v_w_, w_w_, time_)
  SE3<T> t_wp;
  Vec3<T> v_w;
  Vec3<T> w_w;
  T time;
};

// interpolation-buffer.h:50-227 (double storage; T-typed time offset)
struct InterpolationBuffer {
  std::vector<ImuMeas<double>> elements_;
  double start_time_ = -1, end_time_ = -1, average_dt_ = 0;

  void AddElement(const ImuMeas<double>& e) {  // :70-85
    const size_t n = elements_.size();
    double dt = 0;
    if (n > 0) dt = e.time - elements_.back().time;
    average_dt_ = (average_dt_ * n + dt) / (n + 1);
    elements_.push_back(e);
    end_time_ = e.time;
    start_time_ = elements_.front().time;
  }
  template <class T>
  static ImuMeas<T> Cast(const ImuMeas<double>& m) {
    return {{T(m.w.x), T(m.w.y), T(m.w.z)}, {T(m.a.x), T(m.a.y), T(m.a.z)}, T(m.time)};
  }
  template <class T>
  bool HasElement(double time, const T& dt) const {  // :122-125
    return time >= start_time_ + scalar_of(dt) && time <= end_time_ + scalar_of(dt);
  }
  template <class T>
  void InterpolateElements(size_t ai, size_t bi, const T& time_offset, double time,
                           ImuMeas<T>* out) const {  // :136-155
    const ImuMeas<T> a = Cast<T>(elements_[ai]), b = Cast<T>(elements_[bi]);
    const T t_a = T(elements_[ai].time) + time_offset;
    const T t_b = T(elements_[bi].time) + time_offset;
    const T t_out = T(time);
    const T fraction = (t_out - t_a) / (t_b - t_a);
    const T omf = T(1.0) - fraction;
    out->w = a.w * omf + b.w * fraction;
    out->a = a.a * omf + b.a * fraction;
    out->time = t_out;
  }
  template <class T>
  ImuMeas<T> GetElement(double time, const T& dt, size_t* pIndex) const {  // :160-203
    const double off = scalar_of(dt);
    const size_t n = elements_.size();
    // size_t conversion of a possibly negative double is what the reference does (:166);
    // clamp the scalar first so the conversion is defined.
    double guess_d = (time - start_time_ + off) / average_dt_;
    if (!(guess_d > 0.0)) guess_d = 0.0;
    size_t guess = guess_d >= static_cast<double>(n) ? n - 1 : static_cast<size_t>(guess_d);
    guess = std::min(guess, n - 1);
    ImuMeas<T> result;
    if (elements_[guess].time + off > time) {
      if (guess == 0) {
        result = Cast<T>(elements_.front());
        result.time = result.time + dt;
        *pIndex = guess;
      } else {
        while ((guess - 1) > 0 && elements_[guess - 1].time + off > time) --guess;
        InterpolateElements(guess - 1, guess, dt, time, &result);
        *pIndex = guess - 1;
      }
    } else {
      if (guess == n - 1) {
        *pIndex = guess;
        result = Cast<T>(elements_.back());
        result.time = result.time + dt;
      } else {
        while ((guess + 1) < n && (elements_[guess + 1].time + off) < time) ++guess;
        InterpolateElements(guess, guess + 1, dt, time, &result);
        *pIndex = guess;
      }
    }
    return result;
  }
  template <class T>
  bool GetNext(double max_time, const T& dt, size_t* index, ImuMeas<T>* out) const {  // :100-117
    if (*index + 1 >= elements_.size()) {
      *out = GetElement(max_time, dt, index);
      return false;
    } else if (T(elements_[*index + 1].time) + dt > T(max_time)) {
      *out = GetElement(max_time, dt, index);
      return false;
    } else {
      *out = Cast<T>(elements_[++*index]);
      out->time = out->time + dt;
      return true;
    }
  }
  template <class T>
  void GetRange(double start_time, double end_time, const T& dt,
                std::vector<ImuMeas<T>>* out) const {  // :208-226
    size_t index;
    if (HasElement(start_time, dt)) {
      out->push_back(GetElement(start_time, dt, &index));
      ImuMeas<T> meas;
      while (GetNext(end_time, dt, &index, &meas)) out->push_back(meas);
      out->push_back(meas);
    }
  }
};

// ---------------------------------------------------------------- Jet integrator
// ceres-cost-functions.h:38-56
template <class T>
inline ImuPose<T> IntegratePoseJet(const ImuPose<T>& pose, const T k[9], const T& dt) {
  const Vec3<T> wdt{k[3] * dt, k[4] * dt, k[5] * dt};
  const Quat<T> rv2_v1 = so3_exp(wdt);
  ImuPose<T> y = pose;
  y.t_wp.t = pose.t_wp.t + Vec3<T>{k[0] * dt, k[1] * dt, k[2] * dt};
  y.t_wp.q = qmul(rv2_v1, pose.t_wp.q);  // raw product, no renormalisation (:47-51)
  y.v_w = pose.v_w + Vec3<T>{k[6] * dt, k[7] * dt, k[8] * dt};
  return y;
}
// ceres-cost-functions.h:79-105
template <class T>
inline void GetPoseDerivativeJet(const ImuPose<T>& pose, const Vec3<T>& t_w, const ImuMeas<T>& z0,
                                 const ImuMeas<T>& z1, const Vec3<T>& bg, const Vec3<T>& ba,
                                 const T sf[6], const T& dt, T deriv[9]) {
  const T alpha = (z1.time - (z0.time + dt)) / (z1.time - z0.time);
  const T oma = T(1.0) - alpha;
  const Vec3<T> zg = z0.w * alpha + z1.w * oma;
  const Vec3<T> za = z0.a * alpha + z1.a * oma;
  deriv[0] = pose.v_w.x; deriv[1] = pose.v_w.y; deriv[2] = pose.v_w.z;
  T R[9];
  qmat(pose.t_wp.q, R);  // so3().Adj()
  const Vec3<T> wb{zg.x * sf[0] + bg.x, zg.y * sf[1] + bg.y, zg.z * sf[2] + bg.z};
  const Vec3<T> w = mat_mul(R, wb);
  deriv[3] = w.x; deriv[4] = w.y; deriv[5] = w.z;
  const Vec3<T> ab{za.x * sf[3] + ba.x, za.y * sf[4] + ba.y, za.z * sf[5] + ba.z};
  const Vec3<T> a = qrot(pose.t_wp.q, ab) - t_w;  // so3() * v
  deriv[6] = a.x; deriv[7] = a.y; deriv[8] = a.z;
}
// ceres-cost-functions.h:138-177
template <class T>
inline ImuPose<T> IntegrateImuJet(const ImuPose<T>& pose, const ImuMeas<T>& z0, const ImuMeas<T>& z1,
                                  const Vec3<T>& bg, const Vec3<T>& ba, const T sf[6],
                                  const Vec3<T>& g) {
  if (z1.time == z0.time) return pose;
  const T dt = z1.time - z0.time;
  T k1[9], k2[9], k3[9], k4[9], k[9];
  GetPoseDerivativeJet(pose, g, z0, z1, bg, ba, sf, T(0.0), k1);
  const ImuPose<T> y1 = IntegratePoseJet(pose, k1, dt * T(0.5));
  GetPoseDerivativeJet(y1, g, z0, z1, bg, ba, sf, dt / T(2.0), k2);
  const ImuPose<T> y2 = IntegratePoseJet(pose, k2, dt * T(0.5));
  GetPoseDerivativeJet(y2, g, z0, z1, bg, ba, sf, dt / T(2.0), k3);
  const ImuPose<T> y3 = IntegratePoseJet(pose, k3, dt);
  GetPoseDerivativeJet(y3, g, z0, z1, bg, ba, sf, dt, k4);
  for (int i = 0; i < 9; ++i) k[i] = k1[i] + T(2.0) * k2[i] + T(2.0) * k3[i] + k4[i];
  ImuPose<T> res = IntegratePoseJet(pose, k, dt / T(6.0));
  res.w_w = {k[3], k[4], k[5]};
  res.time = z1.time;
  return res;
}
// ceres-cost-functions.h:199-227
template <class T>
inline ImuPose<T> IntegrateResidualJet(const SE3<T>& t_wp, const Vec3<T>& v_w, const T& time,
                                       const std::vector<ImuMeas<T>>& meas, const Vec3<T>& bg,
                                       const Vec3<T>& ba, const T sf[6], const Vec3<T>& g) {
  ImuPose<T> pose{t_wp, v_w, {T(0.0), T(0.0), T(0.0)}, time};
  const ImuMeas<T>* prev = nullptr;
  for (const ImuMeas<T>& m : meas) {
    if (prev != nullptr) pose = IntegrateImuJet(pose, *prev, m, bg, ba, sf, g);
    prev = &m;
  }
  return pose;
}

// ceres-cost-functions.h:402-484.  w_sqrt is 9x9 row-major; residuals = (r^T W)^T.
template <class T>
inline void ImuResidual(const InterpolationBuffer& buf, double start_time, double end_time,
                        const double* w_sqrt, bool rotation_only_switch, const T* tx2, const T* tx1,
                        const T* tvx2, const T* tvx1, const T* tg, const T* tb, const T* tsf,
                        const T* ttime_offset, T* residuals) {
  const SE3<T> t_wx2 = se3_from(tx2);
  const SE3<T> t_wx1 = se3_from(tx1);
  const Vec3<T> v1{tvx1[0], tvx1[1], tvx1[2]}, v2{tvx2[0], tvx2[1], tvx2[2]};
  const Vec3<T> bg{tb[0], tb[1], tb[2]}, ba{tb[3], tb[4], tb[5]};
  const Vec3<T> g_vector = GetGravityVector<T>(tg, T(gravity()));
  std::vector<ImuMeas<T>> measurements;
  buf.GetRange(start_time, end_time, *ttime_offset, &measurements);
  if (measurements.empty()) {
    for (int i = 0; i < 9; ++i) residuals[i] = T(0.0);
    return;
  }
  const ImuPose<T> end_pose = IntegrateResidualJet(t_wx1, v1, measurements.front().time,
                                                   measurements, bg, ba, tsf, g_vector);
  T r[9];
  se3_log(se3_mul(end_pose.t_wp, se3_inv(t_wx2)), r);
  r[6] = end_pose.v_w.x - v2.x;
  r[7] = end_pose.v_w.y - v2.y;
  r[8] = end_pose.v_w.z - v2.z;
  for (int j = 0; j < 9; ++j) {
    T s = T(0.0);
    for (int i = 0; i < 9; ++i) s = s + r[i] * T(w_sqrt[i * 9 + j]);
    residuals[j] = s;
  }
  if (rotation_only_switch) {
    for (int i = 0; i < 3; ++i) residuals[i] = T(0.0);
    for (int i = 6; i < 9; ++i) residuals[i] = T(0.0);
  }
}

// ---------------------------------------------------------------- loss functions
// Ceres SoftLOneLoss(a) / CauchyLoss(a) (vicalibrator.h:127,133); rho[0..2] = rho, rho', rho''
inline void SoftLOne(double a, double s, double rho[3]) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  const double tmp = std::sqrt(sum);
  rho[0] = 2.0 * b * (tmp - 1.0);
  rho[1] = std::max(std::numeric_limits<double>::min(), 1.0 / tmp);
  rho[2] = -(c * rho[1]) / (2.0 * sum);
}
inline void Cauchy(double a, double s, double rho[3]) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  const double inv = 1.0 / sum;
  rho[0] = b * std::log(sum);
  rho[1] = std::max(std::numeric_limits<double>::min(), inv);
  rho[2] = -c * (inv * inv);
}

}  // namespace vo
#endif
