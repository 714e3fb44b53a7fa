// ORACLE — TEST INFRASTRUCTURE ONLY.
// Stand-in for the pre-1.0 Sophus API the reference uses (SO3Group<T> / SE3Group<T>, include/vicalib/vicalibrator.h:48),
// written from the published formulas on top of the Eigen stand-in: quaternion storage (x,y,z,w),
// SE3 storage [so3 (4) | translation (3)], tangent order (upsilon, omega), exp / log with the small-angle Taylor
// branches at SophusConstants::epsilon(), group products that renormalise the quaternion.
#pragma once
#include <Eigen/Core>
#include <cmath>

namespace Sophus {

template <class S> struct SophusConstants {
  static S epsilon() { return S(1e-10); }
  static S pi() { return S(3.141592653589793238462643383279502884); }
};

template <class S>
class SO3Group {
 public:
  typedef Eigen::Matrix<S, 3, 1> Tangent;
  typedef Eigen::Matrix<S, 3, 1> Point;
  typedef Eigen::Matrix<S, 3, 3> Transformation;
  static const int num_parameters = 4;
  static const int DoF = 3;
  SO3Group() {}
  explicit SO3Group(const Eigen::Quaternion<S>& q) : q_(q) { q_.normalize(); }
  template <class O> SO3Group(const SO3Group<O>& o) : q_(o.unit_quaternion().template cast<S>()) {}  // NOLINT
  const Eigen::Quaternion<S>& unit_quaternion() const { return q_; }
  Eigen::Quaternion<S>& unit_quaternion_nonconst() { return q_; }
  S* data() { return q_.coeffs().data(); }
  const S* data() const { return q_.coeffs().data(); }
  Transformation matrix() const { return q_.toRotationMatrix(); }
  Transformation Adj() const { return matrix(); }
  SO3Group inverse() const {
    SO3Group r;
    r.q_ = q_.conjugate();
    return r;
  }
  SO3Group operator*(const SO3Group& o) const {
    SO3Group r;
    r.q_ = q_ * o.q_;
    r.q_.normalize();
    return r;
  }
  SO3Group& operator*=(const SO3Group& o) {
    *this = *this * o;
    return *this;
  }
  Point operator*(const Point& p) const { return q_._transformVector(p); }
  template <class D, int R, int C> Point operator*(const Eigen::Ops<D, S, R, C>& p) const { return q_._transformVector(Point(p.self())); }
  void setQuaternion(const Eigen::Quaternion<S>& q) {
    q_ = q;
    q_.normalize();
  }
  static SO3Group exp(const Tangent& omega) {
    S theta;
    return expAndTheta(omega, &theta);
  }
  static SO3Group expAndTheta(const Tangent& omega, S* theta) {
    using std::cos;
    using std::sin;
    using std::sqrt;
    const S theta_sq = omega.squaredNorm();
    *theta = sqrt(theta_sq);
    const S half_theta = S(0.5) * (*theta);
    S imag_factor, real_factor;
    if ((*theta) < SophusConstants<S>::epsilon()) {
      const S theta_po4 = theta_sq * theta_sq;
      imag_factor = S(0.5) - S(1.0 / 48.0) * theta_sq + S(1.0 / 3840.0) * theta_po4;
      real_factor = S(1.0) - S(1.0 / 8.0) * theta_sq + S(1.0 / 384.0) * theta_po4;
    } else {
      const S sin_half_theta = sin(half_theta);
      imag_factor = sin_half_theta / (*theta);
      real_factor = cos(half_theta);
    }
    SO3Group r;
    r.q_ = Eigen::Quaternion<S>(real_factor, imag_factor * omega[0], imag_factor * omega[1], imag_factor * omega[2]);
    return r;
  }
  Tangent log() const { return SO3Group::log(*this); }
  static Tangent log(const SO3Group& o) {
    S theta;
    return logAndTheta(o, &theta);
  }
  static Tangent logAndTheta(const SO3Group& o, S* theta) {
    using std::abs;
    using std::atan;
    using std::sqrt;
    const S squared_n = o.q_.vec().squaredNorm();
    const S n = sqrt(squared_n);
    const S w = o.q_.w();
    S two_atan_nbyw_by_n;
    if (n < SophusConstants<S>::epsilon()) {
      const S squared_w = w * w;
      two_atan_nbyw_by_n = S(2.0) / w - S(2.0) * squared_n / (w * squared_w);
    } else {
      if (abs(w) < SophusConstants<S>::epsilon()) {
        if (w > S(0.0)) two_atan_nbyw_by_n = SophusConstants<S>::pi() / n;
        else two_atan_nbyw_by_n = -SophusConstants<S>::pi() / n;
      } else {
        two_atan_nbyw_by_n = S(2.0) * atan(n / w) / n;
      }
    }
    *theta = two_atan_nbyw_by_n * n;
    return o.q_.vec() * two_atan_nbyw_by_n;
  }
  static Transformation hat(const Tangent& v) {
    Transformation m;
    m(0, 0) = S(0.0); m(0, 1) = -v[2]; m(0, 2) = v[1];
    m(1, 0) = v[2]; m(1, 1) = S(0.0); m(1, 2) = -v[0];
    m(2, 0) = -v[1]; m(2, 1) = v[0]; m(2, 2) = S(0.0);
    return m;
  }

 protected:
  Eigen::Quaternion<S> q_;
};

template <class S>
class SE3Group {
 public:
  typedef Eigen::Matrix<S, 6, 1> Tangent;
  typedef Eigen::Matrix<S, 3, 1> Point;
  static const int num_parameters = 7;
  static const int DoF = 6;
  SE3Group() {}
  SE3Group(const SO3Group<S>& so3, const Point& t) : so3_(so3), t_(t) {}
  template <class D, int R, int C> SE3Group(const SO3Group<S>& so3, const Eigen::Ops<D, S, R, C>& t) : so3_(so3), t_(t.self()) {}
  template <class O> SE3Group(const SE3Group<O>& o) : so3_(o.so3()), t_(o.translation().template cast<S>()) {}  // NOLINT
  SO3Group<S>& so3() { return so3_; }
  const SO3Group<S>& so3() const { return so3_; }
  Point& translation() { return t_; }
  const Point& translation() const { return t_; }
  const Eigen::Quaternion<S>& unit_quaternion() const { return so3_.unit_quaternion(); }
  Eigen::Matrix<S, 3, 3> rotationMatrix() const { return so3_.matrix(); }
  S* data() { return so3_.data(); }  // layout [q (4) | t (3)]: members are laid out back to back (checked in the harness)
  const S* data() const { return so3_.data(); }
  SE3Group inverse() const {
    const SO3Group<S> inv = so3_.inverse();
    return SE3Group(inv, inv * (t_ * S(-1.0)));
  }
  SE3Group operator*(const SE3Group& o) const { return SE3Group(so3_ * o.so3_, t_ + so3_ * o.t_); }
  Point operator*(const Point& p) const { return so3_ * p + t_; }
  template <class D, int R, int C> Point operator*(const Eigen::Ops<D, S, R, C>& p) const { return so3_ * Point(p.self()) + t_; }
  Tangent log() const { return SE3Group::log(*this); }
  static Tangent log(const SE3Group& se3) {
    using std::abs;
    using std::tan;
    Tangent upsilon_omega;
    S theta;
    const Eigen::Matrix<S, 3, 1> omega = SO3Group<S>::logAndTheta(se3.so3_, &theta);
    upsilon_omega.template tail<3>() = omega;
    if (abs(theta) < SophusConstants<S>::epsilon()) {
      const Eigen::Matrix<S, 3, 3> Omega = SO3Group<S>::hat(omega);
      const Eigen::Matrix<S, 3, 3> V_inv = Eigen::Matrix<S, 3, 3>::Identity() - S(0.5) * Omega + S(1. / 12.) * (Omega * Omega);
      upsilon_omega.template head<3>() = V_inv * se3.t_;
    } else {
      const Eigen::Matrix<S, 3, 3> Omega = SO3Group<S>::hat(omega);
      const S half_theta = S(0.5) * theta;
      const Eigen::Matrix<S, 3, 3> V_inv =
          (Eigen::Matrix<S, 3, 3>::Identity() - S(0.5) * Omega +
           (S(1.0) - theta / (S(2.0) * tan(half_theta))) / (theta * theta) * (Omega * Omega));
      upsilon_omega.template head<3>() = V_inv * se3.t_;
    }
    return upsilon_omega;
  }
  static SE3Group exp(const Tangent& a) {
    using std::cos;
    using std::sin;
    const Eigen::Matrix<S, 3, 1> omega = a.template tail<3>();
    S theta;
    const SO3Group<S> so3 = SO3Group<S>::expAndTheta(omega, &theta);
    const Eigen::Matrix<S, 3, 3> Omega = SO3Group<S>::hat(omega);
    const Eigen::Matrix<S, 3, 3> Omega_sq = Omega * Omega;
    Eigen::Matrix<S, 3, 3> V;
    if (theta < SophusConstants<S>::epsilon()) {
      V = so3.matrix();
    } else {
      const S theta_sq = theta * theta;
      V = (Eigen::Matrix<S, 3, 3>::Identity() + (S(1.0) - cos(theta)) / (theta_sq) * Omega +
           (theta - sin(theta)) / (theta_sq * theta) * Omega_sq);
    }
    return SE3Group(so3, V * a.template head<3>());
  }

 protected:
  SO3Group<S> so3_;
  Point t_;
};

typedef SO3Group<double> SO3d;
typedef SE3Group<double> SE3d;

}  // namespace Sophus

namespace Eigen {
// Sophus specialises Eigen::Map for its groups; the reference maps Ceres parameter blocks through these
template <class S>
class Map<const Sophus::SO3Group<S> > : public Sophus::SO3Group<S> {
 public:
  explicit Map(const S* p) { this->q_ = Quaternion<S>(p[3], p[0], p[1], p[2]); }  // raw coefficients, no renormalisation
};
template <class S>
class Map<const Sophus::SE3Group<S> > : public Sophus::SE3Group<S> {
 public:
  explicit Map(const S* p) {
    this->so3_ = Map<const Sophus::SO3Group<S> >(p);
    this->t_ = Matrix<S, 3, 1>(p[4], p[5], p[6]);
  }
};
template <class S>
class Map<Sophus::SE3Group<S> > {
 public:
  explicit Map(S* p) : p_(p) {}
  Map& operator=(const Sophus::SE3Group<S>& T) {
    for (int i = 0; i < 4; ++i) p_[i] = T.unit_quaternion().coeffs()[i];
    for (int i = 0; i < 3; ++i) p_[4 + i] = T.translation()[i];
    return *this;
  }

 private:
  S* p_;
};
template <class S>
class Map<Sophus::SO3Group<S> > {
 public:
  explicit Map(S* p) : p_(p) {}
  Map& operator=(const Sophus::SO3Group<S>& R) {
    for (int i = 0; i < 4; ++i) p_[i] = R.unit_quaternion().coeffs()[i];
    return *this;
  }

 private:
  S* p_;
};
}  // namespace Eigen
