// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
//
// Camera-model `Project` bodies.  The reference calls `CameraInt::Project(ray, params, pix)`
// (ceres-cost-functions.h:369) with Calibu's CRTP camera classes instantiated at
// vicalibrator.h:412-453; Calibu (calibu/cam/camera_models_crtp.h) is not vendored and
// not pinned, so the published algorithm is restated here.
// Parameter order fu, fv, u0, v0, <distortion...>   (vicalib-engine.cc:210,220,230,250).
#ifndef VICALIB_ORACLE_CAMERA_MODELS_H_
#define VICALIB_ORACLE_CAMERA_MODELS_H_

#include "dual.h"

namespace vo {

enum CameraModel { kLinear = 0, kFov = 1, kPoly2 = 2, kPoly3 = 3, kKb4 = 4 };

inline int NumIntrinsics(int model) {
  switch (model) {
    case kLinear: return 4;
    case kFov: return 5;
    case kPoly2: return 6;
    case kPoly3: return 7;
    case kKb4: return 8;
  }
  return -1;
}

struct LinearCam {
  static constexpr int K = 4;
  template <class T>
  static void Project(const T* ray, const T* p, T* pix) {
    pix[0] = ray[0] / ray[2];
    pix[1] = ray[1] / ray[2];
    pix[0] = pix[0] * p[0] + p[2];
    pix[1] = pix[1] * p[1] + p[3];
  }
};

struct FovCam {
  static constexpr int K = 5;
  template <class T>
  static T Factor(const T& rad, const T* p) {
    const T w = p[4];
    if (w * w > 1e-5) {
      const T mul2_tanw_by2 = T(2.0) * vo::tan(w / T(2.0));
      const T mul2_tanw_by2_byw = mul2_tanw_by2 / w;
      if (rad * rad < 1e-5) {
        return mul2_tanw_by2_byw;
      } else {
        return vo::atan(rad * mul2_tanw_by2) / (rad * w);
      }
    }
    return T(1.0);
  }
  template <class T>
  static void Project(const T* ray, const T* p, T* pix) {
    pix[0] = ray[0] / ray[2];
    pix[1] = ray[1] / ray[2];
    const T fac = Factor(vo::sqrt(pix[0] * pix[0] + pix[1] * pix[1]), p);
    pix[0] = pix[0] * fac;
    pix[1] = pix[1] * fac;
    pix[0] = pix[0] * p[0] + p[2];
    pix[1] = pix[1] * p[1] + p[3];
  }
};

template <int ORDER>  // 2 -> poly2 (k1,k2), 3 -> poly3 (k1,k2,k3)
struct PolyCam {
  static constexpr int K = 4 + ORDER;
  template <class T>
  static T Factor(const T& rad, const T* p) {
    const T r2 = rad * rad;
    const T r4 = r2 * r2;
    T f = T(1.0) + p[4] * r2 + p[5] * r4;
    if (ORDER == 3) f = f + p[6] * r4 * r2;
    return f;
  }
  template <class T>
  static void Project(const T* ray, const T* p, T* pix) {
    pix[0] = ray[0] / ray[2];
    pix[1] = ray[1] / ray[2];
    const T fac = Factor(vo::sqrt(pix[0] * pix[0] + pix[1] * pix[1]), p);
    pix[0] = pix[0] * fac;
    pix[1] = pix[1] * fac;
    pix[0] = pix[0] * p[0] + p[2];
    pix[1] = pix[1] * p[1] + p[3];
  }
};
using Poly2Cam = PolyCam<2>;
using Poly3Cam = PolyCam<3>;

struct Kb4Cam {
  static constexpr int K = 8;
  template <class T>
  static void Project(const T* ray, const T* p, T* pix) {
    const T xsq_ysq = ray[0] * ray[0] + ray[1] * ray[1];
    const T theta = vo::atan2(vo::sqrt(xsq_ysq), ray[2]);
    const T psi = vo::atan2(ray[1], ray[0]);
    const T th2 = theta * theta;
    const T th3 = th2 * theta;
    const T th5 = th3 * th2;
    const T th7 = th5 * th2;
    const T th9 = th7 * th2;
    const T r = theta + p[4] * th3 + p[5] * th5 + p[6] * th7 + p[7] * th9;
    pix[0] = p[0] * r * vo::cos(psi) + p[2];
    pix[1] = p[1] * r * vo::sin(psi) + p[3];
  }
};

}  // namespace vo
#endif
