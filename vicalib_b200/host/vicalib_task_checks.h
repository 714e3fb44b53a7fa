// Post-calibration success checks of VicalibTask (src/vicalib-task.cc:714-856) on top of the ViCalibrator mirror:
// the drop-in reproduces the reference's pass / fail decision, not only its numbers.
//
//   CameraCalibrationsDiffer   vicalib-task.cc:714-813
//   IMUCalibrationDiffer       vicalib-task.cc:815-835
//   IsSuccessful               vicalib-task.cc:837-863 (VicalibTask::IsSuccessful)
//
// Restated AS WRITTEN, including two quirks of the reference (kept so that the verdicts agree; both are flagged
// in INTEGRATION.md):
//   * the distortion checks compare camera->Type() with "FovCamera" / "Poly3Camera" (:748,:753); Calibu's Type()
//     returns the "calibu_fu_fv_u0_v0_..." strings, so those branches never fire;
//   * IMUCalibrationDiffer reports a difference when |last - current| is BELOW the threshold (:818-833).
#ifndef VICALIB_B200_HOST_VICALIB_TASK_CHECKS_H_
#define VICALIB_B200_HOST_VICALIB_TASK_CHECKS_H_

#include <cmath>
#include <cstdio>
#include <vector>

#include "vicalibrator.h"

namespace visual_inertial_calibration {

// gflags defaults of vicalib-task.cc:26-48 and vicalib-engine.cc:56
struct SuccessThresholds {
  double max_fx_diff = 10.0, max_fy_diff = 10.0, max_cx_diff = 10.0, max_cy_diff = 10.0;
  double max_fov_w_diff = 0.3;
  double max_poly3_diff_k1 = 0.1, max_poly3_diff_k2 = 0.1, max_poly3_diff_k3 = 0.1;
  double max_camera_trans_diff = 0.1;   // metres
  double max_camera_angle_diff = 0.1;   // radians
  double max_imu_gyro_diff = 0.1, max_imu_accel_diff = 0.1;
  double max_reprojection_error = 0.15; // pixels, per camera stream
};

// true when the two calibrations differ by more than the thresholds (vicalib-task.cc:714-813)
inline bool CameraCalibrationsDiffer(const CameraAndPose& last, const CameraAndPose& current, const SuccessThresholds& t) {
  if (last.camera->Type() != current.camera->Type()) return true;  // :717-722
  const std::vector<double>& lp = last.camera->GetParams();
  const std::vector<double>& cp = current.camera->GetParams();
  if (std::fabs(lp[0] - cp[0]) > t.max_fx_diff) return true;       // :729-744
  else if (std::fabs(lp[1] - cp[1]) > t.max_fy_diff) return true;
  else if (std::fabs(lp[2] - cp[2]) > t.max_cx_diff) return true;
  else if (std::fabs(lp[3] - cp[3]) > t.max_cy_diff) return true;
  if (current.camera->Type() == "FovCamera" && std::fabs(lp[4] - cp[4]) > t.max_fov_w_diff) {  // :748-752
    return true;
  } else if (current.camera->Type() == "Poly3Camera") {            // :753-764
    if (std::fabs(lp[4] - cp[4]) > t.max_poly3_diff_k1 || std::fabs(lp[5] - cp[5]) > t.max_poly3_diff_k2 ||
        std::fabs(lp[6] - cp[6]) > t.max_poly3_diff_k3)
      return true;
  }
  // extrinsics: distance between the camera positions (:770-780)
  double d2 = 0;
  for (int i = 0; i < 3; ++i) d2 += (last.T_ck.d[4 + i] - current.T_ck.d[4 + i]) * (last.T_ck.d[4 + i] - current.T_ck.d[4 + i]);
  if (std::sqrt(d2) > t.max_camera_trans_diff) return true;
  // orientation: x / y / z angles of last.R^-1 * current.R, computed in float like the reference (:786-803)
  double Rl[9], Rc[9], D[9];
  last.T_ck.rotation(Rl);
  current.T_ck.rotation(Rc);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) D[i * 3 + j] = Rl[0 * 3 + i] * Rc[0 * 3 + j] + Rl[1 * 3 + i] * Rc[1 * 3 + j] + Rl[2 * 3 + i] * Rc[2 * 3 + j];
  const float angle_x = static_cast<float>(std::atan2(D[7], D[8]));
  const float root = static_cast<float>(std::sqrt(D[7] * D[7] + D[8] * D[8]));
  const float angle_y = static_cast<float>(std::atan2(-1 * D[6], root));
  const float angle_z = static_cast<float>(std::atan2(D[3], D[0]));
  if (std::fabs(angle_x) > t.max_camera_angle_diff || std::fabs(angle_y) > t.max_camera_angle_diff ||
      std::fabs(angle_z) > t.max_camera_angle_diff)
    return true;
  return false;
}

// vicalib-task.cc:815-835, comparison direction as written
inline bool IMUCalibrationDiffer(const Vector6d& last, const Vector6d& current, const SuccessThresholds& t) {
  double diff[6];
  for (int i = 0; i < 6; ++i) diff[i] = last[i] - current[i];
  if (std::fabs(diff[0]) < t.max_imu_gyro_diff || std::fabs(diff[1]) < t.max_imu_gyro_diff || std::fabs(diff[2]) < t.max_imu_gyro_diff)
    return true;
  if (std::fabs(diff[3]) < t.max_imu_accel_diff || std::fabs(diff[4]) < t.max_imu_accel_diff || std::fabs(diff[5]) < t.max_imu_accel_diff)
    return true;
  return false;
}

// VicalibTask::IsSuccessful (vicalib-task.cc:837-863): every stream's reprojection rmse under its maximum and, with
// -has_initial_guess, the new calibration close to the one the run was initialised with.
inline bool IsSuccessful(ViCalibrator& calibrator, const std::vector<double>& max_reproj_errors, bool has_initial_guess,
                         const std::vector<CameraAndPose>& input_cameras, const Vector6d& input_imu_biases,
                         const SuccessThresholds& t = SuccessThresholds()) {
  const std::vector<double> errors = calibrator.GetCameraProjRMSE();
  for (size_t ii = 0; ii < max_reproj_errors.size() && ii < errors.size(); ++ii)
    if (errors[ii] > max_reproj_errors[ii]) return false;
  if (has_initial_guess) {
    for (size_t i = 0; i < input_cameras.size(); ++i)
      if (CameraCalibrationsDiffer(input_cameras[i], calibrator.GetCamera(i), t)) return false;
    if (IMUCalibrationDiffer(input_imu_biases, calibrator.GetBiases(), t)) return false;
  }
  return true;
}

}  // namespace visual_inertial_calibration
#endif  // VICALIB_B200_HOST_VICALIB_TASK_CHECKS_H_
