// IMU residual family on the device (SwitchedFullImuCostFunction, ceres-cost-functions.h:379-490)
// and the block-tridiagonal frame chain it induces.  First-stage stubs: the vision-only path is
// complete; inertial terms are wired next.
#pragma once
#include "vc_internal.h"

namespace vc {

inline int imu_not_ready(vcgpu_handle* h) {
  h->err = "inertial terms are not implemented on the device yet";
  return VCGPU_ERR_INVALID;
}
inline void imu_free(vcgpu_handle*) {}
inline int imu_prepare(vcgpu_handle* h) { return h->dp.inertial ? imu_not_ready(h) : VCGPU_OK; }
inline int imu_evaluate(vcgpu_handle* h, int, bool, int* n_cost) { *n_cost = 0; return imu_not_ready(h); }
inline int imu_accumulate(vcgpu_handle* h, int) { return imu_not_ready(h); }
inline const double* imu_cost_part(vcgpu_handle*) { return nullptr; }
inline int imu_chain_solve(vcgpu_handle* h, int, const double*) { return imu_not_ready(h); }
inline int imu_chain_backsub(vcgpu_handle* h, int, const double*) { return imu_not_ready(h); }
inline int imu_update_weights(vcgpu_handle* h, int) { return imu_not_ready(h); }
inline int imu_eval_hook(vcgpu_handle* h, double*, double*) { return imu_not_ready(h); }

}  // namespace vc
