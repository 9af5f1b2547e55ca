// eigen_utils::lie — the host-side Lie-group helpers the registration wrappers call
// (I/utils/eigen_utils.hpp:909-943 se3_exp, :991-1034 se3_log), evaluated by libspx so that they agree bit for
// bit with what the device kernels compute.  The reference's device-side linear algebra helpers of the same header
// live inside the CUDA kernels here (csrc/spx_math.cuh) and have no host face.
#pragma once

#include "spx.h"
#include "sycl_points/points/types.hpp"
#include "sycl_points/utils/sycl_utils.hpp"

namespace sycl_points {
namespace eigen_utils {
namespace lie {

/// twist [rx ry rz tx ty tz] of a rigid transform (rotation first)
inline Eigen::Vector<float, 6> se3_log(const Eigen::Isometry3f& transform) {
    Eigen::Vector<float, 6> r;
    detail::spx_check(spx_se3_log(transform.data(), r.data()));
    return r;
}

inline Eigen::Matrix4f se3_exp(const Eigen::Vector<float, 6>& twist) {
    Eigen::Matrix4f T;
    detail::spx_check(spx_se3_exp(twist.data(), T.data()));
    return T;
}

}  // namespace lie
}  // namespace eigen_utils
}  // namespace sycl_points
