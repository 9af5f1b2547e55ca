// registration::RegistrationResult — I/algorithms/registration/result.hpp:13-28
#pragma once

#include <cstdint>
#include <limits>
#include <memory>

#include "sycl_points/compat/eigen_lite.hpp"

namespace sycl_points {
namespace algorithms {
namespace registration {

struct RegistrationResult {
    using Ptr = std::shared_ptr<RegistrationResult>;

    Eigen::Isometry3f T = Eigen::Isometry3f::Identity();
    bool converged = false;
    size_t iterations = 0;  // index of the last iteration run (registration.hpp:815), not a count
    Eigen::Matrix<float, 6, 6> H = Eigen::Matrix<float, 6, 6>::Zero();
    Eigen::Vector<float, 6> b = Eigen::Vector<float, 6>::Zero();
    float error = std::numeric_limits<float>::max();
    Eigen::Matrix<float, 6, 6> H_raw = Eigen::Matrix<float, 6, 6>::Zero();
    Eigen::Vector<float, 6> b_raw = Eigen::Vector<float, 6>::Zero();
    float error_raw = std::numeric_limits<float>::max();
    uint32_t inlier = 0;
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

}  // namespace registration
}  // namespace algorithms
}  // namespace sycl_points
