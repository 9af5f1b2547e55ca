// registration::LinearizedResult — I/algorithms/registration/linearized_result.hpp:12-23
#pragma once

#include <cstdint>
#include <limits>

#include "sycl_points/compat/eigen_lite.hpp"

namespace sycl_points {
namespace algorithms {
namespace registration {

struct LinearizedResult {
    Eigen::Matrix<float, 6, 6> H = Eigen::Matrix<float, 6, 6>::Zero();
    Eigen::Vector<float, 6> b = Eigen::Vector<float, 6>::Zero();
    float error = std::numeric_limits<float>::max();
    uint32_t inlier = 0;
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

}  // namespace registration
}  // namespace algorithms
}  // namespace sycl_points
