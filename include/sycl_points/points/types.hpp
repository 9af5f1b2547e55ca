// Point / covariance / transform types — I/points/types.hpp:11-50, same names and memory layout.
#pragma once

#include <vector>

#include "sycl_points/compat/eigen_lite.hpp"
#include "sycl_points/utils/sycl_utils.hpp"

namespace sycl_points {

using PointType = Eigen::Vector4f;        // xyz1
using Covariance = Eigen::Matrix4f;       // 3x3 in the top-left block, column-major, 4th row/col zero
using Normal = Eigen::Vector4f;           // xyz0
using RGBType = Eigen::Vector4f;
using TransformMatrix = Eigen::Matrix4f;  // column-major
using TimestampOffset = float;

static_assert(sizeof(PointType) == 16 && sizeof(Covariance) == 64, "layout must match float[4] / float[16]");

using PointContainerCPU = std::vector<PointType, Eigen::aligned_allocator<PointType>>;
using PointContainerShared = shared_vector<PointType>;
using CovarianceContainerCPU = std::vector<Covariance, Eigen::aligned_allocator<Covariance>>;
using CovarianceContainerShared = shared_vector<Covariance>;
using NormalContainerCPU = std::vector<Normal, Eigen::aligned_allocator<Normal>>;
using NormalContainerShared = shared_vector<Normal>;
using RGBContainerCPU = std::vector<RGBType, Eigen::aligned_allocator<RGBType>>;
using RGBContainerShared = shared_vector<RGBType>;
using IntensityContainerCPU = std::vector<float, Eigen::aligned_allocator<float>>;
using IntensityContainerShared = shared_vector<float>;
using TimestampContainerCPU = std::vector<TimestampOffset>;
using TimestampContainerShared = shared_vector<TimestampOffset>;

}  // namespace sycl_points
