// lidar_odometry::MotionPredictor — I/pipeline/motion_predictor.hpp:20-84.  Without an IMU no gyro / IMU-SE3 candidate is
// ever supplied, so every mode resolves to the LiDAR constant-velocity prediction — exactly what the reference does when
// the candidates are empty (:64-76).
#pragma once

#include <algorithm>
#include <cctype>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>

#include "sycl_points/pipeline/adaptive_motion_predictor.hpp"

namespace sycl_points {
namespace pipeline {
namespace lidar_odometry {

enum class MotionPredictionMode { LIDAR_CV = 0, GYRO_LIDAR_CV, IMU_SE3 };

namespace detail {
struct ModeName {
    MotionPredictionMode mode;
    const char* name;
};
inline constexpr ModeName kModeNames[] = {{MotionPredictionMode::LIDAR_CV, "LIDAR_CV"},
                                          {MotionPredictionMode::GYRO_LIDAR_CV, "GYRO_LIDAR_CV"},
                                          {MotionPredictionMode::IMU_SE3, "IMU_SE3"}};
}  // namespace detail

/// case-insensitive; throws std::runtime_error for an unknown name (motion_predictor.hpp:26-33)
inline MotionPredictionMode MotionPredictionMode_from_string(const std::string& str) {
    std::string upper(str);
    for (char& c : upper) c = (char)std::toupper((unsigned char)c);
    for (const auto& m : detail::kModeNames)
        if (upper == m.name) return m.mode;
    throw std::runtime_error("[MotionPredictionMode_from_string] Invalid motion prediction mode '" + str + "'");
}
/// motion_predictor.hpp:35-45
inline std::string MotionPredictionMode_to_string(const MotionPredictionMode mode) {
    for (const auto& m : detail::kModeNames)
        if (mode == m.mode) return m.name;
    throw std::runtime_error("[MotionPredictionMode_to_string] Invalid motion prediction mode");
}

struct MotionPredictionCandidates {
    std::optional<Eigen::Matrix3f> gyro_delta_rotation_lidar;
    std::optional<Eigen::Isometry3f> imu_se3_pose;
};

class MotionPredictor {
public:
    using Ptr = std::shared_ptr<MotionPredictor>;

    struct Params : AdaptiveMotionPredictor::Params {
        MotionPredictionMode mode = MotionPredictionMode::GYRO_LIDAR_CV;
    };

    explicit MotionPredictor(const Params& params) : params_(params), lidar_cv_predictor_(params) {}

    Eigen::Isometry3f predict(const Eigen::Vector3f& linear_velocity, const Eigen::AngleAxisf& angular_velocity,
                              const Eigen::Isometry3f& odom, float dt,
                              const algorithms::registration::RegistrationResult::Ptr& reg_result, bool registrated,
                              const MotionPredictionCandidates& candidates = {}) {
        if (params_.mode == MotionPredictionMode::IMU_SE3 && candidates.imu_se3_pose) return *candidates.imu_se3_pose;
        Eigen::Isometry3f prediction =
            lidar_cv_predictor_.predict(linear_velocity, angular_velocity, odom, dt, reg_result, registrated);
        if (params_.mode == MotionPredictionMode::GYRO_LIDAR_CV && candidates.gyro_delta_rotation_lidar) {
            Eigen::Isometry3f relative = odom.inverse() * prediction;
            relative.linear() = *candidates.gyro_delta_rotation_lidar;
            prediction = odom * relative;
        }
        return prediction;
    }

    EIGEN_MAKE_ALIGNED_OPERATOR_NEW

private:
    Params params_;
    AdaptiveMotionPredictor lidar_cv_predictor_;
};

}  // namespace lidar_odometry
}  // namespace pipeline
}  // namespace sycl_points
