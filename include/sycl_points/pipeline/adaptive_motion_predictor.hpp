// lidar_odometry::AdaptiveMotionPredictor — I/pipeline/adaptive_motion_predictor.hpp:17-143: constant-velocity
// prediction, damped along the directions the previous registration constrained well.  Host code, as in the reference.
#pragma once

#include <algorithm>
#include <cmath>
#include <iostream>
#include <memory>
#include <optional>

#include "sycl_points/algorithms/registration/result.hpp"

namespace sycl_points {
namespace pipeline {
namespace lidar_odometry {

namespace detail {
/// smallest eigenvalue of a symmetric 3x3 (the reference asks Eigen::SelfAdjointEigenSolver for it): cyclic Jacobi in fp64
inline float min_eigenvalue_sym3(const Eigen::Matrix3f& A) {
    double a[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * ((double)A(i, j) + (double)A(j, i));
    for (int sweep = 0; sweep < 32; ++sweep) {
        if (a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2] < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double th = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double x = a[k][p], y = a[k][q];
                    a[k][p] = c * x - s * y;
                    a[k][q] = s * x + c * y;
                }
                for (int k = 0; k < 3; ++k) {
                    const double x = a[p][k], y = a[q][k];
                    a[p][k] = c * x - s * y;
                    a[q][k] = s * x + c * y;
                }
            }
    }
    return (float)std::min(a[0][0], std::min(a[1][1], a[2][2]));
}
}  // namespace detail

class AdaptiveMotionPredictor {
public:
    using Ptr = std::shared_ptr<AdaptiveMotionPredictor>;

    struct Params {
        struct AdaptiveAxis {
            float factor_min = 0.2f;
            float factor_max = 1.0f;
            float min_eigenvalue_low = 1.0f;
            float min_eigenvalue_high = 10.0f;
        };
        struct Adaptive {
            AdaptiveAxis rotation = {0.2f, 1.0f, 5.0f, 10.0f};
            AdaptiveAxis translation;
        };
        bool verbose = false;
        float velocity_ema_alpha = 1.0f;  // 1 = no smoothing
        Adaptive adaptive;
    };

    explicit AdaptiveMotionPredictor(const Params& params) : params_(params) {}

    /// :50-137
    Eigen::Isometry3f predict(const Eigen::Vector3f& linear_velocity, const Eigen::AngleAxisf& angular_velocity,
                              const Eigen::Isometry3f& odom, float dt,
                              const algorithms::registration::RegistrationResult::Ptr& reg_result, bool registrated) {
        float rot_factor = params_.adaptive.rotation.factor_max;
        float trans_factor = params_.adaptive.translation.factor_max;
        if (registrated && reg_result && reg_result->inlier > 0) {
            const Eigen::Matrix3f Hr = reg_result->H_raw.template block<3, 3>(0, 0);
            const Eigen::Matrix3f Ht = reg_result->H_raw.template block<3, 3>(3, 3);
            rot_factor = factor(detail::min_eigenvalue_sym3(Hr) / reg_result->inlier, params_.adaptive.rotation);
            trans_factor = factor(detail::min_eigenvalue_sym3(Ht) / reg_result->inlier, params_.adaptive.translation);
            if (params_.verbose)
                std::cout << "[motion predictor] rot: factor=" << rot_factor << ", trans: factor=" << trans_factor << std::endl;
        }
        const float vel_alpha = params_.velocity_ema_alpha;
        const Eigen::Vector3f ang_vec = angular_velocity.axis() * angular_velocity.angle();
        linear_velocity_smooth_ = linear_velocity_smooth_.has_value()
                                      ? Eigen::Vector3f(linear_velocity * vel_alpha + linear_velocity_smooth_.value() * (1.0f - vel_alpha))
                                      : linear_velocity;
        angular_velocity_smooth_ = angular_velocity_smooth_.has_value()
                                       ? Eigen::Vector3f(ang_vec * vel_alpha + angular_velocity_smooth_.value() * (1.0f - vel_alpha))
                                       : ang_vec;
        const Eigen::Vector3f lin_vel = linear_velocity_smooth_.value();
        const float ang_norm = angular_velocity_smooth_.value().norm();
        const Eigen::AngleAxisf ang_vel = ang_norm > 1e-6f
                                              ? Eigen::AngleAxisf(ang_norm, Eigen::Vector3f(angular_velocity_smooth_.value() / ang_norm))
                                              : Eigen::AngleAxisf::Identity();
        const Eigen::Vector3f delta_trans = lin_vel * dt;
        const Eigen::Matrix3f R = odom.rotation();
        const Eigen::Vector3f predicted_trans = odom.translation() + R * Eigen::Vector3f(delta_trans * trans_factor);
        const Eigen::Matrix3f predicted_rot = R * Eigen::AngleAxisf(ang_vel.angle() * dt * rot_factor, ang_vel.axis()).toRotationMatrix();
        Eigen::Isometry3f init_T = Eigen::Isometry3f::Identity();
        init_T.translation() = predicted_trans;
        init_T.linear() = predicted_rot;
        return init_T;
    }

    EIGEN_MAKE_ALIGNED_OPERATOR_NEW

private:
    static float factor(float min_eig_ratio, const Params::AdaptiveAxis& ax) {  // :63-73
        const float score = std::clamp((min_eig_ratio - ax.min_eigenvalue_low) /
                                           std::max(ax.min_eigenvalue_high - ax.min_eigenvalue_low, 1e-6f),
                                       0.0f, 1.0f);
        return ax.factor_max * (1.0f - score) + ax.factor_min * score;
    }

    Params params_;
    std::optional<Eigen::Vector3f> linear_velocity_smooth_;
    std::optional<Eigen::Vector3f> angular_velocity_smooth_;  // rotation vector [rad/s]
};

}  // namespace lidar_odometry
}  // namespace pipeline
}  // namespace sycl_points
