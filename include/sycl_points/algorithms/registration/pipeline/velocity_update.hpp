// pipeline::VelocityUpdateAligner — I/algorithms/registration/pipeline/velocity_update.hpp:16-104: under a
// constant-velocity model, deskew the source with the current pose estimate, align, and repeat.
#pragma once

#include <algorithm>
#include <iostream>
#include <memory>

#include "sycl_points/algorithms/deskew/relative_pose_deskew.hpp"
#include "sycl_points/algorithms/registration/pipeline/aligner.hpp"

namespace sycl_points {
namespace algorithms {
namespace registration {
namespace pipeline {

class VelocityUpdateAligner {
public:
    using Ptr = std::shared_ptr<VelocityUpdateAligner>;

    VelocityUpdateAligner(RegistrationAligner aligner, size_t velocity_update_iter, bool verbose = false)
        : aligner_(std::move(aligner)), velocity_update_iter_(velocity_update_iter), verbose_(verbose) {}
    VelocityUpdateAligner(const Registration::Ptr& registration, size_t velocity_update_iter, bool verbose = false)
        : VelocityUpdateAligner(make_registration_aligner(registration), velocity_update_iter, verbose) {}

    /// options.prev_pose / options.dt give the motion model; a source without timestamps is aligned as it is
    RegistrationResult align(const PointCloudShared& source, const PointCloudShared& target,
                             const knn::KNNBase& target_knn,
                             const TransformMatrix& initial_guess = TransformMatrix::Identity(),
                             const Registration::ExecutionOptions& options = Registration::ExecutionOptions()) const {
        RegistrationResult result;
        result.T.matrix() = initial_guess;
        if (source.size() == 0) return result;

        this->deskewed_pc_ = std::make_shared<PointCloudShared>(source.queue);  // never aliases the source
        if (!source.has_timestamps()) {
            if (this->verbose_) std::cout << "deskew skipped: source has no timestamps" << std::endl;
            *this->deskewed_pc_ = source;
            return this->aligner_(*this->deskewed_pc_, target, target_knn, result.T.matrix(), options);
        }

        const size_t levels = std::max<size_t>(1, this->velocity_update_iter_);
        for (size_t it = 0; it < levels; ++it) {
            if (this->verbose_) {
                const auto twist = eigen_utils::lie::se3_log(Eigen::Isometry3f(options.prev_pose).inverse() * result.T);
                std::cout << "deskewed[" << it << "]: angle=" << twist.head<3>().norm()
                          << ", dist=" << twist.tail<3>().norm() << std::endl;
            }
            deskew::deskew_point_cloud_constant_velocity(source, *this->deskewed_pc_,
                                                         Eigen::Isometry3f(options.prev_pose), result.T, options.dt);
            result = this->aligner_(*this->deskewed_pc_, target, target_knn, result.T.matrix(), options);
        }
        return result;
    }

    const PointCloudShared::Ptr get_deskewed_point_cloud() const { return this->deskewed_pc_; }

    RegistrationAligner make_aligner() const {
        return [this](const PointCloudShared& source, const PointCloudShared& target, const knn::KNNBase& target_knn,
                      const TransformMatrix& initial_guess, const Registration::ExecutionOptions& options) {
            return this->align(source, target, target_knn, initial_guess, options);
        };
    }

private:
    RegistrationAligner aligner_;
    size_t velocity_update_iter_ = 1;
    bool verbose_ = false;
    mutable PointCloudShared::Ptr deskewed_pc_ = nullptr;
};

}  // namespace pipeline
}  // namespace registration
}  // namespace algorithms
}  // namespace sycl_points
