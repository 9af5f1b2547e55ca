// registration::RegistrationPipeline — I/algorithms/registration/registration_pipeline.hpp:17-151:
// optional random subsampling of the source (default ON, 1000 points, persistent mt19937(1234)),
// then the wrapper chain RobustAligner -> VelocityUpdateAligner -> Registration::align (:99-117).
#pragma once

#include <memory>
#include <stdexcept>

#include "sycl_points/algorithms/filter/preprocess_filter.hpp"
#include "sycl_points/algorithms/registration/pipeline/aligner.hpp"
#include "sycl_points/algorithms/registration/pipeline/robust.hpp"
#include "sycl_points/algorithms/registration/pipeline/velocity_update.hpp"
#include "sycl_points/algorithms/registration/registration.hpp"
#include "sycl_points/algorithms/registration/registration_pipeline_params.hpp"

namespace sycl_points {
namespace algorithms {
namespace registration {

class RegistrationPipeline {
public:
    using Ptr = std::shared_ptr<RegistrationPipeline>;

    /// any callable with the aligner signature (the reference's tests pass lambdas)
    RegistrationPipeline(pipeline::RegistrationAligner aligner,
                         const RegistrationPipelineParams& pipeline_params = RegistrationPipelineParams())
        : pipeline_params_(pipeline_params), aligner_(std::move(aligner)) {
        this->wrap_aligner();
    }
    RegistrationPipeline(const Registration::Ptr& registration,
                         const RegistrationPipelineParams& pipeline_params = RegistrationPipelineParams())
        : RegistrationPipeline(pipeline::make_registration_aligner(registration), pipeline_params) {
        this->registration_ = registration;
    }
    RegistrationPipeline(const sycl_utils::DeviceQueue& queue,
                         const RegistrationPipelineParams& pipeline_params = RegistrationPipelineParams())
        : RegistrationPipeline(std::make_shared<Registration>(queue, pipeline_params.registration), pipeline_params) {}

    RegistrationResult align(const PointCloudShared& source, const PointCloudShared& target,
                             const knn::KNNBase& target_knn,
                             const TransformMatrix& initial_guess = TransformMatrix::Identity(),
                             const Registration::ExecutionOptions& options = Registration::ExecutionOptions()) const {
        this->update_registration_input(source);
        return this->aligner_(*this->registration_input_pc_, target, target_knn, initial_guess, options);
    }

    const Registration::Ptr& registration() const { return this->registration_; }

    void compute_icp_robust_weights(const PointCloudShared& target, const knn::KNNBase& target_knn,
                                    const TransformMatrix& pose, float robust_scale, shared_vector<float>& out) const {
        if (this->registration_ == nullptr)
            throw std::runtime_error(
                "[RegistrationPipeline::compute_icp_robust_weights] Registration backend is not available.");
        const auto source = this->get_deskewed_point_cloud();
        if (source == nullptr)
            throw std::runtime_error(
                "[RegistrationPipeline::compute_icp_robust_weights] Registration input point cloud is not available.");
        this->registration_->compute_icp_robust_weights(*source, target, target_knn, pose, robust_scale, out);
    }

    const PointCloudShared* get_registration_input_point_cloud() const { return this->registration_input_pc_.get(); }
    /// the deskewed source of the most recent align(); the registration input when velocity update is off
    const PointCloudShared::Ptr get_deskewed_point_cloud() const {
        if (this->velocity_update_pipeline_ != nullptr) return this->velocity_update_pipeline_->get_deskewed_point_cloud();
        return this->registration_input_pc_;
    }

    float get_inlier_ratio(const RegistrationResult& result) const {
        const auto* input = this->get_registration_input_point_cloud();
        if (input && input->size() > 0) return static_cast<float>(result.inlier) / static_cast<float>(input->size());
        return 0.0f;
    }

private:
    void wrap_aligner() {
        // for each robust scale: for each deskew update: align(...)   (:99-117)
        if (this->pipeline_params_.velocity_update.enable) {
            this->velocity_update_pipeline_ = std::make_shared<pipeline::VelocityUpdateAligner>(
                this->aligner_, this->pipeline_params_.velocity_update.iter, this->pipeline_params_.registration.verbose);
            this->aligner_ = this->velocity_update_pipeline_->make_aligner();
        }
        if (this->pipeline_params_.robust.auto_scale) {
            this->robust_pipeline_ = std::make_shared<pipeline::RobustAligner>(this->aligner_, this->pipeline_params_);
            this->aligner_ = this->robust_pipeline_->make_aligner();
        }
    }

    /// registration_pipeline.hpp:127-140
    void update_registration_input(const PointCloudShared& source) const {
        if (this->preprocess_filter_ == nullptr || this->registration_input_pc_ == nullptr) {
            this->preprocess_filter_ = std::make_shared<filter::PreprocessFilter>(source.queue);
            this->registration_input_pc_ = std::make_shared<PointCloudShared>(source.queue);
        }
        const auto& rs = this->pipeline_params_.random_sampling;
        if (rs.enable && source.size() > rs.num) {
            if (rs.use_intensities && source.has_intensity()) {
                this->preprocess_filter_->mixed_random_sampling(source, *this->registration_input_pc_,
                                                                *source.intensities, rs.num, rs.weighted_ratio);
            } else {
                this->preprocess_filter_->random_sampling(source, *this->registration_input_pc_, rs.num);
            }
        } else {
            *this->registration_input_pc_ = source;
        }
    }

    Registration::Ptr registration_;
    RegistrationPipelineParams pipeline_params_;
    pipeline::RobustAligner::Ptr robust_pipeline_ = nullptr;
    pipeline::VelocityUpdateAligner::Ptr velocity_update_pipeline_ = nullptr;
    pipeline::RegistrationAligner aligner_;
    mutable filter::PreprocessFilter::Ptr preprocess_filter_ = nullptr;
    mutable PointCloudShared::Ptr registration_input_pc_ = nullptr;
};

}  // namespace registration
}  // namespace algorithms
}  // namespace sycl_points
