// pipeline::RobustAligner — I/algorithms/registration/pipeline/robust.hpp:17-125: runs the wrapped
// aligner once per robust-scale level, scale_{l+1} = scale_l * (min/init)^(1/(L-1)), each level
// seeded with the previous level's pose.
#pragma once

#include <cmath>
#include <iostream>
#include <memory>

#include "sycl_points/algorithms/registration/pipeline/aligner.hpp"
#include "sycl_points/algorithms/registration/registration_pipeline_params.hpp"

namespace sycl_points {
namespace algorithms {
namespace registration {
namespace pipeline {

class RobustAligner {
public:
    using Ptr = std::shared_ptr<RobustAligner>;

    RobustAligner(RegistrationAligner aligner, const RegistrationPipelineParams& params)
        : aligner_(std::move(aligner)), params_(params.registration), pipeline_params_(params.robust) {}

    RegistrationResult align(const PointCloudShared& source, const PointCloudShared& target,
                             const knn::KNNBase& target_knn,
                             const TransformMatrix& initial_guess = TransformMatrix::Identity(),
                             const Registration::ExecutionOptions& options = Registration::ExecutionOptions()) const {
        RegistrationResult result;
        result.T.matrix() = initial_guess;
        if (source.size() == 0) return result;

        const auto& pp = this->pipeline_params_;
        const bool fixed = options.robust_scale > 0.0f || options.rotation_robust_scale > 0.0f;
        bool auto_scaling = !fixed && this->params_.robust.type != robust::RobustLossType::NONE && pp.auto_scale;
        auto bad_range = [](float lo, float hi) { return lo <= 0.0f || lo >= hi; };
        if (auto_scaling && bad_range(pp.min_scale, pp.init_scale)) {
            std::cout << "[Caution] `pipeline.robust.min_scale` must be greater than zero and less than "
                         "`pipeline.robust.init_scale`."
                      << std::endl;
            auto_scaling = false;
        }
        if (auto_scaling && bad_range(pp.rotation_min_scale, pp.rotation_init_scale)) {
            std::cout << "[Caution] `pipeline.robust.rotation_min_scale` must be greater than zero and less than "
                         "`pipeline.robust.rotation_init_scale`."
                      << std::endl;
            auto_scaling = false;
        }
        if (auto_scaling && pp.auto_scaling_iter == 0) {
            std::cout << "[Caution] `pipeline.robust.auto_scaling_iter` must be greater than zero. Disable auto scaling."
                      << std::endl;
            auto_scaling = false;
        }
        const size_t levels = auto_scaling ? std::max<size_t>(1, pp.auto_scaling_iter) : 1;
        auto factor = [&](float lo, float hi) {
            return levels > 1 ? std::pow(lo / hi, 1.0f / static_cast<float>(levels - 1)) : 1.0f;
        };
        float scale = options.robust_scale > 0.0f ? options.robust_scale
                                                  : (auto_scaling ? pp.init_scale : this->params_.robust.default_scale);
        float rot_scale = options.rotation_robust_scale > 0.0f
                              ? options.rotation_robust_scale
                              : (auto_scaling ? pp.rotation_init_scale
                                              : this->params_.rotation_constraint.robust.default_scale);
        const float f = factor(pp.min_scale, pp.init_scale);
        const float rf = factor(pp.rotation_min_scale, pp.rotation_init_scale);
        for (size_t level = 0; level < levels; ++level) {
            if (auto_scaling && this->params_.verbose) std::cout << "Robust scale: " << scale << std::endl;
            auto level_options = options;
            level_options.robust_scale = scale;
            level_options.rotation_robust_scale = rot_scale;
            result = this->aligner_(source, target, target_knn, result.T.matrix(), level_options);
            scale *= f;
            rot_scale *= rf;
        }
        return result;
    }

    RegistrationAligner make_aligner() const {
        return [this](const PointCloudShared& source, const PointCloudShared& target, const knn::KNNBase& target_knn,
                      const TransformMatrix& initial_guess, const Registration::ExecutionOptions& options) {
            return this->align(source, target, target_knn, initial_guess, options);
        };
    }

private:
    RegistrationAligner aligner_;
    RegistrationParams params_;
    RegistrationPipelineParams::Robust pipeline_params_;
};

}  // namespace pipeline
}  // namespace registration
}  // namespace algorithms
}  // namespace sycl_points
