// registration::Registration — I/algorithms/registration/registration.hpp:88-965 over libspx.
//
// With a knn::KDTree as `target_knn` the whole align() runs on the device (spx_registration_align:
// Gauss-Newton as one cooperative kernel, LM / dog-leg with one host decision per trial step).
// With any other KNNBase (user subclasses, as in T/test_registration_pipeline.cpp:16-61) the
// reference's host loop is kept: correspondences from the injected object, linearise / error on
// the device (spx_linearize / spx_error), the 6x6 step on the host (spx_solve_6x6, spx_se3_exp,
// spx_dogleg_step).
#pragma once

#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <tuple>

#include "sycl_points/algorithms/feature/covariance.hpp"
#include "sycl_points/algorithms/knn/kdtree.hpp"
#include "sycl_points/algorithms/knn/knn.hpp"
#include "sycl_points/algorithms/registration/linearized_result.hpp"
#include "sycl_points/algorithms/registration/registration_params.hpp"
#include "sycl_points/algorithms/registration/result.hpp"
#include "sycl_points/points/point_cloud.hpp"

namespace sycl_points {
namespace algorithms {
namespace registration {

class Registration {
public:
    using Ptr = std::shared_ptr<Registration>;

    /// registration.hpp:92-100
    struct ExecutionOptions {
        float robust_scale;
        float rotation_robust_scale;
        float dt;
        TransformMatrix prev_pose;
        ExecutionOptions()
            : robust_scale(-1.0f), rotation_robust_scale(-1.0f), dt(0.1f), prev_pose(TransformMatrix::Identity()) {}
    };

    /// registration.hpp:105-114
    Registration(const sycl_utils::DeviceQueue& queue, const RegistrationParams& params = RegistrationParams())
        : params_(params), queue_(queue) {
        const spx_registration_params p = to_c(params);
        detail::spx_check(spx_registration_create(queue.handle(), &p, &this->handle_));
        const spx_registration_addons a = addons_of(params);  // registration.hpp:112-113
        detail::spx_check(spx_registration_set_addons(this->handle_, &a));
    }

    /// registration.hpp:124-126: arms the MAP prior of the next align from the previous result and the predicted pose
    void set_map_prior_state(const RegistrationResult& prev_result, const Eigen::Isometry3f& T_pred) {
        spx_registration_result R{};
        for (int i = 0; i < 16; ++i) R.T[i] = prev_result.T.matrix().data()[i];
        for (int i = 0; i < 36; ++i) R.H_raw[i] = prev_result.H_raw.data()[i];
        R.error_raw = prev_result.error_raw;
        R.inlier = prev_result.inlier;
        detail::spx_check(spx_registration_set_map_prior_state(this->handle_, &R, T_pred.matrix().data(), nullptr, nullptr));
    }
    Registration(const sycl_utils::DeviceQueue& queue, const RegistrationFactorParams& params)
        : Registration(queue, RegistrationParams(params)) {}
    ~Registration() {
        if (this->handle_) spx_registration_destroy(this->handle_);
    }
    Registration(const Registration&) = delete;
    Registration& operator=(const Registration&) = delete;

    const RegistrationParams& params() const { return this->params_; }

    /// registration.hpp:201-276
    RegistrationResult align(const PointCloudShared& source, const PointCloudShared& target,
                             const knn::KNNBase& target_knn,
                             const TransformMatrix& initial_guess = TransformMatrix::Identity(),
                             const ExecutionOptions& options = ExecutionOptions()) {
        RegistrationResult result;
        result.T.matrix() = initial_guess;
        if (source.size() == 0) return result;
        this->validate_params(source, target);

        const auto* tree = dynamic_cast<const knn::KDTree*>(&target_knn);
        if (tree == nullptr || tree->handle() == nullptr) return this->align_injected(source, target, target_knn, result, options);

        this->prefetch(source, target);
        spx_registration_params p = to_c(this->params_);
        if (options.rotation_robust_scale > 0.0f) p.rotation_constraint_robust_scale = options.rotation_robust_scale;  // :219-221
        detail::spx_check(spx_registration_set_params(this->handle_, &p));
        spx_registration_result R;
        detail::spx_check(spx_registration_align(
            this->handle_, fptr(source.points_ptr()), source.has_cov() ? fptr(source.covs_ptr()) : nullptr, source.size(),
            fptr(target.points_ptr()), target.has_cov() ? fptr(target.covs_ptr()) : nullptr,
            target.has_normal() ? fptr(target.normals_ptr()) : nullptr, target.size(), tree->handle(),
            initial_guess.data(), options.robust_scale, &R, nullptr));
        from_c(R, result);
        return result;
    }

    /// registration.hpp:279-294
    void compute_icp_robust_weights(const PointCloudShared& source, const PointCloudShared& target,
                                    const knn::KNNBase& target_knn, const TransformMatrix& pose, float robust_scale,
                                    shared_vector<float>& out) const {
        const size_t N = source.size();
        out.resize(N);
        if (N == 0) return;
        target_knn.nearest_neighbor_search_async(source, this->neighbors_, {}, pose).wait_and_throw();
        this->prefetch(source, target);
        this->queue_.set_accessed_by_device(out.data(), N);
        const float mc = this->params_.max_correspondence_distance;
        detail::spx_check(spx_robust_weights(
            this->queue_.handle(), (int)this->params_.reg_type, (int)this->effective_loss(), fptr(source.points_ptr()),
            source.has_cov() ? fptr(source.covs_ptr()) : nullptr, N, fptr(target.points_ptr()),
            target.has_cov() ? fptr(target.covs_ptr()) : nullptr,
            target.has_normal() ? fptr(target.normals_ptr()) : nullptr, this->neighbors_.indices->data(),
            this->neighbors_.distances->data(), pose.data(), mc * mc, robust_scale, out.data()));
        detail::spx_check(spx_queue_sync(this->queue_.handle()));
    }

    /// registration.hpp:312-323: the raw linearisation, degenerate-regularised relative to `initial_pose`
    LinearizedResult compute_linearized_result(const PointCloudShared& source, const PointCloudShared& target,
                                               const knn::KNNBase& target_knn, const TransformMatrix& pose,
                                               const TransformMatrix& initial_pose,
                                               const ExecutionOptions& options = ExecutionOptions()) {
        target_knn.nearest_neighbor_search_async(source, this->neighbors_, {}, pose).wait_and_throw();
        LinearizedResult lin = this->linearize(source, target, pose, this->scale_of(options));
        if (this->params_.degenerate_reg.type != DegenerateRegularizationType::none) {
            const spx_registration_addons a = addons_of(this->params_);
            detail::spx_check(spx_degenerate_regularize(&a, lin.H.data(), lin.b.data(), lin.inlier, pose.data(),
                                                        initial_pose.data()));
        }
        return lin;
    }
    LinearizedResult compute_linearized_result(const PointCloudShared& source, const PointCloudShared& target,
                                               const knn::KNNBase& target_knn, const TransformMatrix& pose,
                                               const ExecutionOptions& options = ExecutionOptions()) {
        return this->compute_linearized_result(source, target, target_knn, pose, pose, options);
    }

    /// registration.hpp:350-359
    std::tuple<float, uint32_t> compute_error_frozen(const PointCloudShared& source, const PointCloudShared& target,
                                                     const TransformMatrix& pose,
                                                     const ExecutionOptions& options = ExecutionOptions()) const {
        return this->error_at(source, target, pose, this->scale_of(options));
    }

private:
    static const float* fptr(const void* p) { return static_cast<const float*>(p); }

    static spx_registration_addons addons_of(const RegistrationParams& P) {
        spx_registration_addons a;
        spx_default_registration_addons(&a);
        a.degenerate_type = P.degenerate_reg.type == DegenerateRegularizationType::nl_reg ? 1 : 0;
        a.rot_eigenvalue_threshold = P.degenerate_reg.rot_eigenvalue_threshold;
        a.trans_eigenvalue_threshold = P.degenerate_reg.trans_eigenvalue_threshold;
        a.base_factor = P.degenerate_reg.base_factor;
        a.map_prior_enabled = P.map_prior.enabled ? 1 : 0;
        a.rot_vel_sigma = P.map_prior.rot_vel_sigma;
        a.trans_vel_sigma = P.map_prior.trans_vel_sigma;
        a.rot_base_sigma = P.map_prior.rot_base_sigma;
        a.trans_base_sigma = P.map_prior.trans_base_sigma;
        return a;
    }

    static spx_registration_params to_c(const RegistrationParams& P) {
        spx_registration_params p;
        spx_default_registration_params(&p);
        p.reg_type = (int32_t)P.reg_type;
        p.robust_loss = (int32_t)P.robust.type;
        p.optimization_method = (int32_t)P.optimization_method;
        p.max_iterations = (int32_t)P.max_iterations;
        p.max_correspondence_distance = P.max_correspondence_distance;
        p.robust_default_scale = P.robust.default_scale;
        p.criteria_translation = P.criteria.translation;
        p.criteria_rotation = P.criteria.rotation;
        p.gn_lambda = P.gn.lambda;
        p.lm_max_inner_iterations = (int32_t)P.lm.max_inner_iterations;
        p.lm_lambda_factor = P.lm.lambda_factor;
        p.lm_init_lambda = P.lm.init_lambda;
        p.lm_max_lambda = P.lm.max_lambda;
        p.lm_min_lambda = P.lm.min_lambda;
        p.dogleg_initial_trust_region_radius = P.dogleg.initial_trust_region_radius;
        p.dogleg_min_trust_region_radius = P.dogleg.min_trust_region_radius;
        p.dogleg_max_trust_region_radius = P.dogleg.max_trust_region_radius;
        p.dogleg_eta1 = P.dogleg.eta1;
        p.dogleg_eta2 = P.dogleg.eta2;
        p.dogleg_gamma_decrease = P.dogleg.gamma_decrease;
        p.dogleg_gamma_increase = P.dogleg.gamma_increase;
        p.genz_planarity_threshold = P.genz.planarity_threshold;
        p.rotation_constraint_enable = P.rotation_constraint.enable ? 1 : 0;
        p.rotation_constraint_weight = P.rotation_constraint.weight;
        p.rotation_constraint_robust_scale = P.rotation_constraint.robust.default_scale;
        return p;
    }
    static void from_c(const spx_registration_result& R, RegistrationResult& out) {
        std::memcpy(out.T.matrix().data(), R.T, sizeof(R.T));
        out.converged = R.converged != 0;
        out.iterations = (size_t)R.iterations;
        std::memcpy(out.H.data(), R.H, sizeof(R.H));  // symmetric: row- and column-major coincide
        std::memcpy(out.b.data(), R.b, sizeof(R.b));
        out.error = R.error;
        std::memcpy(out.H_raw.data(), R.H_raw, sizeof(R.H_raw));
        std::memcpy(out.b_raw.data(), R.b_raw, sizeof(R.b_raw));
        out.error_raw = R.error_raw;
        out.inlier = R.inlier;
    }

    /// registration.hpp:129-193
    void validate_params(const PointCloudShared& source, const PointCloudShared& target) {
        if (this->params_.reg_type == RegType::POINT_TO_PLANE && !target.has_normal()) {
            if (!target.has_cov())
                throw std::runtime_error(
                    "[Registration::validate_params] Normal vector or covariance matrices of target must be "
                    "pre-computed before performing Point-to-Plane ICP matching.");
            std::cout << "[Caution] Normal vectors for Point-to-Plane ICP are not provided. " << std::endl;
            std::cout << "          Attempting to derive them from pre-computed covariance matrices." << std::endl;
            covariance::extract_normals(target);
        }
        if (this->params_.reg_type == RegType::GICP && (!source.has_cov() || !target.has_cov()))
            throw std::runtime_error(
                "[Registration::validate_params] Covariance matrices of source and target must be pre-computed "
                "before performing GICP matching.");
        if (this->params_.reg_type == RegType::GENZ) {
            if (!target.has_cov())
                throw std::runtime_error(
                    "[Registration::validate_params] Covariance matrices of target must be pre-computed before "
                    "performing GenZ-ICP matching.");
            if (!target.has_normal()) {
                std::cout << "[Caution] Normal vectors for GenZ-ICP are not provided. " << std::endl;
                std::cout << "          Attempting to derive them from pre-computed covariance matrices." << std::endl;
                covariance::extract_normals(target);
            }
            // the stateless C entry points (linearise / error / weights) read the threshold from a thread-local setting
            detail::spx_check(spx_set_genz_planarity_threshold(this->params_.genz.planarity_threshold));
        }
        if (this->params_.reg_type == RegType::POINT_TO_DISTRIBUTION && !target.has_cov())
            throw std::runtime_error(
                "[Registration::validate_params] Covariance matrices of target must be pre-computed before "
                "performing Point-to-Distribution ICP matching.");
        if (this->params_.rotation_constraint.enable) {
            if (!source.has_cov())
                throw std::runtime_error(
                    "[Registration::validate_params] Covariance matrices of source are required for performing "
                    "rotation constraint matching.");
            if (!target.has_cov())
                throw std::runtime_error(
                    "[Registration::validate_params] Covariance matrices of target are required for performing "
                    "rotation constraint matching.");
        }
        // the stateless C entry points (linearise / error / weights) read this add-on from a thread-local setting
        detail::spx_check(spx_set_rotation_constraint(this->params_.rotation_constraint.enable ? 1 : 0,
                                                      this->params_.rotation_constraint.weight,
                                                      this->params_.rotation_constraint.robust.default_scale));
        if (this->params_.robust.type != robust::RobustLossType::NONE && this->params_.robust.default_scale <= 0.0f) {
            std::cout << "[Caution] `robust.default_scale` must be greater than zero. Disable robust loss." << std::endl;
            this->params_.robust.type = robust::RobustLossType::NONE;
        }
    }
    robust::RobustLossType effective_loss() const {
        return (this->params_.robust.type != robust::RobustLossType::NONE && this->params_.robust.default_scale <= 0.0f)
                   ? robust::RobustLossType::NONE
                   : this->params_.robust.type;
    }
    float scale_of(const ExecutionOptions& options) const {  // registration.hpp:217-218
        return options.robust_scale > 0.0f ? options.robust_scale : this->params_.robust.default_scale;
    }
    void prefetch(const PointCloudShared& source, const PointCloudShared& target) const {
        const auto& q = this->queue_;
        q.set_accessed_by_device(source.points_ptr(), source.size());
        q.set_accessed_by_device(target.points_ptr(), target.size());
        if (source.has_cov()) q.set_accessed_by_device(source.covs_ptr(), source.size());
        if (target.has_cov()) q.set_accessed_by_device(target.covs_ptr(), target.size());
        if (target.has_normal()) q.set_accessed_by_device(target.normals_ptr(), target.size());
    }

    LinearizedResult linearize(const PointCloudShared& source, const PointCloudShared& target, const TransformMatrix& T,
                               float scale) const {
        LinearizedResult out;
        this->prefetch(source, target);
        const float mc = this->params_.max_correspondence_distance;
        detail::spx_check(spx_linearize(
            this->queue_.handle(), (int)this->params_.reg_type, (int)this->effective_loss(), fptr(source.points_ptr()),
            source.has_cov() ? fptr(source.covs_ptr()) : nullptr, source.size(), fptr(target.points_ptr()),
            target.has_cov() ? fptr(target.covs_ptr()) : nullptr,
            target.has_normal() ? fptr(target.normals_ptr()) : nullptr, this->neighbors_.indices->data(),
            this->neighbors_.distances->data(), T.data(), mc * mc, scale, out.H.data(), out.b.data(), &out.error,
            &out.inlier));
        return out;
    }
    std::tuple<float, uint32_t> error_at(const PointCloudShared& source, const PointCloudShared& target,
                                         const TransformMatrix& T, float scale) const {
        float err = 0.0f;
        uint32_t inl = 0;
        this->prefetch(source, target);
        const float mc = this->params_.max_correspondence_distance;
        detail::spx_check(spx_error(
            this->queue_.handle(), (int)this->params_.reg_type, (int)this->effective_loss(), fptr(source.points_ptr()),
            source.has_cov() ? fptr(source.covs_ptr()) : nullptr, source.size(), fptr(target.points_ptr()),
            target.has_cov() ? fptr(target.covs_ptr()) : nullptr,
            target.has_normal() ? fptr(target.normals_ptr()) : nullptr, this->neighbors_.indices->data(),
            this->neighbors_.distances->data(), T.data(), mc * mc, scale, &err, &inl));
        return {err, inl};
    }

    static bool solve(const Eigen::Matrix<float, 6, 6>& H, const Eigen::Vector<float, 6>& b, float lambda,
                      Eigen::Vector<float, 6>& delta) {
        int ok = 0;
        detail::spx_check(spx_solve_6x6(H.data(), b.data(), lambda, delta.data(), &ok));
        return ok != 0;
    }
    static TransformMatrix apply_step(const TransformMatrix& T, const Eigen::Vector<float, 6>& delta) {
        TransformMatrix E;
        detail::spx_check(spx_se3_exp(delta.data(), E.data()));
        return T * E;  // registration.hpp:814
    }
    bool is_converged(const Eigen::Vector<float, 6>& d) const {  // registration.hpp:407-410
        const float rot = std::sqrt(d(0) * d(0) + d(1) * d(1) + d(2) * d(2));
        const float trans = std::sqrt(d(3) * d(3) + d(4) * d(4) + d(5) * d(5));
        return rot < this->params_.criteria.rotation && trans < this->params_.criteria.translation;
    }

    /// registration.hpp:227-272 + :803-964 with an injected KNNBase
    RegistrationResult align_injected(const PointCloudShared& source, const PointCloudShared& target,
                                      const knn::KNNBase& knn, RegistrationResult result, const ExecutionOptions& options) {
        const auto& P = this->params_;
        const float scale = this->scale_of(options);
        float lambda = P.lm.init_lambda;
        float radius = P.dogleg.initial_trust_region_radius;
        auto clamp_radius = [&](float r) {
            return std::min(std::max(r, P.dogleg.min_trust_region_radius), P.dogleg.max_trust_region_radius);
        };
        auto clamp_lambda = [&](float l) { return std::min(std::max(l, P.lm.min_lambda), P.lm.max_lambda); };
        for (size_t iter = 0; iter < P.max_iterations; ++iter) {
            knn.nearest_neighbor_search_async(source, this->neighbors_, {}, result.T.matrix()).wait_and_throw();
            const LinearizedResult lin = this->linearize(source, target, result.T.matrix(), scale);
            result.H_raw = lin.H;
            result.b_raw = lin.b;
            result.error_raw = lin.error;
            Eigen::Vector<float, 6> delta = Eigen::Vector<float, 6>::Zero();
            if (P.optimization_method == OptimizationMethod::GAUSS_NEWTON) {
                const bool ok = solve(lin.H, lin.b, P.gn.lambda, delta);
                result.converged = ok && this->is_converged(delta);
                result.T.matrix() = apply_step(result.T.matrix(), delta);
                result.iterations = iter;
                result.H = lin.H;
                result.b = lin.b;
                result.error = lin.error;
                result.inlier = lin.inlier;
            } else if (P.optimization_method == OptimizationMethod::LEVENBERG_MARQUARDT) {
                float last_error = std::numeric_limits<float>::max();
                for (size_t inner = 0; inner < P.lm.max_inner_iterations; ++inner) {
                    const bool ok = solve(lin.H, lin.b, lambda, delta);
                    result.converged = ok && this->is_converged(delta);
                    const TransformMatrix trial = apply_step(result.T.matrix(), delta);
                    const auto [new_error, inlier] = this->error_at(source, target, trial, scale);
                    const bool accept = new_error <= lin.error;
                    if (accept || std::fabs(new_error - last_error) <= 1e-6f) {
                        result.converged = this->is_converged(delta);
                        result.T.matrix() = trial;
                        result.error = new_error;
                        result.inlier = inlier;
                        if (accept) lambda = clamp_lambda(lambda / P.lm.lambda_factor);
                        break;
                    }
                    lambda = clamp_lambda(lambda * P.lm.lambda_factor);
                    last_error = new_error;
                }
                result.iterations = iter;
                result.H = lin.H;
                result.b = lin.b;
            } else {
                result.H = lin.H;
                result.b = lin.b;
                result.error = lin.error;
                result.inlier = lin.inlier;
                result.iterations = iter;
                radius = clamp_radius(radius);
                float step_norm = 0.0f, predicted = 0.0f;
                detail::spx_check(spx_dogleg_step(lin.H.data(), lin.b.data(), radius, delta.data(), &step_norm, &predicted));
                if (predicted <= 0.0f) {
                    radius = clamp_radius(radius * P.dogleg.gamma_decrease);
                } else {
                    const TransformMatrix trial = apply_step(result.T.matrix(), delta);
                    const auto [new_error, inlier] = this->error_at(source, target, trial, scale);
                    const float rho = (lin.error - new_error) / predicted;
                    if (rho < P.dogleg.eta1) {
                        radius = clamp_radius(radius * P.dogleg.gamma_decrease);
                    } else {
                        result.converged = this->is_converged(delta);
                        result.T.matrix() = trial;
                        result.error = new_error;
                        result.inlier = inlier;
                        if (rho > P.dogleg.eta2 && step_norm >= radius * 0.99f)
                            radius = clamp_radius(radius * P.dogleg.gamma_increase);
                    }
                }
            }
            if (result.converged) break;
        }
        return result;
    }

    RegistrationParams params_;
    sycl_utils::DeviceQueue queue_;
    spx_registration_t handle_ = nullptr;
    mutable knn::KNNResult neighbors_;  // registration.hpp:365
};

}  // namespace registration
}  // namespace algorithms
}  // namespace sycl_points
