// Registration parameter structs — I/algorithms/registration/registration_params.hpp:17-114, same
// names and defaults, including the solver add-ons' parameter structs (degenerate_regularization.hpp:14-40,
// map_prior.hpp:15-21), which the reference declares in their own headers.
#pragma once

#include <algorithm>
#include <cctype>
#include <cstddef>
#include <stdexcept>
#include <string>

#include "sycl_points/algorithms/registration/factor.hpp"
#include "sycl_points/algorithms/robust/robust.hpp"

namespace sycl_points {
namespace algorithms {
namespace registration {

enum class OptimizationMethod { GAUSS_NEWTON = 0, LEVENBERG_MARQUARDT, POWELL_DOGLEG };

inline OptimizationMethod OptimizationMethod_from_string(const std::string& str) {
    std::string upper(str.size(), '\0');
    std::transform(str.begin(), str.end(), upper.begin(), [](unsigned char c) { return std::toupper(c); });
    if (upper == "GN" || upper == "GAUSS_NEWTON") return OptimizationMethod::GAUSS_NEWTON;
    if (upper == "LM" || upper == "LEVENBERG_MARQUARDT") return OptimizationMethod::LEVENBERG_MARQUARDT;
    if (upper == "DOGLEG" || upper == "POWELL_DOGLEG") return OptimizationMethod::POWELL_DOGLEG;
    throw std::runtime_error("[OptimizationMethod_from_string] Invalid OptimizationMethod str [" + str + "]");
}

struct RegistrationConvergenceCriteria {
    float translation = 1e-3f;  // [m]
    float rotation = 1e-3f;     // [rad]
};

enum class DegenerateRegularizationType { none = 0, nl_reg };  // degenerate_regularization.hpp:14-17

inline DegenerateRegularizationType DegenerateRegularizationType_from_string(const std::string& str) {  // :19-33
    std::string upper(str.size(), '\0');
    std::transform(str.begin(), str.end(), upper.begin(), [](unsigned char c) { return std::toupper(c); });
    if (upper == "NONE") return DegenerateRegularizationType::none;
    if (upper == "NL-REG" || upper == "NL_REG") return DegenerateRegularizationType::nl_reg;
    throw std::runtime_error("[DegenerateRegularizationType_from_string] Invalid DegenerateRegularizationType str [" + str +
                             "]");
}

struct DegenerateRegularizationParams {  // degenerate_regularization.hpp:35-40
    DegenerateRegularizationType type = DegenerateRegularizationType::none;
    float rot_eigenvalue_threshold = 10.0f;
    float trans_eigenvalue_threshold = 1.0f;
    float base_factor = 1.0f;
};

struct MapPriorParams {  // map_prior.hpp:15-21
    bool enabled = false;
    float rot_vel_sigma = 1.0f;
    float trans_vel_sigma = 1.0f;
    float rot_base_sigma = 3.16e-2f;
    float trans_base_sigma = 1e-2f;
};

struct RegistrationFactorParams {
    struct Robust {
        robust::RobustLossType type = robust::RobustLossType::NONE;
        float default_scale = 10.0f;
    };
    struct GenZ {
        float planarity_threshold = 0.2f;
    };
    struct RotationConstraint {
        struct Robust {
            float default_scale = 10.0f;
        };
        bool enable = false;
        float weight = 1.0f;
        Robust robust;
    };

    RegType reg_type = RegType::GICP;
    float max_correspondence_distance = 2.0f;
    Robust robust;
    RotationConstraint rotation_constraint;
    GenZ genz;
    bool verbose = false;
};

struct RegistrationOptimizationParams {
    struct GaussNewton {
        float lambda = 1.0f;
    };
    struct LevenbergMarquardt {
        size_t max_inner_iterations = 10;
        float lambda_factor = 2.0f;
        float init_lambda = 1.0f;
        float max_lambda = 1e3f;
        float min_lambda = 1e-6f;
    };
    struct Dogleg {
        float initial_trust_region_radius = 1.0f;
        float min_trust_region_radius = 1e-4f;
        float max_trust_region_radius = 10.0f;
        float eta1 = 0.25f;
        float eta2 = 0.75f;
        float gamma_decrease = 0.25f;
        float gamma_increase = 2.0f;
    };

    GaussNewton gn;
    LevenbergMarquardt lm;
    Dogleg dogleg;
    OptimizationMethod optimization_method = OptimizationMethod::GAUSS_NEWTON;
};

struct RegistrationParams : public RegistrationFactorParams, public RegistrationOptimizationParams {
    using Criteria = RegistrationConvergenceCriteria;

    RegistrationParams() = default;
    explicit RegistrationParams(const RegistrationFactorParams& factor_params,
                                const RegistrationOptimizationParams& optimization_params = {})
        : RegistrationFactorParams(factor_params), RegistrationOptimizationParams(optimization_params) {}

    size_t max_iterations = 20;
    Criteria criteria;
    DegenerateRegularizationParams degenerate_reg;
    MapPriorParams map_prior;
};

}  // namespace registration
}  // namespace algorithms
}  // namespace sycl_points
