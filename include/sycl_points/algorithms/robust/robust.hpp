// robust::RobustLossType and its string parser — I/algorithms/robust/robust.hpp:14-48.  The weight
// and rho functions themselves (:56-114) run inside libspx's kernels.
#pragma once

#include <algorithm>
#include <cctype>
#include <stdexcept>
#include <string>

namespace sycl_points {
namespace algorithms {
namespace robust {

enum class RobustLossType { NONE, HUBER, TUKEY, CAUCHY, GEMAN_MCCLURE };

inline RobustLossType RobustLossType_from_string(const std::string& str) {
    std::string upper(str.size(), '\0');
    std::transform(str.begin(), str.end(), upper.begin(), [](unsigned char c) { return std::toupper(c); });
    static const std::pair<const char*, RobustLossType> table[] = {
        {"NONE", RobustLossType::NONE},     {"HUBER", RobustLossType::HUBER},
        {"TUKEY", RobustLossType::TUKEY},   {"CAUCHY", RobustLossType::CAUCHY},
        {"GEMAN_MCCLURE", RobustLossType::GEMAN_MCCLURE}};
    for (const auto& [name, value] : table)
        if (upper == name) return value;
    throw std::runtime_error("[RobustLossType_from_string] Invalid RobustLossType str '" + str + "'");
}

}  // namespace robust
}  // namespace algorithms
}  // namespace sycl_points
