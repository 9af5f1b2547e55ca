// registration::RegType and its string parser — I/algorithms/registration/factor.hpp:18-61.  The
// factors (:63-484) run inside libspx's kernels (point-to-point, point-to-plane, point-to-distribution,
// GICP and GenZ).
#pragma once

#include <algorithm>
#include <cctype>
#include <stdexcept>
#include <string>

namespace sycl_points {
namespace algorithms {
namespace registration {

enum class RegType { POINT_TO_POINT = 0, POINT_TO_PLANE, POINT_TO_DISTRIBUTION, GICP, GENZ };

inline RegType RegType_from_string(const std::string& str) {
    std::string upper(str.size(), '\0');
    std::transform(str.begin(), str.end(), upper.begin(), [](unsigned char c) { return std::toupper(c); });
    if (upper == "POINT_TO_POINT") return RegType::POINT_TO_POINT;
    if (upper == "POINT_TO_PLANE") return RegType::POINT_TO_PLANE;
    if (upper == "GICP") return RegType::GICP;
    if (upper == "GENZ") return RegType::GENZ;
    if (upper == "POINT_TO_DISTRIBUTION" || upper == "P2D") return RegType::POINT_TO_DISTRIBUTION;
    throw std::runtime_error("[RegType_from_string] Invalid RegType str '" + str + "'");
}

}  // namespace registration
}  // namespace algorithms
}  // namespace sycl_points
