// algorithms::CoordinateSystem — I/algorithms/common/coordinate_system.hpp:13-27 (REP-103 frames)
#pragma once

#include <algorithm>
#include <cctype>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace sycl_points {
namespace algorithms {

enum class CoordinateSystem : std::uint8_t { LIDAR = 0, CAMERA = 1 };

inline CoordinateSystem coordinate_system_from_string(const std::string& str) {
    std::string upper = str;
    std::transform(str.begin(), str.end(), upper.begin(), [](unsigned char c) { return std::toupper(c); });
    if (upper == "LIDAR") return CoordinateSystem::LIDAR;
    if (upper == "CAMERA") return CoordinateSystem::CAMERA;
    throw std::invalid_argument("Invalid coordinate system: " + str);
}

}  // namespace algorithms
}  // namespace sycl_points
