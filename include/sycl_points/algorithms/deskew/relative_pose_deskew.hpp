// deskew::deskew_point_cloud_constant_velocity — I/algorithms/deskew/relative_pose_deskew.hpp:36-178: one device
// pass (spx_deskew_constant_velocity) that moves every point by se3_exp(tau * se3_log(prev^-1 * cur)), tau the
// point's timestamp offset over the scan duration clamped to [0, 1], and rotates normals / covariances with it.
#pragma once

#include <stdexcept>

#include "sycl_points/points/point_cloud.hpp"
#include "sycl_points/utils/eigen_utils.hpp"
#include "sycl_points/utils/sycl_utils.hpp"

namespace sycl_points {
namespace algorithms {
namespace deskew {

/// @return true when the cloud was deskewed; false when it is empty, has no timestamps or the duration is not
/// positive (relative_pose_deskew.hpp:50-61).  `output_cloud` may be `input_cloud` (in place).
inline bool deskew_point_cloud_constant_velocity(const PointCloudShared& input_cloud, PointCloudShared& output_cloud,
                                                 const Eigen::Isometry3f& previous_relative_pose,
                                                 const Eigen::Isometry3f& current_relative_pose,
                                                 float inter_scan_duration_seconds = -1.0f) {
    if (!input_cloud.queue.ptr || !output_cloud.queue.ptr)
        throw std::runtime_error("[deskew_point_cloud_constant_velocity] SYCL queue is not initialized");
    if (input_cloud.queue.get_device().ordinal() != output_cloud.queue.get_device().ordinal())
        throw std::runtime_error(
            "[deskew_point_cloud_constant_velocity] input_cloud and output_cloud must share the same SYCL context");

    const size_t N = input_cloud.size();
    if (N == 0 || !input_cloud.has_timestamps()) return false;
    const float duration = inter_scan_duration_seconds > 0.0f
                               ? inter_scan_duration_seconds
                               : static_cast<float>((input_cloud.end_time_ms - input_cloud.start_time_ms) * 1e-3);
    if (duration <= 0.0f) return false;

    if (&input_cloud != &output_cloud) {  // :63-97: everything that is not deskewed is carried over
        output_cloud.start_time_ms = input_cloud.start_time_ms;
        output_cloud.end_time_ms = input_cloud.end_time_ms;
        output_cloud.timestamp_offsets->assign(input_cloud.timestamp_offsets->begin(),
                                               input_cloud.timestamp_offsets->end());
        output_cloud.points->resize(N);
        if (input_cloud.has_normal()) output_cloud.normals->resize(N);
        else output_cloud.normals->clear();
        if (input_cloud.has_cov()) output_cloud.covs->resize(N);
        else output_cloud.covs->clear();
        if (input_cloud.has_rgb()) output_cloud.rgb->assign(input_cloud.rgb->begin(), input_cloud.rgb->end());
        else output_cloud.rgb->clear();
        if (input_cloud.has_intensity())
            output_cloud.intensities->assign(input_cloud.intensities->begin(), input_cloud.intensities->end());
        else output_cloud.intensities->clear();
        output_cloud.index_hint = PointCloudShared::IndexHint();
    }

    const Eigen::Vector<float, 6> twist = eigen_utils::lie::se3_log(previous_relative_pose.inverse() * current_relative_pose);
    const auto& q = input_cloud.queue;
    const bool nrm = input_cloud.has_normal(), cov = input_cloud.has_cov();
    q.set_accessed_by_device(input_cloud.points_ptr(), N);
    q.set_accessed_by_device(output_cloud.points_ptr(), N);
    q.set_accessed_by_device(output_cloud.timestamp_offsets_ptr(), N);
    detail::spx_check(spx_deskew_constant_velocity(
        q.handle(), reinterpret_cast<const float*>(input_cloud.points_ptr()),
        nrm ? reinterpret_cast<const float*>(input_cloud.normals_ptr()) : nullptr,
        cov ? reinterpret_cast<const float*>(input_cloud.covs_ptr()) : nullptr, output_cloud.timestamp_offsets_ptr(), N,
        twist.data(), duration, reinterpret_cast<float*>(output_cloud.points_ptr()),
        nrm ? reinterpret_cast<float*>(output_cloud.normals_ptr()) : nullptr,
        cov ? reinterpret_cast<float*>(output_cloud.covs_ptr()) : nullptr));
    detail::spx_check(spx_queue_sync(q.handle()));  // :176 wait_and_throw
    return true;
}

}  // namespace deskew
}  // namespace algorithms
}  // namespace sycl_points
