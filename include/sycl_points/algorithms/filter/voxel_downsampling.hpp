// filter::VoxelGrid — I/algorithms/filter/voxel_downsampling.hpp:14-79.  In the reference only the
// voxel key runs on the device and the sort + per-voxel mean are host code; here the whole
// operation is spx_voxel_downsample(_attrs): device radix sort by (key, index), fp32 mean in that
// order, ascending-key output; the cloud overload also aggregates mean RGB, median intensity and
// mean timestamp offset per voxel (:220-288).  Covariances / normals do not survive (as in the
// reference).
#pragma once

#include <memory>
#include <stdexcept>

#include "sycl_points/points/point_cloud.hpp"

namespace sycl_points {
namespace algorithms {
namespace filter {

class VoxelGrid {
public:
    using Ptr = std::shared_ptr<VoxelGrid>;

    /// :21-29
    VoxelGrid(const sycl_utils::DeviceQueue& queue, const float voxel_size) : queue_(queue), voxel_size_(voxel_size) {
        if (voxel_size <= 0.0f) throw std::invalid_argument("voxel_size must be positive");
    }
    /// :33-40
    void set_voxel_size(const float voxel_size) {
        if (voxel_size <= 0.0f) throw std::invalid_argument("voxel_size must be positive");
        this->voxel_size_ = voxel_size;
    }
    float get_voxel_size() const { return this->voxel_size_; }
    void set_min_voxel_count(const size_t min_voxel_count) { this->min_voxel_count_ = min_voxel_count; }
    size_t get_min_voxel_count() const { return this->min_voxel_count_; }

    /// :50-62
    void downsampling(const PointContainerShared& points, PointContainerShared& result) {
        const size_t N = points.size();
        if (N == 0) {
            result.clear();
            return;
        }
        PointContainerShared out(N);  // a distinct buffer: `result` may alias `points`
        this->queue_.set_accessed_by_device(points.data(), N);
        this->queue_.set_accessed_by_device(out.data(), N);
        size_t m = 0;
        detail::spx_check(spx_voxel_downsample(this->queue_.handle(), reinterpret_cast<const float*>(points.data()), N,
                                               this->voxel_size_, this->min_voxel_count_,
                                               reinterpret_cast<float*>(out.data()), &m));
        out.resize(m);
        result.swap(out);
    }

    /// :64-79, :220-288
    void downsampling(const PointCloudShared& cloud, PointCloudShared& result) {
        const size_t N = cloud.size();
        if (N == 0) {
            result.clear();
            return;
        }
        const bool has_rgb = cloud.has_rgb(), has_intensity = cloud.has_intensity(), has_ts = cloud.has_timestamps();
        const auto& q = this->queue_;
        PointContainerShared out_points(N);
        RGBContainerShared out_rgb(has_rgb ? N : 0);
        IntensityContainerShared out_intensity(has_intensity ? N : 0);
        TimestampContainerShared out_ts(has_ts ? N : 0);
        q.set_accessed_by_device(cloud.points_ptr(), N);
        q.set_accessed_by_device(out_points.data(), N);
        if (has_rgb) q.set_accessed_by_device(cloud.rgb_ptr(), N), q.set_accessed_by_device(out_rgb.data(), N);
        if (has_intensity)
            q.set_accessed_by_device(cloud.intensities_ptr(), N), q.set_accessed_by_device(out_intensity.data(), N);
        if (has_ts) q.set_accessed_by_device(cloud.timestamp_offsets_ptr(), N), q.set_accessed_by_device(out_ts.data(), N);
        size_t m = 0;
        detail::spx_check(spx_voxel_downsample_attrs(
            q.handle(), reinterpret_cast<const float*>(cloud.points_ptr()), N, this->voxel_size_, this->min_voxel_count_,
            has_rgb ? reinterpret_cast<const float*>(cloud.rgb_ptr()) : nullptr,
            has_intensity ? cloud.intensities_ptr() : nullptr, has_ts ? cloud.timestamp_offsets_ptr() : nullptr,
            reinterpret_cast<float*>(out_points.data()), has_rgb ? reinterpret_cast<float*>(out_rgb.data()) : nullptr,
            has_intensity ? out_intensity.data() : nullptr, has_ts ? out_ts.data() : nullptr, &m));
        out_points.resize(m);
        out_rgb.resize(has_rgb ? m : 0);
        out_intensity.resize(has_intensity ? m : 0);
        out_ts.resize(has_ts ? m : 0);
        const double t0 = cloud.start_time_ms, t1 = cloud.end_time_ms;
        result.points->swap(out_points);  // `result` may be `cloud` itself: swap only after the kernels ran
        result.rgb->swap(out_rgb);
        result.intensities->swap(out_intensity);
        result.timestamp_offsets->swap(out_ts);
        result.covs->clear();
        result.normals->clear();
        result.start_time_ms = t0;
        result.end_time_ms = t1;
        // the voxel box of the output: KDTree::build(queue, result) needs no measuring pass (1.85 / 2.4 voxels per
        // cell hold ~3 / ~5 points per occupied cell on LiDAR surfaces)
        result.index_hint = PointCloudShared::IndexHint{};
        if (m > 0 && spx_voxel_last_box(q.handle(), result.index_hint.lo, result.index_hint.hi, nullptr) == SPX_OK) {
            result.index_hint.data = result.points->data();
            result.index_hint.n = m;
            result.index_hint.cell = 1.85f * this->voxel_size_;
            result.index_hint.knn_cell = 2.4f * this->voxel_size_;
        }
    }

private:
    sycl_utils::DeviceQueue queue_;
    float voxel_size_;
    size_t min_voxel_count_ = 1;
};

}  // namespace filter
}  // namespace algorithms
}  // namespace sycl_points
