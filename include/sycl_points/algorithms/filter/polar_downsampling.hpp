// filter::PolarGrid — I/algorithms/filter/polar_downsampling.hpp:111-452: grid filter with cells in (range,
// elevation, azimuth) around the sensor; centroid per cell, mean RGB, median intensity, mean timestamp offset,
// cells with fewer than min_voxel_count points dropped, ascending-key output.  One device pass set
// (spx_polar_downsample_attrs) — the voxel grid's sort-and-aggregate with the polar key.
#pragma once

#include <memory>
#include <stdexcept>

#include "sycl_points/algorithms/common/coordinate_system.hpp"
#include "sycl_points/points/point_cloud.hpp"
#include "sycl_points/utils/sycl_utils.hpp"

namespace sycl_points {
namespace algorithms {
namespace filter {

class PolarGrid {
public:
    using Ptr = std::shared_ptr<PolarGrid>;

    /// :124-137 (angles in radians)
    PolarGrid(const sycl_utils::DeviceQueue& queue, float distance_voxel_size, float elevation_voxel_size,
              float azimuth_voxel_size, CoordinateSystem coord = CoordinateSystem::LIDAR)
        : queue_(queue),
          distance_voxel_size_(distance_voxel_size),
          elevation_voxel_size_(elevation_voxel_size),
          azimuth_voxel_size_(azimuth_voxel_size),
          coord_(coord) {
        if (distance_voxel_size <= 0.0f || elevation_voxel_size <= 0.0f || azimuth_voxel_size <= 0.0f)
            throw std::invalid_argument("voxel sizes must be positive");
    }

    void set_distance_voxel_size(const float size) {
        if (size <= 0.0f) throw std::invalid_argument("distance_voxel_size must be positive");
        this->distance_voxel_size_ = size;
    }
    float get_distance_voxel_size() const { return this->distance_voxel_size_; }
    void set_elevation_voxel_size(const float size) {
        if (size <= 0.0f) throw std::invalid_argument("elevation_voxel_size must be positive");
        this->elevation_voxel_size_ = size;
    }
    float get_elevation_voxel_size() const { return this->elevation_voxel_size_; }
    void set_azimuth_voxel_size(const float size) {
        if (size <= 0.0f) throw std::invalid_argument("azimuth_voxel_size must be positive");
        this->azimuth_voxel_size_ = size;
    }
    float get_azimuth_voxel_size() const { return this->azimuth_voxel_size_; }
    void set_min_voxel_count(const size_t min_voxel_count) { this->min_voxel_count_ = min_voxel_count; }
    size_t get_min_voxel_count() const { return this->min_voxel_count_; }
    void set_coordinate_system(const CoordinateSystem coord) { this->coord_ = coord; }
    CoordinateSystem get_coordinate_system() const { return this->coord_; }

    /// :194-205
    void downsampling(const PointContainerShared& points, PointContainerShared& result) {
        const size_t N = points.size();
        if (N == 0) {
            result.resize(0);
            return;
        }
        PointContainerShared out(N);  // `result` may alias `points`
        this->queue_.set_accessed_by_device(points.data(), N);
        this->queue_.set_accessed_by_device(out.data(), N);
        size_t m = 0;
        detail::spx_check(this->run(reinterpret_cast<const float*>(points.data()), N, nullptr, nullptr, nullptr,
                                    reinterpret_cast<float*>(out.data()), nullptr, nullptr, nullptr, &m));
        out.resize(m);
        result.swap(out);
    }

    /// :211-223, :375-452
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
        detail::spx_check(this->run(reinterpret_cast<const float*>(cloud.points_ptr()), N,
                                    has_rgb ? reinterpret_cast<const float*>(cloud.rgb_ptr()) : nullptr,
                                    has_intensity ? cloud.intensities_ptr() : nullptr,
                                    has_ts ? cloud.timestamp_offsets_ptr() : nullptr,
                                    reinterpret_cast<float*>(out_points.data()),
                                    has_rgb ? reinterpret_cast<float*>(out_rgb.data()) : nullptr,
                                    has_intensity ? out_intensity.data() : nullptr, has_ts ? out_ts.data() : nullptr, &m));
        out_points.resize(m);
        out_rgb.resize(has_rgb ? m : 0);
        out_intensity.resize(has_intensity ? m : 0);
        out_ts.resize(has_ts ? m : 0);
        const double t0 = cloud.start_time_ms, t1 = cloud.end_time_ms;
        result.points->swap(out_points);  // `result` may be `cloud` itself
        result.rgb->swap(out_rgb);
        result.intensities->swap(out_intensity);
        result.timestamp_offsets->swap(out_ts);
        result.covs->clear();
        result.normals->clear();
        result.start_time_ms = t0;
        result.end_time_ms = t1;
        result.index_hint = PointCloudShared::IndexHint{};
    }

private:
    int run(const float* pts, size_t N, const float* rgb, const float* intensity, const float* ts, float* out,
            float* out_rgb, float* out_intensity, float* out_ts, size_t* m) const {
        return spx_polar_downsample_attrs(this->queue_.handle(), pts, N, this->distance_voxel_size_,
                                          this->elevation_voxel_size_, this->azimuth_voxel_size_,
                                          this->coord_ == CoordinateSystem::LIDAR ? SPX_COORD_LIDAR : SPX_COORD_CAMERA,
                                          this->min_voxel_count_, rgb, intensity, ts, out, out_rgb, out_intensity, out_ts, m);
    }

    sycl_utils::DeviceQueue queue_;
    float distance_voxel_size_;
    float elevation_voxel_size_;
    float azimuth_voxel_size_;
    CoordinateSystem coord_;
    size_t min_voxel_count_ = 1;
};

}  // namespace filter
}  // namespace algorithms
}  // namespace sycl_points
