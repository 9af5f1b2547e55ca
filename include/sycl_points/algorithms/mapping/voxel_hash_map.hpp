// mapping::VoxelHashMap — I/algorithms/mapping/voxel_hash_map.hpp:22-1066: the odometry submap.  The table, its
// hashing, rehash / staleness rules and the export live in libspx (spx_voxelmap_*, csrc/spx_voxelmap.cu); this
// class keeps the reference's interface and container handling.
#pragma once

#include <memory>
#include <stdexcept>

#include "sycl_points/points/point_cloud.hpp"

namespace sycl_points {
namespace algorithms {
namespace mapping {

class VoxelHashMap {
public:
    using Ptr = std::shared_ptr<VoxelHashMap>;

    /// :29-36
    VoxelHashMap(const sycl_utils::DeviceQueue& queue, const float voxel_size) : queue_(queue) {
        if (voxel_size <= 0.0f) throw std::invalid_argument("voxel_size must be positive.");
        this->voxel_size_ = voxel_size;
        spx_voxelmap_t h = nullptr;
        detail::spx_check(spx_voxelmap_create(queue.handle(), voxel_size, &h));
        this->map_ = std::shared_ptr<spx_voxelmap_s>(h, [](spx_voxelmap_t p) { spx_voxelmap_destroy(p); });
    }

    /// :40-47
    void set_voxel_size(const float voxel_size) {
        if (voxel_size <= 0.0f) throw std::invalid_argument("voxel_size must be positive.");
        this->voxel_size_ = voxel_size;
        this->push();
    }
    float get_voxel_size() const { return this->voxel_size_; }
    void set_max_staleness(const uint32_t max_staleness) {
        this->max_staleness_ = max_staleness;
        this->push();
    }
    uint32_t get_max_staleness() const { return this->max_staleness_; }
    void set_remove_old_data_cycle(const uint32_t remove_old_data_cycle) {
        this->remove_old_data_cycle_ = remove_old_data_cycle;
        this->push();
    }
    uint32_t get_remove_old_data_cycle() const { return this->remove_old_data_cycle_; }
    void set_rehash_threshold(const float rehash_threshold) {
        this->rehash_threshold_ = rehash_threshold;
        this->push();
    }
    float get_rehash_threshold() const { return this->rehash_threshold_; }
    void set_min_num_point(const uint32_t min_num_point) {
        this->min_num_point_ = min_num_point;
        this->push();
    }
    uint32_t get_min_num_point() const { return this->min_num_point_; }

    /// :83-112
    void clear() { detail::spx_check(spx_voxelmap_clear(this->map_.get())); }

    /// :117-140 — `cloud` in the sensor frame, `sensor_pose` maps it into the map frame
    void add_point_cloud(const PointCloudShared& cloud, const Eigen::Isometry3f& sensor_pose) {
        const size_t N = cloud.size();
        const auto& q = this->queue_;
        if (N > 0) {
            q.set_accessed_by_device(cloud.points_ptr(), N);
            if (cloud.has_cov()) q.set_accessed_by_device(cloud.covs_ptr(), N);
            if (cloud.has_rgb()) q.set_accessed_by_device(cloud.rgb_ptr(), N);
            if (cloud.has_intensity()) q.set_accessed_by_device(cloud.intensities_ptr(), N);
        }
        detail::spx_check(spx_voxelmap_add(
            this->map_.get(), N ? reinterpret_cast<const float*>(cloud.points_ptr()) : nullptr,
            N && cloud.has_cov() ? reinterpret_cast<const float*>(cloud.covs_ptr()) : nullptr,
            N && cloud.has_rgb() ? reinterpret_cast<const float*>(cloud.rgb_ptr()) : nullptr,
            N && cloud.has_intensity() ? cloud.intensities_ptr() : nullptr, N, sensor_pose.matrix().data()));
    }

    /// :146-188
    void downsampling(PointCloudShared& result, const Eigen::Vector3f& center, const float distance = 100.0f) {
        uint64_t voxel_num = 0;
        int has_cov = 0, has_rgb = 0, has_intensity = 0;
        detail::spx_check(spx_voxelmap_info(this->map_.get(), nullptr, &voxel_num, nullptr, &has_cov, &has_rgb,
                                            &has_intensity));
        if (voxel_num == 0) {
            result.clear();
            return;
        }
        const size_t n = voxel_num;
        const auto& q = this->queue_;
        result.resize_points(n);
        result.resize_covs(has_cov ? n : 0);
        result.resize_rgb(has_rgb ? n : 0);
        result.resize_intensities(has_intensity ? n : 0);
        result.index_hint = PointCloudShared::IndexHint{};
        q.set_accessed_by_device(result.points_ptr(), n);
        if (has_cov) q.set_accessed_by_device(result.covs_ptr(), n);
        if (has_rgb) q.set_accessed_by_device(result.rgb_ptr(), n);
        if (has_intensity) q.set_accessed_by_device(result.intensities_ptr(), n);
        const float c[3] = {center.x(), center.y(), center.z()};
        size_t m = 0;
        detail::spx_check(spx_voxelmap_downsample(
            this->map_.get(), c, distance, reinterpret_cast<float*>(result.points_ptr()),
            has_cov ? reinterpret_cast<float*>(result.covs_ptr()) : nullptr,
            has_rgb ? reinterpret_cast<float*>(result.rgb_ptr()) : nullptr,
            has_intensity ? result.intensities_ptr() : nullptr, nullptr, n, &m));
        result.resize_points(m);
        result.resize_covs(has_cov ? m : 0);
        result.resize_rgb(has_rgb ? m : 0);
        result.resize_intensities(has_intensity ? m : 0);
    }

    /// :194-246
    float compute_overlap_ratio(const PointCloudShared& cloud, const Eigen::Isometry3f& sensor_pose) const {
        if (!cloud.points || cloud.points->empty()) return 0.0f;
        const size_t N = cloud.size();
        this->queue_.set_accessed_by_device(cloud.points_ptr(), N);
        float ratio = 0.0f;
        detail::spx_check(spx_voxelmap_overlap_ratio(this->map_.get(), reinterpret_cast<const float*>(cloud.points_ptr()), N,
                                                     sensor_pose.matrix().data(), &ratio));
        return ratio;
    }

    /// :248
    void remove_old_data() { detail::spx_check(spx_voxelmap_remove_old(this->map_.get())); }

    /// number of occupied voxels / slots of the table (not in the reference's public interface; diagnostics)
    size_t size() const {
        uint64_t v = 0;
        detail::spx_check(spx_voxelmap_info(this->map_.get(), nullptr, &v, nullptr, nullptr, nullptr, nullptr));
        return v;
    }
    size_t capacity() const {
        uint64_t c = 0;
        detail::spx_check(spx_voxelmap_info(this->map_.get(), &c, nullptr, nullptr, nullptr, nullptr, nullptr));
        return c;
    }

private:
    void push() {
        detail::spx_check(spx_voxelmap_set_params(this->map_.get(), this->voxel_size_, this->max_staleness_,
                                                  this->remove_old_data_cycle_, this->rehash_threshold_,
                                                  this->min_num_point_));
    }

    sycl_utils::DeviceQueue queue_;
    std::shared_ptr<spx_voxelmap_s> map_;
    float voxel_size_ = 0.0f;
    uint32_t max_staleness_ = 100;
    uint32_t remove_old_data_cycle_ = 10;
    float rehash_threshold_ = 0.7f;
    uint32_t min_num_point_ = 1U;
};

}  // namespace mapping
}  // namespace algorithms
}  // namespace sycl_points
