// PointCloudCPU / PointCloudShared — I/points/point_cloud.hpp:12-476.  Same members and methods;
// the shared containers live in CUDA managed memory (see utils/sycl_utils.hpp).
#pragma once

#include <algorithm>
#include <cstring>
#include <memory>
#include <stdexcept>

#include "sycl_points/points/types.hpp"

namespace sycl_points {

/// point_cloud.hpp:12-70
struct PointCloudCPU {
    using Ptr = std::shared_ptr<PointCloudCPU>;
    using ConstPtr = std::shared_ptr<const PointCloudCPU>;

    std::shared_ptr<PointContainerCPU> points = nullptr;
    std::shared_ptr<CovarianceContainerCPU> covs = nullptr;
    std::shared_ptr<NormalContainerCPU> normals = nullptr;
    std::shared_ptr<RGBContainerCPU> rgb = nullptr;
    std::shared_ptr<IntensityContainerCPU> intensities = nullptr;
    std::shared_ptr<TimestampContainerCPU> timestamp_offsets = nullptr;
    double start_time_ms = 0.0;
    double end_time_ms = 0.0;

    PointCloudCPU() {
        this->points = std::make_shared<PointContainerCPU>();
        this->covs = std::make_shared<CovarianceContainerCPU>();
        this->normals = std::make_shared<NormalContainerCPU>();
        this->rgb = std::make_shared<RGBContainerCPU>();
        this->intensities = std::make_shared<IntensityContainerCPU>();
        this->timestamp_offsets = std::make_shared<TimestampContainerCPU>();
    }
    size_t size() const { return this->points->size(); }
    bool has_cov() const { return this->covs != nullptr && this->covs->size() == this->points->size(); }
    bool has_normal() const { return this->normals != nullptr && this->normals->size() == this->points->size(); }
    bool has_rgb() const { return this->rgb != nullptr && this->rgb->size() == this->points->size(); }
    bool has_intensity() const {
        return this->intensities != nullptr && this->intensities->size() == this->points->size();
    }
    bool has_timestamps() const {
        return this->timestamp_offsets != nullptr && this->timestamp_offsets->size() == this->points->size();
    }

    /// point_cloud.hpp:61-69
    void update_end_time() {
        if (this->timestamp_offsets && !this->timestamp_offsets->empty()) {
            const auto max_offset = *std::max_element(this->timestamp_offsets->begin(), this->timestamp_offsets->end());
            this->end_time_ms = this->start_time_ms + static_cast<double>(max_offset);
        } else {
            this->end_time_ms = this->start_time_ms;
        }
    }
};

/// point_cloud.hpp:73-476
struct PointCloudShared {
    using Ptr = std::shared_ptr<PointCloudShared>;
    using ConstPtr = std::shared_ptr<const PointCloudShared>;

    sycl_utils::DeviceQueue queue;
    std::shared_ptr<PointContainerShared> points = nullptr;
    std::shared_ptr<CovarianceContainerShared> covs = nullptr;
    std::shared_ptr<NormalContainerShared> normals = nullptr;
    std::shared_ptr<RGBContainerShared> rgb = nullptr;
    std::shared_ptr<IntensityContainerShared> intensities = nullptr;
    std::shared_ptr<TimestampContainerShared> timestamp_offsets = nullptr;
    double start_time_ms = 0.0;
    double end_time_ms = 0.0;

    /// spx extension (not in the reference): a cloud that just came out of VoxelGrid::downsampling carries the box
    /// and cell edges KDTree::build needs, so the build skips its measuring pass and host round trip
    /// (spx_index_build_hinted).  Valid only while `data` / `n` still describe *points; transform clears it.  A stale
    /// hint can cost speed, never exactness.
    struct IndexHint {
        const void* data = nullptr;
        size_t n = 0;
        float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        float cell = 0.0f, knn_cell = 0.0f;
        bool matches(const void* d, size_t m) const { return data != nullptr && data == d && n == m && m > 0; }
    };
    mutable IndexHint index_hint;

    PointCloudShared(const sycl_utils::DeviceQueue& q) : queue(q) { this->make_containers(); }

    /// copy a host cloud in (point_cloud.hpp:110-198) and start moving it to the device
    PointCloudShared(const sycl_utils::DeviceQueue& q, const PointCloudCPU& cpu) : queue(q) {
        this->make_containers();
        const size_t N = cpu.size();
        copy_in(*this->points, cpu.points.get(), N, N > 0);
        copy_in(*this->covs, cpu.covs.get(), N, cpu.has_cov());
        copy_in(*this->normals, cpu.normals.get(), N, cpu.has_normal());
        copy_in(*this->rgb, cpu.rgb.get(), N, cpu.has_rgb());
        copy_in(*this->intensities, cpu.intensities.get(), N, cpu.has_intensity());
        copy_in(*this->timestamp_offsets, cpu.timestamp_offsets.get(), N, cpu.has_timestamps());
        this->start_time_ms = cpu.start_time_ms;
        this->end_time_ms = cpu.end_time_ms;
        this->queue.set_accessed_by_device(this->points->data(), N);
    }
    PointCloudShared(const PointCloudShared& other) : PointCloudShared(other.queue, other) {}
    PointCloudShared(const sycl_utils::DeviceQueue& target_queue, const PointCloudShared& other) : queue(target_queue) {
        if (!this->queue.ptr) throw std::runtime_error("[PointCloudShared] target queue is not initialized");
        if (!other.queue.ptr) throw std::runtime_error("[PointCloudShared] source queue is not initialized");
        if (!other.points) throw std::runtime_error("[PointCloudShared] source points are not initialized");
        this->make_containers();
        other.queue.ptr->wait();  // the source may still be written by kernels on its own stream
        const size_t N = other.size();
        copy_shared(*this->points, *other.points, N, true);
        copy_shared(*this->covs, *other.covs, N, other.has_cov());
        copy_shared(*this->normals, *other.normals, N, other.has_normal());
        copy_shared(*this->rgb, *other.rgb, N, other.has_rgb());
        copy_shared(*this->intensities, *other.intensities, N, other.has_intensity());
        copy_shared(*this->timestamp_offsets, *other.timestamp_offsets, N, other.has_timestamps());
        this->start_time_ms = other.start_time_ms;
        this->end_time_ms = other.end_time_ms;
    }
    PointCloudShared& operator=(const PointCloudShared& other) {
        if (this == &other) return *this;
        PointCloudShared tmp(this->queue, other);
        this->swap_containers(tmp);
        return *this;
    }

    size_t size() const { return this->points->size(); }
    bool has_cov() const { return this->covs->size() > 0 && this->covs->size() == this->points->size(); }
    bool has_normal() const { return this->normals->size() > 0 && this->normals->size() == this->points->size(); }
    bool has_rgb() const { return this->rgb->size() > 0 && this->rgb->size() == this->points->size(); }
    bool has_intensity() const {
        return this->intensities->size() > 0 && this->intensities->size() == this->points->size();
    }
    bool has_timestamps() const {
        return this->timestamp_offsets->size() > 0 && this->timestamp_offsets->size() == this->points->size();
    }

    PointType* points_ptr() const { return this->points->data(); }
    Covariance* covs_ptr() const { return this->covs->data(); }
    Normal* normals_ptr() const { return this->normals->data(); }
    RGBType* rgb_ptr() const { return this->rgb->data(); }
    float* intensities_ptr() const { return this->intensities->data(); }
    TimestampOffset* timestamp_offsets_ptr() const { return this->timestamp_offsets->data(); }

    void resize_points(size_t N) const { this->points->resize(N); }
    void resize_covs(size_t N) const { this->covs->resize(N); }
    void resize_normals(size_t N) const { this->normals->resize(N); }
    void resize_rgb(size_t N) const { this->rgb->resize(N); }
    void resize_intensities(size_t N) const { this->intensities->resize(N); }
    void resize_timestamps(size_t N) const { this->timestamp_offsets->resize(N); }
    void reserve_points(size_t N) const { this->points->reserve(N); }
    void reserve_covs(size_t N) const { this->covs->reserve(N); }
    void reserve_normals(size_t N) const { this->normals->reserve(N); }
    void reserve_rgb(size_t N) const { this->rgb->reserve(N); }
    void reserve_intensities(size_t N) const { this->intensities->reserve(N); }
    void reserve_timestamps(size_t N) const { this->timestamp_offsets->reserve(N); }

    void clear() {
        this->points->clear();
        this->covs->clear();
        this->normals->clear();
        this->rgb->clear();
        this->intensities->clear();
        this->timestamp_offsets->clear();
    }

    /// point_cloud.hpp:319-338: attributes survive only if both clouds carry them
    void extend(const PointCloudShared& other) {
        this->queue.ptr->wait();
        other.queue.ptr->wait();
        const bool cov = this->has_cov() && other.has_cov(), nrm = this->has_normal() && other.has_normal();
        const bool col = this->has_rgb() && other.has_rgb(), inten = this->has_intensity() && other.has_intensity();
        this->points->insert(this->points->end(), other.points->begin(), other.points->end());
        append_or_clear(*this->covs, *other.covs, cov);
        append_or_clear(*this->normals, *other.normals, nrm);
        append_or_clear(*this->rgb, *other.rgb, col);
        append_or_clear(*this->intensities, *other.intensities, inten);
        this->timestamp_offsets->clear();
    }
    void operator+=(const PointCloudShared& pc) { this->extend(pc); }

    /// point_cloud.hpp:340-366
    void erase(size_t start_idx, size_t end_idx) {
        this->queue.ptr->wait();
        if (this->has_cov()) this->covs->erase(this->covs->begin() + start_idx, this->covs->begin() + end_idx);
        if (this->has_normal())
            this->normals->erase(this->normals->begin() + start_idx, this->normals->begin() + end_idx);
        if (this->has_rgb()) this->rgb->erase(this->rgb->begin() + start_idx, this->rgb->begin() + end_idx);
        if (this->has_intensity())
            this->intensities->erase(this->intensities->begin() + start_idx, this->intensities->begin() + end_idx);
        if (this->has_timestamps())
            this->timestamp_offsets->erase(this->timestamp_offsets->begin() + start_idx,
                                           this->timestamp_offsets->begin() + end_idx);
        this->points->erase(this->points->begin() + start_idx, this->points->begin() + end_idx);
    }

private:
    void make_containers() {
        this->points = std::make_shared<PointContainerShared>();
        this->covs = std::make_shared<CovarianceContainerShared>();
        this->normals = std::make_shared<NormalContainerShared>();
        this->rgb = std::make_shared<RGBContainerShared>();
        this->intensities = std::make_shared<IntensityContainerShared>();
        this->timestamp_offsets = std::make_shared<TimestampContainerShared>();
    }
    void swap_containers(PointCloudShared& o) {
        std::swap(this->points, o.points);
        std::swap(this->covs, o.covs);
        std::swap(this->normals, o.normals);
        std::swap(this->rgb, o.rgb);
        std::swap(this->intensities, o.intensities);
        std::swap(this->timestamp_offsets, o.timestamp_offsets);
        std::swap(this->start_time_ms, o.start_time_ms);
        std::swap(this->end_time_ms, o.end_time_ms);
    }
    // copies go through the stream (cudaMemcpy) so that device-resident pages are not faulted back
    // to the host just to be copied
    template <typename Dst, typename Src>
    void copy_in(Dst& dst, const Src* src, size_t N, bool enable) const {
        dst.resize(enable ? N : 0);
        if (enable && N > 0) {
            detail::spx_check(spx_memcpy_h2d(this->queue.handle(), dst.data(), src->data(),
                                             N * sizeof(typename Dst::value_type)));
            detail::spx_check(spx_queue_sync(this->queue.handle()));  // the host source may go away
        }
    }
    template <typename V>
    void copy_shared(V& dst, const V& src, size_t N, bool enable) const {
        dst.resize(enable ? N : 0);
        if (enable && N > 0) {
            detail::spx_check(spx_memcpy_d2d(this->queue.handle(), dst.data(), src.data(),
                                             N * sizeof(typename V::value_type)));
            detail::spx_check(spx_queue_sync(this->queue.handle()));
        }
    }
    template <typename V>
    static void append_or_clear(V& dst, const V& src, bool keep) {
        if (keep)
            dst.insert(dst.end(), src.begin(), src.end());
        else
            dst.clear();
    }
};

}  // namespace sycl_points
