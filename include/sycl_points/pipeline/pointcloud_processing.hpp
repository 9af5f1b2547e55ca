// pointcloud_processing::PCProcessor — I/pipeline/pointcloud_processing.hpp:29-205: box filter -> polar grid -> voxel grid
// -> random sampling, k-NN covariances (optionally M-estimated), angle-of-incidence refine filter.  Every per-point
// step runs in libspx; the intensity filters and the IMU deskew of the reference class are not built (refused loudly).
#pragma once

#include <memory>
#include <stdexcept>

#include "sycl_points/algorithms/common/coordinate_system.hpp"
#include "sycl_points/algorithms/feature/covariance.hpp"
#include "sycl_points/algorithms/filter/polar_downsampling.hpp"
#include "sycl_points/algorithms/filter/preprocess_filter.hpp"
#include "sycl_points/algorithms/filter/voxel_downsampling.hpp"
#include "sycl_points/algorithms/knn/kdtree.hpp"
#include "sycl_points/pipeline/odometry_common_params.hpp"
#include "sycl_points/points/point_cloud.hpp"

namespace sycl_points {
namespace pipeline {
namespace pointcloud_processing {

struct ProcessingContext {
    algorithms::knn::KDTree::Ptr tree;
    algorithms::knn::KNNResult knn_result;
};

class PCProcessor {
public:
    using Ptr = std::shared_ptr<PCProcessor>;
    using ConstPtr = std::shared_ptr<const PCProcessor>;

    PCProcessor(const sycl_utils::DeviceQueue& q, const odometry::CommonParameters::Scan& scan_params,
                const odometry::CommonParameters::CovarianceEstimation& covs_params,
                const odometry::CommonParameters::IMU& imu_params = {})
        : queue_(q), scan_params_(scan_params), covs_params_(covs_params) {
        if (imu_params.enable) throw std::runtime_error("[PCProcessor] the IMU paths are not built in libspx");
        this->preprocess_filter_ = std::make_shared<algorithms::filter::PreprocessFilter>(this->queue_);
        if (this->scan_params_.downsampling.voxel.enable)
            this->voxel_filter_ =
                std::make_shared<algorithms::filter::VoxelGrid>(this->queue_, this->scan_params_.downsampling.voxel.size);
        if (this->scan_params_.downsampling.polar.enable) {
            const auto cs = algorithms::coordinate_system_from_string(this->scan_params_.downsampling.polar.coord_system);
            const auto& p = this->scan_params_.downsampling.polar;
            this->polar_filter_ = std::make_shared<algorithms::filter::PolarGrid>(this->queue_, p.distance_size,
                                                                                  p.elevation_size, p.azimuth_size, cs);
        }
    }

    /// :130-156 — `src` is never modified
    void prefilter(const PointCloudShared& src, PointCloudShared& dst) const {
        const PointCloudShared* input = &src;
        if (this->scan_params_.preprocess.box_filter.enable) {
            this->preprocess_filter_->box_filter(src, dst, this->scan_params_.preprocess.box_filter.min,
                                                 this->scan_params_.preprocess.box_filter.max);
            input = &dst;
        }
        if (this->scan_params_.downsampling.polar.enable) {
            this->polar_filter_->downsampling(*input, dst);
            input = &dst;
        }
        if (this->scan_params_.downsampling.voxel.enable) {
            this->voxel_filter_->downsampling(*input, dst);
            input = &dst;
        }
        if (input != &dst) dst = src;
        if (this->scan_params_.downsampling.random.enable)
            this->preprocess_filter_->random_sampling(dst, this->scan_params_.downsampling.random.num);
    }

    void random_sampling(const PointCloudShared& src, PointCloudShared& dst, size_t num) const {
        this->preprocess_filter_->random_sampling(src, dst, num);
    }

    ProcessingContext prepare_context(const PointCloudShared& scan) const {
        ProcessingContext ctx;
        ctx.tree = algorithms::knn::KDTree::build(this->queue_, scan);
        return ctx;
    }

    /// :158-171
    void compute_covariances(PointCloudShared& scan, ProcessingContext& ctx) const {
        auto events = ctx.tree->knn_search_async(scan, this->covs_params_.neighbor_num, ctx.knn_result);
        const auto& m = this->covs_params_.m_estimation;
        if (m.enable) {
            events += algorithms::covariance::estimate_robust_async(ctx.knn_result, scan, m.type, m.mad_scale,
                                                                    m.min_robust_scale, m.max_iterations, events.evs);
        } else {
            events += algorithms::covariance::estimate_async(ctx.knn_result, scan, events.evs);
        }
        events.wait_and_throw();
    }

    /// :173-205 — the angle-of-incidence filter; the intensity filters are not built
    void refine_filter(PointCloudShared& scan, const ProcessingContext& /*ctx*/) const {
        const auto& s = this->scan_params_;
        if (s.preprocess.angle_incidence_filter.enable)
            this->preprocess_filter_->angle_incidence_filter(scan, scan, s.preprocess.angle_incidence_filter.min_angle,
                                                             s.preprocess.angle_incidence_filter.max_angle);
        if (scan.has_intensity() && ((s.intensity_correction.enable && !s.enhanced_reflectivity.enable) ||
                                     s.intensity_gaussian.enable || s.intensity_local_mean_norm.enable))
            throw std::runtime_error(
                "[PCProcessor::refine_filter] the intensity filters are not built in libspx: disable "
                "scan.intensity_correction / intensity_gaussian / intensity_local_mean_norm or drop the intensities");
    }

private:
    sycl_utils::DeviceQueue queue_;
    algorithms::filter::PreprocessFilter::Ptr preprocess_filter_ = nullptr;
    algorithms::filter::VoxelGrid::Ptr voxel_filter_ = nullptr;
    algorithms::filter::PolarGrid::Ptr polar_filter_ = nullptr;
    odometry::CommonParameters::Scan scan_params_;
    odometry::CommonParameters::CovarianceEstimation covs_params_;
};

}  // namespace pointcloud_processing
}  // namespace pipeline
}  // namespace sycl_points
