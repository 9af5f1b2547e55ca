// submapping::Submap — I/pipeline/submapping.hpp:20-249 with the VOXEL_HASH_MAP map type: keyframe decision, the
// VoxelHashMap submap, the target index and the covariances / normals the registration factor needs.
#pragma once

#include <cmath>
#include <limits>
#include <memory>
#include <stdexcept>
#include <vector>

#include "sycl_points/algorithms/common/transform.hpp"
#include "sycl_points/algorithms/feature/covariance.hpp"
#include "sycl_points/algorithms/filter/preprocess_filter.hpp"
#include "sycl_points/algorithms/knn/kdtree.hpp"
#include "sycl_points/algorithms/mapping/voxel_hash_map.hpp"
#include "sycl_points/algorithms/registration/registration_params.hpp"
#include "sycl_points/algorithms/registration/result.hpp"
#include "sycl_points/pipeline/odometry_common_params.hpp"

namespace sycl_points {
namespace pipeline {
namespace submapping {

class Submap {
public:
    using Ptr = std::shared_ptr<Submap>;
    using ConstPtr = std::shared_ptr<const Submap>;
    using OdometryCommonParams = odometry::CommonParameters;
    using SubmapMapType = odometry::SubmapMapType;

    const auto& get_last_keyframe_pose() const { return this->last_keyframe_pose_; }
    const auto& get_keyframe_poses() const { return this->keyframe_poses_; }
    const auto& get_submap_kdtree() const { return *this->submap_tree_; }
    const PointCloudShared& get_submap_point_cloud() const { return *this->submap_pc_ptr_; }
    const PointCloudShared& get_last_keyframe_point_cloud() const { return *this->last_keyframe_pc_; }

    Submap(const sycl_utils::DeviceQueue& queue, const OdometryCommonParams& params) : queue_(queue) {
        if (params.submap.map_type != SubmapMapType::VOXEL_HASH_MAP)
            throw std::runtime_error(
                "[Submap] only submap.map_type = VOXEL_HASH_MAP is built in libspx (OccupancyGridMap is out of scope)");
        this->last_keyframe_pc_ = std::make_shared<PointCloudShared>(this->queue_);
        this->submap_pc_ptr_ = std::make_shared<PointCloudShared>(this->queue_);
        this->submap_pc_tmp_ = std::make_shared<PointCloudShared>(this->queue_);
        this->submap_params_ = params.submap;
        this->cov_params_ = params.covariance_estimation;
        this->reg_params_ = params.registration;
        this->last_keyframe_pose_ = params.pose.initial;
        this->last_keyframe_time_ = -1.0;
        this->keyframe_poses_.clear();
        this->keyframe_poses_.push_back(params.pose.initial);
        this->preprocess_filter_ = std::make_shared<algorithms::filter::PreprocessFilter>(this->queue_);
        this->submap_voxel_ =
            std::make_shared<algorithms::mapping::VoxelHashMap>(this->queue_, this->submap_params_.voxel_size);
    }

    /// :90-99
    void add_first_frame(const PointCloudShared& cloud, double timestamp, const Eigen::Isometry3f& current_pose) {
        this->last_keyframe_pose_ = current_pose;
        if (this->keyframe_poses_.empty()) {
            this->keyframe_poses_.push_back(current_pose);
        } else {
            this->keyframe_poses_.front() = current_pose;
        }
        this->build_submap(cloud, current_pose, true);
        this->last_keyframe_time_ = timestamp;
    }

    /// :109-134
    bool add_frame(const PointCloudShared& preprocessed_cloud,
                   const algorithms::registration::RegistrationResult& reg_result, float inlier_ratio, double timestamp,
                   shared_vector_ptr<float> random_sampling_weights = nullptr) {
        if (this->submap_params_.keyframe.inlier_ratio_threshold > 0.0f &&
            inlier_ratio <= this->submap_params_.keyframe.inlier_ratio_threshold)
            return false;  // registration failed
        if (this->is_keyframe(reg_result, timestamp)) {
            this->last_keyframe_pose_ = reg_result.T;
            this->last_keyframe_time_ = timestamp;
            this->keyframe_poses_.push_back(reg_result.T);
            this->build_submap(preprocessed_cloud, reg_result.T, false, random_sampling_weights);
            return true;
        }
        return false;
    }

    EIGEN_MAKE_ALIGNED_OPERATOR_NEW

private:
    sycl_utils::DeviceQueue queue_;
    OdometryCommonParams::Submap submap_params_;
    OdometryCommonParams::CovarianceEstimation cov_params_;
    OdometryCommonParams::Registration reg_params_;
    algorithms::knn::KNNResult knn_result_;
    double last_keyframe_time_ = -1.0;
    Eigen::Isometry3f last_keyframe_pose_;
    std::vector<Eigen::Isometry3f, Eigen::aligned_allocator<Eigen::Isometry3f>> keyframe_poses_;
    algorithms::filter::PreprocessFilter::Ptr preprocess_filter_ = nullptr;
    algorithms::mapping::VoxelHashMap::Ptr submap_voxel_ = nullptr;
    algorithms::knn::KDTree::Ptr submap_tree_ = nullptr;
    PointCloudShared::Ptr last_keyframe_pc_ = nullptr;  // sensor frame
    PointCloudShared::Ptr submap_pc_ptr_ = nullptr;     // odom / world frame
    PointCloudShared::Ptr submap_pc_tmp_ = nullptr;

    /// :157-175 — far enough, turned enough, or long enough since the last keyframe
    bool is_keyframe(const algorithms::registration::RegistrationResult& reg_result, double timestamp) const {
        const auto& k = this->submap_params_.keyframe;
        const Eigen::Isometry3f moved = this->last_keyframe_pose_.inverse() * reg_result.T;
        if (moved.translation().norm() >= k.distance_threshold) return true;
        constexpr float kRadToDeg = 180.0f / 3.14159265358979323846f;
        if (std::fabs(Eigen::AngleAxisf(moved.rotation()).angle()) * kRadToDeg >= k.angle_threshold_degrees) return true;
        // (no keyframe time yet: the reference treats the elapsed time as unbounded)
        return !(this->last_keyframe_time_ > 0.0) || timestamp - this->last_keyframe_time_ >= k.time_threshold_seconds;
    }

    /// :177-212 — keyframe cloud (sampled) -> voxel map -> boxed export -> target index + what the factor needs
    void build_submap(const PointCloudShared& cloud, const Eigen::Isometry3f& current_pose, bool is_first_frame,
                      shared_vector_ptr<float> random_sampling_weights = nullptr) {
        this->sample_keyframe_cloud(cloud, random_sampling_weights);
        this->submap_voxel_->add_point_cloud(*this->last_keyframe_pc_, current_pose);
        this->submap_voxel_->downsampling(*this->submap_pc_tmp_, current_pose.translation(),
                                          this->submap_params_.max_distance_range);
        if (is_first_frame) {
            // the first target is the whole preprocessed scan in the odom frame, not its voxel-map image
            *this->submap_pc_ptr_ = algorithms::transform::transform_copy(cloud, current_pose.matrix());
        } else if (this->submap_pc_tmp_->size() >= this->reg_params_.min_num_points) {
            this->submap_pc_ptr_.swap(this->submap_pc_tmp_);
        }
        this->submap_tree_ = algorithms::knn::KDTree::build(this->queue_, *this->submap_pc_ptr_);
        this->prepare_target_attributes();
    }

    /// robust-ICP-weighted mixed sampling when the caller brought weights for exactly this cloud, uniform otherwise
    void sample_keyframe_cloud(const PointCloudShared& cloud, const shared_vector_ptr<float>& weights) {
        const size_t num = this->submap_params_.point_random_sampling_num;
        if (weights != nullptr && weights->size() == cloud.size()) {
            this->preprocess_filter_->mixed_random_sampling(cloud, *this->last_keyframe_pc_, *weights, num,
                                                            this->submap_params_.weighted_sampling_ratio);
            return;
        }
        this->preprocess_filter_->random_sampling(cloud, *this->last_keyframe_pc_, num);
    }

    /// :214-247 — normals and / or covariances of the submap, whichever the registration factor reads; the k-NN search
    /// runs at most once and only when something needs it
    void prepare_target_attributes() {
        using algorithms::registration::RegType;
        const RegType factor = this->reg_params_.factor.reg_type;
        const bool wants_normals = factor == RegType::POINT_TO_PLANE || factor == RegType::GENZ;
        const bool wants_covs = factor == RegType::GICP || factor == RegType::POINT_TO_DISTRIBUTION ||
                                factor == RegType::GENZ || this->reg_params_.factor.rotation_constraint.enable;
        PointCloudShared& target = *this->submap_pc_ptr_;
        const bool has_covs = target.has_cov();  // (the voxel map exports them when the keyframes carried them)
        sycl_utils::events knn_done, pending;
        bool searched = false;
        const auto neighbours = [&]() -> const std::vector<sycl::event>& {
            if (!searched) {
                knn_done = this->submap_tree_->knn_search_async(target, this->cov_params_.neighbor_num, this->knn_result_);
                searched = true;
            }
            return knn_done.evs;
        };
        if (wants_normals) {
            const auto& deps = neighbours();
            pending += has_covs ? algorithms::covariance::extract_normals_async(target, deps)
                                : algorithms::covariance::estimate_normals_async(this->knn_result_, target, deps);
        }
        if (wants_covs && !has_covs) pending += algorithms::covariance::estimate_async(this->knn_result_, target, neighbours());
        pending.wait_and_throw();
    }
};

}  // namespace submapping
}  // namespace pipeline
}  // namespace sycl_points
