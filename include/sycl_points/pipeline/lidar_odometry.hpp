// lidar_odometry::LiDAROdometryPipeline — I/pipeline/lidar_odometry.hpp:27-621, LiDAR-only:
//   process(scan, t): prefilter -> covariances -> refine filter -> [first frame: seed the submap] -> motion prediction
//                     -> (MAP prior) -> RegistrationPipeline::align against the submap -> keyframe decision + submap
//                     update -> velocity / odometry update.
// The IMU paths of the reference class (preintegration, IMU deskew, initial alignment) are not built: imu.enable must
// stay false (the constructor refuses otherwise).
#pragma once

#include <chrono>
#include <cmath>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "sycl_points/algorithms/deskew/relative_pose_deskew.hpp"
#include "sycl_points/algorithms/registration/registration_pipeline.hpp"
#include "sycl_points/pipeline/lidar_odometry_params.hpp"
#include "sycl_points/pipeline/motion_predictor.hpp"
#include "sycl_points/pipeline/pointcloud_processing.hpp"
#include "sycl_points/pipeline/submapping.hpp"
#include "sycl_points/points/point_cloud.hpp"

namespace sycl_points {
namespace pipeline {
namespace lidar_odometry {
using LidarOdometryParams = lidar_odometry::Parameters;

class LiDAROdometryPipeline {
public:
    using Ptr = std::shared_ptr<LiDAROdometryPipeline>;
    using ConstPtr = std::shared_ptr<const LiDAROdometryPipeline>;

    enum class ResultType : std::int8_t {
        success = 0,
        first_frame,
        waiting_initial_alignment,
        error = 100,
        old_timestamp,
        small_number_of_points
    };

    LiDAROdometryPipeline(const LidarOdometryParams& params) : params_(params) {
        if (params.imu.enable)
            throw std::runtime_error("[LiDAR Odometry] the IMU paths are not built in libspx: set imu.enable = false");
        this->initialize();
    }

    auto get_device_queue() const { return this->queue_ptr_; }
    const auto& get_error_message() const { return this->error_message_; }
    const auto& get_current_processing_time() const { return this->current_processing_time_; }
    const auto& get_total_processing_times() const { return this->total_processing_times_; }
    const auto& get_odom() const { return this->odom_; }
    const auto& get_prev_odom() const { return this->prev_odom_; }
    const auto& get_last_keyframe_pose() const { return this->submap_->get_last_keyframe_pose(); }
    const auto& get_keyframe_poses() const { return this->submap_->get_keyframe_poses(); }
    const PointCloudShared& get_preprocessed_point_cloud() const { return *this->preprocessed_pc_; }
    const PointCloudShared& get_submap_point_cloud() const { return this->submap_->get_submap_point_cloud(); }
    const PointCloudShared& get_last_keyframe_point_cloud() const { return this->submap_->get_last_keyframe_point_cloud(); }
    const PointCloudShared* get_registration_input_point_cloud() const {
        return this->registration_pipeline_->get_registration_input_point_cloud();
    }
    const auto& get_registration_result() const { return *this->reg_result_; }

    /// :115-298
    ResultType process(const PointCloudShared::Ptr scan, double timestamp) {
        this->error_message_.clear();
        if (this->last_frame_time_ > 0.0) {
            const float dt = static_cast<float>(timestamp - this->last_frame_time_);
            if (dt > 0.0f) {
                this->dt_ = dt;
            } else {
                this->error_message_ = "old timestamp";
                return ResultType::old_timestamp;
            }
        }
        this->clear_current_processing_time();
        double dt_preprocessing = 0.0;
        if (!this->guarded("preprocess", [&] { dt_preprocessing = timed([&] { this->pc_processor_->prefilter(*scan, *this->preprocessed_pc_); }); }))
            return ResultType::error;
        if (!this->guarded("compute_covariances", [&] {
                this->add_delta_time(ProcessName::compute_covariances, timed([&] { this->compute_covariances(); }));
            }))
            return ResultType::error;
        if (!this->guarded("refine_filter", [&] {
                dt_preprocessing += timed([&] { this->pc_processor_->refine_filter(*this->preprocessed_pc_, this->processing_ctx_); });
                this->add_delta_time(ProcessName::preprocessing, dt_preprocessing);
            }))
            return ResultType::error;
        if (this->preprocessed_pc_->size() <= this->params_.registration.min_num_points) {
            this->error_message_ = "point cloud size is too small";
            return ResultType::small_number_of_points;
        }
        if (this->is_first_frame_) {
            if (!this->guarded("build_submap (first frame)",
                               [&] { this->submap_->add_first_frame(*this->preprocessed_pc_, timestamp, this->odom_); }))
                return ResultType::error;
            this->is_first_frame_ = false;
            this->last_frame_time_ = timestamp;
            return ResultType::first_frame;
        }
        if (!this->guarded("registration", [&] {
                double dt_reg = 0.0;
                dt_reg = timed([&] { *this->reg_result_ = this->registration(); });
                this->add_delta_time(ProcessName::registration, dt_reg);
            }))
            return ResultType::error;
        if (!this->guarded("submapping", [&] {
                this->add_delta_time(ProcessName::build_submap, timed([&] { this->submapping(*this->reg_result_, timestamp); }));
            }))
            return ResultType::error;
        if (this->params_.lo.pipeline.velocity_update.enable)  // :265-270
            algorithms::deskew::deskew_point_cloud_constant_velocity(*this->preprocessed_pc_, *this->preprocessed_pc_,
                                                                     this->odom_, this->reg_result_->T, this->dt_);
        this->prev_odom_ = this->odom_;
        this->odom_ = this->reg_result_->T;
        this->last_frame_time_ = timestamp;
        const Eigen::Isometry3f delta_pose = this->prev_odom_.inverse() * this->odom_;
        const Eigen::AngleAxisf delta_angle_axis(delta_pose.rotation());
        this->linear_velocity_ = delta_pose.translation() / this->dt_;
        this->angular_velocity_ = Eigen::AngleAxisf(delta_angle_axis.angle() / this->dt_, delta_angle_axis.axis());
        this->registrated_ = true;
        return ResultType::success;
    }

    EIGEN_MAKE_ALIGNED_OPERATOR_NEW

private:
    sycl_utils::DeviceQueue::Ptr queue_ptr_ = nullptr;
    PointCloudShared::Ptr preprocessed_pc_ = nullptr;  // sensor frame
    bool is_first_frame_ = true;
    pointcloud_processing::ProcessingContext processing_ctx_;
    shared_vector_ptr<float> icp_weights_ = nullptr;
    pointcloud_processing::PCProcessor::Ptr pc_processor_ = nullptr;
    algorithms::registration::RegistrationPipeline::Ptr registration_pipeline_ = nullptr;
    bool registrated_ = false;
    algorithms::registration::RegistrationResult::Ptr reg_result_ = nullptr;
    Eigen::Vector3f linear_velocity_;     // [m/s] in the previous LiDAR body frame
    Eigen::AngleAxisf angular_velocity_;  // [rad/s]
    Eigen::Isometry3f prev_odom_;
    Eigen::Isometry3f odom_;
    submapping::Submap::Ptr submap_ = nullptr;
    double last_frame_time_ = -1.0;  // [s]
    float dt_ = -1.0f;               // [s]
    Parameters params_;
    MotionPredictor::Ptr motion_predictor_ = nullptr;
    std::string error_message_;

    enum class ProcessName { preprocessing = 0, compute_covariances, registration, build_submap };
    const std::map<ProcessName, std::string> pn_map_ = {
        {ProcessName::preprocessing, "1. preprocessing"},
        {ProcessName::compute_covariances, "2. compute covariances"},
        {ProcessName::registration, "3. registration"},
        {ProcessName::build_submap, "4. build submap"},
    };
    std::map<std::string, double> current_processing_time_;
    std::map<std::string, std::vector<double>> total_processing_times_;

    template <typename F>
    static double timed(F&& f) {  // [us], like time_utils::measure_execution
        const auto t0 = std::chrono::steady_clock::now();
        f();
        return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    }
    template <typename F>
    bool guarded(const char* stage, F&& f) {  // the reference reports a failing stage and returns `error` (:143-176)
        try {
            f();
            return true;
        } catch (const std::exception& e) {
            this->error_message_ = std::string(stage) + ": " + e.what();
            std::cerr << "[LiDAR Odometry] " << this->error_message_ << std::endl;
            return false;
        }
    }
    void clear_current_processing_time() {
        this->current_processing_time_.clear();
        for (const auto& kv : pn_map_) this->current_processing_time_[kv.second] = 0.0;
    }
    void add_delta_time(ProcessName name, double dt) {
        this->total_processing_times_[pn_map_.at(name)].push_back(dt);
        this->current_processing_time_[pn_map_.at(name)] = dt;
    }

    void initialize() {  // :380-477 without the IMU blocks
        this->queue_ptr_ = std::make_shared<sycl_utils::DeviceQueue>(sycl::device(sycl::default_selector_v));
        this->icp_weights_ = std::make_shared<shared_vector<float>>();
        this->preprocessed_pc_ = std::make_shared<PointCloudShared>(*this->queue_ptr_);
        this->odom_ = this->params_.pose.initial;
        this->prev_odom_ = this->params_.pose.initial;
        this->linear_velocity_ = Eigen::Vector3f::Zero();
        this->angular_velocity_ = Eigen::AngleAxisf::Identity();
        this->pc_processor_ = std::make_shared<pointcloud_processing::PCProcessor>(
            *this->queue_ptr_, this->params_.scan, this->params_.covariance_estimation, this->params_.imu);
        this->submap_ = std::make_shared<submapping::Submap>(*this->queue_ptr_, this->params_);
        this->registration_pipeline_ = std::make_shared<algorithms::registration::RegistrationPipeline>(
            *this->queue_ptr_, this->params_.make_registration_pipeline_params());
        this->reg_result_ = std::make_shared<algorithms::registration::RegistrationResult>();
        this->registrated_ = false;
        for (const auto& kv : pn_map_) this->total_processing_times_[kv.second] = {};
        this->motion_predictor_ = std::make_shared<MotionPredictor>(this->params_.motion_prediction);
    }

    void compute_covariances() {  // :515-530
        using algorithms::registration::RegType;
        const bool needs_covs = this->params_.registration.factor.reg_type == RegType::GICP ||
                                this->params_.registration.factor.rotation_constraint.enable ||
                                this->params_.scan.preprocess.angle_incidence_filter.enable;
        if (!needs_covs) return;
        this->processing_ctx_ = this->pc_processor_->prepare_context(*this->preprocessed_pc_);
        this->pc_processor_->compute_covariances(*this->preprocessed_pc_, this->processing_ctx_);
    }

    algorithms::registration::RegistrationResult registration() {  // :544-597
        const Eigen::Isometry3f init_T = this->motion_predictor_->predict(
            this->linear_velocity_, this->angular_velocity_, this->odom_, this->dt_, this->reg_result_, this->registrated_);
        if (this->registrated_) this->registration_pipeline_->registration()->set_map_prior_state(*this->reg_result_, init_T);
        algorithms::registration::Registration::ExecutionOptions options;
        options.dt = this->dt_;
        options.prev_pose = this->odom_.matrix();
        return this->registration_pipeline_->align(*this->preprocessed_pc_, this->submap_->get_submap_point_cloud(),
                                                   this->submap_->get_submap_kdtree(), init_T.matrix(), options);
    }

    void submapping(const algorithms::registration::RegistrationResult& reg_result, double timestamp) {  // :599-621
        const auto reg_pc_ptr = this->registration_pipeline_->get_deskewed_point_cloud();
        if (reg_pc_ptr == nullptr)
            throw std::runtime_error("[LiDAR Odometry] get_deskewed_point_cloud() returned nullptr unexpectedly.");
        shared_vector_ptr<float> icp_weights = nullptr;
        if (reg_pc_ptr->size() > this->params_.submap.point_random_sampling_num) {
            const float robust_scale = this->params_.lo.pipeline.robust.auto_scale
                                           ? this->params_.lo.pipeline.robust.min_scale
                                           : this->params_.registration.factor.robust.default_scale;
            this->registration_pipeline_->compute_icp_robust_weights(this->submap_->get_submap_point_cloud(),
                                                                     this->submap_->get_submap_kdtree(),
                                                                     reg_result.T.matrix(), robust_scale, *this->icp_weights_);
            icp_weights = this->icp_weights_;
        }
        const float inlier_ratio = this->registration_pipeline_->get_inlier_ratio(reg_result);
        this->submap_->add_frame(*reg_pc_ptr, reg_result, inlier_ratio, timestamp, icp_weights);
    }
};

}  // namespace lidar_odometry
}  // namespace pipeline
}  // namespace sycl_points
