// LiDAR odometry on a directory of scans through the C++ facade (pipeline::lidar_odometry::LiDAROdometryPipeline over
// libspx): the LiDAR-only loop of the reference's ROS node, without ROS.
//   example_lidar_odometry <dir> <n_frames> [dt_seconds]     reads <dir>/scan_000.ply ... and prints one pose per frame
// Settings = tests/test_gpu_odometry.py::make_params (voxel 0.4 m, GICP + Huber against a 0.5 m VoxelHashMap submap).
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "sycl_points/io/point_cloud_reader.hpp"
#include "sycl_points/pipeline/lidar_odometry.hpp"

int main(int argc, char** argv) {
    if (argc < 3) {
        std::cerr << "usage: example_lidar_odometry <dir> <n_frames> [dt_seconds]" << std::endl;
        return 2;
    }
    const std::string dir = argv[1];
    const int n_frames = std::atoi(argv[2]);
    const double dt = argc > 3 ? std::atof(argv[3]) : 0.1;
    namespace lo = sycl_points::pipeline::lidar_odometry;
    lo::Parameters P;
    P.submap.map_type = sycl_points::pipeline::odometry::SubmapMapType::VOXEL_HASH_MAP;
    P.submap.voxel_size = 0.5f;
    P.submap.point_random_sampling_num = 6000;
    P.submap.max_distance_range = 60.0f;
    P.submap.keyframe.distance_threshold = 0.3f;
    P.scan.downsampling.polar.enable = false;
    P.scan.downsampling.voxel.enable = true;
    P.scan.downsampling.voxel.size = 0.4f;
    P.scan.downsampling.random.num = 8000;
    P.scan.preprocess.box_filter.min = 1.0f;
    P.scan.preprocess.box_filter.max = 80.0f;
    P.registration.factor.robust.type = sycl_points::algorithms::robust::RobustLossType::HUBER;
    P.registration.factor.robust.default_scale = 1.0f;
    P.registration_sampling.num = 3000;
    // initial pose: <dir>/pose0.txt (row-major 4x4) when present
    if (FILE* f = std::fopen((dir + "/pose0.txt").c_str(), "r")) {
        Eigen::Matrix4f M = Eigen::Matrix4f::Identity();
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                float v = 0.0f;
                if (std::fscanf(f, "%f", &v) == 1) M(i, j) = v;
            }
        std::fclose(f);
        P.pose.initial = Eigen::Isometry3f(M);
    }
    try {
        lo::LiDAROdometryPipeline pipe(P);
        const auto queue = pipe.get_device_queue();
        for (int k = 0; k < n_frames; ++k) {
            char name[64];
            std::snprintf(name, sizeof(name), "/scan_%03d.ply", k);
            const auto cpu = sycl_points::PointCloudReader::readFile(dir + name, false, false);
            auto scan = std::make_shared<sycl_points::PointCloudShared>(*queue, cpu);
            const auto rc = pipe.process(scan, dt * k);
            const Eigen::Matrix4f T = pipe.get_odom().matrix();
            std::printf("frame %d rc %d pose", k, (int)rc);
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) std::printf(" %.9g", T(i, j));
            std::printf("\n");
            if (rc != lo::LiDAROdometryPipeline::ResultType::success && rc != lo::LiDAROdometryPipeline::ResultType::first_frame) {
                std::cerr << "process() failed: " << pipe.get_error_message() << std::endl;
                return 1;
            }
        }
        std::printf("keyframes %zu submap %zu\n", pipe.get_keyframe_poses().size(), pipe.get_submap_point_cloud().size());
        for (const auto& kv : pipe.get_total_processing_times()) {
            double s = 0.0;
            for (size_t i = kv.second.size() > 5 ? 5 : 0; i < kv.second.size(); ++i) s += kv.second[i];
            const size_t cnt = kv.second.size() > 5 ? kv.second.size() - 5 : kv.second.size();
            std::printf("%-26s mean %.1f us over %zu frames\n", kv.first.c_str(), cnt ? s / cnt : 0.0, cnt);
        }
    } catch (const std::exception& e) {
        std::cerr << "exception: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
