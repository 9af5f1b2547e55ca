// Scan-to-scan registration through the sycl_points C++ API on libspx (B200).  Same pipeline and
// settings as the reference's cpp/examples/example_registration.cpp (box filter 0.5-50 m, 0.25 m
// voxel grid, k = 10 covariances + normals, GICP / LM / Geman-McClure, robust scale 10 -> 2.5 in
// 3 levels, 1000-point random source sampling), written against the facade headers only — user
// code for the reference compiles against include/sycl_points unchanged.
//
//   g++ -std=c++20 -O2 -Iinclude examples/example_registration.cpp -Lsycl_points_b200 -lspx
//       -Wl,-rpath,$PWD/sycl_points_b200 -o example_registration
//   ./example_registration source.ply target.ply [loops] [T_target_source.txt]
#include <chrono>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>

#include "sycl_points/algorithms/feature/covariance.hpp"
#include "sycl_points/algorithms/filter/preprocess_filter.hpp"
#include "sycl_points/algorithms/filter/voxel_downsampling.hpp"
#include "sycl_points/algorithms/knn/kdtree.hpp"
#include "sycl_points/algorithms/registration/registration_pipeline.hpp"
#include "sycl_points/io/point_cloud_reader.hpp"

namespace sp = sycl_points;
namespace reg = sycl_points::algorithms::registration;

struct StageTimer {
    std::map<std::string, double> total_us;
    std::chrono::steady_clock::time_point t0;
    void start() { t0 = std::chrono::steady_clock::now(); }
    void stop(const std::string& stage, bool record) {
        const auto t1 = std::chrono::steady_clock::now();
        if (record) total_us[stage] += std::chrono::duration<double, std::micro>(t1 - t0).count();
        t0 = t1;
    }
};

int main(int argc, char** argv) {
    if (argc < 3) {
        std::cerr << "usage: " << argv[0] << " source.ply target.ply [loops=100] [T_target_source.txt]" << std::endl;
        return 2;
    }
    const size_t loops = argc > 3 ? std::stoul(argv[3]) : 100;
    const size_t warm_up = std::min<size_t>(10, loops);
    const sp::PointCloudCPU source_points = sp::PointCloudReader::readFile(argv[1], false, false);
    const sp::PointCloudCPU target_points = sp::PointCloudReader::readFile(argv[2], false, false);

    sycl::device dev(sp::sycl_utils::device_selector::default_selector_v);
    sp::sycl_utils::DeviceQueue queue(dev);
    queue.print_device_info();

    reg::RegistrationPipelineParams param;
    param.registration.max_iterations = 10;
    param.registration.max_correspondence_distance = 2.0f;
    param.registration.optimization_method = reg::OptimizationMethod::LEVENBERG_MARQUARDT;
    param.registration.robust.type = sp::algorithms::robust::RobustLossType::GEMAN_MCCLURE;
    param.registration.robust.default_scale = 10.0f;
    param.registration.reg_type = reg::RegType::GICP;
    param.robust.auto_scale = true;
    param.robust.init_scale = 10.0f;
    param.robust.min_scale = 2.5f;
    param.robust.rotation_init_scale = 5.0f;
    param.robust.rotation_min_scale = 2.5f;
    param.robust.auto_scaling_iter = 3;

    reg::RegistrationPipeline pipeline(queue, param);
    sp::algorithms::filter::VoxelGrid voxel_grid(queue, 0.25f);
    sp::algorithms::filter::PreprocessFilter filter(queue);
    const size_t k = 10;

    StageTimer timer;
    reg::RegistrationResult result;
    size_t n_src = 0, n_tgt = 0;
    for (size_t i = 0; i < loops + warm_up; ++i) {
        const bool rec = i >= warm_up;
        timer.start();
        sp::PointCloudShared source(queue, source_points), target(queue, target_points);
        timer.stop("1. to PointCloudShared", rec);

        filter.box_filter(source, 0.5f, 50.0f);
        filter.box_filter(target, 0.5f, 50.0f);
        sp::PointCloudShared source_ds(queue), target_ds(queue);
        voxel_grid.downsampling(source, source_ds);
        voxel_grid.downsampling(target, target_ds);
        timer.stop("2. Downsampling", rec);

        const auto source_tree = sp::algorithms::knn::KDTree::build(queue, source_ds);
        const auto target_tree = sp::algorithms::knn::KDTree::build(queue, target_ds);
        timer.stop("3. KDTree build", rec);

        const auto source_nn = source_tree->knn_search(source_ds, k);
        const auto target_nn = target_tree->knn_search(target_ds, k);
        timer.stop("4. KDTree kNN Search", rec);

        sp::algorithms::covariance::estimate_async(source_nn, source_ds).wait_and_throw();
        sp::algorithms::covariance::estimate_async(target_nn, target_ds).wait_and_throw();
        timer.stop("5. compute Covariances", rec);

        sp::algorithms::covariance::estimate_normals_async(source_nn, source_ds).wait_and_throw();
        sp::algorithms::covariance::estimate_normals_async(target_nn, target_ds).wait_and_throw();
        timer.stop("6. compute Normals", rec);

        result = pipeline.align(source_ds, target_ds, *target_tree, sp::TransformMatrix::Identity());
        timer.stop("7. Registration", rec);
        n_src = source_ds.size();
        n_tgt = target_ds.size();
    }

    std::cout << "points after box filter + voxel grid: source " << n_src << ", target " << n_tgt << "\n";
    std::cout << "T =\n" << result.T.matrix() << "\n";
    std::cout << "converged " << result.converged << ", iterations " << result.iterations << ", inlier " << result.inlier
              << ", error " << result.error << "\n\n";
    double total = 0.0;
    for (const auto& [stage, us] : timer.total_us) {
        std::cout << std::setw(26) << stage + ": " << std::setw(9) << std::fixed << std::setprecision(2)
                  << us / loops << " us\n";
        total += us / loops;
    }
    std::cout << std::setw(26) << "TOTAL: " << std::setw(9) << total << " us" << std::endl;

    if (argc > 4) {  // sanity against a ground-truth pose file (4x4, row-major text)
        std::ifstream f(argv[4]);
        sp::TransformMatrix gt;
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 4; ++c) f >> gt(r, c);
        float dt = 0.0f;
        for (int r = 0; r < 3; ++r) dt += (gt(r, 3) - result.T.matrix()(r, 3)) * (gt(r, 3) - result.T.matrix()(r, 3));
        std::cout << "translation error vs ground truth: " << std::sqrt(dt) << " m" << std::endl;
    }
    return 0;
}
