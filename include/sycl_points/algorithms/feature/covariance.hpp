// covariance::estimate_async / estimate_normals_async / extract_normals — 
// I/algorithms/feature/covariance.hpp:260-311,417-506 (launchers; the kernels are in libspx).
#pragma once

#include <stdexcept>

#include "sycl_points/algorithms/knn/knn.hpp"
#include "sycl_points/algorithms/robust/robust.hpp"
#include "sycl_points/points/point_cloud.hpp"

namespace sycl_points {
namespace algorithms {
namespace covariance {

namespace detail_spx {
inline void wait_all(const std::vector<sycl::event>& depends) {
    for (const auto& e : depends) e.wait();
}
}  // namespace detail_spx

/// covariance.hpp:260-292
inline sycl_utils::events estimate_async(const sycl_utils::DeviceQueue& queue, const knn::KNNResult& neighbors,
                                         const PointContainerShared& points, CovarianceContainerShared& covs,
                                         const std::vector<sycl::event>& depends = std::vector<sycl::event>()) {
    const size_t N = points.size();
    if (covs.size() != N) covs.resize(N);
    sycl_utils::events events;
    if (N == 0) return events;
    detail_spx::wait_all(depends);
    queue.set_accessed_by_device(points.data(), N);
    queue.set_accessed_by_device(covs.data(), N);
    queue.set_accessed_by_device(neighbors.indices->data(), neighbors.indices->size());
    detail::spx_check(spx_covariance(queue.handle(), reinterpret_cast<const float*>(points.data()), N,
                                     neighbors.indices->data(), (int)neighbors.k,
                                     reinterpret_cast<float*>(covs.data())));
    events += queue.checkpoint();
    return events;
}

/// covariance.hpp:294-297
inline sycl_utils::events estimate_async(const knn::KNNResult& neighbors, const PointCloudShared& points,
                                         const std::vector<sycl::event>& depends = std::vector<sycl::event>()) {
    return estimate_async(points.queue, neighbors, *points.points, *points.covs, depends);
}

/// covariance.hpp:305-311
inline sycl_utils::events estimate_async(const knn::KNNBase& knn, const PointCloudShared& points,
                                         const size_t k_correspondences,
                                         const std::vector<sycl::event>& depends = std::vector<sycl::event>()) {
    knn::KNNResult neighbors;
    auto knn_events = knn.knn_search_async(points, k_correspondences, neighbors, depends);
    auto ev = estimate_async(neighbors, points, knn_events.evs);
    ev.add_resource(neighbors.indices);
    ev.add_resource(neighbors.distances);
    return ev;
}

/// covariance.hpp:323-373 — M-estimated covariances (kernel: covariance.hpp:97-134,143-250)
inline sycl_utils::events estimate_robust_async(const sycl_utils::DeviceQueue& queue, const knn::KNNResult& neighbors,
                                                const PointContainerShared& points, CovarianceContainerShared& covs,
                                                robust::RobustLossType robust_type = robust::RobustLossType::CAUCHY,
                                                float mad_scale = 1.0f, float min_robust_scale = 1.0f,
                                                size_t robust_max_iterations = 1,
                                                const std::vector<sycl::event>& depends = std::vector<sycl::event>()) {
    if (neighbors.k > 64) throw std::runtime_error("[covariance::estimate_robust_async] neighbor K is too large. MAX_K is 64");
    const size_t N = points.size();
    if (covs.size() != N) covs.resize(N);
    sycl_utils::events events;
    if (N == 0) return events;
    detail_spx::wait_all(depends);
    queue.set_accessed_by_device(points.data(), N);
    queue.set_accessed_by_device(covs.data(), N);
    queue.set_accessed_by_device(neighbors.indices->data(), neighbors.indices->size());
    detail::spx_check(spx_covariance_robust(queue.handle(), reinterpret_cast<const float*>(points.data()), N,
                                            neighbors.indices->data(), (int)neighbors.k, (int)robust_type, mad_scale,
                                            min_robust_scale, (int)robust_max_iterations,
                                            reinterpret_cast<float*>(covs.data())));
    events += queue.checkpoint();
    return events;
}

/// covariance.hpp:383-390
inline sycl_utils::events estimate_robust_async(const knn::KNNResult& neighbors, const PointCloudShared& points,
                                                robust::RobustLossType robust_type = robust::RobustLossType::CAUCHY,
                                                float mad_scale = 1.0f, float min_robust_scale = 1.0f,
                                                size_t robust_max_iterations = 1,
                                                const std::vector<sycl::event>& depends = std::vector<sycl::event>()) {
    return estimate_robust_async(points.queue, neighbors, *points.points, *points.covs, robust_type, mad_scale,
                                 min_robust_scale, robust_max_iterations, depends);
}

/// covariance.hpp:400-410
inline sycl_utils::events estimate_robust_async(const knn::KNNBase& knn, const PointCloudShared& points,
                                                const size_t k_correspondences,
                                                robust::RobustLossType robust_type = robust::RobustLossType::CAUCHY,
                                                float mad_scale = 1.0f, float min_robust_scale = 1.0f,
                                                size_t robust_max_iterations = 1,
                                                const std::vector<sycl::event>& depends = std::vector<sycl::event>()) {
    knn::KNNResult neighbors;
    auto knn_events = knn.knn_search_async(points, k_correspondences, neighbors, depends);
    auto ev = estimate_robust_async(neighbors, points, robust_type, mad_scale, min_robust_scale, robust_max_iterations,
                                    knn_events.evs);
    ev.add_resource(neighbors.indices);
    ev.add_resource(neighbors.distances);
    return ev;
}

/// covariance.hpp:417-445
inline sycl_utils::events estimate_normals_async(const knn::KNNResult& neighbors, const PointCloudShared& points,
                                                 const std::vector<sycl::event>& depends = std::vector<sycl::event>()) {
    const size_t N = points.size();
    if (points.normals->size() != N) points.resize_normals(N);
    sycl_utils::events events;
    if (N == 0) return events;
    detail_spx::wait_all(depends);
    const auto& queue = points.queue;
    queue.set_accessed_by_device(points.points_ptr(), N);
    queue.set_accessed_by_device(points.normals_ptr(), N);
    queue.set_accessed_by_device(neighbors.indices->data(), neighbors.indices->size());
    detail::spx_check(spx_normals(queue.handle(), reinterpret_cast<const float*>(points.points_ptr()), N,
                                  neighbors.indices->data(), (int)neighbors.k,
                                  reinterpret_cast<float*>(points.normals_ptr())));
    events += queue.checkpoint();
    return events;
}

/// covariance.hpp:453-459
inline sycl_utils::events estimate_normals_async(const knn::KNNBase& knn, const PointCloudShared& points,
                                                 const size_t k_correspondences,
                                                 const std::vector<sycl::event>& depends = std::vector<sycl::event>()) {
    knn::KNNResult neighbors;
    auto knn_events = knn.knn_search_async(points, k_correspondences, neighbors, depends);
    auto ev = estimate_normals_async(neighbors, points, knn_events.evs);
    ev.add_resource(neighbors.indices);
    ev.add_resource(neighbors.distances);
    return ev;
}

/// covariance.hpp:465-495
inline sycl_utils::events extract_normals_async(const PointCloudShared& points,
                                                const std::vector<sycl::event>& depends = std::vector<sycl::event>()) {
    const size_t N = points.size();
    if (!points.has_cov()) throw std::runtime_error("[covariance::extract_normals_async] covariances not computed");
    if (points.normals->size() != N) points.resize_normals(N);
    sycl_utils::events events;
    if (N == 0) return events;
    detail_spx::wait_all(depends);
    const auto& queue = points.queue;
    queue.set_accessed_by_device(points.points_ptr(), N);
    queue.set_accessed_by_device(points.covs_ptr(), N);
    queue.set_accessed_by_device(points.normals_ptr(), N);
    detail::spx_check(spx_normals_from_covs(queue.handle(), reinterpret_cast<const float*>(points.points_ptr()),
                                            reinterpret_cast<const float*>(points.covs_ptr()), N,
                                            reinterpret_cast<float*>(points.normals_ptr())));
    events += queue.checkpoint();
    return events;
}

/// covariance.hpp:500-503
inline void extract_normals(const PointCloudShared& points,
                            const std::vector<sycl::event>& depends = std::vector<sycl::event>()) {
    extract_normals_async(points, depends).wait_and_throw();
}

}  // namespace covariance
}  // namespace algorithms
}  // namespace sycl_points
