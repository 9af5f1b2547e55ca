// knn::knn_search_bruteforce — I/algorithms/knn/bruteforce.hpp:24-96 (synchronous, no transT).
#pragma once

#include "sycl_points/algorithms/knn/result.hpp"
#include "sycl_points/points/point_cloud.hpp"

namespace sycl_points {
namespace algorithms {
namespace knn {

inline KNNResult knn_search_bruteforce(const sycl_utils::DeviceQueue& queue, const PointCloudShared& queries,
                                       const PointCloudShared& targets, const size_t k) {
    KNNResult result;
    result.allocate(queue, queries.size(), k);
    if (queries.size() == 0) return result;
    queue.set_accessed_by_device(queries.points_ptr(), queries.size());
    queue.set_accessed_by_device(targets.points_ptr(), targets.size());
    queue.set_accessed_by_device(result.indices->data(), result.indices->size());
    queue.set_accessed_by_device(result.distances->data(), result.distances->size());
    detail::spx_check(spx_knn_bruteforce(queue.handle(), reinterpret_cast<const float*>(queries.points_ptr()),
                                         queries.size(), reinterpret_cast<const float*>(targets.points_ptr()),
                                         targets.size(), (int)k, nullptr, result.indices->data(),
                                         result.distances->data()));
    detail::spx_check(spx_queue_sync(queue.handle()));  // bruteforce.hpp:93 wait_and_throw
    return result;
}

}  // namespace knn
}  // namespace algorithms
}  // namespace sycl_points
