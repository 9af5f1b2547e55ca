// knn::KNNBase — I/algorithms/knn/knn.hpp:14-61.  User code may subclass it (the reference's tests
// inject host-side KNNs: T/test_registration_pipeline.cpp:16-61); Registration::align then drives
// the generic linearise path instead of the fused index path.
#pragma once

#include <vector>

#include "sycl_points/algorithms/knn/result.hpp"
#include "sycl_points/points/point_cloud.hpp"

namespace sycl_points {
namespace algorithms {
namespace knn {

class KNNBase {
public:
    virtual ~KNNBase() = default;

    virtual sycl_utils::events knn_search_async(const PointCloudShared& queries, const size_t k, KNNResult& result,
                                                const std::vector<sycl::event>& depends = std::vector<sycl::event>(),
                                                const TransformMatrix& transT = TransformMatrix::Identity()) const = 0;

    KNNResult knn_search(const PointCloudShared& queries, const size_t k,
                         const std::vector<sycl::event>& depends = std::vector<sycl::event>(),
                         const TransformMatrix& transT = TransformMatrix::Identity()) const {
        KNNResult result;
        knn_search_async(queries, k, result, depends, transT).wait_and_throw();
        return result;
    }

    sycl_utils::events nearest_neighbor_search_async(
        const PointCloudShared& queries, KNNResult& result,
        const std::vector<sycl::event>& depends = std::vector<sycl::event>(),
        const TransformMatrix& transT = TransformMatrix::Identity()) const {
        return knn_search_async(queries, 1, result, depends, transT);
    }

    void nearest_neighbor_search(const PointCloudShared& queries, KNNResult& result,
                                 const std::vector<sycl::event>& depends = std::vector<sycl::event>(),
                                 const TransformMatrix& transT = TransformMatrix::Identity()) const {
        nearest_neighbor_search_async(queries, result, depends, transT).wait_and_throw();
    }
};

}  // namespace knn
}  // namespace algorithms
}  // namespace sycl_points
