// knn::KNNResult — I/algorithms/knn/result.hpp:12-34: row-major [query][k] int32 indices and float
// SQUARED distances, ascending; unfilled slots -1 / FLT_MAX.
#pragma once

#include <limits>
#include <memory>

#include "sycl_points/utils/sycl_utils.hpp"

namespace sycl_points {
namespace algorithms {
namespace knn {

struct KNNResult {
    using Ptr = std::shared_ptr<KNNResult>;

    shared_vector_ptr<int32_t> indices = nullptr;
    shared_vector_ptr<float> distances = nullptr;
    size_t query_size;
    size_t k;
    KNNResult() : query_size(0), k(0) {}

    /// result.hpp:21-27.  The fill values are written by the search kernels themselves (every slot
    /// of every row is initialised on the device), so the arrays are not touched on the host here.
    void allocate(const sycl_utils::DeviceQueue& queue, size_t query_size = 0, size_t k = 0) {
        this->query_size = query_size;
        this->k = k;
        this->indices = std::make_shared<shared_vector<int32_t>>(*queue.ptr);
        this->distances = std::make_shared<shared_vector<float>>(*queue.ptr);
        this->indices->resize(query_size * k);
        this->distances->resize(query_size * k);
    }
    void resize(size_t query_size = 0, size_t k = 0) {
        this->query_size = query_size;
        this->k = k;
        this->indices->resize(query_size * k);
        this->distances->resize(query_size * k);
    }
};

}  // namespace knn
}  // namespace algorithms
}  // namespace sycl_points
