// knn::Octree — I/algorithms/knn/octree.hpp:29-127 behind the same interface.  Like knn::KDTree here, the structure
// is libspx's GPU-resident exact grid index: any exact index satisfies the KNNBase contract, and
// the answer is the brute-force one with ties broken by index.  `resolution` and `max_points_per_node` steer the
// reference's node splitting; the grid chooses its own cell edge from the data and only records them.
#pragma once

#include <memory>
#include <stdexcept>

#include "sycl_points/algorithms/knn/kdtree.hpp"

namespace sycl_points {
namespace algorithms {
namespace knn {

class Octree : public KNNBase {
public:
    using Ptr = std::shared_ptr<Octree>;

    /// octree.hpp:92-95
    Octree(const sycl_utils::DeviceQueue& queue, float resolution, size_t max_points_per_node)
        : queue_(queue), resolution_(resolution), max_points_per_node_(max_points_per_node) {
        if (!(resolution > 0.0f)) throw std::invalid_argument("[Octree] resolution must be positive");
    }

    /// octree.hpp:98-99
    static Ptr build(const sycl_utils::DeviceQueue& queue, const PointCloudShared& points, float resolution,
                     size_t max_points_per_node = 32) {
        auto tree = std::make_shared<Octree>(queue, resolution, max_points_per_node);
        tree->index_ = KDTree::build(queue, points);
        return tree;
    }

    /// octree.hpp:107-110,599-630 (k <= 100 as there)
    sycl_utils::events knn_search_async(const PointCloudShared& queries, const size_t k, KNNResult& result,
                                        const std::vector<sycl::event>& depends = std::vector<sycl::event>(),
                                        const TransformMatrix& transT = TransformMatrix::Identity()) const override {
        if (k == 0) throw std::runtime_error("[Octree::knn_search_async] `k` must be positive");
        if (k > 100) throw std::runtime_error("[Octree::knn_search_async] `k` is too large. not support.");
        if (!this->index_) throw std::runtime_error("[Octree::knn_search_async] Octree is not built");
        return this->index_->knn_search_async(queries, k, result, depends, transT);
    }

    [[nodiscard]] float resolution() const { return this->resolution_; }
    [[nodiscard]] size_t max_points_per_node() const { return this->max_points_per_node_; }
    [[nodiscard]] size_t size() const { return this->index_ ? this->index_->size() : 0; }

    /// octree.hpp:121,276-380: drop the points whose flag is not INCLUDE_FLAG, re-number the others by `indices`
    void remove_nodes_by_flags(const shared_vector<uint8_t>& flags, const shared_vector<int32_t>& indices) {
        if (flags.size() != indices.size())
            throw std::runtime_error("[Octree::remove_nodes_by_flags] flags and indices must have the same size");
        if (!this->index_) return;
        this->index_->remove_nodes_by_flags(flags, indices);
    }

private:
    sycl_utils::DeviceQueue queue_;
    float resolution_;
    size_t max_points_per_node_;
    KDTree::Ptr index_;
};

}  // namespace knn
}  // namespace algorithms
}  // namespace sycl_points
