// knn::KDTree — I/algorithms/knn/kdtree.hpp:142-280 behind the same interface.  The structure is
// NOT a KD-tree: it is libspx's GPU-resident exact grid index (spx_index), built on the device
// (the reference builds on one host thread, kdtree.hpp:292-413).  The contract is the brute-force
// answer with ties broken by index — stricter than the reference's, whose far
// stack truncates at 16 entries and whose tie order follows traversal (kdtree.hpp:127,437,538).
#pragma once

#include <algorithm>

#include <memory>
#include <stdexcept>

#include "sycl_points/algorithms/knn/knn.hpp"

namespace sycl_points {
namespace algorithms {
namespace knn {

class KDTree : public KNNBase {
public:
    using Ptr = std::shared_ptr<KDTree>;
    sycl_utils::DeviceQueue queue;

    KDTree(const sycl_utils::DeviceQueue& q) : queue(q) {}
    ~KDTree() override {
        if (index_) spx_index_destroy(index_);
    }
    KDTree(const KDTree&) = delete;
    KDTree& operator=(const KDTree&) = delete;

    /// kdtree.hpp:165-169.  leaf_threshold has no meaning for the grid index and is ignored.
    static KDTree::Ptr build(const sycl_utils::DeviceQueue& q, const PointContainerShared& points,
                             size_t /*leaf_threshold*/ = 16) {
        auto tree = std::make_shared<KDTree>(q);
        q.set_accessed_by_device(points.data(), points.size());
        detail::spx_check(spx_index_build(q.handle(), reinterpret_cast<const float*>(points.data()), points.size(),
                                          0.0f, &tree->index_));
        tree->size_ = points.size();
        return tree;
    }
    /// kdtree.hpp:177-180
    static KDTree::Ptr build(const sycl_utils::DeviceQueue& q, const PointCloudShared& cloud,
                             size_t leaf_threshold = 16) {
        const auto& h = cloud.index_hint;
        if (h.matches(cloud.points->data(), cloud.points->size())) {  // fresh from the voxel grid: no measuring pass
            auto tree = std::make_shared<KDTree>(q);
            q.set_accessed_by_device(cloud.points->data(), cloud.points->size());
            detail::spx_check(spx_index_build_hinted(q.handle(), reinterpret_cast<const float*>(cloud.points->data()),
                                                     cloud.points->size(), h.lo, h.hi, h.cell, h.knn_cell, &tree->index_));
            tree->size_ = cloud.points->size();
            return tree;
        }
        return KDTree::build(q, *cloud.points, leaf_threshold);
    }

    /// kdtree.hpp:191-197 (MAX_K / MAX_DEPTH are compile-time capacities of the SYCL kernel; ignored)
    template <size_t MAX_K = 20, size_t MAX_DEPTH = 32>
    sycl_utils::events knn_search_async(const PointType* queries, const size_t query_size, const size_t k,
                                        KNNResult& result,
                                        const std::vector<sycl::event>& depends = std::vector<sycl::event>(),
                                        const TransformMatrix& transT = TransformMatrix::Identity()) const {
        return this->search(queries, query_size, k, result, depends, transT);
    }

    /// kdtree.hpp:203-224
    sycl_utils::events knn_search_async(const PointCloudShared& queries, const size_t k, KNNResult& result,
                                        const std::vector<sycl::event>& depends = std::vector<sycl::event>(),
                                        const TransformMatrix& transT = TransformMatrix::Identity()) const override {
        if (k > 100) throw std::runtime_error("[KDTree::knn_search_async] `k` is too large. not support.");
        return this->search(queries.points_ptr(), queries.size(), k, result, depends, transT);
    }

    /// kdtree.hpp:236-242 (MAX_K / MAX_DEPTH: compile-time capacities of the SYCL kernel; ignored)
    template <size_t MAX_K = 20, size_t MAX_DEPTH = 32>
    sycl_utils::events radius_search_async(const PointType* queries, const size_t query_size, const size_t max_k,
                                           const float radius, KNNResult& result,
                                           const std::vector<sycl::event>& depends = std::vector<sycl::event>(),
                                           const TransformMatrix& transT = TransformMatrix::Identity()) const {
        return this->radius(queries, query_size, max_k, radius, result, depends, transT);
    }

    /// kdtree.hpp:251-280: the max_k nearest targets within `radius`, the rest -1 / FLT_MAX
    sycl_utils::events radius_search_async(const PointCloudShared& queries, const size_t max_k, const float radius,
                                           KNNResult& result,
                                           const std::vector<sycl::event>& depends = std::vector<sycl::event>(),
                                           const TransformMatrix& transT = TransformMatrix::Identity()) const {
        if (max_k > 100) throw std::runtime_error("[KDTree::radius_search_async] `max_k` is too large. not support.");
        return this->radius(queries.points_ptr(), queries.size(), max_k, radius, result, depends, transT);
    }

    /// kdtree.hpp:282-284,721-760: drop the points whose flag is not INCLUDE_FLAG and re-number the others by
    /// `indices` (old -> new, -1 for the removed ones; as FilterByFlags::calculate_indices produces them)
    void remove_nodes_by_flags(const shared_vector<uint8_t>& flags, const shared_vector<int32_t>& indices) {
        if (flags.size() != indices.size())
            throw std::runtime_error("[KDTree::remove_nodes_by_flags_impl] flags and indices must have the same size.");
        if (index_ == nullptr || flags.empty()) return;
        size_t kept = 0;
        this->queue.set_accessed_by_host(indices.data(), indices.size());
        for (size_t i = 0; i < indices.size(); ++i)
            if (indices[i] >= 0) kept = std::max(kept, (size_t)indices[i] + 1);
        this->queue.set_accessed_by_device(flags.data(), flags.size());
        this->queue.set_accessed_by_device(indices.data(), indices.size());
        detail::spx_check(spx_index_remove_by_flags(index_, flags.data(), indices.data(), flags.size(), kept));
        detail::spx_check(spx_queue_sync(this->queue.handle()));
        size_ = kept;
    }

    spx_index_t handle() const { return index_; }
    size_t size() const { return size_; }

private:
    sycl_utils::events search(const PointType* queries, size_t query_size, size_t k, KNNResult& result,
                              const std::vector<sycl::event>& depends, const TransformMatrix& transT) const {
        for (const auto& e : depends) e.wait();
        if (result.indices == nullptr || result.distances == nullptr) {
            result.allocate(this->queue, query_size, k);  // kdtree.hpp:446-450
        } else {
            result.resize(query_size, k);
        }
        sycl_utils::events ev;
        if (query_size == 0) return ev;  // kdtree.hpp:429-436
        this->queue.set_accessed_by_device(queries, query_size);
        this->queue.set_accessed_by_device(result.indices->data(), query_size * k);
        this->queue.set_accessed_by_device(result.distances->data(), query_size * k);
        detail::spx_check(spx_index_knn(index_, reinterpret_cast<const float*>(queries), query_size, (int)k,
                                        transT.data(), result.indices->data(), result.distances->data()));
        ev += this->queue.checkpoint();
        return ev;
    }

    sycl_utils::events radius(const PointType* queries, size_t query_size, size_t max_k, float radius, KNNResult& result,
                              const std::vector<sycl::event>& depends, const TransformMatrix& transT) const {
        for (const auto& e : depends) e.wait();
        if (query_size == 0 || max_k == 0) {  // kdtree.hpp:580-587
            if (result.indices == nullptr || result.distances == nullptr) result.allocate(this->queue, 0, 0);
            else result.resize(0, 0);
            return sycl_utils::events();
        }
        if (result.indices == nullptr || result.distances == nullptr) result.allocate(this->queue, query_size, max_k);
        else result.resize(query_size, max_k);
        sycl_utils::events ev;
        this->queue.set_accessed_by_device(queries, query_size);
        this->queue.set_accessed_by_device(result.indices->data(), query_size * max_k);
        this->queue.set_accessed_by_device(result.distances->data(), query_size * max_k);
        detail::spx_check(spx_index_radius(index_, reinterpret_cast<const float*>(queries), query_size, (int)max_k, radius,
                                           transT.data(), result.indices->data(), result.distances->data()));
        ev += this->queue.checkpoint();
        return ev;
    }

    spx_index_t index_ = nullptr;
    size_t size_ = 0;
};

}  // namespace knn
}  // namespace algorithms
}  // namespace sycl_points
