// filter::FilterByFlags — I/algorithms/common/filter_by_flags.hpp:11-100.  Host-side, order-preserving compaction
// of a shared (managed-memory) container by per-element flags, and the old -> new index table; host code in the
// reference too.  (The device filters of this library compact through spx_gather: preprocess_filter.hpp.)
#pragma once

#include <cstdint>
#include <memory>

#include "sycl_points/utils/sycl_utils.hpp"

namespace sycl_points {
namespace algorithms {
namespace filter {

constexpr uint8_t REMOVE_FLAG = 0;
constexpr uint8_t INCLUDE_FLAG = 1;

class FilterByFlags {
public:
    using Ptr = std::shared_ptr<FilterByFlags>;

    FilterByFlags(const sycl_utils::DeviceQueue& queue) : queue_(queue) {}

    /// elements whose flag is INCLUDE_FLAG, in source order (`output` may be `source`)
    template <typename T>
    void filter_by_flags(const shared_vector<T>& source, shared_vector<T>& output,
                         const shared_vector<uint8_t>& flags) const {
        const size_t N = source.size();
        if (N == 0) return;
        output.resize(N);
        this->queue_.set_accessed_by_host(source.data(), N);
        this->queue_.set_accessed_by_host(output.data(), N);
        this->queue_.set_accessed_by_host(flags.data(), N);
        size_t kept = 0;
        for (size_t i = 0; i < N; ++i)
            if (flags[i] == INCLUDE_FLAG) output[kept++] = source[i];
        output.resize(kept);
    }

    template <typename T>
    void filter_by_flags(shared_vector<T>& data, const shared_vector<uint8_t>& flags) const {
        this->filter_by_flags(data, data, flags);
    }

    /// indices[i] = position of element i after the compaction, -1 when it is removed
    void calculate_indices(const shared_vector<uint8_t>& flags, shared_vector<int32_t>& indices) const {
        const size_t N = flags.size();
        if (N == 0) return;
        indices.resize(N);
        this->queue_.set_accessed_by_host(flags.data(), N);
        this->queue_.set_accessed_by_host(indices.data(), N);
        int32_t count = 0;
        for (size_t i = 0; i < N; ++i) indices[i] = (flags[i] == INCLUDE_FLAG) ? count++ : -1;
    }

private:
    sycl_utils::DeviceQueue queue_;
};

}  // namespace filter
}  // namespace algorithms
}  // namespace sycl_points
