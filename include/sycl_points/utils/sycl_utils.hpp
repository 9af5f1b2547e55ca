// sycl_points::sycl_utils over CUDA — replaces I/utils/sycl_utils.hpp (SYCL queue + USM).
//   DeviceQueue   :491-626  -> one in-order CUDA stream (spx_queue) shared by copies of the value
//   events        :234-280  -> list of stream-completion events (+ keep-alive resources)
//   shared_vector :630-635  -> std::vector over CUDA *managed* memory: host-dereferenceable and
//                              device-usable like USM shared; mem_advise hints become prefetches
#pragma once

#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <iterator>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "sycl_points/compat/sycl_shim.hpp"

namespace sycl_points {

namespace sycl_utils {

namespace device_selector {
inline constexpr auto default_selector_v = sycl::default_selector_v;
/// any CUDA device libspx can open is supported (sycl_utils.hpp:398-465 filters SYCL backends)
inline bool is_supported_device(const sycl::device& d) {
    int n = 0;
    return spx_device_count(&n) == SPX_OK && d.ordinal() >= 0 && d.ordinal() < n;
}
inline sycl::device select_device(const std::string& /*vendor*/ = "nvidia", const std::string& /*type*/ = "gpu") {
    return sycl::device(sycl::default_selector_v);
}
}  // namespace device_selector

/// sycl_utils::events — sycl_utils.hpp:234-280
struct events {
    std::vector<sycl::event> evs;
    std::vector<std::shared_ptr<void>> keep_alive;

    void push_back(const sycl::event& event) { this->evs.push_back(event); }
    void wait() {
        for (auto& e : this->evs) e.wait();
        this->clear();
    }
    void wait_and_throw() {
        for (auto& e : this->evs) e.wait_and_throw();
        this->clear();
    }
    void clear() {
        this->evs.clear();
        this->keep_alive.clear();
    }
    void operator+=(const sycl::event& event) { this->evs.push_back(event); }
    void operator+=(const events& e) {
        std::copy(e.evs.begin(), e.evs.end(), std::back_inserter(this->evs));
        std::copy(e.keep_alive.begin(), e.keep_alive.end(), std::back_inserter(this->keep_alive));
    }
    template <typename T>
    void add_resource(const std::shared_ptr<T>& resource) {
        this->keep_alive.emplace_back(resource);
    }
};

inline void print_device_info(const sycl::queue& q) {
    char name[256] = {0};
    int sm = 0, smc = 0, l2 = 0;
    size_t mem = 0;
    detail::spx_check(spx_device_info(q.get_device().ordinal(), name, &sm, &smc, &mem, &l2));
    std::cout << "Device: " << name << " [CUDA sm_" << sm << "]\n"
              << "  multiprocessors: " << smc << "\n"
              << "  global memory: " << (mem >> 20) << " MiB\n"
              << "  L2 cache: " << (l2 >> 20) << " MiB" << std::endl;
}

/// sycl_utils::DeviceQueue — sycl_utils.hpp:491-626
class DeviceQueue {
public:
    using Ptr = std::shared_ptr<DeviceQueue>;
    std::shared_ptr<sycl::queue> ptr = nullptr;

    DeviceQueue(const sycl::device& device) {
        if (!device_selector::is_supported_device(device))
            throw std::runtime_error("[DeviceQueue::DeviceQueue] device ordinal " + std::to_string(device.ordinal()) +
                                     " is not supported.");
        this->ptr = std::make_shared<sycl::queue>(device);
    }

    spx_queue_t handle() const { return this->ptr->handle(); }
    void print_device_info() const { sycl_utils::print_device_info(*this->ptr); }
    sycl::device get_device() const { return this->ptr->get_device(); }
    bool is_cpu() const { return false; }
    bool is_gpu() const { return true; }
    bool is_intel() const { return false; }
    bool is_nvidia() const { return true; }
    bool is_supported_double() const { return true; }
    /// work-group sizing is the library's business on CUDA; kept for source compatibility
    size_t get_work_group_size() const { return 256; }
    void set_work_group_size(size_t) {}
    size_t get_work_group_size_for_parallel_reduction() const { return 256; }
    void set_work_group_size_for_parallel_reduction(size_t) {}
    size_t get_global_size(size_t N) const { return (N + 255) / 256 * 256; }
    size_t get_global_size_for_parallel_reduction(size_t N) const { return get_global_size(N); }

    /// mem_advise hints (sycl_utils.hpp:283-364) -> prefetch of managed memory on the queue's stream
    template <typename T>
    void set_accessed_by_device(T* data_ptr, size_t N) const {
        spx_prefetch(this->handle(), data_ptr, N * sizeof(T), 1);
    }
    template <typename T>
    void clear_accessed_by_device(T*, size_t) const {}
    template <typename T>
    void set_accessed_by_host(T* data_ptr, size_t N) const {
        spx_prefetch(this->handle(), data_ptr, N * sizeof(T), 0);
        spx_queue_sync(this->handle());
    }
    template <typename T>
    void clear_accessed_by_host(T*, size_t) const {}
    template <typename T>
    void set_read_mostly(T*, size_t) const {}
    template <typename T>
    void clear_read_mostly(T*, size_t) const {}

    /// an event that completes when everything enqueued so far has run
    sycl::event checkpoint() const { return sycl::event(*this->ptr); }
};

inline bool is_cpu(const sycl::queue&) { return false; }
inline bool is_gpu(const sycl::queue&) { return true; }
inline bool is_nvidia(const sycl::queue&) { return true; }

}  // namespace sycl_utils

/// allocator of CUDA managed memory; constructible from a sycl::queue like sycl::usm_allocator.
/// Value-initialisation of trivially constructible elements is skipped (resize() of an output array
/// must not touch — and thereby migrate to the host — memory a kernel is about to fill).
template <typename T, size_t Alignment = 0>
struct shared_allocator {
    using value_type = T;
    shared_allocator() = default;
    shared_allocator(const sycl::queue&) {}
    template <typename U, size_t A>
    shared_allocator(const shared_allocator<U, A>&) {}
    template <typename U>
    struct rebind {
        using other = shared_allocator<U, Alignment>;
    };
    T* allocate(size_t n) {
        void* p = nullptr;
        if (n == 0) return nullptr;
        if (spx_malloc_managed(n * sizeof(T), &p) != SPX_OK || !p) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t) { spx_free_managed(p); }
    template <typename U>
    void construct(U* p) {
        if constexpr (!std::is_trivially_default_constructible_v<U>) ::new (static_cast<void*>(p)) U();
    }
    template <typename U, typename... Args>
    void construct(U* p, Args&&... args) {
        ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
    }
    template <typename U, size_t A>
    bool operator==(const shared_allocator<U, A>&) const { return true; }
    template <typename U, size_t A>
    bool operator!=(const shared_allocator<U, A>&) const { return false; }
};
template <typename T>
using shared_vector = std::vector<T, shared_allocator<T, alignof(T)>>;
template <typename T>
using shared_vector_ptr = std::shared_ptr<shared_vector<T>>;

}  // namespace sycl_points
