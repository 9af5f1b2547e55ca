// The few sycl:: names that appear in USER code written against sycl_points (a device, a queue,
// an event, the default selector), re-expressed over the spx C-ABI.  The reference's SYCL queue /
// USM layer (I/utils/sycl_utils.hpp) is replaced, not emulated: there is no handler, no
// parallel_for, no buffer — kernels live in libspx.so.
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "spx.h"

#ifndef SYCL_EXTERNAL
#define SYCL_EXTERNAL
#endif

namespace sycl_points {
namespace detail {
/// throw the exception type the reference throws for this failure class
inline void spx_check(int rc) {
    if (rc == SPX_OK) return;
    const std::string msg = spx_last_error();
    if (rc == SPX_ERR_INVALID_ARGUMENT) {
        if (msg.rfind("voxel_size", 0) == 0) throw std::invalid_argument(msg);  // voxel_downsampling.hpp:23-25
        throw std::runtime_error(msg);
    }
    throw std::runtime_error(msg);
}
}  // namespace detail
}  // namespace sycl_points

namespace sycl {

/// the exception user code catches around device set-up (sycl::exception): CUDA failures surface as
/// std::runtime_error from the C-ABI, so this type is never thrown — it only has to exist and be catchable
class exception : public std::runtime_error {
public:
    explicit exception(const std::string& what) : std::runtime_error(what) {}
};

/// sycl::vec<float, 4> as the reference's host helpers use it (to_sycl_vec / from_sycl_vec): four lanes with accessors
template <typename T, int N>
struct vec {
    T v[N] = {};
    vec() = default;
    vec(T a, T b, T c, T d) : v{a, b, c, d} { static_assert(N == 4, "four-lane constructor"); }
    T& x() { return v[0]; }
    T& y() { return v[1]; }
    T& z() { return v[2]; }
    T& w() { return v[3]; }
    const T& x() const { return v[0]; }
    const T& y() const { return v[1]; }
    const T& z() const { return v[2]; }
    const T& w() const { return v[3]; }
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
};
using float4 = vec<float, 4>;

namespace info {
namespace device {
struct name {};
struct vendor {};
}  // namespace device
}  // namespace info

/// selects a CUDA ordinal: SPX_DEVICE (env) or 0 — the role ONEAPI_DEVICE_SELECTOR plays for the reference
struct default_selector_t {
    int operator()() const {
        const char* e = std::getenv("SPX_DEVICE");
        return e ? std::atoi(e) : 0;
    }
};
inline constexpr default_selector_t default_selector_v{};
struct gpu_selector_t : default_selector_t {};
inline constexpr gpu_selector_t gpu_selector_v{};

class device {
public:
    device() = default;
    explicit device(int ordinal) : ordinal_(ordinal) {}
    template <typename Selector, typename = decltype(std::declval<Selector>()())>
    explicit device(const Selector& s) : ordinal_(s()) {}
    int ordinal() const { return ordinal_; }
    /// device.get_info<sycl::info::device::name>()
    template <typename Tag>
    std::string get_info() const {
        if constexpr (std::is_same_v<Tag, info::device::vendor>) return "NVIDIA Corporation";
        else return this->name();
    }
    bool is_gpu() const { return true; }
    bool is_cpu() const { return false; }
    std::string name() const {
        char buf[256] = {0};
        int sm = 0, smc = 0, l2 = 0;
        size_t mem = 0;
        sycl_points::detail::spx_check(spx_device_info(ordinal_, buf, &sm, &smc, &mem, &l2));
        return buf;
    }

private:
    int ordinal_ = 0;
};

/// in-order queue = one CUDA stream (spx_queue)
class queue {
public:
    explicit queue(const device& d) : dev_(d) {
        spx_queue_t h = nullptr;
        sycl_points::detail::spx_check(spx_queue_create(d.ordinal(), &h));
        h_ = std::shared_ptr<spx_queue_s>(h, [](spx_queue_t p) { spx_queue_destroy(p); });
    }
    spx_queue_t handle() const { return h_.get(); }
    const device& get_device() const { return dev_; }
    void wait() const { sycl_points::detail::spx_check(spx_queue_sync(h_.get())); }
    void wait_and_throw() const { wait(); }
    bool operator==(const queue& o) const { return h_ == o.h_; }

private:
    device dev_;
    std::shared_ptr<spx_queue_s> h_;
};

/// completion of everything enqueued on its queue up to the point the event was made (the queue
/// is in order, so waiting on the stream is exact)
class event {
public:
    event() = default;
    explicit event(const queue& q) : q_(std::make_shared<queue>(q)) {}
    void wait() const {
        if (q_) q_->wait();
    }
    void wait_and_throw() const { wait(); }

private:
    std::shared_ptr<queue> q_;
};

}  // namespace sycl
