// eigen_utils::lie — the host-side Lie-group helpers the registration wrappers call
// (I/utils/eigen_utils.hpp:909-943 se3_exp, :991-1034 se3_log), evaluated by libspx so that they agree bit for
// bit with what the device kernels compute.  The reference's device-side linear algebra helpers of the same header
// live inside the CUDA kernels here (csrc/spx_math.cuh) and have no host face.
#pragma once

#include <array>

#include "spx.h"
#include "sycl_points/points/types.hpp"
#include "sycl_points/utils/sycl_utils.hpp"

namespace sycl_points {
namespace eigen_utils {

/// rows of a 4x4 as four-lane vectors and back (I/utils/eigen_utils.hpp:693-722)
inline sycl::float4 to_sycl_vec(const Eigen::Vector4f& vec) { return {vec[0], vec[1], vec[2], vec[3]}; }
inline std::array<sycl::float4, 4> to_sycl_vec(const Eigen::Matrix4f& mat) {
    std::array<sycl::float4, 4> vecs;
    for (int i = 0; i < 4; ++i) vecs[i] = sycl::float4(mat(i, 0), mat(i, 1), mat(i, 2), mat(i, 3));
    return vecs;
}
inline Eigen::Matrix4f from_sycl_vec(const std::array<sycl::float4, 4>& vecs) {
    Eigen::Matrix4f mat;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) mat(i, j) = vecs[i][j];
    return mat;
}

namespace detail_spd {
/// log / exp of a symmetric 3x3 through its eigen-decomposition, evaluated by the library's device routine (the
/// same arithmetic the voxel map uses, csrc/spx_voxelmap.cu) on a process-wide helper queue
inline Eigen::Matrix3f spd_function(const Eigen::Matrix3f& A, bool is_log, float min_eigenvalue) {
    static spx_queue_t q = [] {
        spx_queue_t h = nullptr;
        detail::spx_check(spx_queue_create(sycl::default_selector_v(), &h));
        return h;
    }();
    void* d = nullptr;
    detail::spx_check(spx_malloc(q, 18 * sizeof(float), &d));
    float* dev = static_cast<float*>(d);
    Eigen::Matrix3f R;
    int rc = spx_memcpy_h2d(q, dev, A.data(), 9 * sizeof(float));
    if (rc == SPX_OK) rc = spx_spd_function(q, dev, 1, is_log ? 1 : 0, min_eigenvalue, dev + 9);
    if (rc == SPX_OK) rc = spx_memcpy_d2h(q, R.data(), dev + 9, 9 * sizeof(float));
    if (rc == SPX_OK) rc = spx_queue_sync(q);
    spx_free(q, d);
    detail::spx_check(rc);
    return R;
}
}  // namespace detail_spd

/// I/utils/eigen_utils.hpp:646-659
inline Eigen::Matrix3f log_spd_3x3(const Eigen::Matrix3f& A, const float min_eigenvalue = 1e-6f) {
    return detail_spd::spd_function(A, true, min_eigenvalue);
}
/// I/utils/eigen_utils.hpp:664-677
inline Eigen::Matrix3f exp_spd_3x3(const Eigen::Matrix3f& A) { return detail_spd::spd_function(A, false, 0.0f); }

namespace lie {

/// twist [rx ry rz tx ty tz] of a rigid transform (rotation first)
inline Eigen::Vector<float, 6> se3_log(const Eigen::Isometry3f& transform) {
    Eigen::Vector<float, 6> r;
    detail::spx_check(spx_se3_log(transform.data(), r.data()));
    return r;
}

inline Eigen::Matrix4f se3_exp(const Eigen::Vector<float, 6>& twist) {
    Eigen::Matrix4f T;
    detail::spx_check(spx_se3_exp(twist.data(), T.data()));
    return T;
}

}  // namespace lie
}  // namespace eigen_utils
}  // namespace sycl_points
